"""Cross-seed donor pool for `fo` (BASELINE config C5): the one exchange step of the hot path.

Every GPU samples D windows (<= stride bytes) from its own corpus shard (eb200_sample_donors, or `sample_windows` below for
host data), the per-GPU pools are all-gathered (NCCL over NVLink on the GPU box, gloo in the CPU tests) and each case's
sed_fuse_old starts out remembering one of the world_size * D windows instead of its own block
(reference hook: remember/1, src/erlamsa_mutations.erl:404-427; pool semantics: include/erlamsa_b200.h)."""

DEFAULT_DONORS = 4096
DEFAULT_STRIDE = 2048   # the reference's re-chunk size (src/erlamsa.hrl:48)


def window_of(d, n_blobs, n_donors, blob_len, stride):
    """-> (blob index, start, length) of donor window d: the rule of eb_sample_donors_kernel"""
    b = d * n_blobs // n_donors
    wlen = min(blob_len(b), stride)
    start = ((d * 2654435761) & 0xFFFFFFFF) % (blob_len(b) - wlen + 1)
    return b, start, wlen


def sample_windows(blobs, n_donors, stride=DEFAULT_STRIDE):
    """host restatement of the sampling kernel (tests, CPU legs): list of n_donors byte strings"""
    out = []
    for d in range(n_donors):
        b, start, wlen = window_of(d, len(blobs), n_donors, lambda i: len(blobs[i]), stride)
        out.append(blobs[b][start:start + wlen])
    return out


def all_gather_pool(pool, lens):
    """pool: uint8 tensor [D, stride], lens: int32 tensor [D] (same device) -> ([W*D, stride], [W*D]) in rank order.
    One collective each; with NCCL the bytes cross NVLink (8 x 4096 x 2048 B = 64 MiB per batch at 8 GPUs)."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return pool, lens
    w = dist.get_world_size()
    gp = torch.empty((w * pool.shape[0], pool.shape[1]), dtype=pool.dtype, device=pool.device)
    gl = torch.empty((w * lens.shape[0],), dtype=lens.dtype, device=lens.device)
    dist.all_gather_into_tensor(gp, pool.contiguous())
    dist.all_gather_into_tensor(gl, lens.contiguous())
    return gp, gl
