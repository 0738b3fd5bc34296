"""Multi-GPU sharding of the case loop. Cases are independent (each draws its own seed from the parent stream by
index -- reference src/erlamsa_main.erl:179 -- and the reference's own --workers splits the range the same way,
src/erlamsa_main.erl:90-108), so a corpus is cut into contiguous windows of case ids, one per rank, with no
data-path collective. Results are identical for any GPU count because the RNG is keyed by the global case id."""


def shard_window(n_cases, rank, world, first_case=1):
    """-> (first_case_of_rank, n_cases_of_rank): contiguous, balanced to within one case."""
    base, rem = divmod(n_cases, world)
    cnt = base + (1 if rank < rem else 0)
    start = rank * base + min(rank, rem)
    return first_case + start, cnt
