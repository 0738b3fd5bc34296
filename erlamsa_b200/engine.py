"""Engine: one native context (one CUDA device) + batch submission from Python bytes or device tensors."""
import ctypes as C

from . import _native as N
from .options import make_opts


class EngineError(RuntimeError):
    def __init__(self, code, detail=""):
        self.code = code
        msg = N.lib().eb200_strerror(code).decode()
        super().__init__("erlamsa_b200: %s (%d)%s" % (msg, code, (": " + detail) if detail else ""))


class Engine:
    def __init__(self, device=0):
        self._ctx = C.c_void_p()
        rc = N.lib().eb200_init(device, C.byref(self._ctx))
        if rc != 0:
            self._ctx = None
            raise EngineError(rc)
        self.device = device
        self.last_stats = None

    def close(self):
        if self._ctx:
            N.lib().eb200_shutdown(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _err(self, rc):
        detail = N.lib().eb200_last_cuda_error(self._ctx).decode() if rc == -1 else ""
        return EngineError(rc, detail)

    def fuzz_batch(self, blobs, opts=None, n_cases=None, want_meta=True):
        """Host path (what the NIF does): list of bytes in, list of bytes out.
        Case I = first_case + k mutates blobs[(I-1) % len(blobs)]."""
        o = opts if isinstance(opts, N.Opts) else make_opts(opts)
        if n_cases is None:
            n_cases = len(blobs)
        data = b"".join(blobs)
        off = (C.c_uint64 * (len(blobs) + 1))()
        acc = 0
        for i, b in enumerate(blobs):
            off[i] = acc
            acc += len(b)
        off[len(blobs)] = acc
        buf = C.create_string_buffer(data, len(data) + 1)
        out_p = C.c_void_p()
        out_off = (C.c_uint64 * (n_cases + 1))()
        out_len = (C.c_uint64 * max(n_cases, 1))()
        meta = (N.Meta * max(n_cases, 1))() if want_meta else None
        st = N.Stats()
        rc = N.lib().eb200_fuzz_batch(self._ctx, C.byref(o), C.cast(buf, C.c_void_p), off, len(blobs), n_cases,
                                      C.byref(out_p), out_off, out_len, meta, C.byref(st))
        if rc != 0:
            raise self._err(rc)
        self.last_stats = st
        try:
            total = max([out_off[k] + out_len[k] for k in range(n_cases)] + [0])
            # (ctypes.string_at takes a C int size: not usable past 2 GiB)
            raw = bytes((C.c_char * total).from_address(out_p.value)) if total else b""
        finally:
            N.lib().eb200_free(out_p)
        outs = [raw[out_off[k]:out_off[k] + out_len[k]] for k in range(n_cases)]
        return outs, (list(meta)[:n_cases] if want_meta else None)

    def sample_donors(self, d_data, d_off, n_blobs, n_donors, stride, d_pool, d_len, stream=0):
        """Config C5: fill a donor pool (device addresses) with n_donors windows of the device-resident corpus."""
        rc = N.lib().eb200_sample_donors(self._ctx, d_data, d_off, n_blobs, n_donors, stride, d_pool, d_len, stream or None)
        if rc != 0:
            raise self._err(rc)

    def fuzz_batch_device(self, opts, d_data, d_off, n_blobs, data_bytes, n_cases, d_out, out_capacity,
                          d_out_off, d_out_len, d_meta=0, stream=0):
        """Device path: all arguments are raw device addresses (ints), e.g. torch tensors' data_ptr()."""
        o = opts if isinstance(opts, N.Opts) else make_opts(opts)
        st = N.Stats()
        rc = N.lib().eb200_fuzz_batch_device(self._ctx, C.byref(o), d_data, d_off, n_blobs, data_bytes, n_cases,
                                             d_out, out_capacity, d_out_off, d_out_len, d_meta or None,
                                             stream or None, C.byref(st))
        if rc != 0:
            raise self._err(rc)
        self.last_stats = st
        return st

    def submit_device(self, opts, d_data, d_off, n_blobs, data_bytes, n_cases, d_out, out_capacity, d_out_off, d_out_len, d_meta=0):
        """Asynchronous device path (eb200_submit_device): queues the batch on one of the context's lanes and returns a ticket
        at once. Every buffer must stay alive, and must not be shared with another batch in flight, until collect(ticket)."""
        o = opts if isinstance(opts, N.Opts) else make_opts(opts)
        t = C.c_void_p()
        rc = N.lib().eb200_submit_device(self._ctx, C.byref(o), d_data, d_off, n_blobs, data_bytes, n_cases,
                                         d_out, out_capacity, d_out_off, d_out_len, d_meta or None, C.byref(t))
        if rc != 0:
            raise self._err(rc)
        return t

    def collect(self, ticket):
        """Blocks until the batch behind `ticket` is complete; returns its stats (raises what the synchronous call would have)."""
        st = N.Stats()
        rc = N.lib().eb200_collect(self._ctx, ticket, C.byref(st))
        if rc != 0:
            raise self._err(rc)
        self.last_stats = st
        return st
