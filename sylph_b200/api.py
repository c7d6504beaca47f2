"""Host-side mirror of the reference's operator interface for the two hot paths, over the C ABI.

Names follow the reference (extract_markers, sketch_genome, sketch_sequences, get_stats/query/
profile); the only change is granularity: every call takes a *batch* (flat base buffer + record
offsets) because a GPU call per record / per pair would be pure launch overhead.

Inputs may be numpy arrays (host memory, copied inside the call) or torch CUDA tensors (device
memory, zero-copy).  torch is plumbing only (device buffers / streams); it is imported lazily.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import SEM_AVX2, SEM_SCALAR, AniRow, ContainParams, Survivor, SylphError  # noqa: F401

SURVIVOR_DTYPE = np.dtype([("hash", "<u8"), ("rec", "<u4"), ("pos", "<u4")])
ANI_ROW_DTYPE = np.dtype([
    ("sample", "<u4"), ("genome", "<u4"), ("lambda_status", "<u4"), ("ci_valid", "<u4"),
    ("contain", "<u8"), ("glen", "<u8"), ("kmers_lost", "<i8"), ("naive_ani", "<f8"),
    ("final_est_ani", "<f8"), ("final_est_cov", "<f8"), ("mean_cov", "<f8"), ("median_cov", "<f8"),
    ("lambda", "<f8"), ("ci", "<f8", (4,)), ("rel_abund", "<f8"), ("seq_abund", "<f8"), ("reserved", "<f8")])
assert ANI_ROW_DTYPE.itemsize == 144


def _is_torch(x):
    return type(x).__module__.startswith("torch")


def _arg(x, dtype):
    """-> (mem, pointer, n, keepalive)"""
    if x is None:
        return None, None, 0, None
    if _is_torch(x):
        if not x.is_cuda:
            x = x.numpy()
        else:
            import torch
            want = {np.uint8: torch.uint8, np.uint64: torch.int64, np.uint32: torch.int32}[dtype]
            if x.dtype not in (want, getattr(torch, np.dtype(dtype).name, want)):
                raise TypeError("expected tensor of %s-compatible dtype, got %s" % (np.dtype(dtype).name, x.dtype))
            if not x.is_contiguous():
                x = x.contiguous()
            return _lib.MEM_DEVICE, C.c_void_p(x.data_ptr()), x.numel(), x
    a = np.ascontiguousarray(x, dtype=dtype)
    return _lib.MEM_HOST, a.ctypes.data_as(C.c_void_p), a.size, a


class Context:
    """One CUDA device + one stream (syl_ctx)."""

    def __init__(self, device=0, stream=None):
        self._h = C.c_void_p()
        L = _lib.lib()
        _lib.check(L.syl_ctx_create(int(device), C.c_void_p(stream) if stream else None, C.byref(self._h)))
        self.device = device

    def close(self):
        if self._h:
            _lib.lib().syl_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        _lib.check(_lib.lib().syl_ctx_sync(self._h))

    @property
    def launches(self):
        return int(_lib.lib().syl_ctx_launch_count(self._h))

    # ---- (1) seeding -------------------------------------------------------------------------
    def extract_markers_batch(self, bases, rec_off, k=31, c=200, sem=SEM_AVX2, with_pos=False, cap=None,
                              out=None):
        """Batched extract_markers / extract_markers_positions (src/sketch.rs:53-93).
        Returns a numpy structured array (hash, rec, pos) in unspecified order, or — when `out`
        is a torch CUDA tensor of >= cap*16 bytes — the survivor count (survivors stay on device)."""
        L = _lib.lib()
        mem_b, pb, nb, kb = _arg(bases, np.uint8)
        mem_o, po, no, ko = _arg(rec_off, np.uint64)
        if mem_b != mem_o:
            raise ValueError("bases and rec_off must live in the same memory space")
        n_rec = no - 1
        n_out = C.c_uint64(0)
        if out is not None:
            assert mem_b == _lib.MEM_DEVICE and _is_torch(out)
            cap = out.numel() * out.element_size() // 16
            _lib.check(L.syl_seed_batch(self._h, mem_b, pb, nb, po, n_rec, k, c, sem, int(with_pos),
                                        C.c_void_p(out.data_ptr()), cap, C.byref(n_out)))
            return n_out.value
        if cap is None:
            cap = max(1024, int(nb / c * 1.3) + 4096)
        while True:
            if mem_b == _lib.MEM_DEVICE:
                import torch
                dbuf = torch.empty(cap * 2, dtype=torch.int64, device=bases.device)
                pout = C.c_void_p(dbuf.data_ptr())
            else:
                hbuf = np.empty(cap, dtype=SURVIVOR_DTYPE)
                pout = hbuf.ctypes.data_as(C.c_void_p)
            rc = L.syl_seed_batch(self._h, mem_b, pb, nb, po, n_rec, k, c, sem, int(with_pos), pout, cap,
                                  C.byref(n_out))
            if rc == _lib.SYL_ERR_CAPACITY:
                cap = n_out.value + 16
                continue
            _lib.check(rc)
            n = n_out.value
            if mem_b == _lib.MEM_DEVICE:
                return dbuf[: 2 * n].cpu().numpy().view(SURVIVOR_DTYPE).copy()
            return hbuf[:n].copy()
