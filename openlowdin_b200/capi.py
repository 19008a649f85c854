"""ctypes binding of include/lowdin_it.h (liblowdin_itgpu.so).  No fallback: a missing library raises."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
CONV_C, CONV_E = 0, 1
GEN_HASH = 1

_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
_f64pf = np.ctypeslib.ndpointer(np.float64, flags="F_CONTIGUOUS")

# every symbol include/lowdin_it.h and include/lowdin_it_host.h declare (checked by tests/test_abi.py)
ABI_SYMBOLS = [
    "lowdin_it_create", "lowdin_it_destroy", "lowdin_it_last_error", "lowdin_it_set_species", "lowdin_it_ao_begin",
    "lowdin_it_ao_push_stacks", "lowdin_it_ao_end", "lowdin_it_ao_set_generator", "lowdin_it_transform",
    "lowdin_it_result_count", "lowdin_it_download_pairs", "lowdin_it_download_quads", "lowdin_it_transform_stream",
    "lowdin_it_stream_num_passes", "lowdin_it_transform_all", "lowdin_it_transform_inter_all",
    "lowdin_it_comm_unique_id", "lowdin_it_comm_init", "lowdin_it_timers", "lowdin_it_kernel_bench",
    "lowdin_it_set_profiling", "lowdin_it_set_option", "lowdin_it_kernel_stats", "lowdin_it_debug_gemm", "lowdin_it_debug_expand",
]


class LowdinITError(RuntimeError):
    pass


def lib_path() -> str:
    return os.path.join(HERE, "liblowdin_itgpu.so")


_lib = None


def load():
    """Load liblowdin_itgpu.so.  Raises (never falls back) when the CUDA library has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    p = lib_path()
    if not os.path.exists(p):
        raise LowdinITError(f"{p} not built: run `python -m openlowdin_b200.build` (nvcc, sm_100a). "
                            "There is no CPU fallback.")
    L = C.CDLL(p)
    H = C.c_void_p
    L.lowdin_it_create.argtypes = [C.c_int, C.POINTER(H)]
    L.lowdin_it_destroy.argtypes = [H]
    L.lowdin_it_last_error.argtypes = [H]
    L.lowdin_it_last_error.restype = C.c_char_p
    L.lowdin_it_set_species.argtypes = [H, C.c_int, C.c_int, _f64pf, C.c_int, C.c_int]
    L.lowdin_it_ao_begin.argtypes = [H, C.c_int, C.c_int, C.c_int]
    L.lowdin_it_ao_push_stacks.argtypes = [H, _i32p, _i32p, _i32p, _i32p, _f64p, C.c_int64]
    L.lowdin_it_ao_end.argtypes = [H]
    L.lowdin_it_ao_set_generator.argtypes = [H, C.c_int, C.c_int, C.c_int, C.c_uint64]
    L.lowdin_it_transform.argtypes = [H, C.c_int, C.c_int, _i32p, C.c_int, C.c_int, C.c_double]
    L.lowdin_it_result_count.argtypes = [H, C.POINTER(C.c_int64)]
    L.lowdin_it_download_pairs.argtypes = [H, _i64p, _i64p, _f64p]
    L.lowdin_it_download_quads.argtypes = [H, _i32p, _i32p, _i32p, _i32p, _f64p]
    L.lowdin_it_transform_stream.argtypes = [H, C.c_int, C.c_int, _i32p, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int,
                                             C.c_void_p, C.c_void_p, C.c_double, _f64p]
    L.lowdin_it_stream_num_passes.argtypes = [H, C.c_int, C.c_int, _i32p, C.c_int, C.c_int, C.POINTER(C.c_int),
                                              C.POINTER(C.c_int)]
    L.lowdin_it_transform_all.argtypes = [_f64pf, _f64p, C.c_int]
    L.lowdin_it_transform_inter_all.argtypes = [_f64pf, _f64pf, _f64p, C.c_int, C.c_int]
    L.lowdin_it_comm_unique_id.argtypes = [C.c_char_p]
    L.lowdin_it_comm_init.argtypes = [H, C.c_int, C.c_int, C.c_char_p]
    L.lowdin_it_timers.argtypes = [H, _f64p]
    L.lowdin_it_kernel_bench.argtypes = [H, C.c_int, C.c_int64, C.c_int64, C.c_int64, C.c_int, C.POINTER(C.c_double),
                                         C.POINTER(C.c_double)]
    L.lowdin_it_set_profiling.argtypes = [H, C.c_int]
    L.lowdin_it_set_option.argtypes = [H, C.c_int, C.c_int64]
    L.lowdin_it_kernel_stats.argtypes = [H, _f64p, _f64p, _f64p]
    L.lowdin_it_debug_gemm.argtypes = [H, _f64p, _f64p, _f64p, C.c_int, C.c_int, C.c_int]
    L.lowdin_it_debug_expand.argtypes = [H, C.c_int, C.c_int, C.c_int64, C.c_int, _f64p]
    _lib = L
    return L


class Transformer:
    """One device context (lowdin_it_handle)."""

    def __init__(self, device: int = 0):
        self.L = load()
        self.h = C.c_void_p()
        if self.L.lowdin_it_create(device, C.byref(self.h)):
            raise LowdinITError(self.L.lowdin_it_last_error(None).decode())
        self.n = {}

    def close(self):
        if self.h:
            self.L.lowdin_it_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc:
            raise LowdinITError(self.L.lowdin_it_last_error(self.h).decode())

    def set_species(self, slot, Cm):
        Cm = np.asfortranarray(Cm, dtype=np.float64)
        self._ck(self.L.lowdin_it_set_species(self.h, slot, Cm.shape[0], Cm, Cm.shape[0], Cm.shape[1]))
        self.n[slot] = Cm.shape[0]

    def upload_ao(self, a, b, p, q, r, s, v, swapped=False, stack=30000):
        """Push the AO list in stacks, the last one carrying the p=-1 terminator like a .ints file."""
        self._ck(self.L.lowdin_it_ao_begin(self.h, a, b, int(swapped)))
        n = len(v)
        for o in range(0, max(n, 1), stack):
            e = min(n, o + stack)
            pp = np.zeros(stack, np.int32); qq = np.zeros(stack, np.int32)
            rr = np.zeros(stack, np.int32); ss = np.zeros(stack, np.int32); vv = np.zeros(stack)
            pp[:e - o] = p[o:e]; qq[:e - o] = q[o:e]; rr[:e - o] = r[o:e]; ss[:e - o] = s[o:e]; vv[:e - o] = v[o:e]
            if e - o < stack:
                pp[e - o] = -1
            self._ck(self.L.lowdin_it_ao_push_stacks(self.h, pp, qq, rr, ss, vv, stack))
        self._ck(self.L.lowdin_it_ao_end(self.h))

    def set_generator(self, a, b, seed, kind=GEN_HASH):
        self._ck(self.L.lowdin_it_ao_set_generator(self.h, a, b, kind, seed))

    def transform(self, a, b, win, conv, symmetric=False, tol=1e-10):
        w = np.ascontiguousarray(win, dtype=np.int32)
        self._ck(self.L.lowdin_it_transform(self.h, a, b, w, conv, int(symmetric), tol))
        cnt = C.c_int64()
        self._ck(self.L.lowdin_it_result_count(self.h, C.byref(cnt)))
        n = cnt.value
        v = np.zeros(n)
        if conv == CONV_E:
            ij = np.zeros(n, np.int64); kl = np.zeros(n, np.int64)
            self._ck(self.L.lowdin_it_download_pairs(self.h, ij, kl, v))
            return ij, kl, v
        o = [np.zeros(n, np.int32) for _ in range(4)]
        self._ck(self.L.lowdin_it_download_quads(self.h, *o, v))
        return (*o, v)

    def num_passes(self, a, b, win, conv, occ_batch=0):
        w = np.ascontiguousarray(win, dtype=np.int32)
        npass, used = C.c_int(), C.c_int()
        self._ck(self.L.lowdin_it_stream_num_passes(self.h, a, b, w, conv, occ_batch, C.byref(npass), C.byref(used)))
        return npass.value, used.value

    def transform_stream(self, a, b, win, conv, tol=1e-10, occ_batch=0, first_pass=0, n_passes=0, epsA=None, epsB=None,
                         lam=2.0):
        w = np.ascontiguousarray(win, dtype=np.int32)
        sums = np.zeros(4)
        ea = np.ascontiguousarray(epsA, dtype=np.float64) if epsA is not None else None
        eb = np.ascontiguousarray(epsB, dtype=np.float64) if epsB is not None else None
        self._ck(self.L.lowdin_it_transform_stream(
            self.h, a, b, w, conv, tol, occ_batch, first_pass, n_passes,
            ea.ctypes.data if ea is not None else None, eb.ctypes.data if eb is not None else None, lam, sums))
        return sums

    def comm_init(self, rank, nranks, uid: bytes):
        self._ck(self.L.lowdin_it_comm_init(self.h, rank, nranks, uid))

    def timers(self):
        t = np.zeros(8)
        self._ck(self.L.lowdin_it_timers(self.h, t))
        return dict(ao_upload=t[0], first_half=t[1], exchange=t[2], second_half=t[3], consume=t[4], download=t[5],
                    flops=t[6], launches=int(t[7]))

    CATEGORIES = ("expand1", "q1", "q2", "expand2", "q3", "q4", "consume", "exchange")

    OPT_WORKSPACE_BYTES, OPT_CHUNK_COLS = 1, 2

    def set_option(self, option, value):
        self._ck(self.L.lowdin_it_set_option(self.h, option, int(value)))

    def set_profiling(self, on=True):
        self._ck(self.L.lowdin_it_set_profiling(self.h, int(on)))

    def kernel_stats(self):
        ms, cnt, work = np.zeros(8), np.zeros(8), np.zeros(8)
        self._ck(self.L.lowdin_it_kernel_stats(self.h, ms, cnt, work))
        return {c: dict(ms=ms[i], launches=int(cnt[i]), work=work[i]) for i, c in enumerate(self.CATEGORIES)}

    def kernel_bench(self, kind, m, n, k, iters=10):
        ms, chk = C.c_double(), C.c_double()
        self._ck(self.L.lowdin_it_kernel_bench(self.h, kind, m, n, k, iters, C.byref(ms), C.byref(chk)))
        return ms.value, chk.value

    def debug_gemm(self, A, B):
        A = np.ascontiguousarray(A); B = np.ascontiguousarray(B)
        out = np.zeros((A.shape[0], B.shape[0]))
        self._ck(self.L.lowdin_it_debug_gemm(self.h, A, B, out, A.shape[0], B.shape[0], A.shape[1]))
        return out

    def debug_expand(self, a, b, slab0, nb):
        n = self.n[a]
        X = np.zeros((nb, n, n))
        self._ck(self.L.lowdin_it_debug_expand(self.h, a, b, slab0, nb, X))
        return X


def unique_id() -> bytes:
    buf = C.create_string_buffer(128)
    if load().lowdin_it_comm_unique_id(buf):
        raise LowdinITError(load().lowdin_it_last_error(None).decode())
    return buf.raw


def transform_all(coeff, ints):
    """Drop-in for c_integrals_transform_all (IntTransfD.h:66): in-place on `ints` (D packing)."""
    L = load()
    coeff = np.asfortranarray(coeff, dtype=np.float64)
    if L.lowdin_it_transform_all(coeff, ints, coeff.shape[0]):
        raise LowdinITError("lowdin_it_transform_all failed: " + L.lowdin_it_last_error(None).decode())
    return ints


def transform_inter_all(coeff, ocoeff, ints):
    """Drop-in for c_integrals_transform_inter_all (IntTransfD.h:68)."""
    L = load()
    coeff = np.asfortranarray(coeff, dtype=np.float64)
    ocoeff = np.asfortranarray(ocoeff, dtype=np.float64)
    if L.lowdin_it_transform_inter_all(coeff, ocoeff, ints, coeff.shape[0], ocoeff.shape[0]):
        raise LowdinITError("lowdin_it_transform_inter_all failed: " + L.lowdin_it_last_error(None).decode())
    return ints
