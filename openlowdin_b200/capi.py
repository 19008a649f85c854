"""ctypes binding of include/lowdin_it.h (liblowdin_itgpu.so).  No fallback: a missing library raises."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
CONV_C, CONV_E = 0, 1
GEN_HASH, GEN_FOLD, GEN_RANKK = 1, 2, 3

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
    "lowdin_it_comm_unique_id", "lowdin_it_comm_init", "lowdin_it_shard_plan", "lowdin_it_exchanged_offset", "lowdin_it_slab_owner", "lowdin_it_slab_local", "lowdin_it_slab_global", "lowdin_it_timers", "lowdin_it_kernel_bench",
    "lowdin_it_set_profiling", "lowdin_it_set_option", "lowdin_it_kernel_stats", "lowdin_it_debug_gemm", "lowdin_it_debug_expand",
    "lowdin_it_ao_push_blocks", "lowdin_it_ao_set_rankk", "lowdin_it_ao_materialize", "lowdin_it_comm_init_local",
    "lowdin_it_debug_first_half", "lowdin_it_debug_first_quarter", "lowdin_it_result_segments", "lowdin_it_group_transform",
    "lowdin_it_group_result_count", "lowdin_it_group_download_pairs", "lowdin_it_group_download_quads",
    "lowdin_it_transform_stream_sink", "lowdin_it_occ_batch_model", "lowdin_it_exchange_is_dma", "lowdin_it_set_basis", "lowdin_it_basis_norma", "lowdin_it_ao_compute", "lowdin_it_ao_download",
]


class LowdinITError(RuntimeError):
    pass


class Shell(C.Structure):
    """lowdin_it_shell: one Cartesian shell as LibintInterface::add_shell receives it (Libint2Iface.cpp:83)."""
    _fields_ = [("l", C.c_int), ("nprim", C.c_int), ("first_prim", C.c_int), ("origin", C.c_double * 3)]


def pack_shells(shells):
    """[(l, origin(3), exponents, coefficients), ...] -> (Shell array, exponents, coefficients, number of Cartesian functions)."""
    arr = (Shell * len(shells))()
    ex, co, nbf = [], [], 0
    for i, (l, origin, e, c) in enumerate(shells):
        assert len(e) == len(c)
        arr[i] = Shell(l, len(e), len(ex), (C.c_double * 3)(*origin))
        ex += list(e); co += list(c)
        nbf += (l + 1) * (l + 2) // 2
    return arr, np.array(ex, dtype=np.float64), np.array(co, dtype=np.float64), nbf


class Block(C.Structure):
    """lowdin_it_block: one dense block of MO integrals handed to a host sink."""
    _fields_ = [("conv", C.c_int), ("nslots", C.c_int), ("n_second", C.c_int), ("n_first", C.c_int), ("orb_second0", C.c_int),
                ("orb_first0", C.c_int), ("second_is_conv_first", C.c_int), ("slot_a", C.POINTER(C.c_int32)),
                ("slot_b", C.POINTER(C.c_int32)), ("values", C.POINTER(C.c_double))]


SINK_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(Block))


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
    L.lowdin_it_transform_stream_sink.argtypes = [H, C.c_int, C.c_int, _i32p, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int,
                                                  C.c_void_p, C.c_void_p, C.c_double, _f64p, SINK_FN, C.c_void_p]
    L.lowdin_it_stream_num_passes.argtypes = [H, C.c_int, C.c_int, _i32p, C.c_int, C.c_int, C.POINTER(C.c_int),
                                              C.POINTER(C.c_int)]
    L.lowdin_it_transform_all.argtypes = [_f64pf, _f64p, C.c_int]
    L.lowdin_it_transform_inter_all.argtypes = [_f64pf, _f64pf, _f64p, C.c_int, C.c_int]
    L.lowdin_it_comm_unique_id.argtypes = [C.c_char_p]
    L.lowdin_it_comm_init.argtypes = [H, C.c_int, C.c_int, C.c_char_p]
    L.lowdin_it_shard_plan.argtypes = [C.c_int, _i32p, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int, _i32p, C.POINTER(C.c_int64),
                                       C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    L.lowdin_it_exchanged_offset.argtypes = [C.c_int64] * 5 + [C.c_int, C.c_int]
    L.lowdin_it_exchanged_offset.restype = C.c_int64
    L.lowdin_it_slab_owner.argtypes = [C.c_int64, C.c_int, C.c_int]
    L.lowdin_it_slab_local.argtypes = [C.c_int64, C.c_int, C.c_int]
    L.lowdin_it_slab_local.restype = C.c_int64
    L.lowdin_it_slab_global.argtypes = [C.c_int64, C.c_int, C.c_int, C.c_int]
    L.lowdin_it_slab_global.restype = C.c_int64
    L.lowdin_it_timers.argtypes = [H, _f64p]
    L.lowdin_it_kernel_bench.argtypes = [H, C.c_int, C.c_int64, C.c_int64, C.c_int64, C.c_int, C.POINTER(C.c_double),
                                         C.POINTER(C.c_double)]
    L.lowdin_it_set_profiling.argtypes = [H, C.c_int]
    L.lowdin_it_set_option.argtypes = [H, C.c_int, C.c_int64]
    L.lowdin_it_kernel_stats.argtypes = [H, _f64p, _f64p, _f64p]
    L.lowdin_it_debug_gemm.argtypes = [H, _f64p, _f64p, _f64p, C.c_int, C.c_int, C.c_int]
    L.lowdin_it_debug_expand.argtypes = [H, C.c_int, C.c_int, C.c_int64, C.c_int, _f64p]
    L.lowdin_it_ao_push_blocks.argtypes = [H, C.c_void_p, C.c_int64, C.c_int]
    L.lowdin_it_ao_set_rankk.argtypes = [H, C.c_int, C.c_int, C.c_int, _f64p, C.c_void_p]
    L.lowdin_it_ao_materialize.argtypes = [H, C.c_int, C.c_int]
    L.lowdin_it_comm_init_local.argtypes = [C.POINTER(H), C.c_int]
    L.lowdin_it_exchange_is_dma.argtypes = [H]
    L.lowdin_it_occ_batch_model.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int64, C.c_double, C.c_int]
    L.lowdin_it_set_basis.argtypes = [H, C.c_int, C.c_int, C.POINTER(Shell), _f64p, _f64p]
    L.lowdin_it_basis_norma.argtypes = [H, C.c_int, _f64p]
    L.lowdin_it_ao_compute.argtypes = [H, C.c_int, C.c_int]
    L.lowdin_it_ao_download.argtypes = [H, C.c_int, C.c_int, _f64p, C.c_int64]
    L.lowdin_it_result_segments.argtypes = [H, C.POINTER(C.c_int64), C.c_void_p, C.c_void_p]
    L.lowdin_it_group_transform.argtypes = [C.POINTER(H), C.c_int, C.c_int, C.c_int, _i32p, C.c_int, C.c_int, C.c_double]
    L.lowdin_it_group_result_count.argtypes = [C.POINTER(H), C.c_int, C.POINTER(C.c_int64)]
    L.lowdin_it_group_download_pairs.argtypes = [C.POINTER(H), C.c_int, _i64p, _i64p, _f64p]
    L.lowdin_it_group_download_quads.argtypes = [C.POINTER(H), C.c_int, _i32p, _i32p, _i32p, _i32p, _f64p]
    L.lowdin_it_debug_first_quarter.argtypes = [H, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int, _f64p]
    L.lowdin_it_debug_first_half.argtypes = [H, C.c_int, C.c_int, _i32p, C.c_int, C.c_double, C.c_int64, C.c_int, C.c_void_p,
                                             C.POINTER(C.c_int64)]
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

    def upload_ao_blocks(self, a, b, raw, stack, swapped=False):
        """Push raw .ints bytes (blocks of int32 p[S],q[S],r[S],s[S]; float64 v[S]) -- what a file read returns."""
        raw = np.ascontiguousarray(np.frombuffer(raw, np.uint8) if not isinstance(raw, np.ndarray) else raw.view(np.uint8))
        assert raw.size % (24 * stack) == 0, "not a whole number of stacks"
        self._ck(self.L.lowdin_it_ao_begin(self.h, a, b, int(swapped)))
        self._ck(self.L.lowdin_it_ao_push_blocks(self.h, raw.ctypes.data, raw.size // (24 * stack), stack))
        self._ck(self.L.lowdin_it_ao_end(self.h))

    def set_generator(self, a, b, seed, kind=GEN_HASH):
        self._ck(self.L.lowdin_it_ao_set_generator(self.h, a, b, kind, seed))

    def set_rankk(self, a, b, La, Lb=None):
        """Kind K: La [K][M_a] pair vectors (xy order) of the K symmetric factor matrices; Lb for species b (inter)."""
        La = np.ascontiguousarray(La, dtype=np.float64)
        Lb = np.ascontiguousarray(Lb, dtype=np.float64) if Lb is not None else None
        self._ck(self.L.lowdin_it_ao_set_rankk(self.h, a, b, La.shape[0], La, Lb.ctypes.data if Lb is not None else None))

    def materialize(self, a, b):
        self._ck(self.L.lowdin_it_ao_materialize(self.h, a, b))

    def set_basis(self, slot, shells):
        """shells: [(l, origin, exponents, coefficients), ...] (see pack_shells); the species must be set with nao = number of functions."""
        arr, ex, co, nbf = pack_shells(shells)
        self._ck(self.L.lowdin_it_set_basis(self.h, slot, len(shells), arr, ex, co))
        return nbf

    def basis_norma(self, slot):
        out = np.zeros(self.n[slot])
        self._ck(self.L.lowdin_it_basis_norma(self.h, slot, out))
        return out

    def compute_ao(self, a, b):
        """Row f4: the AO integrals of the pair evaluated on the device into the stored tensor (no .ints stream)."""
        self._ck(self.L.lowdin_it_ao_compute(self.h, a, b))

    def download_ao(self, a, b):
        Ma, Mb = self.n[a] * (self.n[a] + 1) // 2, self.n[b] * (self.n[b] + 1) // 2
        out = np.zeros(Ma * (Ma + 1) // 2 if a == b else Ma * Mb)
        self._ck(self.L.lowdin_it_ao_download(self.h, a, b, out, out.size))
        return out if a == b else out.reshape(Mb, Ma)

    def debug_first_quarter(self, a, b, f_first, nf, slab0, nslabs):
        out = np.zeros((nf, nslabs, self.n[a]))
        self._ck(self.L.lowdin_it_debug_first_quarter(self.h, a, b, f_first, nf, slab0, nslabs, out))
        return out

    def debug_first_half(self, a, b, win, conv, slab0, nslabs, tol=1e-10):
        w = np.ascontiguousarray(win, dtype=np.int32)
        npairs = C.c_int64()
        self._ck(self.L.lowdin_it_debug_first_half(self.h, a, b, w, conv, tol, slab0, nslabs, None, C.byref(npairs)))
        out = np.zeros((npairs.value, nslabs))
        self._ck(self.L.lowdin_it_debug_first_half(self.h, a, b, w, conv, tol, slab0, nslabs, out.ctypes.data, C.byref(npairs)))
        return out

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

    def result_segments(self):
        """(pair_index, kept): the window pairs (convention-order index) behind this rank's list and the entries each contributed."""
        n = C.c_int64()
        self._ck(self.L.lowdin_it_result_segments(self.h, C.byref(n), None, None))
        idx, kept = np.zeros(n.value, np.int64), np.zeros(n.value, np.int64)
        self._ck(self.L.lowdin_it_result_segments(self.h, C.byref(n), idx.ctypes.data, kept.ctypes.data))
        return idx, kept

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

    def transform_stream_sink(self, a, b, win, conv, sink, tol=1e-10, occ_batch=0, first_pass=0, n_passes=0, epsA=None, epsB=None, lam=2.0):
        """transform_stream with every dense result block also delivered to sink(slot_a, slot_b, values[nslots, n_second, n_first],
        block) on the host (numpy views valid during the call)."""
        w = np.ascontiguousarray(win, dtype=np.int32)
        sums = np.zeros(4)
        ea = np.ascontiguousarray(epsA, dtype=np.float64) if epsA is not None else None
        eb = np.ascontiguousarray(epsB, dtype=np.float64) if epsB is not None else None
        err = []

        def cb(_user, bp):
            try:
                b_ = bp.contents
                n = b_.nslots
                vals = np.ctypeslib.as_array(b_.values, (n, b_.n_second, b_.n_first))
                sink(np.ctypeslib.as_array(b_.slot_a, (n,)), np.ctypeslib.as_array(b_.slot_b, (n,)), vals, b_)
                return 0
            except BaseException as e:  # noqa: BLE001
                err.append(e)
                return 1
        fn = SINK_FN(cb)
        rc = self.L.lowdin_it_transform_stream_sink(
            self.h, a, b, w, conv, tol, occ_batch, first_pass, n_passes,
            ea.ctypes.data if ea is not None else None, eb.ctypes.data if eb is not None else None, lam, sums, fn, None)
        if err:
            raise err[0]
        self._ck(rc)
        return sums

    def comm_init(self, rank, nranks, uid: bytes):
        self._ck(self.L.lowdin_it_comm_init(self.h, rank, nranks, uid))

    def exchange_is_dma(self):
        return bool(self.L.lowdin_it_exchange_is_dma(self.h))

    def timers(self):
        t = np.zeros(8)
        self._ck(self.L.lowdin_it_timers(self.h, t))
        return dict(ao_upload=t[0], first_half=t[1], exchange=t[2], second_half=t[3], consume=t[4], download=t[5],
                    flops=t[6], launches=int(t[7]))

    CATEGORIES = ("expand1", "q1", "q2", "expand2", "q3", "q4", "consume", "exchange")

    OPT_WORKSPACE_BYTES, OPT_CHUNK_COLS, OPT_Q1_VARIANT, OPT_BENCH_GEN, OPT_GEMM_VARIANT, OPT_SPLIT_ROW_TAIL, OPT_FRAG_PERM = 1, 2, 3, 4, 5, 6, 7
    OPT_ASYNC_PUSH, OPT_STAGING_BYTES, OPT_Q3_RED, OPT_AO_LIST, OPT_SLAB_BLOCK_LOG, OPT_Q1_DEBUG, OPT_STORED_FUSED, OPT_OVERLAP_EXCHANGE, OPT_SINK_BLOCK_BYTES, OPT_GEMM_TALL, OPT_Q3_TWO_CTA, OPT_EXCHANGE_DMA = 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19
    DEFAULT_Q1_VARIANT, DEFAULT_GEMM_VARIANT, DEFAULT_FRAG_PERM = 5, 2, 1  # library defaults (it_api.cu); tests restore them after forcing a variant

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


def local_group(transformers):
    """Make the given Transformers (one process, any devices) ranks 0..n-1 of an in-process group; each must then be driven
    by its own thread making the same collective calls."""
    arr = (C.c_void_p * len(transformers))(*[t.h for t in transformers])
    if load().lowdin_it_comm_init_local(arr, len(transformers)):
        raise LowdinITError(load().lowdin_it_last_error(None).decode())


def group_transform(transformers, a, b, win, conv, symmetric=False, tol=1e-10):
    """lowdin_it_group_transform + the merged download: what a one-process host gets from a group of GPUs."""
    L = load()
    arr = (C.c_void_p * len(transformers))(*[t.h for t in transformers])
    w = np.ascontiguousarray(win, dtype=np.int32)
    if L.lowdin_it_group_transform(arr, len(transformers), a, b, w, conv, int(symmetric), tol):
        raise LowdinITError(L.lowdin_it_last_error(transformers[0].h).decode())
    cnt = C.c_int64()
    L.lowdin_it_group_result_count(arr, len(transformers), C.byref(cnt))
    n = cnt.value
    v = np.zeros(n)
    if conv == CONV_E:
        ij, kl = np.zeros(n, np.int64), np.zeros(n, np.int64)
        if L.lowdin_it_group_download_pairs(arr, len(transformers), ij, kl, v):
            raise LowdinITError(L.lowdin_it_last_error(transformers[0].h).decode())
        return ij, kl, v
    o = [np.zeros(n, np.int32) for _ in range(4)]
    if L.lowdin_it_group_download_quads(arr, len(transformers), *o, v):
        raise LowdinITError(L.lowdin_it_last_error(transformers[0].h).decode())
    return (*o, v)


def run_ranks(transformers, fn):
    """Run fn(rank, transformer) on one thread per rank (collective calls of an in-process group); returns the results."""
    import threading
    out, err = [None] * len(transformers), [None] * len(transformers)

    def work(r):
        try:
            out[r] = fn(r, transformers[r])
        except BaseException as e:  # noqa: BLE001
            err[r] = e
    th = [threading.Thread(target=work, args=(r,)) for r in range(len(transformers))]
    for t in th:
        t.start()
    for t in th:
        t.join()
    for e in err:
        if e is not None:
            raise e
    return out


def shard_plan(fbeg, chunk_base, chunk_width, nranks, rank, log_block=5):
    """(own[nranks+1], wblk, loc_lo, count): the library's division of slots and of the chunk's slabs among ranks."""
    fbeg = np.ascontiguousarray(fbeg, np.int32)
    own = np.zeros(nranks + 1, np.int32)
    w, lo, cnt = C.c_int64(), C.c_int64(), C.c_int64()
    if load().lowdin_it_shard_plan(len(fbeg) - 1, fbeg, chunk_base, chunk_width, nranks, rank, log_block, own, C.byref(w), C.byref(lo),
                                   C.byref(cnt)):
        raise LowdinITError("bad shard_plan arguments")
    return own, w.value, lo.value, cnt.value


def exchanged_offset(row, slab, chunk_base, wblk, rows, nranks, log_block=5):
    return load().lowdin_it_exchanged_offset(row, slab, chunk_base, wblk, rows, nranks, log_block)


def slab_owner(slab, nranks, log_block=5):
    return load().lowdin_it_slab_owner(slab, nranks, log_block)


def slab_local(slab, nranks, log_block=5):
    return load().lowdin_it_slab_local(slab, nranks, log_block)


def slab_global(local, nranks, rank, log_block=5):
    return load().lowdin_it_slab_global(local, nranks, rank, log_block)


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


# ---------------------------------------------------------------------------------------------
# include/lowdin_it_host.h -- host-side mirror of the reference's transformer interface (C and E)
# ---------------------------------------------------------------------------------------------
HOST_SYMBOLS = [
    "lowdin_host_last_error", "lowdin_host_partial_transform", "lowdin_host_windows", "lowdin_host_ints_filename",
    "lowdin_host_write_ints_file", "lowdin_host_read_ints_file", "lowdin_host_write_moint_quads",
    "lowdin_host_write_moint_pairs", "lowdin_host_atomic_to_molecular_one_species",
    "lowdin_host_atomic_to_molecular_two_species", "lowdin_host_plan_program", "lowdin_host_run_program",
    "lowdin_host_write_moint_d_intra", "lowdin_host_write_moint_d_inter", "lowdin_host_wfn_read", "lowdin_host_wfn_append",
    "lowdin_host_wfn_load_species", "lowdin_host_group_atomic_to_molecular", "lowdin_host_write_computed_ints", "lowdin_host_group_run_program",
]


class HostControl(C.Structure):
    _fields_ = [("method", C.c_char), ("partial_transform", C.c_char * 16), ("integral_stack_size", C.c_int),
                ("nfiles", C.c_int), ("ionize_mo", C.c_int), ("pt_transition_operator", C.c_int),
                ("n_ionize_species", C.c_int), ("ionize_species", (C.c_char * 32) * 4), ("scratch_dir", C.c_char * 512),
                ("verbose", C.c_int)]


class HostSpecies(C.Structure):
    _fields_ = [("name", C.c_char * 32), ("id", C.c_int), ("nao", C.c_int), ("occupation", C.c_int),
                ("core_orbitals", C.c_int), ("active_orbitals", C.c_int), ("coeff", C.c_void_p), ("ldc", C.c_int),
                ("ncols", C.c_int)]


class HostTask(C.Structure):
    _fields_ = [("first", C.c_int), ("second", C.c_int), ("win", C.c_int * 8), ("symmetric", C.c_int), ("flops", C.c_double),
                ("rank", C.c_int)]


def host_control(method="C", partial="MP2", stack=30000, nfiles=1, ionize_mo=0, pt_transition_operator=False,
                 ionize_species=(), scratch_dir="", verbose=False):
    c = HostControl()
    c.method = method.encode()
    c.partial_transform = partial.encode()
    c.integral_stack_size, c.nfiles, c.ionize_mo = stack, nfiles, ionize_mo
    c.pt_transition_operator = int(pt_transition_operator)
    sp = [s for s in ionize_species if s and s != "NONE"]
    c.n_ionize_species = len(sp)
    for i, s in enumerate(sp[:4]):
        c.ionize_species[i].value = s.encode()
    c.scratch_dir = scratch_dir.encode()
    c.verbose = int(verbose)
    return c


def host_species(name, sid, nao, occ, core=0, active=0, coeff=None):
    s = HostSpecies()
    s.name = name.encode()
    s.id, s.nao, s.occupation, s.core_orbitals, s.active_orbitals = sid, nao, occ, core, active
    if coeff is not None:
        coeff = np.asfortranarray(coeff, dtype=np.float64)
        s._keep = coeff  # keep the buffer alive
        s.coeff = coeff.ctypes.data
        s.ldc, s.ncols = coeff.shape[0], coeff.shape[1]
    return s


def _host():
    L = load()
    if getattr(L, "_host_ready", False):
        return L
    PC, PS = C.POINTER(HostControl), C.POINTER(HostSpecies)
    L.lowdin_host_last_error.restype = C.c_char_p
    L.lowdin_host_partial_transform.argtypes = [C.c_int] * 4 + [C.c_char_p]
    L.lowdin_host_windows.argtypes = [PC, PS, PS, _i32p, C.POINTER(C.c_int)]
    L.lowdin_host_ints_filename.argtypes = [C.c_int, PS, PS, C.c_char_p, C.POINTER(C.c_int)]
    L.lowdin_host_write_ints_file.argtypes = [C.c_char_p, C.c_int, _i32p, _i32p, _i32p, _i32p, _f64p, C.c_int64]
    L.lowdin_host_read_ints_file.argtypes = [C.c_char_p, C.c_int, _i32p, _i32p, _i32p, _i32p, _f64p, C.c_int64,
                                             C.POINTER(C.c_int64)]
    L.lowdin_host_write_moint_quads.argtypes = [C.c_char_p, C.c_int, _i32p, _i32p, _i32p, _i32p, _f64p, C.c_int64]
    L.lowdin_host_write_moint_pairs.argtypes = [C.c_char_p, C.c_int, _i64p, _i64p, _f64p, C.c_int64]
    L.lowdin_host_atomic_to_molecular_one_species.argtypes = [C.c_void_p, PC, PS, C.POINTER(C.c_int64)]
    L.lowdin_host_atomic_to_molecular_two_species.argtypes = [C.c_void_p, PC, PS, PS, C.POINTER(C.c_int64)]
    L.lowdin_host_group_run_program.argtypes = [C.POINTER(C.c_void_p), C.c_int, PC, PS, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int)]
    L.lowdin_host_write_computed_ints.argtypes = [C.c_void_p, PC, PS, PS, C.c_int, C.c_int, C.POINTER(C.c_int64)]
    L.lowdin_host_group_atomic_to_molecular.argtypes = [C.POINTER(C.c_void_p), C.c_int, PC, PS, PS, C.POINTER(C.c_int64)]
    L.lowdin_host_plan_program.argtypes = [PC, PS, C.c_int, C.c_int, C.POINTER(HostTask), C.c_int, C.POINTER(C.c_int)]
    L.lowdin_host_run_program.argtypes = [C.c_void_p, PC, PS, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int)]
    L.lowdin_host_write_moint_d_intra.argtypes = [C.c_char_p, C.c_int, _f64p, C.POINTER(C.c_int64)]
    L.lowdin_host_write_moint_d_inter.argtypes = [C.c_char_p, C.c_int, C.c_int, _f64p, C.POINTER(C.c_int64)]
    L.lowdin_host_wfn_read.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_void_p, C.c_int64, C.POINTER(C.c_int64)]
    L.lowdin_host_wfn_append.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_int, _f64p, C.c_int64, C.c_int]
    L.lowdin_host_wfn_load_species.argtypes = [C.c_char_p, PS, _f64pf, C.c_void_p]
    L._host_ready = True
    return L


def _hck(rc):
    if rc:
        raise LowdinITError(_host().lowdin_host_last_error().decode())


def host_partial_transform(mp=0, pt=0, en=0, ci_level="NONE"):
    out = C.create_string_buffer(16)
    _hck(_host().lowdin_host_partial_transform(mp, pt, en, int(ci_level == "NONE"), out))
    return out.value.decode()


def host_windows(ctl, a, b=None):
    win = np.zeros(8, np.int32)
    sym = C.c_int()
    _hck(_host().lowdin_host_windows(C.byref(ctl), C.byref(a), C.byref(b) if b is not None else None, win, C.byref(sym)))
    return [int(x) for x in win], bool(sym.value)


def host_ints_filename(tid, a, b=None):
    out = C.create_string_buffer(256)
    sw = C.c_int()
    _hck(_host().lowdin_host_ints_filename(tid, C.byref(a), C.byref(b) if b is not None else None, out, C.byref(sw)))
    return out.value.decode(), bool(sw.value)


def host_write_ints_file(path, stack, p, q, r, s, v):
    a = [np.ascontiguousarray(x, np.int32) for x in (p, q, r, s)]
    _hck(_host().lowdin_host_write_ints_file(path.encode(), stack, *a, np.ascontiguousarray(v, np.float64), len(v)))


def host_read_ints_file(path, stack, cap):
    a = [np.zeros(max(cap, 1), np.int32) for _ in range(4)]
    v = np.zeros(max(cap, 1))
    n = C.c_int64()
    _hck(_host().lowdin_host_read_ints_file(path.encode(), stack, *a, v, cap, C.byref(n)))
    m = min(n.value, cap)
    return tuple(x[:m] for x in a) + (v[:m],), n.value


def host_write_moint_quads(path, stack, p, q, r, s, v):
    a = [np.ascontiguousarray(x, np.int32) for x in (p, q, r, s)]
    _hck(_host().lowdin_host_write_moint_quads(path.encode(), stack, *a, np.ascontiguousarray(v, np.float64), len(v)))


def host_write_moint_pairs(path, stack, ij, kl, v):
    _hck(_host().lowdin_host_write_moint_pairs(path.encode(), stack, np.ascontiguousarray(ij, np.int64),
                                               np.ascontiguousarray(kl, np.int64), np.ascontiguousarray(v, np.float64), len(v)))


def host_transform_one_species(T, ctl, a):
    """T may be None for method D, which runs behind the handle-less transformer-D entry points."""
    n = C.c_int64()
    _hck(_host().lowdin_host_atomic_to_molecular_one_species(T.h if T is not None else None, C.byref(ctl), C.byref(a), C.byref(n)))
    return n.value


def host_transform_two_species(T, ctl, a, b):
    n = C.c_int64()
    _hck(_host().lowdin_host_atomic_to_molecular_two_species(T.h if T is not None else None, C.byref(ctl), C.byref(a), C.byref(b), C.byref(n)))
    return n.value


def host_write_computed_ints(T, ctl, a, b, slot_a, slot_b=None):
    """Row f4 file form: the .ints stream files of a species (b None) or a pair from integrals evaluated on the device."""
    n = C.c_int64()
    _hck(_host().lowdin_host_write_computed_ints(T.h, C.byref(ctl), C.byref(a), C.byref(b) if b is not None else None, slot_a,
                                                 slot_a if slot_b is None else slot_b, C.byref(n)))
    return n.value


def host_group_transform(transformers, ctl, a, b=None):
    """File to file on an in-process group of GPUs: one moint.dat, entries in the single-GPU order."""
    n = C.c_int64()
    arr = (C.c_void_p * len(transformers))(*[t.h for t in transformers])
    _hck(_host().lowdin_host_group_atomic_to_molecular(arr, len(transformers), C.byref(ctl), C.byref(a), C.byref(b) if b is not None else None,
                                                       C.byref(n)))
    return n.value


def _species_array(species):
    arr = (HostSpecies * len(species))()
    for i, sp in enumerate(species):
        C.memmove(C.byref(arr[i]), C.byref(sp), C.sizeof(HostSpecies))
    arr._keep = list(species)  # coefficient buffers stay alive with the array
    return arr


def host_plan_program(ctl, species, nranks=1):
    """The transformer calls of the reference program's species loop and the rank each one runs on.
    Returns a list of dicts: first, second (indices into `species`, call order; second None = one species), win, symmetric,
    flops, rank."""
    arr = _species_array(species)
    cap = len(species) * (len(species) + 1) // 2
    tasks = (HostTask * max(cap, 1))()
    n = C.c_int()
    _hck(_host().lowdin_host_plan_program(C.byref(ctl), arr, len(species), nranks, tasks, cap, C.byref(n)))
    return [dict(first=t.first, second=(t.second if t.second >= 0 else None), win=list(t.win), symmetric=bool(t.symmetric),
                 flops=t.flops, rank=t.rank) for t in tasks[:n.value]]


def host_run_program(T, ctl, species, rank=0, nranks=1):
    """Run this rank's share of the program's calls; returns (integrals written, calls made)."""
    arr = _species_array(species)
    nz, nc = C.c_int64(), C.c_int()
    _hck(_host().lowdin_host_run_program(T.h if T is not None else None, C.byref(ctl), arr, len(species), rank, nranks, C.byref(nz), C.byref(nc)))
    return nz.value, nc.value


def host_group_run_program(transformers, ctl, species):
    """The program's loop on an in-process group of GPUs (every call collective); returns (integrals written, calls made)."""
    arr = _species_array(species)
    hs = (C.c_void_p * len(transformers))(*[t.h for t in transformers])
    nz, nc = C.c_int64(), C.c_int()
    _hck(_host().lowdin_host_group_run_program(hs, len(transformers), C.byref(ctl), arr, len(species), C.byref(nz), C.byref(nc)))
    return nz.value, nc.value


def host_write_moint_d(path, ints, nao, onao=None):
    """Transformer D's one-integral-per-record moint.dat from the in-place result of transform_all / transform_inter_all."""
    n = C.c_int64()
    ints = np.ascontiguousarray(ints, np.float64)
    if onao is None:
        _hck(_host().lowdin_host_write_moint_d_intra(path.encode(), nao, ints, C.byref(n)))
    else:
        _hck(_host().lowdin_host_write_moint_d_inter(path.encode(), nao, onao, ints, C.byref(n)))
    return n.value


def host_wfn_append(path, label, species, values, label_len=30, truncate=False):
    v = np.ascontiguousarray(np.asarray(values, np.float64).reshape(-1, order="F"))
    _hck(_host().lowdin_host_wfn_append(path.encode(), label.encode(), species.encode(), label_len, v, len(v), int(truncate)))


def host_wfn_read(path, label, species):
    n = C.c_int64()
    _hck(_host().lowdin_host_wfn_read(path.encode(), label.encode(), species.encode(), None, 0, C.byref(n)))
    out = np.zeros(max(n.value, 1))
    _hck(_host().lowdin_host_wfn_read(path.encode(), label.encode(), species.encode(), out.ctypes.data, n.value, C.byref(n)))
    return out[:n.value]


def host_wfn_load_species(path, name, sid, nao, occ, core=0, active=0):
    """A HostSpecies with its coefficients (and .eps) read from lowdin.wfn as the transformation program does."""
    sp = host_species(name, sid, nao, occ, core, active)
    coeff = np.zeros((nao, max(nao, occ)), order="F")
    eps = np.zeros(nao)
    _hck(_host().lowdin_host_wfn_load_species(path.encode(), C.byref(sp), coeff, eps.ctypes.data))
    sp._keep, sp.eps = coeff, eps
    return sp


def occ_batch_model(n_first, q_max, nranks, slots_per_first, n_first2, nao2, nao1, nslabs, npairs1, avail_bytes, stored=False):
    """The occupied batch the streaming transform picks (pure host logic, lowdin_it_occ_batch_model)."""
    return load().lowdin_it_occ_batch_model(n_first, q_max, nranks, slots_per_first, n_first2, nao2, nao1, nslabs, npairs1,
                                            float(avail_bytes), int(stored))
