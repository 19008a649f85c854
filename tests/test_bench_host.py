"""CPU tests of bench.py's host-side logic: the numpy AO-list generator against the oracle, the flop model, and the
stored-AO end-to-end leg driven through a stand-in for the device context.  The stand-in checks every argument against the
REAL library's ctypes signatures (the same from_param conversion a real call performs) and answers with the CPU oracle, so
that the Python glue of the leg is exercised here even though no kernel can run without a GPU."""
import ctypes as C
import types

import numpy as np
import pytest

import bench
from openlowdin_b200 import capi


@pytest.mark.parametrize("n", [3, 7, 12])
def test_canonical_ao_list_matches_oracle_list(O, n):
    got = bench.canonical_ao_list(11, n)
    packed = O.hash_packed_intra(11, n)
    ref = O.canonical_list_intra(packed, n)
    for g, r in zip(got, ref):
        assert g.dtype == r.dtype and np.array_equal(g, r)
    assert np.array_equal(O.scatter_intra(*got, n), packed)


def test_splitmix_values_match_oracle(O):
    keys = np.array([0, 1, 2, 12345678901234567, 2 ** 63 + 5, 2 ** 64 - 1], dtype=np.uint64)
    got = bench.splitmix_values(bench.SEED, keys)
    ref = [O.lib().orc_hash_value(bench.SEED, int(k) ^ 0) for k in keys]
    assert np.array_equal(got, np.array(ref))


def test_flop_model_matches_survey_figures():
    # N=1500 MP2 1.70e15, N=500 MP2 7.02e12 (whole transform = occ batches summed)
    for n, want in ((500, 7.02e12), (1500, 1.70e15)):
        occ = n // 10
        assert abs(bench.algorithmic_flops_pass(n, occ, occ) / want - 1.0) < 0.01


class _FakePinned:
    def __init__(self, a):
        self.a = a

    def pin_memory(self):
        return self

    def numpy(self):
        return self.a


class _FakeTorch:
    int32, int64, float64, uint8 = np.int32, np.int64, np.float64, np.uint8

    @staticmethod
    def empty(n, dtype):
        return _FakePinned(np.zeros(n, dtype))

    @staticmethod
    def from_numpy(a):
        return _FakePinned(a)

    class cuda:
        @staticmethod
        def synchronize():
            pass


class _FakeLib:
    """Stands in for the loaded library: argument conversion by the real argtypes, behaviour by the oracle."""

    def __init__(self, O, real):
        self.O, self.real, self.calls = O, real, []
        self.C, self.list, self.gen, self.res = None, None, None, None

    def _conv(self, name, args):
        at = getattr(self.real, name).argtypes
        assert len(at) == len(args), name
        for t, a in zip(at, args):
            t.from_param(a)  # raises exactly when a real call would
        self.calls.append(name)

    def lowdin_it_set_species(self, *a):
        self._conv("lowdin_it_set_species", a)
        self.C = np.array(a[3], order="F")
        return 0

    def lowdin_it_ao_begin(self, *a):
        self._conv("lowdin_it_ao_begin", a)
        self.list, self.gen = [[] for _ in range(5)], None
        return 0

    def lowdin_it_ao_push_stacks(self, *a):
        self._conv("lowdin_it_ao_push_stacks", a)
        p, n = a[1], a[6]
        m = 0
        while m < n and p[m] != -1:
            m += 1
        for dst, src in zip(self.list, a[1:6]):
            dst.append(np.array(src[:m]))
        return 0

    def lowdin_it_ao_push_blocks(self, *a):
        self._conv("lowdin_it_ao_push_blocks", a)
        nblk, S = a[2], a[3]
        raw = np.ctypeslib.as_array(C.cast(a[1], C.POINTER(C.c_uint8)), (nblk * 24 * S,)).reshape(nblk, 24 * S)
        idx = [raw[:, 4 * S * k:4 * S * (k + 1)].view(np.int32).ravel() for k in range(4)]
        v = raw[:, 16 * S:].view(np.float64).ravel()
        m = int(np.argmax(idx[0] == -1))
        assert idx[0][m] == -1
        for dst, src in zip(self.list, idx + [v]):
            dst.append(np.array(src[:m]))
        return 0

    def lowdin_it_ao_end(self, *a):
        self._conv("lowdin_it_ao_end", a)
        return 0

    def lowdin_it_ao_set_generator(self, *a):
        self._conv("lowdin_it_ao_set_generator", a)
        self.gen, self.list = a[4], None
        return 0

    def lowdin_it_transform(self, *a):
        self._conv("lowdin_it_transform", a)
        n = self.C.shape[0]
        if self.gen is not None:
            packed = self.O.hash_packed_intra(self.gen, n)
        else:
            packed = self.O.scatter_intra(*[np.concatenate(x) for x in self.list], n)
        self.res = self.O.transform_e_intra(self.C, packed, [int(x) for x in a[3]])
        return 0

    def lowdin_it_result_count(self, *a):
        self._conv("lowdin_it_result_count", a)
        a[1]._obj.value = len(self.res[2])
        return 0

    def lowdin_it_download_pairs(self, *a):
        self._conv("lowdin_it_download_pairs", a)
        for dst, src in zip(a[1:], self.res):
            dst[:len(src)] = src
        return 0

    def lowdin_it_timers(self, *a):
        self._conv("lowdin_it_timers", a)
        return 0

    def lowdin_it_destroy(self, *a):
        return 0

    def lowdin_it_last_error(self, *a):
        return b"fake"


def test_stored_ao_e2e_leg_with_stand_in_context(O, monkeypatch):
    real = capi.load()
    fake = _FakeLib(O, real)

    class FakeTransformer(capi.Transformer):
        def __init__(self, device=0):
            self.L, self.h, self.n = fake, C.c_void_p(1), {}

    ol = types.SimpleNamespace(Transformer=FakeTransformer)
    n, occ = 8, 2
    out = bench.stored_ao_e2e(_FakeTorch, ol, capi, 0, n, occ, steps=2, push_entries=100)
    M = n * (n + 1) // 2
    total = M * (M + 1) // 2
    assert out["same_index_lists_as_generated"] is True and out["max_abs_diff_vs_generated"] == 0.0
    assert out["h2d_bytes_per_step"] == total * 24 + n * n * 8
    assert out["mo_integrals_kept"] == len(fake.res[2]) > 0 and out["d2h_bytes_per_step"] == 24 * out["mo_integrals_kept"]
    assert out["value"] > 0 and out["steps"] == 2
    # every step pushed the whole list (terminator included) in pieces of at most push_entries
    assert fake.calls.count("lowdin_it_ao_push_stacks") == 3 * -(-(total + 1) // 100)
    ref = O.transform_e_intra(O.random_orthonormal(n, n), O.hash_packed_intra(bench.SEED, n), bench.mp2_window_e(n, occ))
    assert np.array_equal(ref[2], fake.res[2])
    # the host-side comparison leg takes the GPU list (here: the stand-in's) and finds it identical to transformer E's
    cpu = bench.whole_transform_cpu_leg(n, occ, out.pop("_result"))
    assert cpu["max_abs_diff_gpu_vs_transformer_e"] == 0.0 and cpu["mo_integrals_e"] == out["mo_integrals_kept"]
    assert cpu["transformer_e_port_s"] > 0 and cpu["transformer_c_port_s"] > 0
    import json
    json.dumps(out), json.dumps(cpu)          # everything that reaches the bench line is JSON-serialisable


def test_stored_ao_e2e_leg_raw_blocks_mode(O):
    """The same leg pushing the list as the raw bytes of a .ints stream (one lowdin_it_ao_push_blocks call per step)."""
    real = capi.load()
    fake = _FakeLib(O, real)

    class FakeTransformer(capi.Transformer):
        def __init__(self, device=0):
            self.L, self.h, self.n = fake, C.c_void_p(1), {}

    ol = types.SimpleNamespace(Transformer=FakeTransformer)
    n, occ = 7, 2
    out = bench.stored_ao_e2e(_FakeTorch, ol, capi, 0, n, occ, steps=1, mode="blocks", stack_size=37)
    assert out["same_index_lists_as_generated"] is True and out["max_abs_diff_vs_generated"] == 0.0
    assert fake.calls.count("lowdin_it_ao_push_blocks") == 2 and "lowdin_it_ao_push_stacks" not in fake.calls


def test_transformer_d_leg_with_stand_in(O):
    """bench.py's reference-D leg: the reference's own IntTransfD.cpp (oracle/_ref) against a stand-in for the GPU drop-in."""
    if O.ref() is None:
        pytest.skip("oracle/_ref not built (no reference tree)")
    calls = []

    def transform_all(Cm, ints):
        calls.append(ints.shape)
        O.lib().orc_transform_d_intra(np.asfortranarray(Cm), ints, Cm.shape[0])
        return ints

    out = bench.transformer_d_leg(types.SimpleNamespace(transform_all=transform_all), n_full=14)
    n = 14 if "OpenBLAS" in out["reference_blas"] else 60
    M = n * (n + 1) // 2
    assert calls == [(M * (M + 1) // 2,)] * 2                    # one warm-up call, one timed call
    assert out["max_abs_diff"] <= 1e-10 and out["reference_ms"] > 0 and out["gpu_ms"] > 0
    assert out["d2h_bytes"] == 8 * M * (M + 1) // 2


def test_occupied_batch_model_decisions():
    """lowdin_it_occ_batch_model (host logic): the batch sizes the cost model picks for the N_bf = 1500 MP2 window (150 occupied,
    1350 virtuals) with ~157 GB for accumulators + chunk buffers per GPU -- one column group of the fused first quarter per pass
    unless the whole window fits with wide chunks."""
    from openlowdin_b200 import capi
    n, occ, virt = 1500, 150, 1350
    M = n * (n + 1) // 2
    avail = 157e9

    def pick(G, qmax, stored=False):
        return capi.occ_batch_model(occ, qmax, G, virt, occ, n, n, M, M, avail, stored)
    q1 = pick(1, 60)
    assert 48 <= q1 <= 56 and -(-occ // q1) == 3           # one GPU: three passes of one <= 56-column group
    q2 = pick(2, 120)
    assert 48 <= q2 <= 64 and -(-occ // q2) == 3           # two GPUs: NOT 110 + 40 (measured: third quarter at 14 TFLOP/s)
    assert pick(8, 150) == 150                             # eight GPUs: the whole window in one pass (6 chunks)
    assert pick(1, 7) == 7                                 # memory-bound small batches: the largest that fits
    # N_bf = 2000 on one GPU: 25 fit; nothing smaller is better
    n2, occ2, virt2 = 2000, 200, 1800
    M2 = n2 * (n2 + 1) // 2
    assert capi.occ_batch_model(occ2, 25, 1, virt2, occ2, n2, n2, M2, M2, avail, False) in (24, 25)
