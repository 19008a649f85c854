"""Host-side mirror of the reference's transformer interface (include/lowdin_it_host.h): window tables against the
oracle's restatement and literal expectations, file names, .ints stream files, moint.dat record layouts (CPU), and the
file-to-file transformer calls on the GPU against the oracle."""
import itertools
import os
import struct

import numpy as np
import pytest

from openlowdin_b200 import capi
from helpers import assert_lists_match, dense_pairs, dense_quads

MODES = ["MP2", "PT2", "MP2-PT2", "ALL", "ALLACTIVE", "BOUNDS"]


class _MockContext:
    """A handle of tests/mock/libmock_host.so: the host mirror (product source, unchanged) linked against a CPU stand-in of the
    device library that answers with the oracle.  Exercises the mirror's orchestration on CPU; the CUDA path is the `gpu` param."""

    def __init__(self, L):
        import ctypes as C
        self.L, self.h = L, C.c_void_p()
        L.lowdin_it_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
        L.lowdin_it_destroy.argtypes = [C.c_void_p]
        assert L.lowdin_it_create(0, C.byref(self.h)) == 0

    def close(self):
        self.L.lowdin_it_destroy(self.h)


@pytest.fixture(params=["cpu_mock", pytest.param("gpu", marks=pytest.mark.gpu)])
def HT(request, monkeypatch):
    """The device context the file-to-file tests run on: the real CUDA library (-m gpu) or the CPU mock (-m "not gpu")."""
    if request.param == "gpu":
        yield request.getfixturevalue("T")
        return
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "mock"))
    import build_mock
    L = build_mock.load()
    monkeypatch.setattr(capi, "_lib", L)          # every capi.host_* call of the test goes to the mock-linked mirror
    ctx = _MockContext(L)
    yield ctx
    ctx.close()


# ---------------------------------------------------------------------------------------------
# integer host logic (CPU)
# ---------------------------------------------------------------------------------------------
def test_partial_transform_choice(O):
    for mp, pt, en, ci in itertools.product((0, 2), (0, 2, 3), (0, 2), ("NONE", "CISD")):
        assert capi.host_partial_transform(mp, pt, en, ci) == O.partial_transform(mp, pt, en, ci)
    assert capi.host_partial_transform(mp=2) == "MP2"


@pytest.mark.parametrize("mode", MODES)
def test_intra_windows_match_restated_tables(O, mode):
    for n, occ, core, active, mo, pto in itertools.product((19,), (1, 5), (0, 1), (0, 15), (0, 3, 7), (False, True)):
        a = capi.host_species("E-", 1, n, occ, core, active)
        wc, sym = capi.host_windows(capi.host_control("C", mode, ionize_mo=mo, pt_transition_operator=pto), a)
        assert (wc, sym) == O.windows_c_intra(mode, n, occ, core, active, mo, pto)
        we, _ = capi.host_windows(capi.host_control("E", mode, ionize_mo=mo, pt_transition_operator=pto), a)
        assert we == O.windows_e_intra(mode, n, occ, core, active, mo, pto)


@pytest.mark.parametrize("mode", MODES)
def test_inter_windows_match_restated_tables(O, mode):
    for (oa, ob), (ca, cb), (aa, ab), mo, ion, pto in itertools.product(
            ((5, 1), (1, 5)), ((0, 0), (1, 0)), ((0, 0), (15, 40)), (0, 1, 3, 9),
            ((), ("E-",), ("H_1",), ("E-", "H_1")), (False, True)):
        a = capi.host_species("E-", 1, 19, oa, ca, aa)
        b = capi.host_species("H_1", 2, 50, ob, cb, ab)
        kw = dict(core_a=ca, core_b=cb, active_a=aa, active_b=ab, ionize_mo=mo, ionize_species=ion or ("NONE",),
                  name_a="E-", name_b="H_1")
        wc, sym = capi.host_windows(capi.host_control("C", mode, ionize_mo=mo, ionize_species=ion, pt_transition_operator=pto), a, b)
        assert (wc, sym) == O.windows_c_inter(mode, 19, 50, oa, ob, **kw)
        we, _ = capi.host_windows(capi.host_control("E", mode, ionize_mo=mo, ionize_species=ion, pt_transition_operator=pto), a, b)
        assert we == O.windows_e_inter(mode, 19, 50, oa, ob, pt_transition_operator=pto, **kw)


def test_window_literals_from_the_reference():
    e = capi.host_species("E-", 1, 19, 5)
    h = capi.host_species("H_1", 2, 50, 1)
    assert capi.host_windows(capi.host_control("C", "MP2"), e) == ([1, 5, 6, 19, 1, 5, 6, 19], True)      # C.f90:1489-1501
    assert capi.host_windows(capi.host_control("C", "PT2"), e) == ([5, 6, 1, 19, 1, 5, 6, 19], False)     # C.f90:1509-1520
    assert capi.host_windows(capi.host_control("E", "MP2"), e)[0] == [6, 19, 1, 5, 6, 19, 1, 5]           # E.f90:1938-1949
    assert capi.host_windows(capi.host_control("C", "MP2"), e, h) == ([1, 5, 6, 19, 1, 1, 2, 50], True)  # C.f90:1672-1684
    assert capi.host_windows(capi.host_control("E", "PT2"), e, h)[0] == [1, 19, 1, 6, 1, 50, 1, 2]        # E.f90:2135-2145 (first species' core)
    assert capi.host_windows(capi.host_control("E", "MP2-PT2"), e, h)[0] == [1, 19, 1, 6, 1, 19, 1, 2]    # E.f90:2283 quirk: r_u = active of A


def test_ints_file_names():
    ea, eb = capi.host_species("E-ALPHA", 1, 7, 3), capi.host_species("E-BETA", 2, 7, 2)
    hh = capi.host_species("H_1", 3, 5, 1)
    assert capi.host_ints_filename(0, ea) == ("0E-ALPHA.ints", False)
    assert capi.host_ints_filename(3, eb) == ("3E-ALPHA.ints", False)                 # C.f90:241-245
    assert capi.host_ints_filename(1, ea, eb) == ("1E-ALPHA.E-BETA.ints", False)      # C.f90:838-848
    assert capi.host_ints_filename(0, eb, hh) == ("0E-ALPHA.H_1.ints", False)
    assert capi.host_ints_filename(0, hh, eb) == ("0E-ALPHA.H_1.ints", True)          # C.f90:906-915 (reversed call order)
    assert capi.host_ints_filename(2, hh, ea) == ("2E-ALPHA.H_1.ints", True)
    assert capi.host_ints_filename(0, ea, hh) == ("0E-ALPHA.H_1.ints", False)


def test_ints_stream_roundtrip_and_layout(tmp_path):
    rng = np.random.default_rng(1)
    for n, S in ((0, 4), (3, 4), (4, 4), (9, 4), (30, 7)):
        p, q, r, s = (rng.integers(1, 9, n).astype(np.int32) for _ in range(4))
        v = rng.uniform(-1, 1, n)
        path = str(tmp_path / f"0X{n}.ints")
        capi.host_write_ints_file(path, S, p, q, r, s, v)
        nblocks = n // S + 1                                        # a terminator always follows the last entry
        assert os.path.getsize(path) == nblocks * 24 * S             # C.f90:251-253: nblocks = filesize/24/S
        raw = open(path, "rb").read()
        blk = raw[(nblocks - 1) * 24 * S:]
        assert struct.unpack_from("<i", blk, 4 * (n % S))[0] == -1  # p(counter) = -1, Libint2Iface.cpp:395-396
        (gp, gq, gr, gs, gv), cnt = capi.host_read_ints_file(path, S, n + 5)
        assert cnt == n and np.array_equal(gp, p) and np.array_equal(gq, q) and np.array_equal(gr, r) and np.array_equal(gs, s)
        assert np.array_equal(gv, v)


def _records(raw):
    out, o = [], 0
    while o < len(raw):
        (ln,) = struct.unpack_from("<I", raw, o)
        body = raw[o + 4:o + 4 + ln]
        assert struct.unpack_from("<I", raw, o + 4 + ln)[0] == ln   # trailing gfortran marker
        out.append(body)
        o += 8 + ln
    return out


def read_moint_quads(path, S):
    """ReadTransformedIntegrals.f90:225-236 (method C layout)."""
    P, Q, R, Sx, V = [], [], [], [], []
    for body in _records(open(path, "rb").read()):
        assert len(body) == 24 * S
        a = np.frombuffer(body, np.int32, 4 * S).reshape(4, S)
        v = np.frombuffer(body, np.float64, S, 16 * S)
        for i in range(S):
            if a[0, i] == -1:
                return tuple(np.array(x) for x in (P, Q, R, Sx, V))
            P.append(a[0, i]); Q.append(a[1, i]); R.append(a[2, i]); Sx.append(a[3, i]); V.append(v[i])
    raise AssertionError("no terminator record")


def read_moint_pairs(path, S):
    """ReadTransformedIntegrals.f90:302-311 (method E layout)."""
    IJ, KL, V = [], [], []
    for body in _records(open(path, "rb").read()):
        assert len(body) == 24 * S
        a = np.frombuffer(body, np.int64, 2 * S).reshape(2, S)
        v = np.frombuffer(body, np.float64, S, 16 * S)
        for i in range(S):
            if a[0, i] == -1:
                return np.array(IJ, np.int64), np.array(KL, np.int64), np.array(V)
            IJ.append(a[0, i]); KL.append(a[1, i]); V.append(v[i])
    raise AssertionError("no terminator record")


def test_moint_record_layouts(tmp_path):
    rng = np.random.default_rng(2)
    for n, S in ((0, 5), (4, 5), (5, 5), (12, 5)):
        p, q, r, s = (rng.integers(1, 9, n).astype(np.int32) for _ in range(4))
        v = rng.uniform(-1, 1, n)
        path = str(tmp_path / "Xmoint.dat")
        capi.host_write_moint_quads(path, S, p, q, r, s, v)
        assert os.path.getsize(path) == (n // S + 1) * (24 * S + 8)   # one sequential record per stack + terminator stack
        g = read_moint_quads(path, S)
        assert all(np.array_equal(a, b) for a, b in zip(g, (p, q, r, s, v)))
        ij, kl = rng.integers(1, 99, n), rng.integers(1, 99, n)
        capi.host_write_moint_pairs(path, S, ij, kl, v)
        assert os.path.getsize(path) == (n // S + 1) * (24 * S + 8)
        g = read_moint_pairs(path, S)
        assert np.array_equal(g[0], ij) and np.array_equal(g[1], kl) and np.array_equal(g[2], v)


def test_errors_are_reported():
    with pytest.raises(capi.LowdinITError):
        capi.host_windows(capi.host_control("Q", "MP2"), capi.host_species("E-", 1, 5, 2))
    with pytest.raises(capi.LowdinITError):
        capi.host_read_ints_file("/nonexistent/0E-.ints", 4, 4)


# ---------------------------------------------------------------------------------------------
# file -> GPU -> file, against the oracle
# ---------------------------------------------------------------------------------------------
def _split_to_files(tmp_path, name, lst, nfiles, S):
    """lowdin-ints.x writes quartet k to file k % nthreads (Libint2Iface.cpp:282-286, :332)."""
    for t in range(nfiles):
        part = [x[t::nfiles] for x in lst]
        capi.host_write_ints_file(str(tmp_path / f"{t}{name}.ints"), S, *part)


@pytest.mark.parametrize("method,mode", [("C", "MP2"), ("C", "ALL"), ("C", "PT2"), ("E", "MP2"), ("E", "MP2-PT2")])
def test_one_species_file_to_file(O, HT, tmp_path, method, mode):
    T = HT
    n, occ, S, nfiles = 13, 4, 64, 3
    packed = O.hash_packed_intra(61, n)
    Cm = O.random_orthonormal(n, n)
    _split_to_files(tmp_path, "E-", O.canonical_list_intra(packed, n), nfiles, S)
    ctl = capi.host_control(method, mode, stack=S, nfiles=nfiles, scratch_dir=str(tmp_path))
    sp = capi.host_species("E-", 1, n, occ, coeff=Cm)
    cnt = capi.host_transform_one_species(T, ctl, sp)
    path = str(tmp_path / "E-moint.dat")
    if method == "C":
        win, sym = O.windows_c_intra(mode, n, occ)
        ref = O.transform_c_intra(Cm, packed, win, sym)
        got = read_moint_quads(path, S)
        assert len(got[4]) == cnt
        assert np.abs(dense_quads(*got, n, n) - dense_quads(*ref, n, n)).max() <= 1e-10
        assert_lists_match(got[:4], got[4], ref[:4], ref[4])
        e_got = O.mp2_intra_from_quads(*[np.ascontiguousarray(x) for x in got], n, occ, O.synthetic_eps(occ, n))
        e_ref = O.mp2_intra_from_quads(*ref, n, occ, O.synthetic_eps(occ, n))
    else:
        win = O.windows_e_intra(mode, n, occ)
        ref = O.transform_e_intra(Cm, packed, win)
        got = read_moint_pairs(path, S)
        assert len(got[2]) == cnt
        M = O.npairs(n)
        assert np.abs(dense_pairs(*got, M, M) - dense_pairs(*ref, M, M)).max() <= 1e-10
        assert_lists_match(got[:2], got[2], ref[:2], ref[2])
        e_got = O.mp2_intra_from_pairs(*got, n, occ, O.synthetic_eps(occ, n))
        e_ref = O.mp2_intra_from_pairs(*ref, n, occ, O.synthetic_eps(occ, n))
    if mode == "MP2":
        assert abs(e_got - e_ref) <= 1e-9          # downstream energy read back from the file


@pytest.mark.parametrize("method", ["C", "E"])
def test_empty_window_writes_a_terminator_only_file(O, HT, tmp_path, method):
    """A one-function species under MP2 has the virtual window 2..1: the reference writes a moint.dat holding only the terminator
    stack (TransformIntegralsC.f90:450-456, TransformIntegralsE.f90:1263-1268); so must the drop-in, on a FRESH context."""
    n, occ, S = 1, 1, 16
    one = np.array([1], np.int32)
    _split_to_files(tmp_path, "H_1", (one, one, one, one, np.array([0.75])), 1, S)
    ctl = capi.host_control(method, "MP2", stack=S, nfiles=1, scratch_dir=str(tmp_path))
    sp = capi.host_species("H_1", 1, n, occ, coeff=np.ones((1, 1)))
    cnt = capi.host_transform_one_species(HT, ctl, sp)
    assert cnt == 0
    path = str(tmp_path / "H_1moint.dat")
    got = read_moint_quads(path, S) if method == "C" else read_moint_pairs(path, S)
    assert len(got[-1]) == 0
    assert len(list(_records(open(path, "rb").read()))) == 1


@pytest.mark.parametrize("method", ["C", "E"])
def test_two_species_file_to_file_both_call_orders(O, HT, tmp_path, method):
    T = HT
    na, nb, oa, ob, S = 9, 7, 3, 1, 50
    rect = O.hash_rect_inter(17, na, nb)                       # stored (rs_B, pq_A)
    Ca, Cb = O.random_orthonormal(na, 5), O.random_orthonormal(nb, 6)
    lst = O.canonical_list_inter(rect, na, nb)                 # file order (A A|B B), A = lower species id
    _split_to_files(tmp_path, "E-.H_1", lst, 2, S)
    A = capi.host_species("E-", 1, na, oa, coeff=Ca)
    B = capi.host_species("H_1", 2, nb, ob, coeff=Cb)
    ctl = capi.host_control(method, "MP2", stack=S, nfiles=2, scratch_dir=str(tmp_path))
    # forward order
    cnt = capi.host_transform_two_species(T, ctl, A, B)
    if method == "C":
        win, sym = O.windows_c_inter("MP2", na, nb, oa, ob)
        ref = O.transform_c_inter(Ca, Cb, rect, win, sym)
        got = read_moint_quads(str(tmp_path / "E-.H_1moint.dat"), S)
        assert len(got[4]) == cnt
        assert np.abs(dense_quads(*got, na, nb) - dense_quads(*ref, na, nb)).max() <= 1e-10
        # reversed call order (the species with fewer occupied orbitals first, IntegralTransformation.f90:322-334)
        cnt2 = capi.host_transform_two_species(T, ctl, B, A)
        rect_sw = O.scatter_inter(*lst, nb, na, swapped=True)
        win2, sym2 = O.windows_c_inter("MP2", nb, na, ob, oa)
        ref2 = O.transform_c_inter(Cb, Ca, rect_sw, win2, sym2)
        got2 = read_moint_quads(str(tmp_path / "H_1.E-moint.dat"), S)
        assert len(got2[4]) == cnt2
        assert np.abs(dense_quads(*got2, nb, na) - dense_quads(*ref2, nb, na)).max() <= 1e-10
        # same physical integrals, transposed roles
        assert np.abs(dense_quads(*got2, nb, na).transpose(2, 3, 0, 1) - dense_quads(*got, na, nb)).max() <= 1e-10
    else:
        win = O.windows_e_inter("MP2", na, nb, oa, ob)
        ref = O.transform_e_inter(Ca, Cb, rect, win)
        got = read_moint_pairs(str(tmp_path / "E-.H_1moint.dat"), S)
        assert len(got[2]) == cnt
        Ma, Mb = O.npairs(na), O.npairs(nb)
        assert np.abs(dense_pairs(*got, Ma, Mb) - dense_pairs(*ref, Ma, Mb)).max() <= 1e-10
        assert_lists_match(got[:2], got[2], ref[:2], ref[2])


# ---------------------------------------------------------------------------------------------
# the program's species loop and its division among devices (IntegralTransformation.f90:171-355)
# ---------------------------------------------------------------------------------------------
def _apmo_species(coeffs=None):
    """H2O.APMO shapes: electrons N=19 occ 5, two protons N=50 occ 1 each."""
    shapes = [("E-", 19, 5), ("H-A_1", 50, 1), ("H-B_1", 50, 1)]
    return [capi.host_species(nm, i + 1, n, occ, coeff=None if coeffs is None else coeffs[i]) for i, (nm, n, occ) in enumerate(shapes)]


def test_program_plan_order_and_method_c_pair_swap():
    sp = _apmo_species()
    plan_c = capi.host_plan_program(capi.host_control("C", "MP2"), sp)
    # program order: species i, then pairs (i, j>i); C calls with the species of fewer occupied orbitals first (:322-334)
    assert [(t["first"], t["second"]) for t in plan_c] == [(0, None), (1, 0), (2, 0), (1, None), (1, 2), (2, None)]
    plan_e = capi.host_plan_program(capi.host_control("E", "MP2"), sp)
    assert [(t["first"], t["second"]) for t in plan_e] == [(0, None), (0, 1), (0, 2), (1, None), (1, 2), (2, None)]
    for t in plan_c + plan_e:
        assert t["rank"] == 0 and t["flops"] > 0
    # windows are those of lowdin_host_windows for the call order
    assert plan_c[1]["win"] == capi.host_windows(capi.host_control("C", "MP2"), sp[1], sp[0])[0]
    # flop model of the electronic MP2 call: 2 N Q (N+P) M + 2 N S (N+R) n_ij, transformer-E roles
    n, occ, M = 19, 5, 190
    want = 2.0 * n * occ * (n + n - occ) * M + 2.0 * n * occ * (n + n - occ) * (occ * (n - occ))
    assert plan_e[0]["flops"] == want


def test_program_plan_pt2_species_filter():
    sp = _apmo_species()
    # PT2 with IONIZE_SPECIES: only the named species and the pairs that contain it (:176-185, :259-269)
    plan = capi.host_plan_program(capi.host_control("E", "PT2", ionize_species=("H-A_1",)), sp)
    assert [(t["first"], t["second"]) for t in plan] == [(0, 1), (1, None), (1, 2)]
    # without IONIZE_SPECIES every species and pair is transformed, also under PT2
    assert len(capi.host_plan_program(capi.host_control("E", "PT2"), sp)) == 6
    # the filter applies to PT2 only
    assert len(capi.host_plan_program(capi.host_control("E", "MP2-PT2", ionize_species=("H-A_1",)), sp)) == 6


@pytest.mark.parametrize("nranks", [1, 2, 3, 4, 8])
def test_program_plan_assignment_is_balanced_and_complete(nranks):
    sp = _apmo_species()
    plan = capi.host_plan_program(capi.host_control("C", "MP2"), sp, nranks)
    load = np.zeros(nranks)
    for t in plan:
        assert 0 <= t["rank"] < nranks
        load[t["rank"]] += t["flops"]
    total, biggest = load.sum(), max(t["flops"] for t in plan)
    # longest-processing-time-first: no rank exceeds the mean by more than one call
    assert load.max() <= total / nranks + biggest
    if nranks <= len(plan):
        assert (load > 0).all()
    # deterministic: a second evaluation gives the same plan (every rank computes it independently)
    assert plan == capi.host_plan_program(capi.host_control("C", "MP2"), sp, nranks)


def test_program_plan_errors():
    sp = _apmo_species()
    with pytest.raises(capi.LowdinITError):
        capi.host_plan_program(capi.host_control("C", "MP2"), sp, 0)
    with pytest.raises(capi.LowdinITError):
        capi.host_plan_program(capi.host_control("X", "MP2"), sp, 1)


def test_run_program_apmo_shapes_file_to_file(O, HT, tmp_path):
    T = HT
    """Three species (small stand-ins of the H2O.APMO layout), method E, MP2: the whole species loop through
    lowdin_host_run_program, every moint.dat against the oracle."""
    shapes = [("E-", 7, 3), ("H-A_1", 5, 1), ("H-B_1", 4, 1)]
    S = 40
    Cs = [O.random_orthonormal(n, 31 + i) for i, (_, n, _) in enumerate(shapes)]
    sp = [capi.host_species(nm, i + 1, n, occ, coeff=Cs[i]) for i, (nm, n, occ) in enumerate(shapes)]
    intra, inter = {}, {}
    for i, (nm, n, _) in enumerate(shapes):
        intra[i] = O.hash_packed_intra(100 + i, n)
        _split_to_files(tmp_path, nm, O.canonical_list_intra(intra[i], n), 2, S)
        for j in range(i + 1, len(shapes)):
            inter[i, j] = O.hash_rect_inter(200 + 10 * i + j, n, shapes[j][1])
            _split_to_files(tmp_path, f"{nm}.{shapes[j][0]}", O.canonical_list_inter(inter[i, j], n, shapes[j][1]), 2, S)
    ctl = capi.host_control("E", "MP2", stack=S, nfiles=2, scratch_dir=str(tmp_path))
    written, calls = capi.host_run_program(T, ctl, sp)
    assert calls == 6
    total = 0
    for i, (nm, n, occ) in enumerate(shapes):
        ref = O.transform_e_intra(Cs[i], intra[i], O.windows_e_intra("MP2", n, occ))
        got = read_moint_pairs(str(tmp_path / f"{nm}moint.dat"), S)
        M = O.npairs(n)
        assert np.abs(dense_pairs(*got, M, M) - dense_pairs(*ref, M, M)).max() <= 1e-10
        total += len(got[2])
        for j in range(i + 1, len(shapes)):
            nm2, n2, occ2 = shapes[j]
            ref = O.transform_e_inter(Cs[i], Cs[j], inter[i, j], O.windows_e_inter("MP2", n, n2, occ, occ2))
            got = read_moint_pairs(str(tmp_path / f"{nm}.{nm2}moint.dat"), S)
            assert np.abs(dense_pairs(*got, M, O.npairs(n2)) - dense_pairs(*ref, M, O.npairs(n2))).max() <= 1e-10
            total += len(got[2])
    assert total == written
    # a second "rank" of two runs only its share and the shares are disjoint and complete
    plan = capi.host_plan_program(ctl, sp, 2)
    for r in (0, 1):
        _, c = capi.host_run_program(T, ctl, sp, r, 2)
        assert c == sum(1 for t in plan if t["rank"] == r)
    # the one-process group form of the loop (here a group of one handle) writes the same files
    names = sorted(f for f in os.listdir(tmp_path) if f.endswith("moint.dat"))
    before = {f: read_moint_pairs(str(tmp_path / f), S) for f in names}
    for f in names:
        os.remove(tmp_path / f)
    written_g, calls_g = capi.host_group_run_program([T], ctl, sp)
    assert (written_g, calls_g) == (written, 6) and len(names) == 6
    for f in names:   # same records: pair ids in the same order, values to rounding
        a, b = before[f], read_moint_pairs(str(tmp_path / f), S)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.abs(a[2] - b[2]).max() <= 1e-13


# ---------------------------------------------------------------------------------------------
# lowdin.wfn labelled records and transformer D's record layout (CPU; scipy's FortranFile is the independent reader/writer)
# ---------------------------------------------------------------------------------------------
def _write_wfn_like_multiscf(path, blocks, label_len=30):
    """MultiSCF_saveWfn (MultiSCF.f90:1346-1389): write(unit) labels(1); write(unit) labels(2); write(unit) int(size,8);
    write(unit) values -- through scipy.io.FortranFile, i.e. not through the code under test."""
    from scipy.io import FortranFile
    f = FortranFile(str(path), "w")
    for label, species, values in blocks:
        f.write_record(np.frombuffer(label.ljust(label_len).encode(), dtype="S1"))
        f.write_record(np.frombuffer(species.ljust(label_len).encode(), dtype="S1"))
        f.write_record(np.array([np.asarray(values).size], dtype=np.int64))
        f.write_record(np.asarray(values, dtype=np.float64).reshape(-1, order="F"))
    f.close()


def test_wfn_reader_follows_matrix_get_from_file(tmp_path):
    rng = np.random.default_rng(3)
    Ce, Ch = rng.standard_normal((5, 5)), rng.standard_normal((4, 4))
    ee, eh = np.sort(rng.standard_normal(5)), np.sort(rng.standard_normal(4))
    path = tmp_path / "lowdin.wfn"
    _write_wfn_like_multiscf(path, [
        ("EXCHANGE-CORRELATION", "E-", np.zeros((5, 5))), ("EXCHANGE-CORRELATION-ENERGY", "E-", [0.25]),
        ("COEFFICIENTS", "E-", Ce), ("DENSITY", "E-", Ce @ Ce.T), ("ORBITALS", "E-", ee),
        ("COEFFICIENTS", "H_1", Ch), ("DENSITY", "H_1", Ch @ Ch.T), ("ORBITALS", "H_1", eh)])
    assert np.array_equal(capi.host_wfn_read(str(path), "COEFFICIENTS", "E-").reshape(5, 5, order="F"), Ce)
    assert np.array_equal(capi.host_wfn_read(str(path), "COEFFICIENTS", "H_1").reshape(4, 4, order="F"), Ch)   # skips the E- block
    assert np.array_equal(capi.host_wfn_read(str(path), "ORBITALS", "H_1"), eh)
    assert np.array_equal(capi.host_wfn_read(str(path), "EXCHANGE-CORRELATION-ENERGY", "E-"), [0.25])
    # the first-record test is a PREFIX test (Matrix.f90:716-722): the matrix block is found because it comes first
    assert capi.host_wfn_read(str(path), "EXCHANGE-CORRELATION", "E-").size == 25
    sp = capi.host_wfn_load_species(str(path), "H_1", 2, 4, 1)
    assert (sp.ldc, sp.ncols) == (4, 4) and np.array_equal(sp._keep, Ch) and np.array_equal(sp.eps, eh)
    with pytest.raises(capi.LowdinITError, match="End of file"):
        capi.host_wfn_read(str(path), "COEFFICIENTS", "POSITRON")
    with pytest.raises(capi.LowdinITError, match="dimensions of the matrix COEFFICIENTS"):
        capi.host_wfn_load_species(str(path), "H_1", 2, 5, 1)          # asks for 5x5, the file holds 16 values


def test_wfn_writer_produces_gfortran_records(tmp_path):
    from scipy.io import FortranFile
    path = str(tmp_path / "w.wfn")
    Cm = np.arange(12.0).reshape(3, 4)
    capi.host_wfn_append(path, "COEFFICIENTS", "E-ALPHA", Cm, truncate=True)
    capi.host_wfn_append(path, "ORBITALS", "E-ALPHA", [1.0, 2.0, 3.0])
    f = FortranFile(path, "r")
    assert f.read_record("S1").tobytes() == b"COEFFICIENTS".ljust(30)
    assert f.read_record("S1").tobytes() == b"E-ALPHA".ljust(30)
    assert f.read_ints(np.int64)[0] == 12
    assert np.array_equal(f.read_reals(np.float64), Cm.reshape(-1, order="F"))     # column by column (Matrix.f90:599)
    assert f.read_record("S1").tobytes() == b"ORBITALS".ljust(30)
    f.close()
    assert np.array_equal(capi.host_wfn_read(path, "ORBITALS", "E-ALPHA"), [1.0, 2.0, 3.0])


def _read_d_records(path):
    raw = open(path, "rb").read()
    rec = np.frombuffer(raw, dtype=np.dtype([("h", "<u4"), ("i", "<i4", 4), ("v", "<f8"), ("t", "<u4")]))
    assert (rec["h"] == 24).all() and (rec["t"] == 24).all()
    return rec["i"], rec["v"]


def test_moint_d_layout_intra_and_inter(O, tmp_path):
    n, on = 5, 3
    M, oM = O.npairs(n), O.npairs(on)
    ints = np.arange(M * (M + 1) // 2, dtype=np.float64) + 0.5
    path = str(tmp_path / "E-moint.dat")
    cnt = capi.host_write_moint_d(path, ints, n)
    idx, v = _read_d_records(path)
    assert cnt == M * (M + 1) // 2 == len(v) - 1
    assert list(idx[-1]) == [-1, 0, 0, 0] and v[-1] == 0.0                            # TransformIntegralsD.f90:266
    want = [(p, q, r, s) for p in range(1, n + 1) for q in range(1, p + 1) for r in range(1, p + 1)
            for s in range(1, (q if p == r else r) + 1)]                                 # :221-236
    assert [tuple(x) for x in idx[:-1]] == want
    L = O.lib()
    assert all(v[k] == ints[L.orc_d_multi_index(p - 1, q - 1, r - 1, s - 1)] for k, (p, q, r, s) in enumerate(want))
    rect = np.arange(M * oM, dtype=np.float64) - 7.0
    path2 = str(tmp_path / "E-.H_1moint.dat")
    cnt2 = capi.host_write_moint_d(path2, rect, n, on)
    idx2, v2 = _read_d_records(path2)
    assert cnt2 == M * oM == len(v2) - 1 and list(idx2[-1]) == [-1, 0, 0, 0]
    lt = lambda i, j: max(i, j) * (max(i, j) + 1) // 2 + min(i, j)
    k = 0
    for p in range(1, n + 1):                                                            # :447-457
        for q in range(p, n + 1):
            for r in range(1, on + 1):
                for s in range(r, on + 1):
                    assert tuple(idx2[k]) == (p, q, r, s) and v2[k] == rect[lt(p - 1, q - 1) * oM + lt(r - 1, s - 1)]
                    k += 1


# ---------------------------------------------------------------------------------------------
# method D through the host mirror
# ---------------------------------------------------------------------------------------------
def test_method_d_window_tables_and_plan():
    e = capi.host_species("E-", 1, 19, 5)
    h = capi.host_species("H_1", 2, 50, 1)
    # TransformIntegralsD.f90:606-626: the intra table is 0-based; :668-690: the inter table is 1-based
    assert capi.host_windows(capi.host_control("D", "MP2"), e) == ([0, 4, 5, 18, 0, 4, 5, 18], False)
    assert capi.host_windows(capi.host_control("D", "ALL"), e) == ([0, 18, 0, 18, 0, 18, 0, 18], False)
    assert capi.host_windows(capi.host_control("D", "MP2"), e, h) == ([1, 5, 6, 19, 1, 1, 2, 50], False)
    assert capi.host_windows(capi.host_control("D", "BOUNDS"), e, h) == ([1, 19, 1, 19, 1, 50, 1, 50], False)
    plan = capi.host_plan_program(capi.host_control("D", "MP2"), [e, h])
    assert [(t["first"], t["second"]) for t in plan] == [(0, None), (0, 1), (1, None)]
    # D transforms everything whatever the windows say: 8 M N^3 (F_full)
    assert plan[0]["flops"] == 8.0 * 190 * 19 ** 3 and plan[2]["flops"] == 8.0 * 1275 * 50 ** 3


def test_method_d_without_gpu_fails_loudly(O, tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    n = 5
    capi.host_write_ints_file(str(tmp_path / "0E-.ints"), 16, *O.canonical_list_intra(O.hash_packed_intra(1, n), n))
    sp = capi.host_species("E-", 1, n, 2, coeff=O.random_orthonormal(n, 3))
    with pytest.raises(capi.LowdinITError, match="no CUDA device"):
        capi.host_transform_one_species(None, capi.host_control("D", "ALL", stack=16, scratch_dir=str(tmp_path)), sp)


def _read_d_file(path):
    raw = open(path, "rb").read()
    rec = np.frombuffer(raw, dtype=np.dtype([("h", "<u4"), ("i", "<i4", 4), ("v", "<f8"), ("t", "<u4")]))
    assert (rec["h"] == 24).all() and (rec["t"] == 24).all() and rec["i"][-1][0] == -1
    return rec["i"][:-1], rec["v"][:-1]


def test_method_d_file_to_file(O, HT, tmp_path):
    """.ints streams -> ReadIntegrals packing on the host -> lowdin_it_transform_all / _inter_all -> D's record file, against the
    reference's own transformer D (oracle/_ref) on the same packed array."""
    n, on, S = 9, 6, 32
    M, oM = O.npairs(n), O.npairs(on)
    Ca, Cb = O.random_orthonormal(n, 5), O.random_orthonormal(on, 6)
    packed = O.hash_packed_intra(41, n)
    lst = O.canonical_list_intra(packed, n)
    _split_to_files(tmp_path, "E-", lst, 2, S)
    rect = O.hash_rect_inter(42, n, on)
    lsti = O.canonical_list_inter(rect, n, on)
    _split_to_files(tmp_path, "E-.H_1", lsti, 2, S)
    ctl = capi.host_control("D", "ALL", stack=S, nfiles=2, scratch_dir=str(tmp_path))
    A, B = capi.host_species("E-", 1, n, 3, coeff=Ca), capi.host_species("H_1", 2, on, 1, coeff=Cb)
    use_ref = O.ref() is not None
    lt = lambda i, j: np.maximum(i, j) * (np.maximum(i, j) + 1) // 2 + np.minimum(i, j)
    # intra
    cnt = capi.host_transform_one_species(None, ctl, A)
    assert cnt == M * (M + 1) // 2
    eris = np.zeros(M * (M + 1) // 2)
    p, q, r, s, v = [np.asarray(x) for x in lst]
    eris[lt(lt(p - 1, q - 1), lt(r - 1, s - 1))] = v                  # ReadIntegrals_index4Intra
    want = O.transform_d_intra(Ca, eris, use_reference=use_ref)
    idx, val = _read_d_file(str(tmp_path / "E-moint.dat"))
    got = np.zeros_like(want)
    got[lt(lt(idx[:, 0] - 1, idx[:, 1] - 1), lt(idx[:, 2] - 1, idx[:, 3] - 1))] = val
    assert np.abs(got - want).max() <= 1e-10
    # inter
    cnt = capi.host_transform_two_species(None, ctl, A, B)
    assert cnt == M * oM
    er = np.zeros(M * oM)
    p, q, r, s, v = [np.asarray(x) for x in lsti]
    er[lt(p - 1, q - 1) * oM + lt(r - 1, s - 1)] = v                   # ReadIntegrals_index4Inter
    want = O.transform_d_inter(Ca, Cb, er, use_reference=use_ref)
    idx, val = _read_d_file(str(tmp_path / "E-.H_1moint.dat"))
    got = np.zeros_like(want)
    got[lt(idx[:, 0] - 1, idx[:, 1] - 1) * oM + lt(idx[:, 2] - 1, idx[:, 3] - 1)] = val
    assert np.abs(got - want).max() <= 1e-10
