#!/usr/bin/env python
"""bench.py -- four-index AO->MO transform throughput on B200 (BASELINE.json metric).

Workload (config.workload): synthetic kind-H AO integrals (generated on the device from a counter
hash) + random orthonormal coefficients, N_bf basis functions (default 1500),
O = N/10 occupied, MP2 window in transformer-E roles (p,r in virt; q,s in occ).  The (i a|mu nu)
half-transformed block of N=1500 is 1.8 TB, so the transform runs one OCCUPIED BATCH at a time
(lowdin_it_transform_stream); a "step" is one such pass: first half over all M AO-pair slabs for
`occ_batch` occupied orbitals (56 at N=1500 on one B200: 3 passes = the whole transform), third
quarter accumulated chunk by chunk, fourth quarter, results consumed on the device
(count / sums / MP2 pair energy).  GFLOP/s uses the ALGORITHMIC two-half flop count
F = 2 N Qb (N+P) n_pq + 2 N S (N+R) n_ij of SURVEY.md 8d -- no credit for padding or redundancy.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--nbf 1500] [--occ-batch 0] [--impl reference]

N>1: launched by torchrun, one rank per GPU; first half sharded over AO pair slabs, one NCCL
all-to-all, second half sharded over MO pairs (strong scaling: the job is fixed).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEED = 20261017
GEN_KIND = 1


def mp2_window_e(n, occ):
    # TransformIntegralsE.f90:1938-1949: p,r in [occ+1,N]; q,s in [1,occ]
    return [occ + 1, n, 1, occ, occ + 1, n, 1, occ]


def algorithmic_flops_pass(n, occ, qb, nslabs_frac=1.0):
    P = n - occ
    M = n * (n + 1) // 2
    first = 2.0 * n * qb * (n + P) * M * nslabs_frac
    second = 2.0 * n * occ * (n + P) * (P * qb) * nslabs_frac
    return first + second


def random_orthonormal(n, seed):
    q, _ = np.linalg.qr(np.random.default_rng(seed).standard_normal((n, n)))
    return np.asfortranarray(q)


def synthetic_eps(occ, n):
    return np.concatenate([np.linspace(-2.0, -0.5, occ), np.linspace(0.2, 3.0, n - occ)])


def splitmix_values(seed, keys):
    """kind-H AO values: 2u-1, u = (splitmix64(seed ^ key) >> 11) * 2^-53, on a uint64 array of keys."""
    with np.errstate(over="ignore"):
        x = (np.uint64(seed) ^ keys) + np.uint64(0x9E3779B97F4A7C15)
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        x = x ^ (x >> np.uint64(31))
    return 2.0 * ((x >> np.uint64(11)).astype(np.float64) * 2.0 ** -53) - 1.0


def canonical_ao_list(seed, n, out=None):
    """The kind-H intra-species AO tensor of an n-function basis as the .ints list lowdin-ints writes
    (Iterators.cpp:45-77: i>=j, k>=l, (ij)>=(kl), 1-based).  Returns p,q,r,s (int32) and v (float64), M(M+1)/2 entries;
    `out` = five preallocated (e.g. pinned) arrays to fill.  The value is a function of the two 0-based row-wise upper-triangular
    pair ids only, key = max*M+min (the same keys as the device generator).  No value is within 1e-10 of zero in practice, so
    the |v|>1e-10 filter of Libint2Iface.cpp:369 keeps the whole list."""
    M = n * (n + 1) // 2
    ti, tj = np.tril_indices(n)              # lower-triangular compound index t = i(i+1)/2+j -> (i, j), i >= j
    xy = (tj * n - tj * (tj - 1) // 2 + (ti - tj)).astype(np.uint64)   # the same pair's row-wise upper-triangular id
    total = M * (M + 1) // 2
    if out is None:
        out = [np.empty(total, np.int32) for _ in range(4)] + [np.empty(total, np.float64)]
    p, q, r, s, v = out
    o = 0
    rows = max(1, (1 << 22) // M)            # blocks of about 4 M entries
    for a in range(0, M, rows):
        b = min(M, a + rows)
        T1, T2 = np.tril_indices(b, 0, b)    # (ij) in [0,b), (kl) <= (ij)
        keep = T1 >= a
        T1, T2 = T1[keep], T2[keep]
        m = len(T1)
        p[o:o + m] = ti[T1] + 1; q[o:o + m] = tj[T1] + 1; r[o:o + m] = ti[T2] + 1; s[o:o + m] = tj[T2] + 1
        k1, k2 = xy[T1], xy[T2]
        v[o:o + m] = splitmix_values(seed, np.maximum(k1, k2) * np.uint64(M) + np.minimum(k1, k2))
        o += m
    assert o == total
    return p, q, r, s, v


def raw_ints_blocks(lst, total, S, out):
    """The five list arrays as the bytes of a .ints stream (blocks of int32 p[S],q[S],r[S],s[S]; float64 v[S], the last block
    carrying p = -1 after its last entry; Libint2Iface.cpp:3414-3426) into the uint8 buffer `out` of (total // S + 1) * 24 S bytes."""
    nblk = total // S + 1
    blk = out[:nblk * 24 * S].reshape(nblk, 24 * S)
    for k in range(4):
        dst = blk[:, 4 * S * k:4 * S * (k + 1)].view(np.int32)
        full = (total // S) * S
        dst[:total // S] = lst[k][:full].reshape(-1, S)
        dst[total // S] = 0
        dst[total // S, :total - full] = lst[k][full:total]
    dst = blk[:, 16 * S:].view(np.float64)
    full = (total // S) * S
    dst[:total // S] = lst[4][:full].reshape(-1, S)
    dst[total // S] = 0.0
    dst[total // S, :total - full] = lst[4][full:total]
    blk[:, :4 * S].view(np.int32)[total // S, total - full] = -1
    return nblk


def stored_ao_e2e(torch, ol, capi, dev_index, n, occ, steps, push_entries=1 << 20, mode="stacks", stack_size=30000):
    """End to end through the C ABI with HOST buffers, stored AO integrals (the reference's own data flow): per step
    coefficients + the whole canonical AO list go host->device in the .ints stack layout (lowdin_it_ao_begin /
    _push_stacks / _end -> scatter), the transform runs (lowdin_it_transform, E convention, MP2 window) and every kept MO
    integral comes back (lowdin_it_download_pairs).  All host buffers are pinned.  The same transform with the AO values
    generated on the device from the same keys is run once, untimed, as a consistency check of the upload path.
    mode "stacks": the list as five arrays, push_entries per lowdin_it_ao_push_stacks call; mode "blocks": the same list as the
    raw bytes of a .ints file (stacks of stack_size entries) in ONE lowdin_it_ao_push_blocks call."""
    import ctypes as C
    M = n * (n + 1) // 2
    total = M * (M + 1) // 2
    win = np.ascontiguousarray(mp2_window_e(n, occ), dtype=np.int32)
    pin = lambda cnt, dt: torch.empty(cnt, dtype=dt).pin_memory().numpy()
    lst = canonical_ao_list(SEED, n, out=[pin(total + 1, torch.int32) for _ in range(4)] + [pin(total + 1, torch.float64)])
    lst[0][total] = -1                        # terminator of the last stack (C.f90:279-280)
    Cm = random_orthonormal(n, n)
    Cpin = torch.from_numpy(np.ascontiguousarray(Cm.T)).pin_memory().numpy().T   # column-major C(mu,p), pinned
    P = n - occ
    cap = (P * occ) ** 2
    o_ij, o_kl, o_v = pin(cap, torch.int64), pin(cap, torch.int64), pin(cap, torch.float64)
    T = ol.Transformer(dev_index)
    L, h = T.L, T.h
    cnt = C.c_int64()
    if mode == "blocks":
        raw = pin((total // stack_size + 1) * 24 * stack_size, torch.uint8)
        nblk = raw_ints_blocks(lst, total, stack_size, raw)

    host_ms = {"set_species": 0.0, "ao_upload": 0.0, "transform": 0.0, "download": 0.0}   # wall time of the calls as the host sees them

    def step():
        t_a = time.perf_counter()
        T.set_species(0, Cpin)
        t_b = time.perf_counter()
        T._ck(L.lowdin_it_ao_begin(h, 0, 0, 0))
        if mode == "blocks":
            T._ck(L.lowdin_it_ao_push_blocks(h, raw.ctypes.data, nblk, stack_size))
        else:
            for a in range(0, total + 1, push_entries):
                b = min(total + 1, a + push_entries)
                T._ck(L.lowdin_it_ao_push_stacks(h, lst[0][a:b], lst[1][a:b], lst[2][a:b], lst[3][a:b], lst[4][a:b], b - a))
        T._ck(L.lowdin_it_ao_end(h))
        t_c = time.perf_counter()
        T._ck(L.lowdin_it_transform(h, 0, 0, win, capi.CONV_E, 0, 1e-10))
        t_d = time.perf_counter()
        T._ck(L.lowdin_it_result_count(h, C.byref(cnt)))
        if cnt.value > cap:
            raise RuntimeError("more results than window pairs")
        T._ck(L.lowdin_it_download_pairs(h, o_ij, o_kl, o_v))
        t_e = time.perf_counter()
        for k, dt_ in (("set_species", t_b - t_a), ("ao_upload", t_c - t_b), ("transform", t_d - t_c), ("download", t_e - t_d)):
            host_ms[k] += dt_ * 1e3
        return cnt.value

    step()                                    # warm-up (allocations, first-launch costs)
    torch.cuda.synchronize()
    for k in host_ms:
        host_ms[k] = 0.0
    t0 = time.perf_counter()
    for _ in range(steps):
        kept = step()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / steps
    tm = T.timers()
    stored = o_v[:kept].copy()
    # consistency: same keys generated on the device (fused first quarter instead of scatter + expansion)
    T.set_generator(0, 0, SEED, GEN_KIND)
    ij2, kl2, v2 = T.transform(0, 0, win, capi.CONV_E)
    same = (len(v2) == kept) and bool(np.array_equal(ij2, o_ij[:kept])) and bool(np.array_equal(kl2, o_kl[:kept]))
    diff = float(np.abs(v2 - stored).max()) if same and kept else None
    T.close()
    flops = 2.0 * n * occ * (n + P) * M + 2.0 * n * occ * (n + P) * (P * occ)
    return {"value": flops / dt / 1e9, "unit": "GFLOP/s", "steps": steps, "ms_per_step": dt * 1e3,
            "workload": f"N_bf={n} MP2 window O={occ} (C6H6/cc-pVDZ shape when N=120), kind-H AO list of {total} canonical integrals "
                        "pushed from pinned host memory in the .ints stack layout, all kept MO integrals downloaded",
            "push": ("raw .ints bytes, one lowdin_it_ao_push_blocks call" if mode == "blocks"
                     else f"five arrays, {push_entries} entries per lowdin_it_ao_push_stacks call"),
            "upload_gb_per_s": total * 24 / max(tm["ao_upload"], 1e-12) / 1e9,
            "h2d_bytes_per_step": int(total * 24 + n * n * 8), "d2h_bytes_per_step": int(kept * 24), "mo_integrals_kept": int(kept),
            "device_ms": {"ao_upload_scatter": tm["ao_upload"] * 1e3, "first_half": tm["first_half"] * 1e3,
                          "second_half": tm["second_half"] * 1e3, "compaction": tm["consume"] * 1e3, "download": tm["download"] * 1e3},
            "host_ms_per_call": {k: x / steps for k, x in host_ms.items()},
            "same_index_lists_as_generated": same, "max_abs_diff_vs_generated": diff,
            "_result": (o_ij[:kept].copy(), o_kl[:kept].copy(), stored)}


def stored_resident_leg(ol, dev_index, n, all_mode=False, hbm_gbs=6549.4, fp64_peak=None):
    """The STORED-AO kernels at scale: the packed kind-H tensor of an n-function basis (63 GB at N=500) is filled on the device
    (lowdin_it_ao_materialize: as if its list had been uploaded) and the whole MP2-window transform runs from it -- expansion of
    packed rows into dense slabs, the four quarter transforms, the device-side MP2 consumer.  Reports the per-kernel table and the
    agreement of the streamed sums with the generated-source path on the same keys."""
    occ = n // 10
    P, M = n - occ, n * (n + 1) // 2
    win = mp2_window_e(n, occ)
    Cm = random_orthonormal(n, n)
    eps = synthetic_eps(occ, n)
    T = ol.Transformer(dev_index)
    T.set_species(0, Cm)
    T.set_generator(0, 0, SEED, GEN_KIND)
    s_gen = T.transform_stream(0, 0, win, ol.CONV_E, epsA=eps)
    t0 = time.perf_counter()
    T.materialize(0, 0)
    fill_s = time.perf_counter() - t0
    out = {"workload": f"N_bf={n} MP2 window O={occ}, packed AO tensor of {M * (M + 1) // 2} doubles ({M * (M + 1) * 4 / 1e9:.1f} GB) resident in HBM",
           "fill_on_device_s": fill_s}

    def run(w, label):
        T.transform_stream(0, 0, w, ol.CONV_E, epsA=eps if label == "mp2" else None)       # warm-up
        T.set_profiling(True)
        sm = T.transform_stream(0, 0, w, ol.CONV_E, epsA=eps if label == "mp2" else None)
        tm, st = T.timers(), T.kernel_stats()
        T.set_profiling(False)
        sec = tm["first_half"] + tm["exchange"] + tm["second_half"] + tm["consume"]
        kern = {}
        for c, v in st.items():
            if v["launches"]:
                gem = c in ("q1", "q2", "q3", "q4")
                kern[c] = {"ms": round(v["ms"], 3), "launches": v["launches"], ("TFLOP/s" if gem else "GB/s"): v["work"] / (v["ms"] * 1e-3) / (1e12 if gem else 1e9)}
        res = {"value": tm["flops"] / sec / 1e9, "unit": "GFLOP/s", "ms_per_transform": sec * 1e3, "kernels": kern, "gpu_launches": tm["launches"]}
        if "expand1" in kern:
            res["roofline_expand1"] = {"bound": "hbm", "achieved": kern["expand1"]["GB/s"], "peak": hbm_gbs, "unit": "GB/s",
                                       "frac": kern["expand1"]["GB/s"] / hbm_gbs,
                                       "algorithmic_bytes": "8 M read + 8 N^2 written per slab (E.f90:1047-1063)"}
        if "q1" in kern and fp64_peak:
            res["roofline_q1"] = {"bound": "tensor", "achieved": kern["q1"]["TFLOP/s"], "peak": fp64_peak, "unit": "TFLOP/s",
                                  "frac": kern["q1"]["TFLOP/s"] / fp64_peak}
        return res, sm

    out["mp2"], s_st = run(win, "mp2")          # default: unpack fused into the first quarter (q1_load_ws5_kernel), no dense slab in HBM
    scale = np.array([1.0, np.sqrt(s_gen[0] * s_gen[2]), s_gen[2], max(abs(s_gen[3]), 1e-300)])
    out["mp2"]["sums"] = list(map(float, s_st))
    out["mp2"]["vs_generated_source"] = {"count_equal": bool(s_st[0] == s_gen[0]), "max_rel_diff": float((np.abs(s_st - s_gen) / scale)[1:].max())}
    out["mp2"]["first_quarter"] = "fused: packed rows -> shared memory -> DMMA (8 M bytes of HBM reads per slab)"
    T.set_option(T.OPT_STORED_FUSED, 0)        # the two-kernel form: expansion kernel (dense slab to HBM) + DMMA GEMM
    out["mp2_unfused"], s_un = run(win, "mp2")
    T.set_option(T.OPT_STORED_FUSED, 1)
    out["mp2_unfused"]["first_quarter"] = "expansion kernel writes the dense slab (8 M + 8 N^2 bytes per slab), TMA GEMM reads it back"
    out["mp2_unfused"]["vs_fused"] = {"count_equal": bool(s_un[0] == s_st[0]), "max_rel_diff": float((np.abs(s_un - s_st) / scale)[1:].max())}
    if all_mode:
        out["all"], _ = run([1, n] * 4, "all")
    T.close()
    return out


def whole_transform_cpu_leg(n, occ, gpu_result=None):
    """Part of the cpu_baseline leg: the oracle's restatements of the reference's transformers E (one thread, as in the reference)
    and C (OpenMP over p, as in the reference) on the WHOLE stored-AO workload of e2e_stored_ao, timed from packed AO integrals
    in memory to the MO-integral list, and the largest difference between the GPU's list and transformer E's."""
    from oracle import oracle as O
    Cm = O.random_orthonormal(n, n)
    packed = O.hash_packed_intra(SEED, n)
    we = mp2_window_e(n, occ)
    t0 = time.perf_counter()
    rij, rkl, rv = O.transform_e_intra(Cm, packed, we)
    te = time.perf_counter() - t0
    wc, sym = O.windows_c_intra("MP2", n, occ)
    t0 = time.perf_counter()
    rc = O.transform_c_intra(Cm, packed, wc, sym)
    tc = time.perf_counter() - t0
    P, M = n - occ, n * (n + 1) // 2
    flops = 2.0 * n * occ * (n + P) * M + 2.0 * n * occ * (n + P) * (P * occ)
    out = {"workload": f"N_bf={n} MP2 window O={occ}, whole transform from packed AO integrals in memory to the MO-integral list",
           "transformer_e_port_s": te, "transformer_e_port_gflops": flops / te / 1e9, "transformer_e_threads": 1,
           "transformer_c_port_s": tc, "transformer_c_port_gflops_same_flop_count": flops / tc / 1e9,
           "transformer_c_threads": min(os.cpu_count() or 1, occ), "mo_integrals_e": int(len(rv)), "mo_integrals_c": int(len(rc[4]))}
    if gpu_result is not None:
        gij, gkl, gv = gpu_result
        ref = np.zeros((M, M)); ref[rij - 1, rkl - 1] = rv
        got = np.zeros((M, M)); got[gij - 1, gkl - 1] = gv
        out["max_abs_diff_gpu_vs_transformer_e"] = float(np.abs(got - ref).max())
    return out


def transformer_d_leg(ol, n_full=120):
    """The reference's OWN transformer D (oracle/_ref/libref_d.so, compiled from IntTransfD.cpp) timed on the host cores
    beside the drop-in lowdin_it_transform_all on the GPU: same call (coeff, ints in place, nao), same host buffers, full
    transform of a seeded symmetric tensor at the largest of the reference's configurations (C6H6/cc-pVDZ shape, N=120).
    Part of the cpu_baseline leg (the one place the product arm may run oracle/)."""
    from oracle import oracle as O
    R = O.ref()
    if R is None:
        return {"unavailable": "oracle/_ref/libref_d.so not built"}
    blas = O.bind_openblas()
    n = n_full if blas else 60                      # the 15-line dgemm_ shim is slow: smaller sample without OpenBLAS
    M = n * (n + 1) // 2
    eris = np.random.default_rng(n).uniform(-1.0, 1.0, M * (M + 1) // 2)
    Cm = random_orthonormal(n, n)
    ref = eris.copy()
    t0 = time.perf_counter()
    R.c_integrals_transform_all(Cm, ref, n)
    t_ref = time.perf_counter() - t0
    ol.transform_all(Cm, eris.copy())               # warm-up: context, allocations
    got = eris.copy()
    t0 = time.perf_counter()
    ol.transform_all(Cm, got)
    t_gpu = time.perf_counter() - t0
    flops = 8.0 * M * float(n) ** 3                 # what transformer D executes (IntTransfD.cpp:145-178)
    return {"workload": f"N_bf={n} full in-place transform, c_integrals_transform_all(coeff, ints, nao) semantics (IntTransfD.h:66), "
                        f"{M * (M + 1) // 2} packed integrals in and out through host buffers",
            "reference_ms": t_ref * 1e3, "reference_gflops": flops / t_ref / 1e9, "reference_cores": os.cpu_count() or 1,
            "reference_blas": "OpenBLAS dgemm_ (bundled with opencv)" if blas else "naive dgemm_ shim (oracle/blas_shim.c)",
            "gpu_ms": t_gpu * 1e3, "gpu_gflops": flops / t_gpu / 1e9, "max_abs_diff": float(np.abs(got - ref).max()),
            "h2d_bytes": int(eris.nbytes + Cm.nbytes), "d2h_bytes": int(eris.nbytes)}


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        self.idx = gpu_index

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                       "-lms", "200", "-i", str(self.idx)], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2])); pw.append(float(c[3]))
            except ValueError:
                continue
            for nm, val in zip(names, c[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        os.unlink(self.f.name)
        # "under load" = samples drawing more than half of the max power seen
        load = [s for s, p in zip(sm, pw) if pw and p >= 0.5 * max(pw)] or sm
        return {"sm_mhz": statistics.median(load) if load else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def cpu_sample(n, occ, nslabs, nthreads):
    """The oracle's restatement of transformer E's first half (E.f90:1043-1132) on `nslabs` slabs."""
    from oracle import oracle as O
    Cm = O.random_orthonormal(n, n)
    win = mp2_window_e(n, occ)
    t0 = time.perf_counter()
    chk = O.e_first_half_sample(SEED, Cm, win, 0, nslabs, nthreads)
    dt = time.perf_counter() - t0
    flops = 2.0 * n * occ * (n + (n - occ)) * nslabs
    return flops / dt / 1e9, dt, chk


def first_half_parity(T, ol, n, occ, gen, nthreads, nslabs=8):
    """Half-transformed integrals (i a|pq) of the first and the last `nslabs` AO-pair slabs: device (lowdin_it_debug_first_half, the
    kernels of the timed path) against the oracle port of E.f90:1043-1132 on the same slabs."""
    from oracle import oracle as O
    Cm = O.random_orthonormal(n, n)
    win = mp2_window_e(n, occ)
    M = n * (n + 1) // 2
    worst, cnt = 0.0, 0
    for pq0 in (0, M - nslabs):
        got = T.debug_first_half(0, 0, win, ol.CONV_E, pq0, nslabs)
        ref = O.e_first_half_values(SEED, Cm, win, pq0, nslabs, nthreads, gen_kind=gen)
        worst = max(worst, float(np.abs(got - ref).max()))
        cnt += got.size
    return {"max_abs_diff": worst, "values_compared": cnt, "slabs": f"first and last {nslabs} of {M}", "tolerance": 1e-10, "ok": worst <= 1e-10}


def cpu_sample_timed(n, occ, nthreads, target_s):
    """Calibrate on one slab per thread, then run a sample sized for about target_s seconds."""
    v, dt, _ = cpu_sample(n, occ, max(1, nthreads), nthreads)
    per_round = max(dt, 1e-3)
    rounds = int(max(1, min(400, target_s / per_round)))
    nsl = max(1, nthreads) * rounds
    v, dt, _ = cpu_sample(n, occ, nsl, nthreads)
    return v, dt, nsl


def run_reference(args):
    """--impl reference: the reference's CPU transformer (oracle port of transformer E: the reference's
    Fortran cannot be compiled here and its C++ transformer D overflows 32-bit indices past N~300)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n, occ = args.nbf, args.nbf // 10
    nthreads = os.cpu_count() or 1
    _, dt1, _ = cpu_sample(n, occ, max(1, nthreads), nthreads)
    per_step = max(1, nthreads) * int(max(1, min(200, 4.0 / max(dt1, 1e-3))))  # ~4 s of CPU work per step
    for _ in range(args.warmup):
        cpu_sample(n, occ, per_step, nthreads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_sample(n, occ, per_step, nthreads)
    dt = time.perf_counter() - t0
    flops = 2.0 * n * occ * (n + (n - occ)) * per_step * args.steps
    val = flops / dt / 1e9
    sample = (f"first half of transformer E (E.f90:1043-1132, oracle port), {per_step} of {n*(n+1)//2} AO-pair slabs per step, "
              f"full occupied window, {nthreads} threads over independent slabs")
    line = {"impl": "reference", "metric": "4-index transform FP64 GFLOP/s", "value": val, "unit": "GFLOP/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"N_bf={n} MP2 window O={occ} (transformer-E roles), kind-H synthetic AO, random orthonormal C; "
                                   f"step = a bounded sample of {per_step} AO-pair slabs of the same transform on the host cores",
                       "nbf": n, "occ": occ},
            "cpu_baseline": {"value": val, "unit": "GFLOP/s", "cores": nthreads, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def bind_to_gpu_numa_node(torch, local):
    """Pin this process to the CPUs next to its GPU (sysfs local_cpulist of the device's PCI function), so that the pinned host
    buffers it allocates afterwards -- the sink's ring that receives every MO integral -- sit on the GPU's NUMA node: with eight
    ranks of one box delivering 41 GB each, buffers on the far socket make the slowest rank's copies cross the inter-socket link.
    Returns a description, or None when nothing was done (no sysfs entry, one node, restricted cpuset)."""
    try:
        p = torch.cuda.get_device_properties(local)
        bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        cpus = set()
        for part in open(f"/sys/bus/pci/devices/{bdf}/local_cpulist").read().strip().split(","):
            if part:
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if cpus and len(cpus) < len(allowed):
            os.sched_setaffinity(0, cpus)
            return f"{len(cpus)} of {len(allowed)} CPUs, local to {bdf}"
    except Exception:  # noqa: BLE001  (best effort: never fail the bench over placement)
        pass
    return None


def cublas_fp64_peak(torch, dev, n=4096, iters=6):
    a = torch.rand(n, n, dtype=torch.float64, device=dev)
    b = torch.rand(n, n, dtype=torch.float64, device=dev)
    torch.matmul(a, b)
    torch.cuda.synchronize(dev)
    best = 1e30
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b); e1.record()
        torch.cuda.synchronize(dev)
        best = min(best, e0.elapsed_time(e1))
    del a, b
    torch.cuda.empty_cache()
    return 2.0 * n ** 3 / (best * 1e-3) / 1e12


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--nbf", type=int, default=1500)
    ap.add_argument("--occ-batch", type=int, default=0)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the end-to-end leg (profiling runs)")
    ap.add_argument("--stored-nbf", type=int, default=120, help="basis size of the stored-AO end-to-end leg (0 = skip)")
    ap.add_argument("--stored-only", action="store_true", help="run only the stored-AO end-to-end leg and print it (profiling / quick checks)")
    ap.add_argument("--push-mode", default="blocks", help="--stored-only: 'blocks' (raw .ints bytes) or 'stacks' (five arrays)")
    ap.add_argument("--resident-nbf", type=int, default=500, help="basis size of the stored-AO-resident leg (packed tensor filled on the device; 0 = skip)")
    ap.add_argument("--resident-all", action="store_true", help="resident leg: also the ALL (full) window")
    ap.add_argument("--resident-only", action="store_true", help="run only the stored-AO-resident leg and print it")
    ap.add_argument("--gen", type=int, default=GEN_KIND, help="synthetic AO generator: 1 = kind H (splitmix64), 2 = kind F (mul-fold-mul)")
    ap.add_argument("--q1-variant", type=int, default=0, help="fused first-quarter kernel variant (0 = library default)")
    ap.add_argument("--gemm-variant", type=int, default=0, help="quarter-transform GEMM variant (0 = library default)")
    ap.add_argument("--frag-perm", type=int, default=-1, help="fragment-row permutation of the TMA kernels: 0 / 1 (-1 = library default)")
    ap.add_argument("--overlap", type=int, default=-1, help="N>1: all-to-all of chunk c under the first half of chunk c+1: 0 / 1 (-1 = library default)")
    ap.add_argument("--gemm-tall", type=int, default=-1, help="192 x 64 tiles for the second / fourth quarter when the rows are a multiple of 192 plus a few: 0 / 1 (-1 = library default)")
    ap.add_argument("--numa-bind", type=int, default=-1, help="pin the process to the CPUs local to its GPU before allocating pinned buffers: 0 / 1 (-1 = when N > 1)")
    ap.add_argument("--exchange-dma", type=int, default=-1, help="N>1 on one node: all-to-all as peer-to-peer DMA (1) or ncclSend/ncclRecv (0); -1 = library default")
    ap.add_argument("--q3-two-cta", type=int, default=-1, help="third quarter: products with K <= this value as two 4-warp CTAs per SM (0 = off, -1 = library default)")
    ap.add_argument("--q3-red", type=int, default=-1, help="third-quarter accumulation by red.global.add.f64: 0 / 1 (-1 = library default)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import openlowdin_b200 as ol
    from openlowdin_b200 import capi

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.resident_only:
        print(json.dumps({"stored_ao_resident": stored_resident_leg(ol, local, args.resident_nbf or 500, args.resident_all)}))
        return
    if args.stored_only:
        n_st = args.stored_nbf or 120
        res = stored_ao_e2e(torch, ol, capi, local, n_st, max(1, n_st * 21 // 120), max(1, args.steps), mode=args.push_mode)
        res.pop("_result", None)
        print(json.dumps({"e2e_stored_ao": res}))
        return
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus N>1 must be launched with torch.distributed.run (one rank per GPU)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # N > 1: NUMA-local pinned buffers (not at N = 1, where the cpu_baseline leg wants every host core)
    host_affinity = bind_to_gpu_numa_node(torch, local) if (args.numa_bind == 1 or (args.numa_bind < 0 and world > 1)) else None
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    n, occ = args.nbf, args.nbf // 10
    win = mp2_window_e(n, occ)
    Cm = random_orthonormal(n, n)
    eps = synthetic_eps(occ, n)

    fp64_peak = cublas_fp64_peak(torch, dev)  # the FP64 roofline denominator, measured live (cuBLAS DGEMM)

    T = ol.Transformer(local)
    if world > 1:
        uid = [capi.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        T.comm_init(rank, world, uid[0])
    T.set_species(0, Cm)
    if args.q1_variant:
        T.set_option(T.OPT_Q1_VARIANT, args.q1_variant)
    if args.gemm_variant:
        T.set_option(T.OPT_GEMM_VARIANT, args.gemm_variant)
    if args.frag_perm >= 0:
        T.set_option(T.OPT_FRAG_PERM, args.frag_perm)
    if args.q3_red >= 0:
        T.set_option(T.OPT_Q3_RED, args.q3_red)
    if args.overlap >= 0:
        T.set_option(T.OPT_OVERLAP_EXCHANGE, args.overlap)
    if args.gemm_tall >= 0:
        T.set_option(T.OPT_GEMM_TALL, args.gemm_tall)
    if args.q3_two_cta >= 0:
        T.set_option(T.OPT_Q3_TWO_CTA, args.q3_two_cta)
    if args.exchange_dma >= 0:
        T.set_option(T.OPT_EXCHANGE_DMA, args.exchange_dma)
    T.set_generator(0, 0, SEED, args.gen)
    npass, qb = T.num_passes(0, 0, win, ol.CONV_E, args.occ_batch)

    def barrier():
        torch.cuda.synchronize(dev)
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # Watchdog: a pass that does not come back (ranks out of step inside a collective, a wedged kernel) must end the
    # process instead of holding the box until the caller's limit; torchrun then takes the other ranks down.
    import threading
    last_beat = [time.monotonic()]
    limit_s = float(os.environ.get("LOWDIN_BENCH_STEP_LIMIT_S", "900"))

    def watchdog():
        while True:
            time.sleep(5.0)
            if time.monotonic() - last_beat[0] > limit_s:
                sys.stderr.write(f"bench.py watchdog: rank {rank}: no step finished for {limit_s:.0f} s, aborting\n")
                sys.stderr.flush()
                os._exit(3)

    threading.Thread(target=watchdog, daemon=True).start()

    def one_pass(i, with_eps=True):
        out = T.transform_stream(0, 0, win, ol.CONV_E, occ_batch=qb, first_pass=i % npass, n_passes=1,
                                 epsA=eps if with_eps else None)
        last_beat[0] = time.monotonic()
        return out

    golden = {}
    try:
        golden = json.load(open(os.path.join(ROOT, "tests", "golden", "bench_sums.json")))
    except (OSError, ValueError):
        pass
    parity = {}
    if world > 1 and "small" in golden:
        # N>1 ranks: before anything is timed, a small collective transform whose sums the CPU oracle computed offline
        # (tests/golden/bench_sums.json, oracle/make_bench_golden.py): the exchange of THIS communicator, checked.
        g = golden["small"]
        Cs = random_orthonormal(g["n"], g["n"])
        T.set_species(1, Cs)
        T.set_generator(1, 1, g["seed"], 1)
        worst = 0.0
        for cols, qb_small in ((0, 0), (60, 4), (1, 2)):
            T.set_option(T.OPT_CHUNK_COLS, cols)
            sm = torch.tensor(T.transform_stream(1, 1, mp2_window_e(g["n"], g["occ"]), ol.CONV_E, occ_batch=qb_small,
                                                 epsA=synthetic_eps(g["occ"], g["n"])), dtype=torch.float64, device=dev)
            dist.all_reduce(sm)
            sm = sm.tolist()
            worst = max(worst, abs(sm[0] - g["sums"][0]), *[abs(a - b) for a, b in zip(sm[1:], g["sums"][1:])])
        T.set_option(T.OPT_CHUNK_COLS, 0)
        parity["small_collective_transform"] = {"n": g["n"], "ranks": world, "max_abs_diff_vs_oracle_golden": worst, "ok": worst <= 1e-9}
        if worst > 1e-9:
            raise SystemExit(f"bench.py: the {world}-rank transform of the small golden case differs from the oracle by {worst}")

    pass_sums = {}
    for i in range(args.warmup):
        pass_sums[i % npass] = one_pass(i)

    # ---------------- timed region: `value` (inputs resident in HBM) ----------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    T.set_profiling(True)
    barrier()
    t0 = time.perf_counter()
    dev_s, flops, launches = 0.0, 0.0, 0
    for i in range(args.steps):
        pass_sums[(args.warmup + i) % npass] = one_pass(args.warmup + i)
        tm = T.timers()
        dev_s += tm["first_half"] + tm["exchange"] + tm["second_half"] + tm["consume"]
        flops += tm["flops"]
        launches += tm["launches"]
    barrier()
    wall_s = time.perf_counter() - t0
    stats = T.kernel_stats()
    exchange_dma = T.exchange_is_dma()
    T.set_profiling(False)
    clocks = sampler.stop() if rank == 0 else None

    # ---------------- e2e: through the C ABI with host buffers every step ----------------
    # Every step re-uploads what a caller of this size owns on the host (coefficients from pinned memory, orbital
    # energies) and reads the reduced result back; the 5 TB AO tensor of N=1500 cannot exist on a host, so its values
    # stay a pure function of the canonical index evaluated where they are consumed.
    # the same passes as the timed region above, at most one whole transform (npass passes)
    e2e_steps = 0 if args.no_e2e else max(1, min(args.steps, npass))
    if e2e_steps:
        # the sink's pinned two-slot ring, allocated once, outside the timed region (like `pinned` below); a block holds at least the
        # pairs of one occupied orbital: (N-O)^2 O doubles
        T.set_option(T.OPT_SINK_BLOCK_BYTES, max(256 << 20, (n - occ) ** 2 * occ * 8))
    pinned = torch.from_numpy(np.ascontiguousarray(Cm.T)).pin_memory()  # column-major C(mu,p) == row-major C^T
    Cpin = pinned.numpy().T
    barrier()
    t1 = time.perf_counter()
    e2e_flops = 0.0
    sunk = [0, 0.0]                      # bytes of MO integrals that reached the host, a checksum over a sample of them
    sink_note = None

    def host_sink(sa, sb, vals, blk):
        sunk[0] += vals.nbytes
        sunk[1] += float(vals.reshape(-1)[::65521].sum())

    for i in range(e2e_steps):
        T.set_species(0, Cpin)           # H2D: coefficients
        T.set_generator(0, 0, SEED, args.gen)
        # H2D: orbital energies; D2H: EVERY MO integral of the pass, block by block into pinned host memory, + the reduced sums
        try:
            T.transform_stream_sink(0, 0, win, ol.CONV_E, host_sink, occ_batch=qb, first_pass=(args.warmup + i) % npass, n_passes=1, epsA=eps)
        except ol.LowdinITError as e:   # no room for the sink's second block buffer beside this occupied batch: reduced sums only
            sink_note = f"sink unavailable at this occupied batch ({e}); reduced sums only"
            one_pass(args.warmup + i)
        last_beat[0] = time.monotonic()
        e2e_flops += T.timers()["flops"]
    barrier()
    e2e_s = max(time.perf_counter() - t1, 1e-9)

    if dist is not None:
        t = torch.tensor([dev_s, wall_s, e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_s, wall_s, e2e_s = t.tolist()
        f = torch.tensor([flops, e2e_flops, float(launches), float(sunk[0])], dtype=torch.float64, device=dev)
        dist.all_reduce(f, op=dist.ReduceOp.SUM)
        flops, e2e_flops, launches, sunk[0] = f.tolist()

    # results of the passes that ran (every pass of the transform when warmup + steps >= npass), summed over ranks
    covered = sorted(pass_sums)
    tot = np.sum([pass_sums[k] for k in covered], axis=0)
    if dist is not None:
        tt = torch.tensor(tot, dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.SUM)
        tot = np.array(tt.tolist())
    parity["passes_covered"] = f"{len(covered)} of {npass}"
    parity["sums"] = {"count": tot[0], "sum": tot[1], "sum_sq": tot[2], "mp2_pair_energy": tot[3]}
    key = f"n{n}_gen{args.gen}"
    if len(covered) == npass and key in golden:
        ref = np.array(golden[key]["sums"])
        scale = np.array([1.0, np.sqrt(ref[0] * ref[2]), ref[2], max(abs(ref[3]), 1e-300)])
        rel = np.abs(tot - ref) / scale
        parity["whole_transform_vs_reference_sums"] = {"reference": golden[key]["sums"], "source": golden[key].get("source"),
                                                       "count_equal": bool(tot[0] == ref[0]), "max_rel_diff": float(rel[1:].max()),
                                                       "ok": bool(tot[0] == ref[0] and rel[1:].max() <= 1e-9)}

    if rank == 0:
        value = flops / dev_s / 1e9
        gemm_cats = ("q1", "q2", "q3", "q4")
        dom = max(stats, key=lambda c: stats[c]["ms"])
        st = stats[dom]
        kname = {"q1": "q1_gen_ws5_kernel (slab generation + first quarter)", "q2": "dgemm_tma_kernel<EpiScatterH> (second quarter)",
                 "q3": "dgemm_tma_kernel<EpiAccT> (third quarter, chunked)", "q4": "dgemm_tma_kernel<EpiOut> (fourth quarter)",
                 "expand1": "expand_block_kernel (first half)", "expand2": "expand_block_kernel (second half)",
                 "consume": "reduce_block_kernel", "exchange": "NCCL all-to-all"}.get(dom, dom)
        traffic = None
        try:  # dram bytes per launch of the dominant kernel from the committed ncu --set full capture of this workload
            tr = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
            traffic = tr.get(f"n{n}", {}).get(dom)
        except (OSError, ValueError):
            pass
        if dom in gemm_cats:
            achieved = st["work"] / (st["ms"] * 1e-3) / 1e12
            roof = {"bound": "tensor", "kernel": kname, "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s",
                    "frac": achieved / fp64_peak, "traffic": traffic, "launches": st["launches"],
                    "avg_launch_ms": st["ms"] / max(st["launches"], 1),
                    "peak_source": "cuBLAS DGEMM 4096^3 best of 6, measured live by bench.py (MEASURED_PEAKS.json has no FP64 entry)"}
        else:
            peaks = {}
            try:
                peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
            except (OSError, ValueError):
                pass
            hbm = peaks.get("hbm_gbs", 6650.0)
            achieved = st["work"] / (st["ms"] * 1e-3) / 1e9
            roof = {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": hbm, "unit": "GB/s", "frac": achieved / hbm,
                    "traffic": traffic, "launches": st["launches"], "avg_launch_ms": st["ms"] / max(st["launches"], 1),
                    "peak_source": "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6.65 TB/s"}
        kernels = {}
        for c, s_ in stats.items():
            if s_["launches"] == 0:
                continue
            rate = s_["work"] / (s_["ms"] * 1e-3)
            kernels[c] = {"ms": round(s_["ms"], 3), "launches": s_["launches"],
                          ("TFLOP/s" if c in gemm_cats else "GB/s"): rate / (1e12 if c in gemm_cats else 1e9)}
        line = {"metric": "4-index transform FP64 GFLOP/s", "value": value, "unit": "GFLOP/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": dev_s / args.steps * 1e3, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": f"N_bf={n} MP2 window O={occ} (transformer-E roles), kind-H synthetic AO generated on device, "
                                       f"random orthonormal C; step = one occupied-batch pass of {qb} occupied orbitals "
                                       f"({npass} passes = the whole transform)",
                           "nbf": n, "occ": occ, "gen": args.gen, "occ_batch": qb, "passes_per_transform": npass,
                           "host_affinity": host_affinity,
                           "transform_wall_s_at_this_rate": (2.0 * n * occ * (n + (n - occ)) * (n * (n + 1) // 2 + (n - occ) * occ)) / max(flops / dev_s, 1e-9),
                           "exchange": ("none (one GPU)" if world == 1 else "peer-to-peer DMA over cudaIpc-mapped chunk buffers" if exchange_dma
                                        else "grouped ncclSend/ncclRecv"),
                           "l2": "working set larger than L2 (each pass re-streams the third-quarter accumulators and chunk buffers, tens of GB)",
                           "flops_per_step": flops / args.steps, "fp64_pct_of_cublas_dgemm_per_gpu": 100.0 * value / 1e3 / fp64_peak / world,
                           "fp64_pct_of_nominal_37tf_per_gpu": 100.0 * value / 37000.0 / world, "cublas_dgemm_tflops_measured": fp64_peak, "wall_ms_per_step": wall_s / args.steps * 1e3},
                "roofline": roof, "kernels": kernels,
                "e2e": {"value": (e2e_flops / e2e_s / 1e9) if e2e_steps else None, "unit": "GFLOP/s", "steps": e2e_steps,
                        "h2d_bytes_per_step": int(n * n * 8 + n * 8) * world, "d2h_bytes_per_step": int(sunk[0] / max(e2e_steps, 1)) + 32 * world,
                        "note": "C-ABI calls with host buffers (lowdin_it_set_species, lowdin_it_transform_stream_sink): coefficients (pinned) + orbital "
                                "energies up; EVERY MO integral of the pass comes down as dense blocks into pinned host memory while the transform "
                                "runs, + the reduced sums.  AO values are generated on the device from the canonical index (a 5 TB host tensor "
                                "cannot exist); the stored-AO flow is measured by e2e_stored_ao / stored_ao_resident"
                                + (f" [{sink_note}]" if sink_note else "")},
                "gpu_launches": int(launches), "clocks": clocks, "parity": parity}
        last_beat[0] = time.monotonic() + 3600.0   # the CPU legs below are bounded by their own sampling, not by the watchdog
        if world == 1 and not args.no_cpu_baseline:
            nthreads = os.cpu_count() or 1
            v, dt, nsl = cpu_sample_timed(n, occ, nthreads, 10.0)
            line["cpu_baseline"] = {"value": v, "unit": "GFLOP/s", "cores": nthreads, "kind": "port",
                                    "sample": f"first half of transformer E (oracle port of E.f90:1043-1132) on {nsl} of {n*(n+1)//2} "
                                              f"AO-pair slabs, full occupied window, {dt:.1f} s"}
            try:   # the same CPU port as the checker of the device's first half at THIS size (first and last slabs of the tensor)
                line["cpu_baseline"]["first_half_parity"] = first_half_parity(T, ol, n, occ, args.gen, nthreads)
            except Exception as e:
                line["cpu_baseline"]["first_half_parity"] = {"error": f"{type(e).__name__}: {e}"}
    T.close()
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            try:   # the reference's own C++ transformer D beside its GPU drop-in (same call, same buffers)
                line["cpu_baseline"]["reference_transformer_d"] = transformer_d_leg(ol)
            except Exception as e:
                line["cpu_baseline"]["reference_transformer_d"] = {"error": f"{type(e).__name__}: {e}"}
        if world == 1 and args.resident_nbf > 0 and not args.no_e2e:
            try:
                line["stored_ao_resident"] = stored_resident_leg(ol, local, args.resident_nbf, args.resident_all, fp64_peak=fp64_peak)
            except Exception as e:
                line["stored_ao_resident"] = {"error": f"{type(e).__name__}: {e}"}
        if world == 1 and args.stored_nbf > 0 and not args.no_e2e:
            # second end-to-end figure: STORED AO integrals through the whole upload -> transform -> download ABI
            occ_st = max(1, args.stored_nbf * 21 // 120)
            gpu_result = None
            try:
                line["e2e_stored_ao"] = stored_ao_e2e(torch, ol, capi, local, args.stored_nbf, occ_st, 3, mode="blocks")
                gpu_result = line["e2e_stored_ao"].pop("_result", None)
                five = stored_ao_e2e(torch, ol, capi, local, args.stored_nbf, occ_st, 2, mode="stacks")
                line["e2e_stored_ao"]["five_array_push"] = {k: five[k] for k in ("value", "ms_per_step", "push", "upload_gb_per_s")}
            except Exception as e:  # never lose the main line to the secondary leg
                line["e2e_stored_ao"] = {"value": None, "error": f"{type(e).__name__}: {e}"}
            if not args.no_cpu_baseline and args.stored_nbf <= 160:
                try:   # the same workload on the host: restated transformers E and C, and GPU-vs-E parity of the stored path
                    line["cpu_baseline"]["port_whole_transform"] = whole_transform_cpu_leg(args.stored_nbf, occ_st, gpu_result)
                except Exception as e:
                    line["cpu_baseline"]["port_whole_transform"] = {"error": f"{type(e).__name__}: {e}"}
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
