"""Model of the library's memory plan (pick_occ_batch and the chunk sizing of run_passes in csrc/it_api.cu) for the MP2
window of the scaling configurations: occupied batch, passes, third-quarter accumulators, chunk width, buffers per rank.
No GPU needed; used to check that a sizing change still fits 180 GB before spending GPU time on it.
  python scripts/plan_model.py [free_GB]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from openlowdin_b200 import capi  # noqa: E402


def ceil_div(a, b):
    return -(-a // b)


def model(N, G, free=178e9, ws=1 << 30):
    occ, virt = N // 10, N - N // 10
    nf, nf2, ns2, ld = occ, occ, virt, N + (N % 2)
    spf, per_out, M = virt, virt * occ * 8.0, N * (N + 1) // 2
    out_need = max(min(4e9, spf * per_out * nf), min(2, nf) * spf * per_out)          # pick_occ_batch
    avail = 0.92 * free - 2 * ws - out_need - (1 << 30)
    t3_per_f = spf * nf2 * ld * 8.0 / G
    hc_per_f = spf * 8.0 * min(M, 8 * N) * (3.0 / G if G > 1 else 1.0)
    qmax = int(max(1, min(nf, avail / (t3_per_f + hc_per_f))))
    # the batch itself comes from the library's cost model (first + third quarter; a pure host function of the shared library)
    qb = capi.occ_batch_model(nf, qmax, G, spf, nf2, N, N, M, M, avail, False)
    nslots = qb * virt                                                               # run_passes, first pass
    nmine = max((qb * (r + 1) // G - qb * r // G) * virt for r in range(G))
    t3 = nmine * nf2 * ld * 8
    out2 = max(min(4e9, nmine * per_out), spf * per_out)
    avail2 = 0.92 * (free - t3) - out2 - 2 * ws - (1 << 29)
    per_col = (nslots * 8.0 / G + 2.0 * nmine * 8.0) if G > 1 else nslots * 8.0
    width = min(int(avail2 / per_col), M)
    wblk = ceil_div(width, G)
    H = nslots * (wblk if G > 1 else width) * 8
    H2 = nmine * wblk * G * 8 if G > 1 else 0
    return dict(N=N, G=G, occ_batch=qb, passes=ceil_div(nf, qb), T3_GB=round(t3 / 1e9, 1), chunk_cols=width,
                rows_in_first_chunk=width // N, chunks=ceil_div(M, max(width, 1)), H_GB=round(H / 1e9, 1), H2_GB=round(H2 / 1e9, 1),
                total_GB=round((t3 + H + H2 + out2 + 2 * ws) / 1e9, 1))


if __name__ == "__main__":
    free = float(sys.argv[1]) * 1e9 if len(sys.argv) > 1 else 178e9
    for N in (500, 1000, 1500, 2000):
        for G in (1, 2, 4, 8):
            print(model(N, G, free))
