"""Round-1 GPU probe: FP64 ceilings (cuBLAS DGEMM via torch, the kernel to beat) and our kernels alone.
Writes gpurun_out/probe.json.  Not part of the product; measurement helper."""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import openlowdin_b200 as ol  # noqa: E402

out = {}
dev = torch.device("cuda:0")
out["gpu"] = torch.cuda.get_device_name(0)


def cublas_dgemm(n, iters=10):
    a = torch.rand(n, n, dtype=torch.float64, device=dev)
    b = torch.rand(n, n, dtype=torch.float64, device=dev)
    torch.matmul(a, b)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    # sustained: back to back for ~3 s
    t0 = time.time(); cnt = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    while time.time() - t0 < 3.0:
        for _ in range(5):
            torch.matmul(a, b); cnt += 1
        torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize()
    sus = e0.elapsed_time(e1) / cnt
    return 2.0 * n ** 3 / best / 1e9, 2.0 * n ** 3 / sus / 1e9


for n in (4096, 8192):
    b, s = cublas_dgemm(n)
    out[f"cublas_dgemm_{n}_tflops_burst"] = b
    out[f"cublas_dgemm_{n}_tflops_sustained"] = s
    print("cublas dgemm", n, b, s, flush=True)

T = ol.Transformer(0)
shapes = [(8192, 8192, 8192), (4096, 4096, 4096), (1350, 4096, 1500), (33000, 8, 1500), (33000, 16, 1500), (33000, 150, 1500),
          (1350, 3300, 1500), (16384, 32, 1500), (16384, 64, 1500), (16384, 80, 1500), (16384, 128, 1500)]
for (m, n, k) in shapes:
    ms, _ = T.kernel_bench(1, m, n, k, iters=5)
    tf = 2.0 * m * n * k / (ms * 1e-3) / 1e12
    out[f"dmma_gemm_{m}x{n}x{k}_tflops"] = tf
    print("dmma gemm", m, n, k, "ms", ms, "TF/s", tf, flush=True)

for (nb, nslab, kind) in [(1500, 16, 2), (1500, 16, 0), (1500, 16, 1), (500, 64, 2), (500, 64, 0)]:
    try:
        ms, _ = T.kernel_bench(0, nb, nslab, kind, iters=5)
    except Exception as e:  # packed tensor for N=1500 does not fit
        print("expand", nb, nslab, kind, "skipped:", e)
        continue
    M = nb * (nb + 1) // 2
    alg = nslab * (8.0 * M + 8.0 * nb * nb)
    out[f"expand_n{nb}_kind{kind}_GBs"] = alg / (ms * 1e-3) / 1e9
    print("expand", nb, nslab, kind, "ms", ms, "GB/s", alg / (ms * 1e-3) / 1e9, flush=True)

json.dump(out, open("gpurun_out/probe.json", "w"), indent=1)
