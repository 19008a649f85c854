"""Device-to-host bandwidth into pinned memory, one rank alone and all ranks at once (what bounds the e2e leg's host sink when
several GPUs of one box deliver their MO integrals together).  Run under torchrun; rank 0 prints one JSON line."""
import json
import os
import time

import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
nbytes = 2 << 30
src = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
dst = torch.empty(nbytes, dtype=torch.uint8).pin_memory()


def run(active):
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 0
    if active:
        while time.perf_counter() - t0 < 1.5:
            dst.copy_(src, non_blocking=True)
            torch.cuda.synchronize()
            n += 1
    dt = time.perf_counter() - t0
    t = torch.tensor([n * nbytes / dt / 1e9 if active else 0.0], dtype=torch.float64, device="cuda")
    dist.all_reduce(t)
    return t.item()


dst.copy_(src); torch.cuda.synchronize()
alone = run(rank == 0)
together = run(True)
if rank == 0:
    print(json.dumps({"ranks": world, "d2h_gb_per_s_one_rank_alone": alone, "d2h_gb_per_s_all_ranks_sum": together}))
dist.destroy_process_group()
