"""N>1 path on CPU (gloo, world_size 2): the library's division of work (lowdin_it_shard_plan: block-cyclic slabs, slots by
blocks of first-contracted indices) and the layout the all-to-all leaves behind (lowdin_it_exchanged_offset) drive a numpy
emulation of the two-half transform; the ranks'
combined result must equal the oracle's.  The arithmetic here is numpy (test scaffolding); what is under test is the
host-side sharding logic that the CUDA path uses unchanged."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n, occ, rows_per_chunk, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from openlowdin_b200 import capi
        from oracle import oracle as O
        M = O.npairs(n)
        packed = O.hash_packed_intra(5, n)
        Cm = np.asarray(O.random_orthonormal(n, n))
        sq = O.packed_to_square(packed, M)
        xy = O.pair_table(n)
        V = n - occ
        # MP2 window, transformer-E roles: slots (a virt, i occ), numbered f-major (f = i)
        fbeg = np.arange(occ + 1, dtype=np.int32) * V
        LOGB = 1                                                 # blocks of two slabs: several blocks per chunk at these sizes
        own, _, _, _ = capi.shard_plan(fbeg, 0, 0, world, rank, LOGB)
        mine = own[rank + 1] - own[rank]
        Hmine = np.zeros((mine, M))
        Cv, Co = Cm[:, occ:], Cm[:, :occ]
        for p0 in range(0, n, rows_per_chunk):
            p1 = min(n, p0 + rows_per_chunk)
            base, width = xy[p0, p0], sum(n - p for p in range(p0, p1))
            own, wblk, loc_lo, cnt = capi.shard_plan(fbeg, base, width, world, rank, LOGB)
            wblk = max(wblk, 1)
            # first half of my slabs of the chunk (consecutive LOCAL slabs), for ALL slots
            Hc = np.zeros((occ * V, wblk))
            for c in range(cnt):
                z = capi.slab_global(loc_lo + c, world, rank, LOGB)
                assert base <= z < base + width and capi.slab_owner(z, world, LOGB) == rank and capi.slab_local(z, world, LOGB) == loc_lo + c
                X = sq[z][xy]                                    # dense slab (mu,nu)
                T2 = Cv.T @ X @ Co                               # [a][i]
                Hc[:, c] = T2.T.reshape(-1)                      # slot = i*V + a
            # exchange: rows own[g]:own[g+1] go to rank g (grouped send/recv like ncclSend/ncclRecv)
            recv = [torch.zeros(mine * wblk, dtype=torch.float64) for _ in range(world)]
            reqs = []
            for g in range(world):
                blk = torch.from_numpy(np.ascontiguousarray(Hc[own[g]:own[g + 1]]).reshape(-1))
                if g == rank:
                    recv[g].copy_(blk)
                else:
                    reqs.append(dist.isend(blk, g))
                    reqs.append(dist.irecv(recv[g], g))
            for r in reqs:
                r.wait()
            flat = torch.cat(recv).numpy()                       # [g][slot_local][wblk]
            for s in range(mine):
                for z in range(base, base + width):
                    Hmine[s, z] = flat[capi.exchanged_offset(s, z, base, wblk, mine, world, LOGB)]
        # second half on my slots
        out = {}
        for s in range(mine):
            Y = Hmine[s][xy]
            out[own[rank] + s] = Cv.T @ Y @ Co                   # [b][j]
        q.put((rank, int(own[rank]), int(own[rank + 1]), {k: v.copy() for k, v in out.items()}))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n,occ,rows", [(8, 3, 2), (7, 2, 4), (6, 4, 6)])
def test_two_rank_sharding_reproduces_the_oracle(O, n, occ, rows):
    world, port = 2, 29611 + n
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, occ, rows, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # slot ranges tile [0, occ*V) without overlap
    res.sort()
    V = n - occ
    assert res[0][1] == 0 and res[0][2] == res[1][1] and res[1][2] == occ * V
    packed = O.hash_packed_intra(5, n)
    Cm = O.random_orthonormal(n, n)
    ij, kl, v = O.transform_e_intra(Cm, packed, O.windows_e_intra("MP2", n, occ))
    M = O.npairs(n)
    ref = np.zeros((M, M))
    ref[ij - 1, kl - 1] = v
    xy = O.pair_table(n)
    for _, _, _, out in res:
        for slot, blk in out.items():
            i, a = slot // V, occ + slot % V
            for b in range(V):
                for j in range(occ):
                    want = ref[xy[i, a], xy[j, occ + b]]
                    got = blk[b, j]
                    assert abs(got - want) <= 1e-10 or abs(want) == 0.0 and abs(got) <= 1.1e-10


def test_shard_plan_properties():
    from openlowdin_b200 import capi
    fbeg = np.array([0, 3, 7, 7, 12, 20], np.int32)
    base, width = 37, 101
    for G in (1, 2, 3, 5, 8):
        for logB in (0, 2, 5):
            cover, offs, wmax = [], set(), 0
            for r in range(G):
                own, wblk, lo, cnt = capi.shard_plan(fbeg, base, width, G, r, logB)
                assert own[0] == 0 and own[G] == 20 and all(own[k] <= own[k + 1] for k in range(G))
                assert all(o in fbeg for o in own)               # whole f-blocks only: exchange partners stay together
                assert 0 <= cnt <= wblk
                wmax = max(wmax, cnt)
                mine = [capi.slab_global(lo + c, G, r, logB) for c in range(cnt)]
                assert all(capi.slab_owner(z, G, logB) == r and capi.slab_local(z, G, logB) == lo + c for c, z in enumerate(mine))
                cover += mine
            assert wblk == wmax
            assert sorted(cover) == list(range(base, base + width))      # the ranks' slabs tile the chunk exactly
            rows = 4
            for z in range(base, base + width):                           # the exchanged layout has one place per (row, slab)
                for row in range(rows):
                    o = capi.exchanged_offset(row, z, base, max(wblk, 1), rows, G, logB)
                    assert 0 <= o < G * rows * max(wblk, 1) and o not in offs
                    offs.add(o)
    assert capi.exchanged_offset(2, 7, 5, 5, 4, 1, 0) == 2 * 5 + 2


def _plan_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from openlowdin_b200 import capi
        shapes = [("E-", 19, 5), ("H-A_1", 50, 1), ("H-B_1", 50, 1)]
        sp = [capi.host_species(nm, i + 1, n, occ) for i, (nm, n, occ) in enumerate(shapes)]
        plan = capi.host_plan_program(capi.host_control("C", "MP2"), sp, world)
        mine = [(t["first"], t["second"]) for t in plan if t["rank"] == rank]
        work = torch.tensor([sum(t["flops"] for t in plan if t["rank"] == rank)], dtype=torch.float64)
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)         # no data-path collective is needed; this one is the test's own check
        dist.all_reduce(work)
        if rank == 0:
            q.put((gathered, float(work[0]), [(t["first"], t["second"]) for t in plan], sum(t["flops"] for t in plan)))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_species_pair_calls_are_divided_among_ranks_without_overlap():
    world, port = 2, 29655
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_plan_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    gathered, work, program, total = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    flat = [c for part in gathered for c in part]
    assert sorted(flat, key=str) == sorted(program, key=str) and len(set(flat)) == len(program) == 6
    assert all(len(part) > 0 for part in gathered)
    assert abs(work - total) <= 1e-6 * total
