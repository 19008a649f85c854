"""Multi-GPU parity check (run under torchrun, one rank per GPU): first half sharded over AO-pair slabs,
NCCL all-to-all per chunk, second half on the slot owners; the summed results must equal the CPU oracle's.
Used by tests/test_multi_gpu.py and by hand:  torchrun --nproc-per-node 2 scripts/mgpu_check.py"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import openlowdin_b200 as ol  # noqa: E402
from openlowdin_b200 import capi  # noqa: E402
from oracle import oracle as O  # noqa: E402


def species_pairs_across_ranks(rank, world, local, dev):
    """SURVEY 8e, "species pairs are scheduled across devices": the calls of the program's species loop are divided among
    the ranks (lowdin_host_run_program, no communicator); every moint.dat is then checked against the oracle by rank 0."""
    import tempfile
    shapes = [("E-", 7, 3), ("H-A_1", 5, 1), ("H-B_1", 4, 1)]
    S = 40
    d = [tempfile.mkdtemp(prefix="lowdin_it_mgpu_") if rank == 0 else None]
    dist.broadcast_object_list(d, src=0)
    d = d[0]
    Cs = [O.random_orthonormal(n, 31 + i) for i, (_, n, _) in enumerate(shapes)]
    intra = {i: O.hash_packed_intra(100 + i, n) for i, (_, n, _) in enumerate(shapes)}
    inter = {(i, j): O.hash_rect_inter(200 + 10 * i + j, shapes[i][1], shapes[j][1]) for i in range(3) for j in range(i + 1, 3)}
    if rank == 0:
        for i, (nm, n, _) in enumerate(shapes):
            capi.host_write_ints_file(os.path.join(d, f"0{nm}.ints"), S, *O.canonical_list_intra(intra[i], n))
            for j in range(i + 1, 3):
                capi.host_write_ints_file(os.path.join(d, f"0{nm}.{shapes[j][0]}.ints"), S, *O.canonical_list_inter(inter[i, j], n, shapes[j][1]))
    dist.barrier()
    sp = [capi.host_species(nm, i + 1, n, occ, coeff=Cs[i]) for i, (nm, n, occ) in enumerate(shapes)]
    ctl = capi.host_control("E", "MP2", stack=S, nfiles=1, scratch_dir=d)
    T2 = ol.Transformer(local)
    written, calls = capi.host_run_program(T2, ctl, sp, rank, world)
    T2.close()
    tot = torch.tensor([float(written), float(calls)], dtype=torch.float64, device=dev)
    dist.all_reduce(tot)
    dist.barrier()
    good = True
    if rank == 0:
        import struct

        def read_pairs(path):  # E record layout: int64 ij[S], kl[S], real64 v[S] per record, terminator ij = -1
            raw = open(path, "rb").read()
            ij, kl, v, o = [], [], [], 0
            while o < len(raw):
                (ln,) = struct.unpack_from("<I", raw, o)
                a = np.frombuffer(raw, np.int64, S, o + 4); b = np.frombuffer(raw, np.int64, S, o + 4 + 8 * S)
                c = np.frombuffer(raw, np.float64, S, o + 4 + 16 * S)
                m = int(np.argmax(a == -1)) if (a == -1).any() else S
                ij.append(a[:m]); kl.append(b[:m]); v.append(c[:m])
                o += ln + 8
            return np.concatenate(ij), np.concatenate(kl), np.concatenate(v)
        count = 0
        for i, (nm, n, occ) in enumerate(shapes):
            M = O.npairs(n)
            ref = O.transform_e_intra(Cs[i], intra[i], O.windows_e_intra("MP2", n, occ))
            got = read_pairs(os.path.join(d, f"{nm}moint.dat"))
            good = good and np.abs(O.pairs_to_dense(*got, M, M) - O.pairs_to_dense(*ref, M, M)).max() <= 1e-10
            count += len(got[2])
            for j in range(i + 1, 3):
                nm2, n2, occ2 = shapes[j]
                ref = O.transform_e_inter(Cs[i], Cs[j], inter[i, j], O.windows_e_inter("MP2", n, n2, occ, occ2))
                got = read_pairs(os.path.join(d, f"{nm}.{nm2}moint.dat"))
                good = good and np.abs(O.pairs_to_dense(*got, M, O.npairs(n2)) - O.pairs_to_dense(*ref, M, O.npairs(n2))).max() <= 1e-10
                count += len(got[2])
        good = bool(good) and count == int(tot[0].item()) and int(tot[1].item()) == 6
        print(f"species pairs across {world} ranks: 6 calls, {count} integrals written -> {'ok' if good else 'MISMATCH'}", flush=True)
    return good


def collective_download(T, rank, world, local):
    """lowdin_it_transform on the communicator with a STORED tensor (every rank is pushed the whole list and keeps the rows it
    owns), per-rank downloads, merged by window-pair segments on rank 0 into ONE moint.dat: the same entries in the same order as
    the file a single GPU writes (TransformIntegralsE.f90:1242-1268 record order), values equal to rounding (the ranks' GEMM batches
    have other shapes than the single GPU's, so the last bit may differ; byte identity is reported, not required)."""
    import tempfile
    n, occ, S = 17, 5, 64
    packed = O.hash_packed_intra(808, n)
    Cm = O.random_orthonormal(n, n)
    lst = O.canonical_list_intra(packed, n)
    win = O.windows_e_intra("MP2", n, occ)
    T.set_option(T.OPT_CHUNK_COLS, 45)
    T.set_species(0, Cm)
    T.upload_ao(0, 0, *lst, stack=S)
    ij, kl, v = T.transform(0, 0, win, ol.CONV_E)
    idx, kept = T.result_segments()
    T.set_option(T.OPT_CHUNK_COLS, 0)
    parts = [None] * world
    dist.all_gather_object(parts, (idx, kept, ij, kl, v))
    good = True
    if rank == 0:
        segs = []
        for (pi, pk, a, b, c) in parts:
            off = np.concatenate([[0], np.cumsum(pk)])
            segs += [(int(pi[t]), a[off[t]:off[t + 1]], b[off[t]:off[t + 1]], c[off[t]:off[t + 1]]) for t in range(len(pi))]
        segs.sort(key=lambda e: e[0])
        mij, mkl, mv = (np.concatenate([e[k] for e in segs]) for k in (1, 2, 3))
        d = tempfile.mkdtemp(prefix="lowdin_it_mgpu_dl_")
        capi.host_write_moint_pairs(os.path.join(d, "merged.dat"), S, mij, mkl, mv)
        T1 = ol.Transformer(local)
        T1.set_species(0, Cm)
        T1.upload_ao(0, 0, *lst, stack=S)
        one = T1.transform(0, 0, win, ol.CONV_E)
        T1.close()
        capi.host_write_moint_pairs(os.path.join(d, "one.dat"), S, *one)
        same_keys = np.array_equal(mij, one[0]) and np.array_equal(mkl, one[1])
        diff = float(np.abs(mv - one[2]).max()) if same_keys else float("inf")
        identical = open(os.path.join(d, "merged.dat"), "rb").read() == open(os.path.join(d, "one.dat"), "rb").read()
        ref = O.transform_e_intra(Cm, packed, win)
        M = O.npairs(n)
        vs_oracle = float(np.abs(O.pairs_to_dense(mij, mkl, mv, M, M) - O.pairs_to_dense(*ref, M, M)).max())
        good = same_keys and diff <= 1e-12 and vs_oracle <= 1e-10
        print(f"collective download on {world} ranks: {len(mv)} integrals, same order as 1 GPU: {same_keys}, max|delta| vs 1 GPU {diff:.1e}, "
              f"vs oracle {vs_oracle:.1e}, moint.dat byte-identical: {identical} -> {'ok' if good else 'MISMATCH'}", flush=True)
    return good


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    T = ol.Transformer(local)
    uid = [capi.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    T.comm_init(rank, world, uid[0])
    ok = True
    # intra, MP2 window with exchange energy; inter MP2
    cases = [("intra", 24, 6, 31337), ("intra", 19, 5, 7), ("inter", (21, 16), (5, 2), 99)]
    for kind, n, occ, seed in cases:
        if kind == "intra":
            packed = O.hash_packed_intra(seed, n)
            Cm = O.random_orthonormal(n, n)
            eps = O.synthetic_eps(occ, n)
            T.set_species(0, Cm)
            T.set_generator(0, 0, seed)
            win = O.windows_e_intra("MP2", n, occ)
            rij, rkl, rv = O.transform_e_intra(Cm, packed, win)
            want = np.array([len(rv), rv.sum(), (rv * rv).sum(), O.mp2_intra_from_pairs(rij, rkl, rv, n, occ, eps, lam=2.0)])
            run = lambda qb: T.transform_stream(0, 0, win, ol.CONV_E, occ_batch=qb, epsA=eps, lam=2.0)  # noqa: E731
        else:
            (na, nb), (oa, ob) = n, occ
            rect = O.hash_rect_inter(seed, na, nb)
            Ca, Cb = O.random_orthonormal(na, 3), O.random_orthonormal(nb, 4)
            ea, eb = O.synthetic_eps(oa, na), O.synthetic_eps(ob, nb)
            T.set_species(0, Ca); T.set_species(1, Cb)
            T.set_generator(0, 1, seed)
            win = O.windows_e_inter("MP2", na, nb, oa, ob)
            rij, rkl, rv = O.transform_e_inter(Ca, Cb, rect, win)
            want = np.array([len(rv), rv.sum(), (rv * rv).sum(),
                             O.mp2_inter_from_pairs(rij, rkl, rv, na, nb, oa, ob, ea, eb, charge_a=1.0, charge_b=1.0, lam_a=1.0, lam_b=1.0)])
            run = lambda qb: T.transform_stream(0, 1, win, ol.CONV_E, occ_batch=qb, epsA=ea, epsB=eb)  # noqa: E731
        for cols, qb in ((0, 0), (60, 4), (1, 2), (200, 3)):
            T.set_option(T.OPT_CHUNK_COLS, cols)
            s = torch.tensor(run(qb), dtype=torch.float64, device=dev)
            dist.all_reduce(s)
            got = s.cpu().numpy()
            good = got[0] == want[0] and np.all(np.abs(got[1:] - want[1:]) <= 1e-9)
            ok = ok and bool(good)
            if rank == 0:
                print(f"{kind} n={n} chunk_cols={cols} occ_batch={qb}: got {got} want {want} -> {'ok' if good else 'MISMATCH'}", flush=True)
    T.set_option(T.OPT_CHUNK_COLS, 0)
    flag = torch.tensor([1.0 if collective_download(T, rank, world, local) else 0.0], dtype=torch.float64, device=dev)
    dist.broadcast(flag, 0)
    ok = ok and bool(flag.item() == 1.0)
    T.close()
    ok = species_pairs_across_ranks(rank, world, local, dev) and ok
    dist.destroy_process_group()
    if rank == 0:
        print("MGPU_CHECK_OK" if ok else "MGPU_CHECK_FAILED", flush=True)
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
