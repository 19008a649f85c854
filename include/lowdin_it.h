/*
 * lowdin_it.h -- C ABI of liblowdin_itgpu.so, the B200 (sm_100a) four-index AO->MO
 * two-particle integral transformation for openLOWDIN.
 *
 * This is the drop-in boundary for the reference's integral-transformation stage
 * (src/integralsTransformation).  A Fortran host binds these entry points with
 * ISO_C_BINDING exactly as TransformIntegralsD.f90:55-96 binds the reference's own
 * C++ transformer (IntTransfD.h:57-82); see INTEGRATION.md for the shim.
 *
 * Conventions (identical to the reference's):
 *   - coefficient matrices are COLUMN-MAJOR, C(mu,p), leading dimension ldc
 *     (TransformIntegralsD.f90:192-196);
 *   - windows are 1-based inclusive {p_l,p_u,q_l,q_u,r_l,r_u,s_l,s_u}
 *     (TransformIntegralsC.f90:1456-1612, TransformIntegralsE.f90:1916-2062);
 *   - AO integrals arrive in the reference's .ints stack layout
 *     (int32 p[],q[],r[],s[]; double v[]; 1-based; terminator p=-1;
 *     Libint2Iface.cpp:3414-3426, TransformIntegralsC.f90:251-298);
 *   - pair ids are the 1-based row-wise upper-triangular numbering xy(p,q)
 *     (TransformIntegralsC.f90:214-221 == IndexMap_tensorR2ToVectorB, IndexMap.f90:249-265).
 *
 * Every function returns 0 on success, non-zero on error; lowdin_it_last_error()
 * gives the message (the Fortran shim raises Exception ERROR with it, mirroring
 * TransformIntegralsC.f90:2030-2044).  All pointers are HOST pointers unless named d_*.
 * The library has no CPU fallback: with no usable CUDA device every call fails.
 */
#ifndef LOWDIN_IT_H
#define LOWDIN_IT_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct lowdin_it_ctx *lowdin_it_handle;

/* Output convention of a transform. */
#define LOWDIN_IT_CONV_C 0 /* transformer C: quads (p,q,r,s), `symmetric` skip rule  (TransformIntegralsC.f90:345-442) */
#define LOWDIN_IT_CONV_E 1 /* transformer E: pair ids (ij,kl), j<=i, l<=k, half-drop (TransformIntegralsE.f90:1043-1260) */

/* Synthetic AO generators (benchmark inputs). */
#define LOWDIN_IT_GEN_HASH 1 /* kind H: value = 2u-1, u = (splitmix64(seed ^ key) >> 11) * 2^-53 */
#define LOWDIN_IT_GEN_FOLD 2 /* kind F: same, with mulfold64(x) = ((x*K1) ^ ((x*K1)>>32)) * K2 instead of splitmix64 */
#define LOWDIN_IT_GEN_RANKK 3 /* kind K: (mu nu|lam sig) = sum_{k<K} La^k[mu nu] Lb^k[lam sig], K <= 8 (lowdin_it_ao_set_rankk) */

/* ---- lifetime --------------------------------------------------------------------- */
/* Replaces the per-call malloc/free of IntTransfD.cpp:131-143: device buffers live in the handle. */
int lowdin_it_create(int device, lowdin_it_handle *out);
int lowdin_it_destroy(lowdin_it_handle h);
const char *lowdin_it_last_error(lowdin_it_handle h); /* h may be NULL: last error of create() or of the handle-less transformer-D entry points */

/* ---- inputs ----------------------------------------------------------------------- */
/* Coefficients of one species into slot (0..7).  Replaces the coeff(nao,nao) copy +
 * c_loc(coeff) hand-off of TransformIntegralsD.f90:189-196, :211-214. */
int lowdin_it_set_species(lowdin_it_handle h, int slot, int nao, const double *C_colmajor, int ldc, int ncols);

/* AO integrals of one species pair, pushed in the reference's stack layout.  Replaces the
 * loaders TransformIntegralsC.f90:231-298 (intra), :852-972 (inter) and
 * ReadIntegrals.f90:23-99.  slotB == slotA means intra-species.  `swapped` != 0 restates the
 * reversed-pair branch TransformIntegralsC.f90:906-972: the stacks are (B B|A A). */
int lowdin_it_ao_begin(lowdin_it_handle h, int slotA, int slotB, int swapped);
/* n entries as five arrays.  The terminator test (p = -1 ends the stream of THIS call, C.f90:279-280), the index
 * range check and the scatter all run on the device; the call returns when its host buffers have been copied
 * (they may be reused), not when the scatter is done.  An entry with an index outside the basis makes
 * lowdin_it_ao_end fail (the AO set stays unusable). */
int lowdin_it_ao_push_stacks(lowdin_it_handle h, const int32_t *p, const int32_t *q, const int32_t *r,
                             const int32_t *s, const double *v, int64_t n);
/* The same entries as RAW .ints bytes: nblocks blocks of `int32 p[S],q[S],r[S],s[S]; real64 v[S]` (24 S bytes each,
 * Libint2Iface.cpp:3414-3426; what `read(unit) pp,qq,rr,ss,shellIntegrals` consumes, C.f90:258-262), e.g. a whole file
 * or a memory-mapped range of it: one host-to-device copy per staging buffer, decoded on the device. */
int lowdin_it_ao_push_blocks(lowdin_it_handle h, const void *blocks, int64_t nblocks, int stack_size);
int lowdin_it_ao_end(lowdin_it_handle h);
/* Instead of an upload: slabs generated on the device on the fly (no N^4/8 array). */
int lowdin_it_ao_set_generator(lowdin_it_handle h, int slotA, int slotB, int kind, uint64_t seed);
/* Kind K: rank-K separable tensor, the one synthetic input whose MO integrals have a closed form
 * (p q|r s) = sum_k (Ca^T La^k Ca)[p,q] (Cb^T Lb^k Cb)[r,s], i.e. an O(K N^3) oracle at any N.  La: host [K][M_a] values of the
 * K symmetric matrices at the pairs (mu<=nu) in xy order; Lb likewise for species B (ignored for slotB == slotA). */
int lowdin_it_ao_set_rankk(lowdin_it_handle h, int slotA, int slotB, int K, const double *La, const double *Lb);
/* ---- row f4 of SURVEY.md section 8: the AO integrals evaluated on the device -----------------------------------------
 * The basis of a species as the reference hands it to libint2, one call per shell of LibintInterface::add_shell(alpha, coeff,
 * origin, l, nprim) (Libint2Iface.cpp:83-130): Cartesian shells (components in libint2's order: lx = l..0, ly = l-lx..0),
 * contraction coefficients of unit-normalised primitives (libint2::Shell::renorm), every function divided by the square root of
 * its self overlap (`norma`).  first_prim indexes `exponents` / `coefficients`.  l <= 3.  The species must have been set
 * (lowdin_it_set_species) with nao = the number of Cartesian functions of the shells. */
typedef struct { int l, nprim, first_prim; double origin[3]; } lowdin_it_shell;
int lowdin_it_set_basis(lowdin_it_handle h, int slot, int nshells, const lowdin_it_shell *shells, const double *exponents,
                        const double *coefficients);
int lowdin_it_basis_norma(lowdin_it_handle h, int slot, double *norma /* [nao] */);
/* (p q|r s) of the pair (slotA, slotB) evaluated on the device into the stored AO tensor (packed / rectangular / the rows this rank
 * owns on a communicator), as if the `.ints` streams of LibintInterface::compute_2body_disk (Libint2Iface.cpp:219-416; slotA ==
 * slotB) or ::compute_coupling_disk (:930-1110) had been uploaded: raw values <= 1e-10 are dropped (:369, :1053), the others scaled
 * by norma.  The reference's density-weighted Schwarz screening (:300-330) is not applied: every integral is evaluated. */
int lowdin_it_ao_compute(lowdin_it_handle h, int slotA, int slotB);
/* The stored AO tensor back on the host (one rank): packed M(M+1)/2 doubles (intra, row lo holds hi = lo..M-1) or [M_b][M_a]. */
int lowdin_it_ao_download(lowdin_it_handle h, int slotA, int slotB, double *out, int64_t capacity);

/* Turn a generated AO set into a STORED one on the device (packed M(M+1)/2 / rectangular M_b x M_a), as if its whole
 * list had been uploaded: the stored-AO kernels at sizes whose list no host could hold (bench leg, tests). */
int lowdin_it_ao_materialize(lowdin_it_handle h, int slotA, int slotB);

/* ---- the transform ---------------------------------------------------------------- */
/* One species (slotB==slotA) or species pair.  Replaces
 * TransformIntegralsC_atomicToMolecularOfOneSpecie / OfTwoSpecies (TransformIntegralsC.f90:141, :728)
 * and the E versions (TransformIntegralsE.f90:153, :1285) up to, not including, the file write.
 * conv selects whose semantics are reproduced; `symmetric` is C's flag (ignored for E);
 * drop_tol is the reference's 1e-10 (C.f90:418, E.f90:1113, :1242). */
int lowdin_it_transform(lowdin_it_handle h, int slotA, int slotB, const int win[8], int conv, int symmetric,
                        double drop_tol);
/* On a communicator lowdin_it_transform is COLLECTIVE and every rank ends up with the integrals of ITS window pairs (first pairs
 * (i,j) / (p,q) are divided among the ranks): lowdin_it_result_count and the downloads are then per rank, each list in the
 * reference's loop order.  lowdin_it_result_segments tells how the lists interleave: this rank holds `npairs` window pairs, the
 * one with convention-order index pair_index[t] (0-based, counted over all window pairs: ijmap order for E, the (p,q) loop for C)
 * contributed the next kept[t] entries of its list.  Taking the pairs in increasing index from whichever rank holds them gives
 * the single-GPU list, entry for entry (what the lowdin_it_group_* calls below do for a one-process host). */
int lowdin_it_result_count(lowdin_it_handle h, int64_t *count);
int lowdin_it_result_segments(lowdin_it_handle h, int64_t *npairs, int64_t *pair_index, int64_t *kept);
/* E record content (TransformIntegralsE.f90:1244-1250): pair ids + value, reference loop order. */
int lowdin_it_download_pairs(lowdin_it_handle h, int64_t *ij, int64_t *kl, double *v);
/* C record content (TransformIntegralsC.f90:420-425): p,q,r,s + value, p,q,r,s loop order. */
int lowdin_it_download_quads(lowdin_it_handle h, int32_t *p, int32_t *q, int32_t *r, int32_t *s, double *v);

/* Large-N streaming form: same transform, executed one occupied batch at a time
 * (occ_batch values of the FIRST contracted index per pass, 0 = library picks from free HBM),
 * results consumed on the device instead of being stored: sums[0]=count(|x|>tol),
 * sums[1]=sum x, sums[2]=sum x^2, sums[3]=MP2-like pair energy sum_x x(lambda x - x_exch)/den
 * when eps (host, length nao of species A then B) is given, else 0.
 * first_pass/n_passes select a sub-range of the occupied batches (n_passes<=0: all). */
int lowdin_it_transform_stream(lowdin_it_handle h, int slotA, int slotB, const int win[8], int conv,
                               double drop_tol, int occ_batch, int first_pass, int n_passes,
                               const double *epsA, const double *epsB, double lambda, double sums[4]);
/* The same, and every dense block of results also goes to the HOST: the MO integrals of N_bf = 1500 (328 GB for the MP2 window)
 * fit neither the device nor one download, so they leave block by block while the transform runs.  A block holds the integrals of
 * `nslots` first pairs: values[slot][ks][kf] = (a_slot b_slot | r s) with r, s the orbitals orb_second0 + ks, orb_first0 + kf in the
 * order (r, s) when second_is_conv_first != 0, else (s, r); slot_a/slot_b are the first pair's orbital numbers in convention order
 * ((i,j) for E, (p,q) for C).  No threshold is applied: the consumer drops |x| <= 1e-10 as it writes its records
 * (TransformIntegralsE.f90:1242).  `values` is pinned memory owned by the library, valid during the callback, which runs on the
 * calling thread while the next block is being computed and copied; a non-zero return aborts the transform. */
typedef struct lowdin_it_block {
  int conv, nslots, n_second, n_first, orb_second0, orb_first0, second_is_conv_first;
  const int32_t *slot_a, *slot_b;
  const double *values;
} lowdin_it_block;
typedef int (*lowdin_it_sink_fn)(void *user, const lowdin_it_block *block);
int lowdin_it_transform_stream_sink(lowdin_it_handle h, int slotA, int slotB, const int win[8], int conv, double drop_tol,
                                    int occ_batch, int first_pass, int n_passes, const double *epsA, const double *epsB,
                                    double lambda, double sums[4], lowdin_it_sink_fn sink, void *user);
/* On a communicator (lowdin_it_comm_init, nranks > 1) lowdin_it_transform_stream is COLLECTIVE: every rank makes the same
 * call with the same arguments.  With occ_batch == 0 so is lowdin_it_stream_num_passes (the ranks agree on the batch that fits
 * the rank with the least free memory). */
int lowdin_it_stream_num_passes(lowdin_it_handle h, int slotA, int slotB, const int win[8], int conv,
                                int occ_batch, int *n_passes, int *occ_batch_used);

/* ---- transformer-D compatible entry points ---------------------------------------- */
/* Same arguments and in-place semantics as c_integrals_transform_all /
 * c_integrals_transform_inter_all (IntTransfD.h:66-68, IntTransfD.cpp:125-181, :245-323):
 * full transform of the lower-triangular 0-based packed tensor ERIS[ij(ij+1)/2+kl]
 * (inter: ERIS[ij*om+kl]), but returning a status and using 64-bit sizes. */
int lowdin_it_transform_all(const double *coeff, double *ints, int nao);
int lowdin_it_transform_inter_all(const double *coeff, const double *ocoeff, double *ints, int nao, int onao);

/* ---- multi-GPU (one process per GPU; first half sharded over AO pair slabs,
 *      NCCL all-to-all, second half sharded over MO pairs) ----------------------------- */
/* The division of work the library uses (no device needed; the CPU multi-rank test drives it).
 * First half: the AO-pair slabs are distributed BLOCK-CYCLICALLY, blocks of 2^log_block consecutive slabs, block b on rank
 * b % nranks (LOWDIN_IT_OPT_SLAB_BLOCK_LOG, default 5); a rank numbers its own slabs consecutively ("local" slab number), which is
 * also its row in a stored AO tensor uploaded on a communicator (each rank keeps only the full M-vectors of its own slabs).
 * Second half: own[r]..own[r+1] = slots of rank r (fbeg[f] = first slot of first-contracted index f, nfb+1 entries).
 * For the chunk of slabs [chunk_base, chunk_base+chunk_width): `rank` computes its local slabs [loc_lo, loc_lo+count);
 * wblk = the largest count over the ranks = row stride of the exchanged blocks.  After the all-to-all the owner of a slot holds
 * element (its local slot `row`, global slab `slab`) at lowdin_it_exchanged_offset(). */
int lowdin_it_shard_plan(int nfb, const int *fbeg, int64_t chunk_base, int64_t chunk_width, int nranks, int rank, int log_block,
                         int *own, int64_t *wblk, int64_t *loc_lo, int64_t *count);
/* 1 when the all-to-all of this handle's communicator runs as peer-to-peer DMA (LOWDIN_IT_OPT_EXCHANGE_DMA and the node link was set up), else 0 */
int lowdin_it_exchange_is_dma(lowdin_it_handle h);
/* The occupied batch lowdin_it_transform_stream picks when occ_batch == 0, as a pure function (host logic, no device): n_first
 * first-window values, at most q_max per pass (memory), slots_per_first window pairs per first-window value, second-half window
 * sizes n_first2 / basis nao2, first-pair basis nao1, nslabs AO-pair slabs of npairs1 doubles, avail_bytes of device memory for the
 * third-quarter accumulators + chunk buffers of one rank, stored != 0 when the AO tensor is re-read by every pass. */
int lowdin_it_occ_batch_model(int n_first, int q_max, int nranks, int64_t slots_per_first, int n_first2, int nao2, int nao1, int64_t nslabs,
                              int64_t npairs1, double avail_bytes, int stored);
int lowdin_it_slab_owner(int64_t slab, int nranks, int log_block);
int64_t lowdin_it_slab_local(int64_t slab, int nranks, int log_block);
int64_t lowdin_it_slab_global(int64_t local_slab, int nranks, int rank, int log_block);
int64_t lowdin_it_exchanged_offset(int64_t row, int64_t slab, int64_t chunk_base, int64_t wblk, int64_t rows, int nranks, int log_block);
int lowdin_it_comm_unique_id(char id[128]);
int lowdin_it_comm_init(lowdin_it_handle h, int rank, int nranks, const char id[128]);
/* The same collective semantics for handles of ONE process (rank r = handles[r], each then driven by its own host thread):
 * the all-to-all is a set of direct peer copies ordered by CUDA events.  The handles may share a device, which is how the
 * multi-rank division of work is parity-tested on a one-GPU box. */
int lowdin_it_comm_init_local(lowdin_it_handle *handles, int nranks);
/* What a ONE-PROCESS host (the reference's transformation program is one process) calls on such a group: the transform runs on
 * all handles at once (one internal host thread per handle); the downloads return the merged list in the reference's order, as a
 * single GPU would.  Inputs (lowdin_it_set_species, the lowdin_it_ao_* upload) go to every handle of the group; each keeps only
 * the rows of the AO tensor it owns (2/nranks of the packed tensor per GPU). */
int lowdin_it_group_transform(lowdin_it_handle *handles, int nranks, int slotA, int slotB, const int win[8], int conv, int symmetric,
                              double drop_tol);
int lowdin_it_group_result_count(lowdin_it_handle *handles, int nranks, int64_t *count);
int lowdin_it_group_download_pairs(lowdin_it_handle *handles, int nranks, int64_t *ij, int64_t *kl, double *v);
int lowdin_it_group_download_quads(lowdin_it_handle *handles, int nranks, int32_t *p, int32_t *q, int32_t *r, int32_t *s, double *v);

/* ---- tuning ----------------------------------------------------------------------- */
#define LOWDIN_IT_OPT_WORKSPACE_BYTES 1 /* size of each slab-batch workspace (default 1 GiB) */
#define LOWDIN_IT_OPT_CHUNK_COLS 2      /* cap on AO-pair columns per chunk of the half-transformed block (0 = from free HBM) */
#define LOWDIN_IT_OPT_Q1_VARIANT 3      /* fused generation + first quarter: 1 = shared-memory ring, 2 = L1 path without barriers, 3 = warp-specialised (8 generator + 8 DMMA warps, TMA), 4 = 4 vectorised generator warps + 8 DMMA warps with register double-buffering, 5 = 256-row tiles with 32-row DMMA warps and setmaxnreg register rebalancing */
#define LOWDIN_IT_OPT_BENCH_GEN 4       /* generator kind used by lowdin_it_kernel_bench kind 2 */
#define LOWDIN_IT_OPT_GEMM_VARIANT 5    /* quarter-transform GEMM: 1 = cp.async ring + block barrier, 2 = TMA + mbarrier, persistent */
#define LOWDIN_IT_OPT_SPLIT_ROW_TAIL 6  /* TMA GEMM: 1 (default) = the <= 80-row tail of a few-rows x many-columns product runs as a second, operand-swapped launch instead of a padded 128-row tile */
#define LOWDIN_IT_OPT_FRAG_PERM 7       /* TMA kernels: 1 = fragment rows permuted so that the 128-bit shared loads of a quarter-warp are conflict-free (default 1; 0 = the unpermuted mapping, bit-identical results) */
#define LOWDIN_IT_OPT_ASYNC_PUSH 8      /* 1 = the caller leaves pushed host buffers untouched until lowdin_it_ao_end: pushes return without waiting for their copies */
#define LOWDIN_IT_OPT_STAGING_BYTES 9   /* size of each of the two device staging buffers of the AO upload (default 96 MiB) */
#define LOWDIN_IT_OPT_AO_LIST 11        /* 1 = the uploads that follow keep the canonical AO list on the device as it comes (16 bytes per stored integral, no
                                         * M(M+1)/2 dense tensor); the first quarter is then LIST-DRIVEN: every integral is scattered with its <= 4 images into the
                                         * quarter-transformed slabs (the DIRECT first quarter of Libint2Iface.cpp:793-853 / TransformIntegralsC.f90:545-558) */
#define LOWDIN_IT_OPT_SLAB_BLOCK_LOG 12 /* log2 of the block of consecutive AO-pair slabs one rank owns in the block-cyclic first half (default 5); set before uploading */
#define LOWDIN_IT_OPT_EXCHANGE_DMA 19     /* NCCL communicator whose ranks share one node: 1 (default) = the all-to-all between the halves as peer-to-peer DMA pulls over cudaIpc-mapped chunk buffers (no SMs taken from the persistent compute kernels), 0 = grouped ncclSend/ncclRecv; set alike on every rank */
#define LOWDIN_IT_OPT_Q3_TWO_CTA 18       /* third quarter: accumulating products whose K (pair rows of the chunk) is <= this value run as two independent 4-warp CTAs per SM, so that one tile's read-modify-write epilogue overlaps the other's DMMAs; 0 = off */
#define LOWDIN_IT_OPT_GEMM_TALL 17        /* TMA GEMM: 1 = 192 x 64 tiles for few-rows x many-columns products whose row count is a multiple of 192 plus <= 16 (1350 = 7 x 192 + 6) */
#define LOWDIN_IT_OPT_SINK_BLOCK_BYTES 16 /* size of one dense block handed to the host sink (default 256 MiB; at least one first-contracted index worth of pairs is always sent); setting it allocates the pinned two-slot ring at once */
#define LOWDIN_IT_OPT_OVERLAP_EXCHANGE 15 /* N > 1 ranks: two sets of chunk buffers, the all-to-all of chunk c on its own stream under the first half of chunk c + 1: 0 = never, 1 (default) = when the pass needs at most ~12 chunks (the second buffer set makes chunks 3/4 as wide), 2 = always */
#define LOWDIN_IT_OPT_STORED_FUSED 14   /* stored AO tensors: 1 (default) = the packed rows feed the first quarter's DMMA warps directly (unpack fused, no dense slab in HBM), 0 = expansion kernel + DMMA GEMM */
#define LOWDIN_IT_OPT_Q1_DEBUG 13       /* probe switches of the warp-specialised first quarter (timing experiments only; results are wrong when set) */
#define LOWDIN_IT_OPT_Q3_RED 10         /* third-quarter accumulation into T3: 0 = read-modify-write epilogue staged through shared memory, 1 = one red.global.add.f64 per element */
int lowdin_it_set_option(lowdin_it_handle h, int option, int64_t value);

/* ---- instrumentation -------------------------------------------------------------- */
/* out[0]=AO upload+scatter, [1]=first half, [2]=exchange, [3]=second half, [4]=compaction/consume,
 * [5]=download, [6]=algorithmic flops of the last transform, [7]=kernels launched by it.
 * Times in seconds (CUDA events on the library's stream). */
int lowdin_it_timers(lowdin_it_handle h, double out[8]);
/* Per-kernel-category device timing of the transforms that follow (CUDA event pairs around every
 * launch on the library's stream).  Categories: 0 slab expansion (1st half), 1 first quarter,
 * 2 second quarter + H scatter, 3 slab expansion (2nd half), 4 third quarter, 5 fourth quarter,
 * 6 consumer, 7 exchange.  work[] = algorithmic flops (GEMMs) or bytes (expansion, consumer). */
int lowdin_it_set_profiling(lowdin_it_handle h, int on);
int lowdin_it_kernel_stats(lowdin_it_handle h, double ms[8], double launches[8], double work[8]);
/* Stand-alone kernel entry points used by tests and bench.py to time one kernel on device
 * buffers owned by the handle. kind: 0 = slab expansion, 1 = DGEMM (DMMA) m x n x k. */
int lowdin_it_kernel_bench(lowdin_it_handle h, int kind, int64_t m, int64_t n, int64_t k, int iters,
                           double *ms_per_launch, double *check);
/* Parity-test access to single kernels: C[m][n] = sum_k A[m][k] B[n][k] on host arrays through the
 * DMMA kernel; dense expansion of nb slabs of an uploaded / generated AO set to host X[nb][n][n]. */
int lowdin_it_debug_gemm(lowdin_it_handle h, const double *A, const double *B, double *C, int m, int n, int k);
int lowdin_it_debug_expand(lowdin_it_handle h, int slotA, int slotB, int64_t slab0, int nb, double *X);
/* First half only (TransformIntegralsE.f90:1043-1132) of AO-pair slabs [slab0, slab0+nslabs): out[k][z] for the k-th window
 * pair in convention order (E: ijmap order) -- the oracle check of the first half at sizes no CPU transforms whole. */
/* First quarter only: out[f][z][mu] = sum_nu AO(slab0+z; mu nu) C(nu, f_first+f), f < nf, z < nslabs (host, nf*nslabs*nao doubles). */
int lowdin_it_debug_first_quarter(lowdin_it_handle h, int slotA, int slotB, int f_first, int nf, int64_t slab0, int nslabs, double *out);
int lowdin_it_debug_first_half(lowdin_it_handle h, int slotA, int slotB, const int win[8], int conv, double drop_tol,
                               int64_t slab0, int nslabs, double *out, int64_t *npairs);

#ifdef __cplusplus
}
#endif
#endif /* LOWDIN_IT_H */
