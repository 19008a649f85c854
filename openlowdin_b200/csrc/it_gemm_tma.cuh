// it_gemm_tma.cuh -- FP64 tensor-core GEMM with a TMA + mbarrier producer/consumer pipeline (sm_100a).
//
//   C[m][n] = sum_k A[m][k] B[n][k]        A: [M][K], B: [N][K], both K-contiguous (the quarter transforms of
//                                          TransformIntegralsE.f90:1081-1110, :1213-1239 in this layout)
//
// Differences from dgemm_tn_kernel (it_kernels.cuh):
//   * ONE lane (lane 0 of warp 0) issues cp.async.bulk.tensor (TMA) loads of 128-byte-swizzled [rows][16 doubles]
//     boxes into a STAGES-deep shared-memory ring, each load signalling the stage's "full" mbarrier; no thread computes
//     a per-element global address and the warps never meet at a block-wide barrier: each warp waits on the stage's
//     mbarrier, runs its DMMAs and arrives on the stage's "empty" mbarrier, which the producing lane polls without
//     blocking.  (A dedicated producer warp would make the CTA 288 threads, which the register file allocates as 384
//     and caps at 168 registers per thread -- the 128x128 tile needs more -- so the producer rides on a consumer warp.)
//   * the kernel is persistent (one CTA per SM, tiles walked m-fastest so that concurrently processed tiles share the
//     streamed operand in L2) and the ring keeps filling with the next tile's k-tiles while the warps are in the epilogue;
//   * fragments are read with 128-bit shared loads: lane (grp,tig) reads the 16-byte chunk (4h+tig) of row grp, i.e.
//     k = 8h+2tig and 8h+2tig+1, and feeds the two values to two DMMAs.  Both operands use the same k permutation,
//     so the sum over k is unchanged.  With the 128B swizzle (chunk ^= row & 7) a warp's 512 bytes hit every bank
//     group exactly four times: the 4-wavefront minimum, no padding.
// K tails and row tails are zero-filled by TMA (out-of-bounds box elements), so there is no predication in the loop.
#pragma once
#include <cuda.h>
#include "it_kernels.cuh"

namespace lowdin {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok;
}
// Non-blocking probe (try_wait may suspend the thread for a system-dependent time; test_wait never does).
__device__ __forceinline__ uint32_t mbar_test_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok;
}
// Bounded wait: a protocol error becomes a kernel fault (reported through the C ABI) instead of a hung device.  The bound
// is ELAPSED TIME (20 s of %globaltimer, looked at every 2^16 failed polls), not a poll count: a legitimately long stall
// (time-slicing under MPS, a profiler replay, a debugger) must not kill the context.
__device__ __forceinline__ uint64_t global_timer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  uint64_t t0 = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 0xffffu) == 0) {
      const uint64_t now = global_timer_ns();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 20000000000ull) __trap();
    }
  }
}
// Consumer release of a ring stage.  The stage's fragments were read with ld.shared; ptxas schedules the mbarrier arrive right behind
// the ISSUE of the last of those loads, ahead of the DMMAs that wait for their data (SASS of q1_load_ws5_kernel: LDS, LDS, LDS,
// SYNCS.ARRIVE, DMMA ...).  A shared load that is still in flight when the producer is released can then be overtaken by the
// producer's refill of the stage: seen on B200 with the loader warps' global loads backing up the memory pipe -- about one
// 32-row x 24-column piece of one tile in 10^4, always the columns of the last-loaded fragments.  The fence makes every lane's
// loads of the stage complete before the warp's arrive.
__device__ __forceinline__ void release_stage(uint32_t empty_bar, int lane) {
  __threadfence_block();
  __syncwarp();
  if (lane == 0) mbar_arrive(empty_bar);
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
               "l"(map), "r"(bar), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap *map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

struct TmaGemmShape {
  int M, N, K;
};

// Fragment-row permutation (PERM template parameter of the kernels below; LOWDIN_IT_OPT_FRAG_PERM).
// A 128-bit shared load is served one quarter-warp (8 lanes, 128 bytes) per cycle.  With lane (grp,tig) on row grp of an 8-row
// group, the lanes of quarter-warp p sit on rows 2p and 2p+1; under the 128-byte swizzle (chunk ^= row & 7) both rows put
// their four chunks into the SAME half of the 128-byte line, i.e. every quarter-warp has a 2-way bank conflict and every
// fragment load takes 8 wavefronts instead of 4 (ncu: 50 % excessive shared wavefronts, profiles/r01e_ncu_full_q1ws_n1500).
// rho(grp) = (grp >> 1) | ((grp & 1) << 2) puts the two lane groups of a quarter-warp on rows p and p + 4, whose swizzle keys
// differ in bit 2: their chunks fall into different halves of the line and the conflict is gone.  The MMA then computes the
// 8x8 tile with its rows AND columns permuted by rho: accumulator element [0]/[1] of lane (grp,tig) is output element
// (m = rho(grp), n = rho(2 tig) = tig / rho(2 tig + 1) = tig + 4) instead of (grp, 2 tig / 2 tig + 1); the epilogues index
// accordingly.  Shared-memory contents, TMA maps and the k order are unchanged, so results are bit-identical.
template <bool PERM>
__device__ __forceinline__ int frag_row(int grp) { return PERM ? ((grp >> 1) | ((grp & 1) << 2)) : grp; }
template <bool PERM>
__device__ __forceinline__ int frag_col0(int tig) { return PERM ? tig : 2 * tig; }      // column of accumulator element [0]
template <bool PERM>
__device__ __forceinline__ int frag_col1(int tig) { return PERM ? tig + 4 : 2 * tig + 1; }  // column of accumulator element [1]

constexpr int TMA_STAGE_LDM = 34;  // staged epilogue: [n][34] doubles per warp (32 rows + padding)
template <int BM, int BN, int STAGES>
constexpr size_t tma_gemm_smem_bytes(int warps = 0, int tn = 0) {
  // ring + barriers + alignment slack (+ the per-warp staging tiles of a row-coalescing epilogue)
  return (size_t)STAGES * (BM + BN) * 128 + 2 * STAGES * 8 + 1024 + (size_t)warps * tn * 8 * TMA_STAGE_LDM * 8;
}

// OCC = CTAs resident per SM: 2 for the short-K accumulating products of the third quarter, whose read-modify-write epilogue is as
// long as their main loop -- two independent CTAs of 4 warps per SM let one tile's epilogue run under the other's DMMAs.
template <int BM, int BN, int WM, int WN, int STAGES, class Epi, bool PERM = false, int OCC = 1>
__global__ void __launch_bounds__(WM *WN * 32, OCC)
    dgemm_tma_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, TmaGemmShape g, Epi epi) {
  constexpr int BK = 16;
  constexpr int NCW = WM * WN;  // warps; all of them consume, lane 0 of warp 0 also produces
  constexpr int TM = BM / WM / 8, TN = BN / WN / 8;
  constexpr uint32_t A_BYTES = BM * 128, B_BYTES = BN * 128, STAGE_BYTES = A_BYTES + B_BYTES;
  static_assert(BM % (WM * 8) == 0 && BN % (WN * 8) == 0, "tile/warp mismatch");
  static_assert(BM <= 256 && BN <= 256, "TMA box rows");
  static_assert(A_BYTES % 1024 == 0 && B_BYTES % 1024 == 0, "swizzle atom alignment");
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *ring = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t ring_u = smem_u32(ring);
  const uint32_t bars = ring_u + STAGES * STAGE_BYTES;  // full[s] at bars + 8 s, empty[s] at bars + 8 (STAGES + s)
  double *staging = reinterpret_cast<double *>(ring + STAGES * STAGE_BYTES + 2 * STAGES * 8);  // Epi::kRowCoalesced only

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int KT = (g.K + BK - 1) / BK;
  const int tiles_m = (g.M + BM - 1) / BM, tiles_n = (g.N + BN - 1) / BN;
  const int ntiles = tiles_m * tiles_n;
  // k-tiles this CTA goes through, over all of its output tiles (blockIdx.x, +gridDim.x, ...)
  const int my_tiles = (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const uint32_t total_it = (uint32_t)my_tiles * (uint32_t)KT;

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bars + 8 * s, 1);               // the producer's arrive.expect_tx
      mbar_init(bars + 8 * (STAGES + s), NCW);  // one arrive per warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tma_prefetch_desc(&mapA);
    tma_prefetch_desc(&mapB);
  }
  __syncthreads();

  // ---- producer side (lane 0 of warp 0): load number p of this CTA's k-tile sequence -> ring slot p % STAGES ----
  const bool producer = (tid == 0);
  uint32_t prod_it = 0;
  auto issue = [&](uint32_t p) {
    const int tile = (int)blockIdx.x + (int)(p / (uint32_t)KT) * (int)gridDim.x;
    const int kt = (int)(p % (uint32_t)KT);
    const int tm = tile % tiles_m, tn = tile / tiles_m;
    const uint32_t s = p % STAGES, full = bars + 8 * s;
    mbar_expect_tx(full, STAGE_BYTES);
    tma_load_2d(ring_u + s * STAGE_BYTES, &mapA, full, kt * BK, tm * BM);
    tma_load_2d(ring_u + s * STAGE_BYTES + A_BYTES, &mapB, full, kt * BK, tn * BN);
  };
  // the slot of load p was last used by load p - STAGES: it is free once every warp has arrived for that one
  auto slot_free_parity = [&](uint32_t p) { return ((p / STAGES) & 1u) ^ 1u; };
  if (producer)
    for (; prod_it < total_it && prod_it < (uint32_t)STAGES; ++prod_it) issue(prod_it);  // fresh slots

  const int wm = warp / WN, wn = warp % WN;
  const int tig = lane & 3;
  const int grp = frag_row<PERM>(lane >> 2);           // row of the 8-row group this lane's fragments come from
  const uint32_t off0 = (uint32_t)((tig ^ grp) << 4);  // swizzled chunk of k-half 0; half 1 is off0 ^ 64
  const uint8_t *a_base = ring + (wm * TM * 8 + grp) * 128;
  const uint8_t *b_base = ring + A_BYTES + (wn * TN * 8 + grp) * 128;

  uint32_t it = 0;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int tm = tile % tiles_m, tn = tile / tiles_m;
    const int m0 = tm * BM, n0 = tn * BN;
    double acc[TM][TN][2];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
      for (int j = 0; j < TN; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    for (int kt = 0; kt < KT; ++kt, ++it) {
      const uint32_t s = it % STAGES, ph = (it / STAGES) & 1u;
      if (producer) {
        // the load this iteration needs must be in flight before waiting for it (blocks only if a warp lags a whole ring)
        while (prod_it <= it) {
          mbar_wait(bars + 8 * (STAGES + prod_it % STAGES), slot_free_parity(prod_it));
          issue(prod_it);
          ++prod_it;
        }
      }
      mbar_wait(bars + 8 * s, ph);
      const uint8_t *as = a_base + s * STAGE_BYTES, *bs = b_base + s * STAGE_BYTES;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const uint32_t off = h ? (off0 ^ 64u) : off0;
        double2 a[TM], b[TN];
#pragma unroll
        for (int i = 0; i < TM; ++i) a[i] = *reinterpret_cast<const double2 *>(as + i * 8 * 128 + off);
#pragma unroll
        for (int j = 0; j < TN; ++j) b[j] = *reinterpret_cast<const double2 *>(bs + j * 8 * 128 + off);
#pragma unroll
        for (int i = 0; i < TM; ++i)
#pragma unroll
          for (int j = 0; j < TN; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i].x, b[j].x);
#pragma unroll
        for (int i = 0; i < TM; ++i)
#pragma unroll
          for (int j = 0; j < TN; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i].y, b[j].y);
      }
      release_stage(bars + 8 * (STAGES + s), lane);
      if (producer) {
        // refill, without blocking, every slot that all warps have released (runs ahead into the next tile)
        while (prod_it < total_it && prod_it < it + 1u + (uint32_t)STAGES) {
          if (!mbar_test_wait(bars + 8 * (STAGES + prod_it % STAGES), slot_free_parity(prod_it))) break;
          issue(prod_it);
          ++prod_it;
        }
      }
    }

    if constexpr (Epi::kRowCoalesced && TM == 4) {
      // Output elements that are consecutive in m are consecutive in memory (third-quarter accumulators, mu contiguous):
      // turn the warp's 32 x (8 TN) accumulator tile in shared memory so that lane l owns row l, then read-modify-write
      // column by column with 256-byte coalesced accesses instead of 64-byte pieces per lane group.
      double *st = staging + warp * (TN * 8 * TMA_STAGE_LDM);
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) {
          if constexpr (PERM) {
            st[(j * 8 + tig) * TMA_STAGE_LDM + i * 8 + grp] = acc[i][j][0];
            st[(j * 8 + tig + 4) * TMA_STAGE_LDM + i * 8 + grp] = acc[i][j][1];
          } else {
            st[(j * 8 + tig * 2) * TMA_STAGE_LDM + i * 8 + grp] = acc[i][j][0];
            st[(j * 8 + tig * 2 + 1) * TMA_STAGE_LDM + i * 8 + grp] = acc[i][j][1];
          }
        }
      __syncwarp();
      const int m = m0 + wm * 32 + lane;
      const int nb = n0 + wn * TN * 8;
      if (m < g.M) {
        double *row = epi.row_ptr(m);
        const int64_t cs = epi.col_stride();
        // all loads of the warp tile first (the accumulator registers are free by now): 8 TN x 256 bytes in flight per
        // warp, which is what it takes to cover HBM latency with one CTA per SM; then add and store
        double t[TN * 8];
#pragma unroll
        for (int u = 0; u < TN * 8; ++u) t[u] = (nb + u < g.N) ? row[(int64_t)(nb + u) * cs] : 0.0;
#pragma unroll
        for (int u = 0; u < TN * 8; ++u)
          if (nb + u < g.N) row[(int64_t)(nb + u) * cs] = t[u] + st[u * TMA_STAGE_LDM + lane];
      }
      __syncwarp();
    } else {
#pragma unroll
      for (int i = 0; i < TM; ++i) {
        const int m = m0 + wm * TM * 8 + i * 8 + grp;
        if (m >= g.M) continue;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
          if constexpr (PERM) {
            const int n = n0 + wn * TN * 8 + j * 8 + tig;
            if (n < g.N) epi(0, m, n, acc[i][j][0]);
            if (n + 4 < g.N) epi(0, m, n + 4, acc[i][j][1]);
          } else {
            const int n = n0 + wn * TN * 8 + j * 8 + tig * 2;
            if (n < g.N) epi(0, m, n, acc[i][j][0]);
            if (n + 1 < g.N) epi(0, m, n + 1, acc[i][j][1]);
          }
        }
      }
    }
  }
}

// =================================================================================================================
// Warp-specialised fused slab generation + first quarter for GENERATED AO sources (q1 variant 3):
//   T1t[f][z][mu] = sum_nu AO(slab0+z ; pair(mu,nu)) * C(nu, f)                      (E.f90:1047-1090 in one kernel)
// The ncu capture of the single-role kernel (profiles/r01b_ncu_summary.md) shows the FP64 tensor pipe at 67 %: every
// warp alternates ~50 dependent integer instructions per generated value with its DMMAs, and since a warp issues in
// order, a warp inside its hash chain has no DMMA to offer and a warp waiting for the tensor pipe cannot hash.  Here the
// two jobs run in different warps of a 16-warp CTA (two of each kind per scheduler):
//   generators (warps 8-15)  hash the A operand (the 128 x 16 slab tile of the k-tile, as unscaled integers-in-doubles,
//                            see bits_to_unscaled) and store it into a STAGES-deep shared-memory ring in the same
//                            128-byte-swizzled layout TMA produces; one lane also issues the TMA load of the stage's
//                            tile of the coefficient window scaled by 2^-53 (B operand);
//   consumers  (warps 0-7)   wait on the stage's "full" mbarrier, read fragments with 128-bit shared loads and do
//                            nothing but DMMAs, then arrive on the stage's "empty" mbarrier.
// Persistent: CTAs walk (slab, 128-row block) tiles; the generators run ahead into the next tile during the epilogue.
// =================================================================================================================
// Generated AO value WITHOUT its 2^-53 scale: the integer t = 2 (bits >> 11) - 2^53 as a double (exact, |t| <= 2^53).
// value = t * 2^-53 (bits_to_value); the power of two is folded into the coefficient operand (Species::Cs = C * 2^-53,
// exact), so the DMMA products t * (c 2^-53) round exactly like (t 2^-53) * c and T1t is bit-identical to the other
// variants, while the exponent fix-up (two compares, an add and a select per value) disappears from the generator.
//   t = ((int64)(bits ^ 2^63) >> 10) & ~1 :  bits ^ 2^63 is bits - 2^63 as a signed number, the arithmetic shift gives
//   floor(bits / 2^10) - 2^53 = 2 (bits >> 11) + bit10 - 2^53, and clearing bit 0 removes bit10.
__device__ __forceinline__ double bits_to_unscaled(uint64_t bits) {
  const long long t = ((long long)(bits ^ 0x8000000000000000ull) >> 10) & ~1ll;
  return __ll2double_rn(t);
}
// base(x) = x n - x (x - 1) / 2 - x, so that the 0-based pair id of (lo, hi >= lo) is base(lo) + hi   (C.f90:214-221)
__device__ __forceinline__ uint32_t pair_base(uint32_t x, uint32_t n) { return x * n - ((x * (x - 1u)) >> 1) - x; }

template <int KIND, int GEN>
__device__ __forceinline__ double gen_unscaled(uint32_t slab, uint32_t mu, uint32_t nu, uint32_t base_mu, uint32_t base_nu, uint32_t m32,
                                               uint64_t seed) {
  const uint32_t pair = (nu >= mu) ? base_mu + nu : base_nu + mu;
  uint64_t key;
  if (KIND == SRC_HASH_SYM) {
    const uint32_t a = min(slab, pair), b = max(slab, pair);
    key = (uint64_t)b * (uint64_t)m32 + a;  // m32 = M (pairs per slab vector) < 2^32
  } else {
    key = (uint64_t)pair * (uint64_t)m32 + slab;  // m32 = number of slabs (M_b)
  }
  return bits_to_unscaled(GEN == 2 ? mulfold64(seed ^ key) : splitmix64(seed ^ key));
}

struct Q1WsArgs {
  int64_t slab0;  // first slab of the batch, in the rank's local numbering (slab_global() gives the id the generator hashes)
  int logB, G, rank;
  int bc, nc, nfb;
  uint32_t m32;   // SRC_HASH_SYM: pairs per slab vector (M); SRC_HASH_RECT: number of slabs (M_b)
  uint64_t seed;
  double *T1t;
  int64_t ldt;
  int dbg;  // probe switches (LOWDIN_IT_OPT_Q1_DEBUG; results are wrong when set): 1 = generator warps store zeros instead of hashing,
            // 2 = DMMA warps skip loads and DMMAs, 4 = variant 3 only: roles by scheduler (warps with bit 1 set generate)
};

template <int TN, int STAGES>
constexpr size_t q1_ws_smem_bytes() {
  return (size_t)STAGES * (128 + TN * 8) * 128 + 2 * STAGES * 8 + 1024;
}

template <int TN, int STAGES, int KIND, int GEN, bool PERM = false>
__global__ void __launch_bounds__(512, 1) q1_gen_ws_kernel(const __grid_constant__ CUtensorMap mapB, Q1WsArgs q) {
  constexpr int BK = 16, BM = 128, BN = TN * 8;
  constexpr int NCW = 8, NGW = 8;  // consumer / generator warps
  constexpr uint32_t A_BYTES = BM * 128, B_BYTES = BN * 128, STAGE_BYTES = A_BYTES + B_BYTES;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *ring = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t ring_u = smem_u32(ring);
  const uint32_t bars = ring_u + STAGES * STAGE_BYTES;  // full[s] at bars + 8 s, empty[s] at bars + 8 (STAGES + s)

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tig = lane & 3;
  const int grp = frag_row<PERM>(lane >> 2);  // generators WRITE and consumers READ row grp of each 8-row group (see frag_row)
  const int KT = (q.nc + BK - 1) / BK;
  const int row_blocks = (q.nc + BM - 1) / BM;
  const int ntiles = row_blocks * q.bc;

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bars + 8 * s, NGW + 1);         // one arrive per generator warp + the arrive.expect_tx of the TMA issuer
      mbar_init(bars + 8 * (STAGES + s), NCW);  // one arrive per consumer warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tma_prefetch_desc(&mapB);
  }
  __syncthreads();

  const uint32_t off0 = (uint32_t)((tig ^ grp) << 4);  // swizzled chunk tig of a row with (row & 7) == grp; chunk 4+tig is off0 ^ 64

  // roles: warps 8-15 generate (two of each kind per scheduler); probe switch 4: the warps of schedulers 2 and 3 generate
  const bool split = (q.dbg & 4) != 0;
  const bool is_gen = split ? ((warp & 2) != 0) : (warp >= NCW);
  const int role_idx = split ? (((warp >> 2) << 1) | (warp & 1)) : (warp & 7);  // 0..7 within the role
  if (is_gen) {
    // ============================== generators ==============================
    const int gw = role_idx;
    const uint32_t n = (uint32_t)q.nc;
    uint8_t *a_rows = ring + (gw * 16 + grp) * 128;  // rows 16 gw + grp and + 8
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const int z = tile / row_blocks, rb = tile - z * row_blocks;
      const uint32_t slab = (uint32_t)slab_global(q.slab0 + z, q.logB, q.G, q.rank);
      const uint32_t mu0 = (uint32_t)(rb * BM + gw * 16 + grp);
      const uint32_t base_mu[2] = {pair_base(mu0, n), pair_base(mu0 + 8u, n)};
      for (int kt = 0; kt < KT; ++kt, ++it) {
        const uint32_t s = it % STAGES;
        if (it >= (uint32_t)STAGES) mbar_wait(bars + 8 * (STAGES + s), ((it / STAGES) & 1u) ^ 1u);
        if (gw == 0 && lane == 0) {
          mbar_expect_tx(bars + 8 * s, B_BYTES);
          tma_load_2d(ring_u + s * STAGE_BYTES + A_BYTES, &mapB, bars + 8 * s, kt * BK, 0);
        }
        uint8_t *dst = a_rows + s * STAGE_BYTES;
        const uint32_t nu0 = (uint32_t)(kt * BK + 2 * tig);
        // the lane's four k values of this k-tile (chunks tig and 4 + tig), shared by its two rows
        const uint32_t nus[4] = {nu0, nu0 + 1u, nu0 + 8u, nu0 + 9u};
        uint32_t base_nu[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) base_nu[c] = pair_base(nus[c], n);
#pragma unroll
        for (int rg = 0; rg < 2; ++rg) {
#pragma unroll
          for (int cg = 0; cg < 2; ++cg) {
            const uint32_t mu = mu0 + 8 * rg;
            double2 v;
            if (q.dbg & 1) { v.x = 0.0; v.y = 0.0; }
            else {
              v.x = gen_unscaled<KIND, GEN>(slab, mu, nus[2 * cg], base_mu[rg], base_nu[2 * cg], q.m32, q.seed);
              v.y = gen_unscaled<KIND, GEN>(slab, mu, nus[2 * cg + 1], base_mu[rg], base_nu[2 * cg + 1], q.m32, q.seed);
            }
            *reinterpret_cast<double2 *>(dst + rg * 8 * 128 + (cg ? (off0 ^ 64u) : off0)) = v;
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bars + 8 * s);
      }
    }
    return;
  }

  // ============================== consumers ==============================
  const int cw = role_idx;
  const uint8_t *a_base = ring + (cw * 16 + grp) * 128;
  const uint8_t *b_base = ring + A_BYTES + grp * 128;
  uint32_t it = 0;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int z = tile / row_blocks, rb = tile - z * row_blocks;
    double acc[2][TN][2];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < TN; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    for (int kt = 0; kt < KT; ++kt, ++it) {
      const uint32_t s = it % STAGES;
      mbar_wait(bars + 8 * s, (it / STAGES) & 1u);
      const uint8_t *as = a_base + s * STAGE_BYTES, *bs = b_base + s * STAGE_BYTES;
      if (!(q.dbg & 2))
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const uint32_t off = h ? (off0 ^ 64u) : off0;
        double2 a[2], b[TN];
#pragma unroll
        for (int i = 0; i < 2; ++i) a[i] = *reinterpret_cast<const double2 *>(as + i * 8 * 128 + off);
#pragma unroll
        for (int j = 0; j < TN; ++j) b[j] = *reinterpret_cast<const double2 *>(bs + j * 8 * 128 + off);
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
          for (int j = 0; j < TN; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i].x, b[j].x);
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
          for (int j = 0; j < TN; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i].y, b[j].y);
      }
      release_stage(bars + 8 * (STAGES + s), lane);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int m = rb * BM + cw * 16 + i * 8 + grp;
      if (m >= q.nc) continue;
#pragma unroll
      for (int j = 0; j < TN; ++j) {
        if constexpr (PERM) {
          const int f = j * 8 + tig;
          if (f < q.nfb) q.T1t[((int64_t)f * q.bc + z) * q.ldt + m] = acc[i][j][0];
          if (f + 4 < q.nfb) q.T1t[((int64_t)(f + 4) * q.bc + z) * q.ldt + m] = acc[i][j][1];
        } else {
          const int f = j * 8 + tig * 2;
          if (f < q.nfb) q.T1t[((int64_t)f * q.bc + z) * q.ldt + m] = acc[i][j][0];
          if (f + 1 < q.nfb) q.T1t[((int64_t)(f + 1) * q.bc + z) * q.ldt + m] = acc[i][j][1];
        }
      }
    }
  }
}

// =================================================================================================================
// q1 variant 4: the same two roles, re-balanced after the ncu capture of variant 3 (profiles/r01e_*): there the generator
// warps idle a third of the time on the "empty" barriers while the FP64 tensor pipe is only 75 % busy, because the DMMA
// warps spend ~20 % of their time outside DMMAs (fragment loads whose latency nothing covers, barrier polls) and the
// generator warps run one dependent hash chain at a time (IPC 0.23).  Here
//   * the 8 DMMA warps double-buffer their fragments in registers: the 128-bit shared loads of the next half k-tile are
//     issued BEFORE the 28 DMMAs of the current one (170 registers per thread are available with 384 threads, 128 with 512);
//   * 4 generator warps (one per scheduler) hash 8 values per lane in lock step, stage by stage, so that the eight
//     independent chains fill the integer pipes (32 rows x 16 k per warp and k-tile).
// =================================================================================================================
template <int GEN>
__device__ __forceinline__ void hash8(uint64_t (&x)[8]) {
  if (GEN == 2) {
#pragma unroll
    for (int v = 0; v < 8; ++v) x[v] *= 0x9E3779B97F4A7C15ull;
#pragma unroll
    for (int v = 0; v < 8; ++v) x[v] ^= x[v] >> 32;
#pragma unroll
    for (int v = 0; v < 8; ++v) x[v] *= 0xD6E8FEB86659FD93ull;
  } else {
#pragma unroll
    for (int v = 0; v < 8; ++v) x[v] += 0x9E3779B97F4A7C15ull;
#pragma unroll
    for (int v = 0; v < 8; ++v) x[v] = (x[v] ^ (x[v] >> 30)) * 0xBF58476D1CE4E5B9ull;
#pragma unroll
    for (int v = 0; v < 8; ++v) x[v] = (x[v] ^ (x[v] >> 27)) * 0x94D049BB133111EBull;
#pragma unroll
    for (int v = 0; v < 8; ++v) x[v] ^= x[v] >> 31;
  }
}

template <int TN, int STAGES, int KIND, int GEN, bool PERM = false>
__global__ void __launch_bounds__(384, 1) q1_gen_ws2_kernel(const __grid_constant__ CUtensorMap mapB, Q1WsArgs q) {
  constexpr int BK = 16, BM = 128, BN = TN * 8;
  constexpr int NCW = 8, NGW = 4;  // DMMA warps (16 rows each) / generator warps (32 rows each)
  constexpr uint32_t A_BYTES = BM * 128, B_BYTES = BN * 128, STAGE_BYTES = A_BYTES + B_BYTES;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *ring = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t ring_u = smem_u32(ring);
  const uint32_t bars = ring_u + STAGES * STAGE_BYTES;  // full[s] at bars + 8 s, empty[s] at bars + 8 (STAGES + s)

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tig = lane & 3;
  const int grp = frag_row<PERM>(lane >> 2);
  const int KT = (q.nc + BK - 1) / BK;
  const int row_blocks = (q.nc + BM - 1) / BM;
  const int ntiles = row_blocks * q.bc;
  const int my_tiles = (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const uint32_t total_it = (uint32_t)my_tiles * (uint32_t)KT;

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bars + 8 * s, NGW + 1);         // one arrive per generator warp + the arrive.expect_tx of the TMA issuer
      mbar_init(bars + 8 * (STAGES + s), NCW);  // one arrive per DMMA warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tma_prefetch_desc(&mapB);
  }
  __syncthreads();

  const uint32_t off0 = (uint32_t)((tig ^ grp) << 4);  // swizzled chunk tig of a row with (row & 7) == grp; chunk 4+tig is off0 ^ 64

  if (warp >= NCW) {
    // ============================== generators ==============================
    const int gw = warp - NCW;
    const uint32_t n = (uint32_t)q.nc;
    uint8_t *a_rows = ring + (gw * 32 + grp) * 128;  // rows 32 gw + 8 rg + grp, rg = 0..3
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const int z = tile / row_blocks, rb = tile - z * row_blocks;
      const uint32_t slab = (uint32_t)slab_global(q.slab0 + z, q.logB, q.G, q.rank);
      const uint32_t mu0 = (uint32_t)(rb * BM + gw * 32 + grp);
      uint32_t base_mu[4];
#pragma unroll
      for (int rg = 0; rg < 4; ++rg) base_mu[rg] = pair_base(mu0 + 8u * rg, n);
      for (int kt = 0; kt < KT; ++kt, ++it) {
        const uint32_t s = it % STAGES;
        if (it >= (uint32_t)STAGES) mbar_wait(bars + 8 * (STAGES + s), ((it / STAGES) & 1u) ^ 1u);
        if (gw == 0 && lane == 0) {
          mbar_expect_tx(bars + 8 * s, B_BYTES);
          tma_load_2d(ring_u + s * STAGE_BYTES + A_BYTES, &mapB, bars + 8 * s, kt * BK, 0);
        }
        uint8_t *dst = a_rows + s * STAGE_BYTES;
        const uint32_t nu0 = (uint32_t)(kt * BK + 2 * tig);
        const uint32_t nus[4] = {nu0, nu0 + 1u, nu0 + 8u, nu0 + 9u};  // chunks tig and 4 + tig
        uint32_t base_nu[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) base_nu[c] = pair_base(nus[c], n);
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {  // two row groups (8 values) at a time, all chains advanced together
          uint64_t x[8];
#pragma unroll
          for (int v = 0; v < 8; ++v) {
            const int rg = 2 * hh + (v >> 2), c = v & 3;
            const uint32_t mu = mu0 + 8u * rg;
            const uint32_t pair = (nus[c] >= mu) ? base_mu[rg] + nus[c] : base_nu[c] + mu;
            uint64_t key;
            if (KIND == SRC_HASH_SYM) {
              const uint32_t a = min(slab, pair), b = max(slab, pair);
              key = (uint64_t)b * (uint64_t)q.m32 + a;
            } else {
              key = (uint64_t)pair * (uint64_t)q.m32 + slab;
            }
            x[v] = q.seed ^ key;
          }
          hash8<GEN>(x);
          double d[8];
#pragma unroll
          for (int v = 0; v < 8; ++v) d[v] = bits_to_unscaled(x[v]);
#pragma unroll
          for (int r2 = 0; r2 < 2; ++r2) {
            const int rg = 2 * hh + r2;
            *reinterpret_cast<double2 *>(dst + rg * 8 * 128 + off0) = make_double2(d[4 * r2 + 0], d[4 * r2 + 1]);
            *reinterpret_cast<double2 *>(dst + rg * 8 * 128 + (off0 ^ 64u)) = make_double2(d[4 * r2 + 2], d[4 * r2 + 3]);
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bars + 8 * s);
      }
    }
    return;
  }

  // ============================== DMMA warps ==============================
  if (total_it == 0) return;
  const uint8_t *a_base = ring + (warp * 16 + grp) * 128;
  const uint8_t *b_base = ring + A_BYTES + grp * 128;
  double2 fa[2][2], fb[2][TN];  // [buffer][fragment]
  auto load_frags = [&](int buf, uint32_t stage, uint32_t off) {
    const uint8_t *as = a_base + stage * STAGE_BYTES, *bs = b_base + stage * STAGE_BYTES;
#pragma unroll
    for (int i = 0; i < 2; ++i) fa[buf][i] = *reinterpret_cast<const double2 *>(as + i * 8 * 128 + off);
#pragma unroll
    for (int j = 0; j < TN; ++j) fb[buf][j] = *reinterpret_cast<const double2 *>(bs + j * 8 * 128 + off);
  };
  uint32_t it = 0;
  mbar_wait(bars, 0);  // stage 0 of the first k-tile
  load_frags(0, 0, off0);
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int z = tile / row_blocks, rb = tile - z * row_blocks;
    double acc[2][TN][2];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < TN; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    for (int kt = 0; kt < KT; ++kt, ++it) {
      const uint32_t s = it % STAGES;
      // ---- k-half 0 from buffer 0; meanwhile fetch k-half 1 of the same stage into buffer 1 ----
      load_frags(1, s, off0 ^ 64u);
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) dmma884(acc[i][j][0], acc[i][j][1], fa[0][i].x, fb[0][j].x);
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) dmma884(acc[i][j][0], acc[i][j][1], fa[0][i].y, fb[0][j].y);
      // ---- k-half 1 from buffer 1; meanwhile fetch k-half 0 of the NEXT k-tile (next stage) into buffer 0 ----
      if (it + 1u < total_it) {
        const uint32_t s1 = (it + 1u) % STAGES;
        mbar_wait(bars + 8 * s1, ((it + 1u) / STAGES) & 1u);
        load_frags(0, s1, off0);
      }
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) dmma884(acc[i][j][0], acc[i][j][1], fa[1][i].x, fb[1][j].x);
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) dmma884(acc[i][j][0], acc[i][j][1], fa[1][i].y, fb[1][j].y);
      release_stage(bars + 8 * (STAGES + s), lane);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int m = rb * BM + warp * 16 + i * 8 + grp;
      if (m >= q.nc) continue;
#pragma unroll
      for (int j = 0; j < TN; ++j) {
        if constexpr (PERM) {
          const int f = j * 8 + tig;
          if (f < q.nfb) q.T1t[((int64_t)f * q.bc + z) * q.ldt + m] = acc[i][j][0];
          if (f + 4 < q.nfb) q.T1t[((int64_t)(f + 4) * q.bc + z) * q.ldt + m] = acc[i][j][1];
        } else {
          const int f = j * 8 + tig * 2;
          if (f < q.nfb) q.T1t[((int64_t)f * q.bc + z) * q.ldt + m] = acc[i][j][0];
          if (f + 1 < q.nfb) q.T1t[((int64_t)(f + 1) * q.bc + z) * q.ldt + m] = acc[i][j][1];
        }
      }
    }
  }
}

// =================================================================================================================
// q1 variant 5: 256-row tiles, 32-row DMMA warps, register file re-divided between the two roles.
// Measured on B200 (profiles/r02a_perm_probe.log): the fused first quarter scales with the number of DMMAs a warp runs per set of
// fragment loads -- 28.1 / 27.4 / 25.9 / 25.2 TF/s at 64 / 56 / 48 / 40 window columns, i.e. a pipe utilisation of
// TN / (TN + 2.6): every half k-tile costs a DMMA warp about 2.6 column-tiles' worth of tensor time in fragment loads and barrier
// traffic that its partner warp does not cover.  With 16-row warps the DMMA phase is 2 x 2 TN instructions per 2 + TN loads; a
// 32-row warp runs 4 x 2 TN per 4 + TN loads (0.196 loads per DMMA at TN = 7, the ratio at which the plain GEMM reaches cuBLAS
// speed) -- but needs 8 TN + 44 accumulator/fragment registers, more than the 128 a 512-thread CTA gives each thread.  So the
// roles trade registers (setmaxnreg, sm_90+): the 8 generator warps shrink to 88, the 8 DMMA warps grow to 168.
//   tile = 256 rows x 8 TN columns of one slab;  ring stage = 32 KB (A, generated) + TN KB (B, TMA);  5 stages.
// Fragment rows are permuted (frag_row<true>): conflict-free 128-bit shared accesses on both sides.
// =================================================================================================================
template <int TN, int STAGES>
constexpr size_t q1_ws5_smem_bytes() {
  return (size_t)STAGES * (256 + TN * 8) * 128 + 2 * STAGES * 8 + 1024;
}

template <int TN, int STAGES, int KIND, int GEN>
__global__ void __launch_bounds__(512, 1) q1_gen_ws5_kernel(const __grid_constant__ CUtensorMap mapB, Q1WsArgs q) {
  constexpr int BK = 16, BM = 256, BN = TN * 8;
  constexpr int NCW = 8, NGW = 8;  // DMMA warps (32 rows each) / generator warps (32 rows each)
  constexpr uint32_t A_BYTES = BM * 128, B_BYTES = BN * 128, STAGE_BYTES = A_BYTES + B_BYTES;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *ring = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t ring_u = smem_u32(ring);
  const uint32_t bars = ring_u + STAGES * STAGE_BYTES;  // full[s] at bars + 8 s, empty[s] at bars + 8 (STAGES + s)

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tig = lane & 3;
  const int grp = frag_row<true>(lane >> 2);
  const int KT = (q.nc + BK - 1) / BK;
  const int row_blocks = (q.nc + BM - 1) / BM;
  const int ntiles = row_blocks * q.bc;

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bars + 8 * s, NGW + 1);         // one arrive per generator warp + the arrive.expect_tx of the TMA issuer
      mbar_init(bars + 8 * (STAGES + s), NCW);  // one arrive per DMMA warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tma_prefetch_desc(&mapB);
  }
  __syncthreads();

  const uint32_t off0 = (uint32_t)((tig ^ grp) << 4);  // swizzled chunk tig of a row with (row & 7) == grp; chunk 4+tig is off0 ^ 64

  if (warp >= NCW) {
    // ============================== generators (warpgroups 2 and 3) ==============================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 88;");
    const int gw = warp - NCW;
    const uint32_t n = (uint32_t)q.nc;
    uint8_t *a_rows = ring + (gw * 32 + grp) * 128;  // rows 32 gw + 8 rg + grp, rg = 0..3
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const int z = tile / row_blocks, rb = tile - z * row_blocks;
      const uint32_t slab = (uint32_t)slab_global(q.slab0 + z, q.logB, q.G, q.rank);
      const uint32_t mu0 = (uint32_t)(rb * BM + gw * 32 + grp);
      uint32_t base_mu[4];
#pragma unroll
      for (int rg = 0; rg < 4; ++rg) base_mu[rg] = pair_base(mu0 + 8u * rg, n);
      for (int kt = 0; kt < KT; ++kt, ++it) {
        const uint32_t s = it % STAGES;
        if (it >= (uint32_t)STAGES) mbar_wait(bars + 8 * (STAGES + s), ((it / STAGES) & 1u) ^ 1u);
        if (gw == 0 && lane == 0) {
          mbar_expect_tx(bars + 8 * s, B_BYTES);
          tma_load_2d(ring_u + s * STAGE_BYTES + A_BYTES, &mapB, bars + 8 * s, kt * BK, 0);
        }
        uint8_t *dst = a_rows + s * STAGE_BYTES;
        const uint32_t nu0 = (uint32_t)(kt * BK + 2 * tig);
        const uint32_t nus[4] = {nu0, nu0 + 1u, nu0 + 8u, nu0 + 9u};  // chunks tig and 4 + tig
        uint32_t base_nu[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) base_nu[c] = pair_base(nus[c], n);
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {  // two row groups (8 values) at a time, all chains advanced together
          uint64_t x[8];
#pragma unroll
          for (int v = 0; v < 8; ++v) {
            const int rg = 2 * hh + (v >> 2), c = v & 3;
            const uint32_t mu = mu0 + 8u * rg;
            const uint32_t pair = (nus[c] >= mu) ? base_mu[rg] + nus[c] : base_nu[c] + mu;
            uint64_t key;
            if (KIND == SRC_HASH_SYM) {
              const uint32_t a = min(slab, pair), b = max(slab, pair);
              key = (uint64_t)b * (uint64_t)q.m32 + a;
            } else {
              key = (uint64_t)pair * (uint64_t)q.m32 + slab;
            }
            x[v] = q.seed ^ key;
          }
          if (!(q.dbg & 1)) hash8<GEN>(x);
#pragma unroll
          for (int r2 = 0; r2 < 2; ++r2) {
            const int rg = 2 * hh + r2;
            *reinterpret_cast<double2 *>(dst + rg * 8 * 128 + off0) = make_double2(bits_to_unscaled(x[4 * r2 + 0]), bits_to_unscaled(x[4 * r2 + 1]));
            *reinterpret_cast<double2 *>(dst + rg * 8 * 128 + (off0 ^ 64u)) = make_double2(bits_to_unscaled(x[4 * r2 + 2]), bits_to_unscaled(x[4 * r2 + 3]));
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bars + 8 * s);
      }
    }
    return;
  }

  // ============================== DMMA warps (warpgroups 0 and 1) ==============================
  asm volatile("setmaxnreg.inc.sync.aligned.u32 168;");
  const uint8_t *a_base = ring + (warp * 32 + grp) * 128;
  const uint8_t *b_base = ring + A_BYTES + grp * 128;
  uint32_t it = 0;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int z = tile / row_blocks, rb = tile - z * row_blocks;
    double acc[4][TN][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < TN; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    for (int kt = 0; kt < KT; ++kt, ++it) {
      const uint32_t s = it % STAGES;
      mbar_wait(bars + 8 * s, (it / STAGES) & 1u);
      const uint8_t *as = a_base + s * STAGE_BYTES, *bs = b_base + s * STAGE_BYTES;
      if (!(q.dbg & 2))
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const uint32_t off = h ? (off0 ^ 64u) : off0;
        double2 a[4], b[TN];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = *reinterpret_cast<const double2 *>(as + i * 8 * 128 + off);
#pragma unroll
        for (int j = 0; j < TN; ++j) b[j] = *reinterpret_cast<const double2 *>(bs + j * 8 * 128 + off);
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < TN; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i].x, b[j].x);
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < TN; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i].y, b[j].y);
      }
      release_stage(bars + 8 * (STAGES + s), lane);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int m = rb * BM + warp * 32 + i * 8 + grp;
      if (m >= q.nc) continue;
#pragma unroll
      for (int j = 0; j < TN; ++j) {
        const int f = j * 8 + tig;  // permuted fragment mapping: accumulator [0] is column tig, [1] column tig + 4
        if (f < q.nfb) q.T1t[((int64_t)f * q.bc + z) * q.ldt + m] = acc[i][j][0];
        if (f + 4 < q.nfb) q.T1t[((int64_t)(f + 4) * q.bc + z) * q.ldt + m] = acc[i][j][1];
      }
    }
  }
}

// =================================================================================================================
// Fused first quarter for STORED AO tensors: the packed row is the A operand, no dense slab in HBM.
//   T1t[f][z][mu] = sum_nu AO(slab0+z ; pair(mu,nu)) * C(nu, f)            (E.f90:1047-1090 in one kernel)
// The unpack of E.f90:1047-1063 happens on the way into shared memory: the kernel is q1 variant 5 with the hashing warps
// replaced by LOADER warps.  A loader lane owns rows mu = mu0 + 8 rg and, per k-tile, the four k values of its two 16-byte
// chunks; the element (mu,nu) of the slab is read from where the symmetric packing keeps it:
//     nu >= mu: pair = base(mu) + nu  -- the 4 lanes of a lane group read 2 x 64 contiguous bytes of packed row mu;
//     nu <  mu: pair = base(nu) + mu  -- the 8 lane groups read 64 contiguous bytes of packed row nu (8 consecutive mu);
// so both triangles arrive in whole 32-byte sectors, each element of the M-vector being read twice (once per triangle; the second
// read is an L2 hit), and nothing is written back: per slab 8 M bytes of HBM reads instead of 8 M + 16 N^2 (dense slab written, then
// read by the GEMM).  The source is row `slab` of a rectangular [slab][pair] tensor: the inter-species AO storage, the row-sharded
// intra tensors of a communicator, or -- for the packed intra tensor of one GPU -- the batch of M-vectors that
// complete_rows_kernel (it_kernels.cuh) has just assembled from both halves of the packing (C.f90:264-272).  (Resolving the packing
// per element here instead was measured at 6.5 TF/s at N_bf = 500: one page and one sector per 8-byte element.)
// The loads of k-tile t + 1 are in flight while k-tile t is stored (two register sets).
// =================================================================================================================
struct Q1LoadArgs {
  const double *data;  // [slab][ld]
  int64_t M, ld;       // pairs per slab vector; row stride
  int64_t slab0;
  int bc, nc, nfb;
  double *T1t;
  int64_t ldt;
};

template <int TN, int STAGES>
__global__ void __launch_bounds__(512, 1) q1_load_ws5_kernel(const __grid_constant__ CUtensorMap mapB, Q1LoadArgs q) {
  constexpr int BK = 16, BM = 256, BN = TN * 8;
  constexpr int NCW = 8, NGW = 8;  // DMMA warps (32 rows each) / loader warps (32 rows each)
  constexpr uint32_t A_BYTES = BM * 128, B_BYTES = BN * 128, STAGE_BYTES = A_BYTES + B_BYTES;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *ring = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t ring_u = smem_u32(ring);
  const uint32_t bars = ring_u + STAGES * STAGE_BYTES;  // full[s] at bars + 8 s, empty[s] at bars + 8 (STAGES + s)

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tig = lane & 3;
  const int grp = frag_row<true>(lane >> 2);
  const int KT = (q.nc + BK - 1) / BK;
  const int row_blocks = (q.nc + BM - 1) / BM;
  const int ntiles = row_blocks * q.bc;
  const int my_tiles = (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const uint32_t total_it = (uint32_t)my_tiles * (uint32_t)KT;

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bars + 8 * s, NGW + 1);         // one arrive per loader warp + the arrive.expect_tx of the TMA issuer
      mbar_init(bars + 8 * (STAGES + s), NCW);  // one arrive per DMMA warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tma_prefetch_desc(&mapB);
  }
  __syncthreads();

  const uint32_t off0 = (uint32_t)((tig ^ grp) << 4);  // swizzled chunk tig of a row with (row & 7) == grp; chunk 4+tig is off0 ^ 64

  if (warp >= NCW) {
    // ============================== loaders (warpgroups 2 and 3) ==============================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 88;");
    if (total_it == 0) return;
    const int gw = warp - NCW;
    const uint32_t n = (uint32_t)q.nc;
    uint8_t *a_rows = ring + (gw * 32 + grp) * 128;  // rows 32 gw + 8 rg + grp, rg = 0..3
    // the 16 values of iteration (tile, kt): v = 4 rg + c, c = 0..3 <-> k = kt*16 + {2 tig, 2 tig + 1, 2 tig + 8, 2 tig + 9}
    auto fetch = [&](int tile, int kt, double (&x)[16]) {
      const int z = tile / row_blocks, rb = tile - z * row_blocks;
      const int64_t slab = q.slab0 + z;
      const uint32_t mu0 = (uint32_t)(rb * BM + gw * 32 + grp);
      const uint32_t nu0 = (uint32_t)(kt * BK + 2 * tig);
      const uint32_t nus[4] = {nu0, nu0 + 1u, nu0 + 8u, nu0 + 9u};
      const double *row = q.data + slab * q.ld;
#pragma unroll
      for (int rg = 0; rg < 4; ++rg) {
        const uint32_t mu = mu0 + 8u * rg;
        const uint32_t base_mu = pair_base(mu, n);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const uint32_t nu = nus[c];
          double v = 0.0;
          if (mu < n && nu < n) {
            const uint32_t pair = (nu >= mu) ? base_mu + nu : pair_base(nu, n) + mu;
            v = __ldg(row + pair);
          }
          x[4 * rg + c] = v;
        }
      }
    };
    double cur[16], nxt[16];
    int tile = blockIdx.x, kt = 0;
    fetch(tile, kt, cur);
    for (uint32_t it = 0; it < total_it; ++it) {
      // coordinates of the next iteration; its loads go out before this one's values are stored
      int ntile = tile, nkt = kt + 1;
      if (nkt == KT) { nkt = 0; ntile += gridDim.x; }
      if (it + 1u < total_it) fetch(ntile, nkt, nxt);
      const uint32_t s = it % STAGES;
      if (it >= (uint32_t)STAGES) mbar_wait(bars + 8 * (STAGES + s), ((it / STAGES) & 1u) ^ 1u);
      if (gw == 0 && lane == 0) {
        mbar_expect_tx(bars + 8 * s, B_BYTES);
        tma_load_2d(ring_u + s * STAGE_BYTES + A_BYTES, &mapB, bars + 8 * s, kt * BK, 0);
      }
      uint8_t *dst = a_rows + s * STAGE_BYTES;
#pragma unroll
      for (int rg = 0; rg < 4; ++rg) {
        *reinterpret_cast<double2 *>(dst + rg * 8 * 128 + off0) = make_double2(cur[4 * rg + 0], cur[4 * rg + 1]);
        *reinterpret_cast<double2 *>(dst + rg * 8 * 128 + (off0 ^ 64u)) = make_double2(cur[4 * rg + 2], cur[4 * rg + 3]);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(bars + 8 * s);
#pragma unroll
      for (int v = 0; v < 16; ++v) cur[v] = nxt[v];
      tile = ntile; kt = nkt;
    }
    return;
  }

  // ============================== DMMA warps (warpgroups 0 and 1) ==============================
  asm volatile("setmaxnreg.inc.sync.aligned.u32 168;");
  const uint8_t *a_base = ring + (warp * 32 + grp) * 128;
  const uint8_t *b_base = ring + A_BYTES + grp * 128;
  uint32_t it = 0;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int z = tile / row_blocks, rb = tile - z * row_blocks;
    double acc[4][TN][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < TN; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    for (int kt = 0; kt < KT; ++kt, ++it) {
      const uint32_t s = it % STAGES;
      mbar_wait(bars + 8 * s, (it / STAGES) & 1u);
      const uint8_t *as = a_base + s * STAGE_BYTES, *bs = b_base + s * STAGE_BYTES;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const uint32_t off = h ? (off0 ^ 64u) : off0;
        double2 a[4], b[TN];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = *reinterpret_cast<const double2 *>(as + i * 8 * 128 + off);
#pragma unroll
        for (int j = 0; j < TN; ++j) b[j] = *reinterpret_cast<const double2 *>(bs + j * 8 * 128 + off);
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < TN; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i].x, b[j].x);
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < TN; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i].y, b[j].y);
      }
      release_stage(bars + 8 * (STAGES + s), lane);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int m = rb * BM + warp * 32 + i * 8 + grp;
      if (m >= q.nc) continue;
#pragma unroll
      for (int j = 0; j < TN; ++j) {
        const int f = j * 8 + tig;  // permuted fragment mapping: accumulator [0] is column tig, [1] column tig + 4
        if (f < q.nfb) q.T1t[((int64_t)f * q.bc + z) * q.ldt + m] = acc[i][j][0];
        if (f + 4 < q.nfb) q.T1t[((int64_t)(f + 4) * q.bc + z) * q.ldt + m] = acc[i][j][1];
      }
    }
  }
}

}  // namespace lowdin
