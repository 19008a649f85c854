// it_kernels.cuh -- sm_100a device kernels of the four-index AO->MO transformation.
//
//   expand_block_kernel   packed / rectangular / generated AO slab  ->  dense symmetric N x N (or a block of it)
//   q1_gen_kernel         generated AO slab x coefficient window, the slab produced in registers (no HBM, no smem)
//                         (the "unpack" of TransformIntegralsE.f90:1047-1063, :1623-1630; HBM bound)
//   dgemm_tn_kernel       C[m][n] = sum_k A[m][k] B[n][k]   FP64 tensor cores (DMMA.8x8x4 via
//                         mma.sync.m8n8k4.f64), cp.async multi-stage shared-memory pipeline,
//                         fused epilogues (plain / transposed / window-pair scatter + 1e-10 drop)
//                         = the four quarter transforms (E.f90:1081-1110, :1213-1239)
//   compaction kernels    dense MO block -> (ij,kl,v) / (p,q,r,s,v) lists, |v| > 1e-10, in the
//                         reference's loop order (E.f90:1242-1257, C.f90:418-440)
//   reduce_block_kernel   streaming consumer: count / sum / sum^2 / MP2-like pair energy
//
// Blackwell has no FP64 kind of tcgen05.mma: FP64 tensor throughput is reached through the
// warp-level DMMA only (every mma.sync f64 shape lowers to DMMA.8x8x4 on sm_100a), so this
// is a shared-memory-fed mma.sync kernel, not a TMEM kernel.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace lowdin {

// ---------------------------------------------------------------------------------------------
// AO sources
// ---------------------------------------------------------------------------------------------
enum SrcKind : int {
  SRC_SYM_PACKED = 0,  // intra, C/E layout: row lo holds hi = lo..M-1 at lo*M - lo(lo+1)/2 + hi  (C.f90:223-226, :267-271)
  SRC_RECT = 1,        // data[slab*ld + pair]   (inter AO storage (rs-1)*M_a+pq, C.f90:882; and the half-transformed H)
  SRC_HASH_SYM = 2,    // generated, key = hi*M + lo
  SRC_HASH_RECT = 3,   // generated, key = pair*aux + slab   (inter: PQ*M_b + RS)
  // 4: retired (the divide-per-element blocked layout of round 1; replaced by SRC_RECT_TABLE)
  SRC_LIST = 6,        // the canonical AO list kept as uploaded (it_list.cuh): no dense tensor, list-driven first quarter
  SRC_RECT_TABLE = 7,  // data[tab[pair] + slab*ld], tab = (const int64_t *)data2: the half-transformed chunk as it arrives from the
                       // all-to-all (one [slots][ld] block per sending rank), addressed through a per-chunk column table
  SRC_RANKK = 5        // generated, kind K: (slab | pair) = sum_{k<8} data2[k*aux + slab] * data[k*M + pair]
                       // (rank-8 separable tensor with closed-form MO integrals; intra: data2 == data, aux == M)
};
constexpr int RANKK = 8;  // number of separable terms of kind K (shorter expansions are zero padded)

struct AoSource {
  int kind;
  const double *data;
  int64_t M;    // pairs per slab vector
  int64_t ld;   // SRC_RECT row stride
  int64_t aux;  // SRC_HASH_RECT: number of slabs (M_b)
  uint64_t seed;
  int gen;      // generated sources: 1 = splitmix64 (kind H), 2 = mul-fold-mul (kind F), 3 = rank-K separable (kind K)
  const double *data2;  // SRC_RANKK: pair-vector factors of the SLAB species [RANKK][aux]; SRC_RECT_TABLE: the int64 column table
  // Multi-GPU first half: the slab argument of the kernels is the rank's LOCAL slab number; sources addressed by the global
  // slab id (generated ones) translate it with slab_global().  G == 0 or 1: identity.
  int logB, G, rank;
};

// Block-cyclic distribution of the AO-pair slabs over the ranks of the first half: blocks of 2^logB consecutive slabs, block b on
// rank b % G; a rank numbers its own slabs consecutively (its LOCAL slab number = its row in row-sharded stored tensors).
// Fixed at upload time and independent of the chunking, which is what lets a stored tensor be sharded by rows.
__host__ __device__ __forceinline__ int slab_owner(int64_t slab, int logB, int G) { return (int)((slab >> logB) % G); }
__host__ __device__ __forceinline__ int64_t slab_local(int64_t slab, int logB, int G) {
  return (((slab >> logB) / G) << logB) + (slab & ((1ll << logB) - 1));
}
__host__ __device__ __forceinline__ int64_t slab_global(int64_t local, int logB, int G, int rank) {
  if (G <= 1) return local;
  return ((((local >> logB) * G) + rank) << logB) + (local & ((1ll << logB) - 1));
}
// number of slabs below x that rank r owns (= the local number of the first owned slab >= x)
__host__ __device__ __forceinline__ int64_t slabs_owned_below(int64_t x, int logB, int G, int r) {
  if (G <= 1) return x;
  const int64_t nb = x >> logB, rem = x & ((1ll << logB) - 1);
  return (((nb + G - 1 - r) / G) << logB) + ((nb % G) == r ? rem : 0);
}

__host__ __device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
// kind F ("mul-fold-mul"): two odd 64-bit multiplies around a 32-bit fold -- a bijection of the 64-bit key like
// splitmix64, at a third of the integer instructions (the slab generator is a stand-in for AO production; it must be a
// pure function of the canonical index, not a strong hash).
__host__ __device__ __forceinline__ uint64_t mulfold64(uint64_t x) {
  x *= 0x9E3779B97F4A7C15ull;
  x ^= x >> 32;
  return x * 0xD6E8FEB86659FD93ull;
}
// bits -> value in [-1,1): 2u-1 with u = (bits >> 11) * 2^-53, i.e. (2(bits>>11) - 2^53) * 2^-53, exact in FP64.
// The device form converts the signed integer (XU pipe) and rescales by an exponent subtraction, so that no FP64-pipe
// instruction (the pipe the DMMAs run on) is spent on generation; the value is identical to the host expression.
__host__ __device__ __forceinline__ double bits_to_value(uint64_t bits) {
#ifdef __CUDA_ARCH__
  const long long t = (long long)((bits >> 11) << 1) - (1ll << 53);
  const double d = __ll2double_rn(t);
  const int hi = __double2hiint(d), lo = __double2loint(d);
  return __hiloint2double(t ? hi - (53 << 20) : hi, lo);
#else
  return 2.0 * ((double)(bits >> 11) * (1.0 / 9007199254740992.0)) - 1.0;
#endif
}
__host__ __device__ __forceinline__ double hash_value(int gen, uint64_t seed, uint64_t key) {
  return bits_to_value(gen == 2 ? mulfold64(seed ^ key) : splitmix64(seed ^ key));
}

// 0-based row-wise upper-triangular pair id (== xy(p,q)-1, C.f90:214-221)
__host__ __device__ __forceinline__ int64_t pair0(int64_t i, int64_t j, int64_t n) {
  if (i > j) { int64_t t = i; i = j; j = t; }
  return i * n - (i * (i - 1)) / 2 + (j - i);
}

// ---------------------------------------------------------------------------------------------
// Slab expansion (the unpack of E.f90:1047-1063 / :1623-1630), block form:
//   X[b][row][col] = AO(slab0+b ; pair(r0+row, c0+col) - colbase),  row < nrows, col < ncols, ldx >= ncols (even)
// The full slab is (r0,c0,nrows,ncols) = (0,0,n,n); the chunked second half expands only the rows/columns a
// chunk of AO-pair rows touches.  One CTA per 32x32 tile per slab: tiles of the upper triangle are read along
// the column index (contiguous in the packed row), tiles of the lower triangle along the ROW index (the
// transposed element (col,row) is contiguous in row) and turned in shared memory, so both triangles are read
// with full 32-byte sectors; the tile is written back with 16-byte stores.
// grid = (ceil(ldx/32), ceil(nrows/32), B), block = 256.
// ---------------------------------------------------------------------------------------------
// Source kind as a template parameter (no per-element switch), pair ids in 32-bit arithmetic (n < 65536, so pair ids
// < 2^31 and i*n < 2^32), slab offset hoisted out of the element loop.
__device__ __forceinline__ uint32_t pair0_u32(uint32_t i, uint32_t j, uint32_t n) {  // i <= j
  return i * n - ((i * (i - 1u)) >> 1) + (j - i);
}
template <int KIND>
struct SlabReader {
  const double *base;  // SRC_RECT: row of the slab; SRC_SYM_PACKED / SRC_RECT_BLOCKED: tensor start
  int64_t slab, M;
  uint32_t ld, rows;
  uint64_t seed;
  int gen;
  double coef[KIND == SRC_RANKK ? RANKK : 1];  // SRC_RANKK: the slab's factor of every separable term
  const int64_t *tab;
  __device__ __forceinline__ SlabReader(const AoSource &src, int64_t slab_) : M(src.M), ld((uint32_t)src.ld), rows((uint32_t)src.aux), seed(src.seed), gen(src.gen) {
    // stored sources are addressed by the rank's local slab number, generated ones by the global slab id
    slab = (KIND == SRC_HASH_SYM || KIND == SRC_HASH_RECT || KIND == SRC_RANKK) ? slab_global(slab_, src.logB, src.G, src.rank) : slab_;
    base = (KIND == SRC_RECT || KIND == SRC_RECT_TABLE) ? src.data + slab_ * src.ld : src.data;
    tab = (KIND == SRC_RECT_TABLE) ? reinterpret_cast<const int64_t *>(src.data2) : nullptr;
    if (KIND == SRC_HASH_RECT) M = src.aux;
    if (KIND == SRC_RANKK) {
#pragma unroll
      for (int k = 0; k < RANKK; ++k) coef[k] = __ldg(src.data2 + (int64_t)k * src.aux + slab);
    }
  }
  __device__ __forceinline__ double operator()(int64_t pair) const {
    if (KIND == SRC_RANKK) {
      double v = 0.0;
#pragma unroll
      for (int k = 0; k < RANKK; ++k) v = fma(coef[k], __ldg(base + (int64_t)k * M + pair), v);
      return v;
    }
    if (KIND == SRC_RECT) return __ldg(base + pair);
    if (KIND == SRC_RECT_TABLE) return __ldg(base + __ldg(tab + pair));
    if (KIND == SRC_SYM_PACKED) {
      const int64_t lo = slab < pair ? slab : pair, hi = slab < pair ? pair : slab;
      return __ldg(base + (lo * M - (lo * (lo + 1)) / 2 + hi));
    }
    if (KIND == SRC_HASH_SYM) {
      const int64_t lo = slab < pair ? slab : pair, hi = slab < pair ? pair : slab;
      return hash_value(gen, seed, (uint64_t)(hi * M + lo));
    }
    return hash_value(gen, seed, (uint64_t)(pair * M + slab));  // SRC_HASH_RECT: M holds the number of slabs
  }
};

template <int KIND>
__global__ void __launch_bounds__(256) expand_block_kernel(AoSource src, int64_t slab0, int n, int r0, int nrows, int c0, int ncols,
                                                          int64_t colbase, int ldx, double *__restrict__ X) {
  __shared__ double tile[32][33];
  const int64_t b = blockIdx.z;
  const SlabReader<KIND> rd(src, slab0 + b);
  const int tr0 = blockIdx.y * 32, tc0 = blockIdx.x * 32;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int grow0 = r0 + tr0, gcol0 = c0 + tc0;
  const uint32_t un = (uint32_t)n;
  if (grow0 >= gcol0 + 31) {
    // strictly-lower tile: element (row,col) lives at pair(col,row) = base(col) + row - col  -> lanes along rows
    const int row = grow0 + lane;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int c = warp + 8 * k, col = gcol0 + c;
      double v = 0.0;
      if (tr0 + lane < nrows && tc0 + c < ncols) v = rd((int64_t)pair0_u32((uint32_t)col, (uint32_t)row, un) - colbase);
      tile[lane][c] = v;
    }
  } else {
    const int col = gcol0 + lane;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int r = warp + 8 * k, row = grow0 + r;
      double v = 0.0;
      if (tr0 + r < nrows && tc0 + lane < ncols) {
        const uint32_t lo = (uint32_t)min(row, col), hi = (uint32_t)max(row, col);
        v = rd((int64_t)pair0_u32(lo, hi, un) - colbase);
      }
      tile[r][lane] = v;
    }
  }
  __syncthreads();
  const int cp = (threadIdx.x & 15) * 2, rr = threadIdx.x >> 4;  // 16 column pairs x 16 rows per sweep
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int r = rr + 16 * k;
    if (tr0 + r < nrows && tc0 + cp < ldx) {
      double2 v = make_double2(tile[r][cp], tile[r][cp + 1]);
      *reinterpret_cast<double2 *>(X + ((b * nrows + tr0 + r) * (int64_t)ldx + tc0 + cp)) = v;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// AO list scatter (C.f90:262-273 intra; :879-883 / :947-951 inter).  Everything the reference's reader does per
// entry happens here, on the device: the terminator test (p = -1 ends the stream, C.f90:279-280), the index range
// check and the store to the packed position.  The host only moves bytes.
//   StackView: `nstk` stacks of S entries each; stack t's arrays are p/q/r/s/v + t*stride_{i,v}: one view covers both
//   the five separate arrays of lowdin_it_ao_push_stacks (one stack of n entries) and the raw .ints blocks
//   (int32 p[S] q[S] r[S] s[S]; real64 v[S] per block, Libint2Iface.cpp:3414-3426) of lowdin_it_ao_push_blocks.
//   state[0] = first terminator position seen so far (entries at or after it are ignored), state[1] = 1-based position
//   of an entry with an index outside the basis (0: none).
// ---------------------------------------------------------------------------------------------
struct StackView {
  const int32_t *p, *q, *r, *s;
  const double *v;
  int64_t stride_i, stride_v;  // per stack, in elements of the respective type
  int64_t S, total;            // entries per stack, entries in the view
  int64_t pos0;                // position of the view's first entry in the whole upload (for the error report)
};
struct ScatterDst {
  double *dst;
  int intra, swapped, na, nb;
  // row-sharded storage (multi-GPU): a rank keeps full M-vectors of the slabs it owns; see SlabOwner
  int sharded, logB, G, rank;
};
__device__ __forceinline__ void stack_entry(const StackView &w, int64_t k, int &p, int &q, int &r, int &s, double &v) {
  const int64_t t = k / w.S, e = k - t * w.S;
  p = __ldg(w.p + t * w.stride_i + e); q = __ldg(w.q + t * w.stride_i + e);
  r = __ldg(w.r + t * w.stride_i + e); s = __ldg(w.s + t * w.stride_i + e);
  v = __ldg(w.v + t * w.stride_v + e);
}
// pass 1: position of the first terminator of the view (atomicMin into state[0])
__global__ void __launch_bounds__(256) find_terminator_kernel(StackView w, unsigned long long *__restrict__ state) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= w.total) return;
  const int64_t t = k / w.S, e = k - t * w.S;
  if (__ldg(w.p + t * w.stride_i + e) == -1) atomicMin(state, (unsigned long long)(w.pos0 + k));
}
// pass 2: scatter the entries before the terminator
__global__ void __launch_bounds__(256) scatter_stacks_kernel(StackView w, ScatterDst d, unsigned long long *__restrict__ state) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= w.total) return;
  if ((unsigned long long)(w.pos0 + k) >= state[0]) return;
  int p, q, r, s; double v;
  stack_entry(w, k, p, q, r, s, v);
  const unsigned lim_pq = (unsigned)(d.intra ? d.na : (d.swapped ? d.nb : d.na));
  const unsigned lim_rs = (unsigned)(d.intra ? d.na : (d.swapped ? d.na : d.nb));
  if ((unsigned)(p - 1) >= lim_pq || (unsigned)(q - 1) >= lim_pq || (unsigned)(r - 1) >= lim_rs || (unsigned)(s - 1) >= lim_rs) {
    atomicMin(state + 1, (unsigned long long)(w.pos0 + k + 1));
    return;
  }
  const int64_t Ma = (int64_t)d.na * (d.na + 1) / 2;
  if (d.intra) {
    const int64_t pq = pair0(p - 1, q - 1, d.na), rs = pair0(r - 1, s - 1, d.na);
    if (!d.sharded) {
      const int64_t lo = pq < rs ? pq : rs, hi = pq < rs ? rs : pq;
      d.dst[lo * Ma - (lo * (lo + 1)) / 2 + hi] = v;
    } else {  // full rows of the owned slabs: the entry lands in row pq (if owned) and in row rs (if owned)
      if (slab_owner(pq, d.logB, d.G) == d.rank) d.dst[slab_local(pq, d.logB, d.G) * Ma + rs] = v;
      if (slab_owner(rs, d.logB, d.G) == d.rank) d.dst[slab_local(rs, d.logB, d.G) * Ma + pq] = v;
    }
  } else {
    int64_t slab, pair;  // inter AO storage: slab = pair id on species b, pair = pair id on species a (C.f90:882, :947-951)
    if (!d.swapped) { pair = pair0(p - 1, q - 1, d.na); slab = pair0(r - 1, s - 1, d.nb); }
    else { slab = pair0(p - 1, q - 1, d.nb); pair = pair0(r - 1, s - 1, d.na); }
    if (!d.sharded) d.dst[slab * Ma + pair] = v;
    else if (slab_owner(slab, d.logB, d.G) == d.rank) d.dst[slab_local(slab, d.logB, d.G) * Ma + pair] = v;
  }
}

// Column table of one chunk after the all-to-all: chunk column c is the global slab base + c, computed by rank r = owner(slab) as
// its local slab number loc; rank r's block [rows][ld] starts at r * rows * ld and holds its local slabs from loc_lo[r] on:
//   tab[c] = r * rows * ld + (loc - loc_lo[r])      (offset of row 0; row s adds s * ld)
struct RankStarts { int64_t lo[16]; };
__global__ void column_table_kernel(int64_t base, int64_t width, int logB, int G, int64_t rows, int64_t ld, RankStarts st, int64_t *__restrict__ tab) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= width) return;
  const int64_t slab = base + c;
  const int r = slab_owner(slab, logB, G);
  tab[c] = (int64_t)r * rows * ld + (slab_local(slab, logB, G) - st.lo[r]);
}

// ---------------------------------------------------------------------------------------------
// Step 1 of the two-step unpack of a PACKED intra tensor (E.f90:1047-1063): the full M-vectors of a batch of slabs.
//   R[z][pair] = (slab0 + z | pair) = P[min][max],  z < bc, pair < M,   row stride ldr (even)
// The packing keeps (slab | pair) in row `slab` only for pair >= slab; for pair < slab it lives in row `pair`, column `slab` -- a
// stride-M gather if done slab by slab (one 32-byte sector and one page per 8-byte element: the expansion kernel of round 1 ran at
// 0.38 of the HBM rate because of it).  Here a CTA owns a 32 x 32 tile of (slab, pair): the part with pair >= slab is read along
// pair (row `slab` is contiguous), the part with pair < slab along SLAB (row `pair` holds the 32 slabs of the tile contiguously)
// and turned in shared memory, so every read and every write is a run of whole sectors.  Each packed element is read twice over
// the whole tensor (once as (lo | hi), once as (hi | lo)) and the full square is written once.
// grid = (ceil(M/32), ceil(bc/32)), block = 256.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) complete_rows_kernel(const double *__restrict__ P, int64_t M, int64_t slab0, int bc, int64_t ldr,
                                                            double *__restrict__ R) {
  __shared__ double tile[32][33];  // [slab within tile][pair within tile]
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t p0 = (int64_t)blockIdx.x * 32;            // first pair of the tile
  const int z0 = blockIdx.y * 32;                         // first slab of the tile, relative to slab0
  const int64_t s_lo = slab0 + z0, s_hi = s_lo + 31;      // slab range of the tile (s_hi may exceed the batch)
  // part A: pair >= slab, lanes along pair
  if (p0 + 31 >= s_lo) {
    const int64_t pair = p0 + lane;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int zz = w + 8 * k;
      const int64_t slab = s_lo + zz;
      if (z0 + zz < bc && pair < M && pair >= slab) tile[zz][lane] = __ldg(P + (slab * M - (slab * (slab + 1)) / 2 + pair));
    }
  }
  // part B: pair < slab, lanes along slab (row `pair` of the packed tensor, columns s_lo .. s_lo + 31)
  if (p0 < s_hi) {
    const int64_t slab = s_lo + lane;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int pp = w + 8 * k;
      const int64_t pair = p0 + pp;
      if (z0 + lane < bc && pair < M && pair < slab) tile[lane][pp] = __ldg(P + (pair * M - (pair * (pair + 1)) / 2 + slab));
    }
  }
  __syncthreads();
  const int64_t pair = p0 + lane;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int zz = w + 8 * k;
    if (z0 + zz < bc && pair < M) R[(int64_t)(z0 + zz) * ldr + pair] = tile[zz][lane];
  }
}

// Generated AO set -> stored layout on the device (tests and the stored-AO bench leg at sizes whose list cannot come from a host):
// intra: packed row `slab` holds pairs slab..M-1; inter: rectangular [slab][pair].  grid.x strides over slabs.
template <int KIND>
__global__ void __launch_bounds__(256) materialize_kernel(AoSource src, int intra, int64_t Ma, int64_t nslabs, double *__restrict__ dst) {
  for (int64_t slab = blockIdx.x; slab < nslabs; slab += gridDim.x) {
    const SlabReader<KIND> rd(src, slab);
    if (src.G > 1) {  // row-sharded: `slab` is the rank's local slab number, the row holds the whole M-vector
      double *row = dst + slab * Ma;
      for (int64_t pair = threadIdx.x; pair < Ma; pair += blockDim.x) row[pair] = rd(pair);
    } else if (intra) {
      double *row = dst + (slab * Ma - (slab * (slab + 1)) / 2);
      for (int64_t pair = slab + threadIdx.x; pair < Ma; pair += blockDim.x) row[pair] = rd(pair);
    } else {
      double *row = dst + slab * Ma;
      for (int64_t pair = threadIdx.x; pair < Ma; pair += blockDim.x) row[pair] = rd(pair);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// DGEMM on the FP64 tensor cores
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}
__device__ __forceinline__ void cp_async16(void *smem, const void *gmem, int src_bytes) {
  unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(sa), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N)); }

struct GemmArgs {
  const double *A;  // [M][K] row-major, lda   (K contiguous)
  const double *B;  // [N][K] row-major, ldb   (K contiguous)
  int M, N, K;
  int64_t lda, ldb;
  int64_t strideA, strideB;  // per blockIdx.z
};

// Epilogues -----------------------------------------------------------------------------------
struct EpiPlain {  // C[z][m][n]
  static constexpr bool kSplitRowTail = true;
  static constexpr bool kRowCoalesced = false;
  double *C; int64_t ldc, strideC;
  __device__ __forceinline__ void operator()(int z, int m, int n, double v) const { C[z * strideC + (int64_t)m * ldc + n] = v; }
};
// First quarter: m = (z, mu) over the bc stacked slabs of a batch, n = f (first-contracted window)
//   T1t[f][z][mu]  (mu contiguous: the second quarter reads K-contiguous rows; z inside f so that the
//   second quarter's output columns for one window pair are consecutive AO-pair slabs)
struct EpiQ1 {
  static constexpr bool kSplitRowTail = false;
  static constexpr bool kRowCoalesced = false;
  double *T1t; int nc; int bc; int64_t ldt;
  __device__ __forceinline__ void operator()(int, int m, int n, double v) const {
    int zz = m / nc, mu = m - zz * nc;
    T1t[((int64_t)n * bc + zz) * ldt + mu] = v;
  }
};
// Second quarter of a first half: m = is (second-contracted window), n = (f, z).
// T2 -> H[slot(is,f)][col0+z], keeping only window pairs (slot >= 0) and, for transformer E
// semantics, zeroing |t| <= tol (E.f90:1113).  tol < 0 keeps everything.  Consecutive n are consecutive
// columns of one H row: 64-byte runs per 8x8 accumulator tile.
struct EpiScatterH {
  static constexpr bool kSplitRowTail = true;
  static constexpr bool kRowCoalesced = false;
  double *H; int64_t ldh; int64_t col0; const int32_t *slot; int nf; int bc; double tol;
  __device__ __forceinline__ void operator()(int, int m, int n, double v) const {
    int f = n / bc, zz = n - f * bc;
    int sl = __ldg(slot + (int64_t)m * nf + f);
    if (sl >= 0) H[(int64_t)sl * ldh + col0 + zz] = (fabs(v) > tol) ? v : 0.0;
  }
};
// Third quarter, chunked: m = (z, row) over `nrows` rows of each slot's expanded block, n = kf.
//   T3[slot0+z][kf][roff+row] += v   (accumulated over the AO-pair chunks; mu contiguous for the fourth quarter)
struct EpiAccT {
  static constexpr bool kSplitRowTail = false;
  static constexpr bool kRowCoalesced = true;  // consecutive m (mu within one slot) are consecutive in T3
  double *T3; int nrows; int roff; int nf2; int64_t ldt;
  __device__ __forceinline__ void operator()(int, int m, int n, double v) const {
    int zz = m / nrows, r = m - zz * nrows;
    double *p = T3 + (((int64_t)zz * nf2 + n) * ldt + roff + r);
    *p += v;
  }
  // element (m, n) = row_ptr(m)[n * col_stride()]
  __device__ __forceinline__ double *row_ptr(int m) const {
    int zz = m / nrows, r = m - zz * nrows;
    return T3 + ((int64_t)zz * nf2 * ldt + roff + r);
  }
  __device__ __forceinline__ int64_t col_stride() const { return ldt; }
};
// The same accumulation as ONE fire-and-forget reduction per element (red.global.add.f64, performed by the L2): no load comes
// back to the SM, so the epilogue costs its issue time instead of an HBM round trip, and needs no shared-memory staging -- the 8
// rows of a lane group are 64 consecutive bytes of T3, i.e. whole 32-byte sectors.  Every T3 element receives exactly one add
// per launch (one tile owns it) and launches are stream-ordered, so the result does not depend on the execution order.
struct EpiAccRed {
  static constexpr bool kSplitRowTail = false;
  static constexpr bool kRowCoalesced = false;
  double *T3; int nrows; int roff; int nf2; int64_t ldt;
  __device__ __forceinline__ void operator()(int, int m, int n, double v) const {
    int zz = m / nrows, r = m - zz * nrows;
    atomicAdd(T3 + (((int64_t)zz * nf2 + n) * ldt + roff + r), v);
  }
};
// Fourth quarter: m = ks, n = (z, kf)  ->  OUT[z][ks][kf]
struct EpiOut {
  static constexpr bool kSplitRowTail = true;
  static constexpr bool kRowCoalesced = false;
  double *OUT; int ns2, nf2;
  __device__ __forceinline__ void operator()(int, int m, int n, double v) const {
    int zz = n / nf2, kf = n - zz * nf2;
    OUT[((int64_t)zz * ns2 + m) * nf2 + kf] = v;
  }
};

// Adapter for a GEMM launched with its operands exchanged (C^T = B A^T): element (m, n) of that launch is element
// (n + row_off, m) of the original problem.  Used for the row tail of a tall-skinny-by-wide product (see launch_gemm).
template <class Epi>
struct EpiSwapped {
  static constexpr bool kSplitRowTail = false;
  static constexpr bool kRowCoalesced = false;
  Epi epi; int row_off;
  __device__ __forceinline__ void operator()(int z, int m, int n, double v) const { epi(z, n + row_off, m, v); }
};

template <int BM, int BN, int WM, int WN, int STAGES, class Epi>
__global__ void __launch_bounds__(WM *WN * 32) dgemm_tn_kernel(GemmArgs g, Epi epi) {
  constexpr int BK = 16;        // doubles per k-tile
  constexpr int LDS = BK + 4;   // padded smem row: 160 B -> conflict-free 64-bit fragment loads
  constexpr int NT = WM * WN * 32;
  constexpr int TM = BM / WM / 8, TN = BN / WN / 8;
  static_assert(BM % (WM * 8) == 0 && BN % (WN * 8) == 0, "tile/warp mismatch");
  extern __shared__ __align__(16) double smem[];
  double *As = smem;                        // [STAGES][BM][LDS]
  double *Bs = smem + STAGES * BM * LDS;    // [STAGES][BN][LDS]

  const int z = blockIdx.z;
  const double *__restrict__ A = g.A + z * g.strideA;
  const double *__restrict__ B = g.B + z * g.strideB;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp / WN, wn = warp % WN;
  const int grp = lane >> 2, tig = lane & 3;
  const int KT = (g.K + BK - 1) / BK;

  auto load_tile = [&](int stage, int kt) {
    const int k0 = kt * BK;
    double *as = As + stage * BM * LDS, *bs = Bs + stage * BN * LDS;
#pragma unroll
    for (int c = tid; c < BM * (BK / 2); c += NT) {
      int row = c / (BK / 2), kc = (c % (BK / 2)) * 2;
      int gm = m0 + row, gk = k0 + kc;
      int valid = (gm < g.M) ? min(max(g.K - gk, 0), 2) : 0;
      const double *src = A + (int64_t)(gm < g.M ? gm : 0) * g.lda + (valid ? gk : 0);
      cp_async16(as + row * LDS + kc, src, valid * 8);
    }
#pragma unroll
    for (int c = tid; c < BN * (BK / 2); c += NT) {
      int row = c / (BK / 2), kc = (c % (BK / 2)) * 2;
      int gn = n0 + row, gk = k0 + kc;
      int valid = (gn < g.N) ? min(max(g.K - gk, 0), 2) : 0;
      const double *src = B + (int64_t)(gn < g.N ? gn : 0) * g.ldb + (valid ? gk : 0);
      cp_async16(bs + row * LDS + kc, src, valid * 8);
    }
  };

  double acc[TM][TN][2];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < KT) load_tile(s, s);
    cp_async_commit();
  }
  for (int kt = 0; kt < KT; ++kt) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    {
      int nk = kt + STAGES - 1;
      if (nk < KT) load_tile(nk % STAGES, nk);
      cp_async_commit();
    }
    const double *as = As + (kt % STAGES) * BM * LDS + (wm * TM * 8 + grp) * LDS + tig;
    const double *bs = Bs + (kt % STAGES) * BN * LDS + (wn * TN * 8 + grp) * LDS + tig;
#pragma unroll
    for (int kk = 0; kk < BK; kk += 4) {
      double a[TM], b[TN];
#pragma unroll
      for (int i = 0; i < TM; ++i) a[i] = as[i * 8 * LDS + kk];
#pragma unroll
      for (int j = 0; j < TN; ++j) b[j] = bs[j * 8 * LDS + kk];
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
  }
  cp_async_wait<0>();

#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int m = m0 + wm * TM * 8 + i * 8 + grp;
    if (m >= g.M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int n = n0 + wn * TN * 8 + j * 8 + tig * 2;
      if (n < g.N) epi(z, m, n, acc[i][j][0]);
      if (n + 1 < g.N) epi(z, m, n + 1, acc[i][j][1]);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Fused slab generation + first quarter for GENERATED AO sources (SRC_HASH_*):
//   T1t[f][z][mu] = sum_nu AO(slab0+z ; pair(mu,nu)) * C(nu, f)          (E.f90:1047-1090 in one kernel)
// The dense N x N slab never exists: every lane produces its own DMMA A-fragment element
// (row = mu, k = nu) straight into a register from the counter hash, so neither HBM nor shared memory
// carries the expanded slab.  Only the coefficient window (the B operand, L2 resident) is staged through a
// cp.async shared-memory ring.  CTA = 8 warps x 16 rows = 128 rows of one slab, all BN = 8*TN window columns.
// grid = (ceil(nc/128), bc)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double gen_value(const AoSource &src, uint32_t slab, uint32_t mu, uint32_t nu, uint32_t n) {
  const uint32_t lo = min(mu, nu), hi = max(mu, nu);
  const uint32_t pair = lo * n - ((lo * (lo - 1u)) >> 1) + (hi - lo);  // (lo*(lo-1)) is even; lo=0 wraps to 0*... = 0
  uint64_t key;
  if (src.kind == SRC_HASH_SYM) {
    const uint32_t a = min(slab, pair), b = max(slab, pair);
    key = (uint64_t)b * (uint64_t)src.M + a;
  } else {
    key = (uint64_t)pair * (uint64_t)src.aux + slab;
  }
  return hash_value(src.gen, src.seed, key);
}

// Variant with the coefficient window staged through a cp.async shared-memory ring (block-wide barrier per k-tile).
template <int TN, int STAGES>
__global__ void __launch_bounds__(256) q1_gen_smem_kernel(AoSource src, int64_t slab0, int bc, int nc, const double *__restrict__ Cf,
                                                     int64_t ldc, int nfb, double *__restrict__ T1t, int64_t ldt) {
  constexpr int BN = TN * 8, BK = 16, LDS = BK + 4, NT = 256, TM = 2;
  extern __shared__ __align__(16) double smem[];  // [STAGES][BN][LDS]
  const int z = blockIdx.y;
  const uint32_t slab = (uint32_t)slab_global(slab0 + z, src.logB, src.G, src.rank);
  const int m0 = blockIdx.x * 128;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int grp = lane >> 2, tig = lane & 3;
  const int KT = (nc + BK - 1) / BK;
  const uint32_t n = (uint32_t)nc;

  auto load_b = [&](int stage, int kt) {
    const int k0 = kt * BK;
    double *bs = smem + stage * BN * LDS;
    for (int c = tid; c < BN * (BK / 2); c += NT) {
      int row = c / (BK / 2), kc = (c % (BK / 2)) * 2;
      int gk = k0 + kc;
      int valid = (row < nfb) ? min(max(nc - gk, 0), 2) : 0;
      const double *g = Cf + (int64_t)(row < nfb ? row : 0) * ldc + (valid ? gk : 0);
      cp_async16(bs + row * LDS + kc, g, valid * 8);
    }
  };

  uint32_t mu[TM];
#pragma unroll
  for (int i = 0; i < TM; ++i) mu[i] = (uint32_t)(m0 + warp * 16 + i * 8 + grp);

  double acc[TM][TN][2];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < KT) load_b(s, s);
    cp_async_commit();
  }
  // A fragments of k-tile 0
  double a_nxt[TM][4];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) a_nxt[i][kk] = gen_value(src, slab, min(mu[i], n - 1u), min((uint32_t)(kk * 4 + tig), n - 1u), n);

  for (int kt = 0; kt < KT; ++kt) {
    double a_cur[TM][4];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) a_cur[i][kk] = a_nxt[i][kk];
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    {
      int nk = kt + STAGES - 1;
      if (nk < KT) load_b(nk % STAGES, nk);
      cp_async_commit();
    }
    const double *bs = smem + (kt % STAGES) * BN * LDS + grp * LDS + tig;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      // the next k-tile's fragment for this kk is hashed between the DMMA groups, so that every warp always has
      // tensor work queued behind its integer work (k beyond nc multiplies zero-filled B)
#pragma unroll
      for (int i = 0; i < TM; ++i)
        a_nxt[i][kk] = gen_value(src, slab, min(mu[i], n - 1u), min((uint32_t)((kt + 1) * BK + kk * 4 + tig), n - 1u), n);
      double b[TN];
#pragma unroll
      for (int j = 0; j < TN; ++j) b[j] = bs[j * 8 * LDS + kk * 4];
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) dmma884(acc[i][j][0], acc[i][j][1], a_cur[i][kk], b[j]);
    }
  }
  cp_async_wait<0>();

#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int m = m0 + warp * 16 + i * 8 + grp;
    if (m >= nc) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int f = j * 8 + tig * 2;
      if (f < nfb) T1t[((int64_t)f * bc + z) * ldt + m] = acc[i][j][0];
      if (f + 1 < nfb) T1t[((int64_t)(f + 1) * bc + z) * ldt + m] = acc[i][j][1];
    }
  }
}

// No shared memory and no block-wide barrier: the A fragments come from the hash, the B fragments (coefficient window,
// 8 columns x 32 bytes per load, identical for all warps of all CTAs on the SM) from L1 through the read-only path.
// Each warp therefore runs its own software pipeline (next k-step's hash + loads issued before this k-step's DMMAs)
// and the warps of an SM drift apart, which keeps the FP64 tensor pipe fed while other warps are hashing.
template <int TN>
__global__ void __launch_bounds__(256, 2) q1_gen_kernel(AoSource src, int64_t slab0, int bc, int nc, const double *__restrict__ Cf,
                                                        int64_t ldc, int nfb, double *__restrict__ T1t, int64_t ldt) {
  constexpr int TM = 2;
  const int z = blockIdx.y;
  const uint32_t slab = (uint32_t)slab_global(slab0 + z, src.logB, src.G, src.rank);
  const int m0 = blockIdx.x * 128;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int grp = lane >> 2, tig = lane & 3;
  const uint32_t n = (uint32_t)nc;
  const int KS = (nc + 3) >> 2;

  uint32_t mu[TM];
#pragma unroll
  for (int i = 0; i < TM; ++i) mu[i] = min((uint32_t)(m0 + warp * 16 + i * 8 + grp), n - 1u);
  const double *bp[TN];
  bool bok[TN];
#pragma unroll
  for (int j = 0; j < TN; ++j) {
    const int f = j * 8 + grp;
    bok[j] = f < nfb;
    bp[j] = Cf + (int64_t)(bok[j] ? f : 0) * ldc + tig;
  }

  double acc[TM][TN][2];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  double a_cur[TM], b_cur[TN];
  {
    const bool kok = tig < nc;
#pragma unroll
    for (int i = 0; i < TM; ++i) a_cur[i] = gen_value(src, slab, mu[i], min((uint32_t)tig, n - 1u), n);
#pragma unroll
    for (int j = 0; j < TN; ++j) b_cur[j] = (bok[j] && kok) ? __ldg(bp[j]) : 0.0;
  }
#pragma unroll 2
  for (int ks = 0; ks < KS; ++ks) {
    double a_nxt[TM], b_nxt[TN];
    const int kn = (ks + 1) * 4 + tig;
    const bool kok = kn < nc;  // also false past the last k-step
#pragma unroll
    for (int j = 0; j < TN; ++j) b_nxt[j] = (bok[j] && kok) ? __ldg(bp[j] + (ks + 1) * 4) : 0.0;
#pragma unroll
    for (int i = 0; i < TM; ++i) a_nxt[i] = gen_value(src, slab, mu[i], min((uint32_t)kn, n - 1u), n);
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
      for (int j = 0; j < TN; ++j) dmma884(acc[i][j][0], acc[i][j][1], a_cur[i], b_cur[j]);
#pragma unroll
    for (int i = 0; i < TM; ++i) a_cur[i] = a_nxt[i];
#pragma unroll
    for (int j = 0; j < TN; ++j) b_cur[j] = b_nxt[j];
  }

#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int m = m0 + warp * 16 + i * 8 + grp;
    if (m >= nc) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int f = j * 8 + tig * 2;
      if (f < nfb) T1t[((int64_t)f * bc + z) * ldt + m] = acc[i][j][0];
      if (f + 1 < nfb) T1t[((int64_t)(f + 1) * bc + z) * ldt + m] = acc[i][j][1];
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Result selection / compaction.  Candidates are enumerated in the reference's loop order:
// slot (the (i,j) or (p,q) pair, ijmap order), then the outer index of the second pair, then the inner.
// ---------------------------------------------------------------------------------------------
struct SelectArgs {
  const double *OUT;    // [nslots][ns2][nf2], indexed by slot
  const int32_t *order; // slots in the reference's loop order (convention order -> slot)
  int nslots_batch;
  int ns2, nf2;         // extents of OUT's two trailing dims (second-contracted, first-contracted window)
  int swap2;            // 0: outer loop index of the convention == OUT's ns2 dim; 1: == nf2 dim
  int n_outer, n_inner; // convention loop extents for the second pair (outer, inner)
  int lo_outer, lo_inner;   // first orbital number of the outer / inner loop (1-based)
  int conv;             // LOWDIN_IT_CONV_*
  int symmetric, intra;
  const int32_t *slot_a, *slot_b;   // orbital numbers of the first pair per slot: E: (i,j)  C: (p,q)
  int nA, nB;           // basis sizes for pair ids
  double tol;
};

__device__ __forceinline__ bool select_candidate(const SelectArgs &a, int64_t c, double &v, int &slot, int &o, int &in) {
  const int64_t per = (int64_t)a.n_outer * a.n_inner;
  const int kk = (int)(c / per);
  const int rem = (int)(c - (int64_t)kk * per);
  slot = __ldg(a.order + kk);
  o = rem / a.n_inner + a.lo_outer;
  in = rem % a.n_inner + a.lo_inner;
  const int64_t gs = slot;
  bool keep;
  if (a.conv == 1) keep = (in <= o);  // klmap: l <= k   (E.f90:901-907)
  else {
    keep = !(a.symmetric && in < o);  // s < r skipped   (C.f90:409)
    if (a.intra && a.symmetric && o < a.slot_a[gs]) keep = false;  // r < p skipped (C.f90:394)
  }
  if (!keep) return false;
  const int io = o - a.lo_outer, ii = in - a.lo_inner;
  const int64_t off = a.swap2 ? ((int64_t)ii * a.nf2 + io) : ((int64_t)io * a.nf2 + ii);
  v = a.OUT[(int64_t)slot * a.ns2 * a.nf2 + off];
  return fabs(v) > a.tol;
}

constexpr int SEL_CHUNK = 2048;  // candidates per block
constexpr int SEL_THREADS = 256;

__global__ void __launch_bounds__(SEL_THREADS) select_count_kernel(SelectArgs a, int64_t ncand, unsigned *__restrict__ blockcount) {
  __shared__ unsigned wsum[SEL_THREADS / 32];
  const int64_t base = (int64_t)blockIdx.x * SEL_CHUNK;
  unsigned cnt = 0;
  for (int t = threadIdx.x; t < SEL_CHUNK; t += SEL_THREADS) {
    int64_t c = base + t;
    double v; int s, o, in;
    if (c < ncand && select_candidate(a, c, v, s, o, in)) ++cnt;
  }
  for (int d = 16; d; d >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, d);
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = cnt;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned t = 0;
    for (int w = 0; w < SEL_THREADS / 32; ++w) t += wsum[w];
    blockcount[blockIdx.x] = t;
  }
}

// Single-block exclusive scan of the block counts; *running is the number of entries already emitted
// by earlier batches and is advanced by this batch's total.
__global__ void __launch_bounds__(1024) select_scan_kernel(const unsigned *__restrict__ blockcount, int64_t nblk,
                                                           int64_t *__restrict__ blockoff, unsigned long long *running) {
  __shared__ unsigned long long part[1024];
  __shared__ unsigned long long carry;
  if (threadIdx.x == 0) carry = *running;
  __syncthreads();
  for (int64_t base = 0; base < nblk; base += 1024) {
    int64_t i = base + threadIdx.x;
    unsigned long long v = (i < nblk) ? blockcount[i] : 0ull;
    part[threadIdx.x] = v;
    __syncthreads();
    for (int d = 1; d < 1024; d <<= 1) {
      unsigned long long t = (threadIdx.x >= d) ? part[threadIdx.x - d] : 0ull;
      __syncthreads();
      part[threadIdx.x] += t;
      __syncthreads();
    }
    if (i < nblk) blockoff[i] = (int64_t)(carry + part[threadIdx.x] - v);
    __syncthreads();
    if (threadIdx.x == 1023) carry += part[1023];
    __syncthreads();
  }
  if (threadIdx.x == 0) *running = carry;
}

// Entries kept per window pair: block kk counts the kept candidates of the kk-th pair of `order` (the rank's list is the
// concatenation of these segments, which is how the lists of several ranks are merged into the reference's order).
__global__ void __launch_bounds__(SEL_THREADS) select_segments_kernel(SelectArgs a, unsigned long long *__restrict__ seg) {
  __shared__ unsigned wsum[SEL_THREADS / 32];
  const int64_t per = (int64_t)a.n_outer * a.n_inner, base = (int64_t)blockIdx.x * per;
  unsigned cnt = 0;
  for (int64_t t = threadIdx.x; t < per; t += SEL_THREADS) {
    double v; int s, o, in;
    if (select_candidate(a, base + t, v, s, o, in)) ++cnt;
  }
  for (int d = 16; d; d >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, d);
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = cnt;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long t = 0;
    for (int w = 0; w < SEL_THREADS / 32; ++w) t += wsum[w];
    seg[blockIdx.x] = t;
  }
}

struct EmitArgs {
  int64_t *o_ij, *o_kl;                 // conv E
  int32_t *o_p, *o_q, *o_r, *o_s;       // conv C
  double *o_v;
  int64_t capacity;
  int *overflow;
};

__global__ void __launch_bounds__(SEL_THREADS) select_emit_kernel(SelectArgs a, int64_t ncand, const int64_t *__restrict__ blockoff, EmitArgs e) {
  __shared__ unsigned wbase[SEL_THREADS / 32];
  __shared__ unsigned round_base;
  const int64_t base = (int64_t)blockIdx.x * SEL_CHUNK;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) round_base = 0;
  __syncthreads();
  for (int t0 = 0; t0 < SEL_CHUNK; t0 += SEL_THREADS) {
    const int64_t c = base + t0 + threadIdx.x;
    double v = 0.0; int s = 0, o = 0, in = 0;
    const bool keep = (c < ncand) && select_candidate(a, c, v, s, o, in);
    const unsigned bal = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) wbase[warp] = __popc(bal);
    __syncthreads();
    unsigned pre = 0, tot = 0;
    for (int w = 0; w < SEL_THREADS / 32; ++w) { if (w < warp) pre += wbase[w]; tot += wbase[w]; }
    const unsigned rb = round_base;
    if (keep) {
      const int64_t pos = blockoff[blockIdx.x] + rb + pre + __popc(bal & ((1u << lane) - 1u));
      if (pos < e.capacity) {
        const int64_t gs = s;
        e.o_v[pos] = v;
        if (a.conv == 1) {
          e.o_ij[pos] = pair0(a.slot_a[gs] - 1, a.slot_b[gs] - 1, a.nA) + 1;
          e.o_kl[pos] = pair0(o - 1, in - 1, a.nB) + 1;
        } else {
          e.o_p[pos] = a.slot_a[gs]; e.o_q[pos] = a.slot_b[gs]; e.o_r[pos] = o; e.o_s[pos] = in;
        }
      } else *e.overflow = 1;
    }
    __syncthreads();
    if (threadIdx.x == 0) round_base = rb + tot;
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------
// Streaming consumer: count / sum / sum of squares / pair energy of one OUT batch.
//   energy term (intra, E convention, P==R and Q==S windows): x (lambda x - x_exch) / den with
//   x = (a i|b j) = OUT[slot(a,i)][b][j], x_exch = (b i|a j) = OUT[slot(b,i)][a][j]
//   (closed form of src/MBPT/MPFunctions.f90:450-512); inter: x^2/den (MPFunctions.f90:750-765).
// sums: [0]=count, [1]=sum, [2]=sum sq, [3]=energy   (double atomics, order non-deterministic)
// ---------------------------------------------------------------------------------------------
struct ReduceArgs {
  const double *OUT;                // [nslots][ns2][nf2]: the slots [slot_base, slot_base+nslots) of the pass (whole f-blocks)
  int nslots, ns2, nf2, slot_base;
  const int32_t *slot_s, *slot_f;   // per pass-local slot: index in the second / first window of the first pair
  const int32_t *slot_table;        // [ns1][nfb] -> pass-local slot or -1
  int nfb;
  int orb_s1, orb_f1, orb_s2, orb_f2;  // 1-based orbital number of index 0 along each of the four dims
  const double *epsA, *epsB;        // device orbital energies (null: no energy term)
  int exchange;                     // 1: intra, identical windows -> exchange partner is in OUT
  double lambda, tol;
};

__global__ void __launch_bounds__(256) reduce_block_kernel(ReduceArgs a, double *__restrict__ sums) {
  const int64_t per = (int64_t)a.ns2 * a.nf2;
  const int64_t total = (int64_t)a.nslots * per;
  double cnt = 0, s1 = 0, s2 = 0, en = 0;
  for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < total; c += (int64_t)gridDim.x * blockDim.x) {
    const double x = a.OUT[c];
    if (fabs(x) > a.tol) { cnt += 1.0; s1 += x; s2 += x * x; }
    if (a.epsA) {
      const int slot = (int)(c / per);
      const int rem = (int)(c - (int64_t)slot * per);
      const int k2 = rem / a.nf2, k1 = rem - k2 * a.nf2;
      const int i2 = a.slot_s[slot], i1 = a.slot_f[slot];
      const int o1s = a.orb_s1 + i2, o1f = a.orb_f1 + i1, o2s = a.orb_s2 + k2, o2f = a.orb_f2 + k1;
      const double den = a.epsA[min(o1s, o1f) - 1] + a.epsB[min(o2s, o2f) - 1] - a.epsA[max(o1s, o1f) - 1] - a.epsB[max(o2s, o2f) - 1];
      if (a.exchange) {
        const int xs = a.slot_table[(int64_t)k2 * a.nfb + i1] - a.slot_base;   // slot of (second=k2, first=i1): same f-block
        const double xe = (xs >= 0 && xs < a.nslots) ? a.OUT[(int64_t)xs * per + (int64_t)i2 * a.nf2 + k1] : 0.0;
        en += x * (a.lambda * x - xe) / den;
      } else {
        en += x * x / den;
      }
    }
  }
  __shared__ double sh[4][8];
  for (int d = 16; d; d >>= 1) {
    cnt += __shfl_xor_sync(0xffffffffu, cnt, d); s1 += __shfl_xor_sync(0xffffffffu, s1, d);
    s2 += __shfl_xor_sync(0xffffffffu, s2, d);   en += __shfl_xor_sync(0xffffffffu, en, d);
  }
  if ((threadIdx.x & 31) == 0) { int w = threadIdx.x >> 5; sh[0][w] = cnt; sh[1][w] = s1; sh[2][w] = s2; sh[3][w] = en; }
  __syncthreads();
  if (threadIdx.x < 4) {
    double t = 0;
    for (int w = 0; w < 8; ++w) t += sh[threadIdx.x][w];
    atomicAdd(sums + threadIdx.x, t);
  }
}

// dst[i] = src[i] * factor (factor a power of two: exact).  Builds Species::Cs = C * 2^-53 for the warp-specialised first quarter.
__global__ void scale_copy_kernel(const double *__restrict__ src, double *__restrict__ dst, int64_t n, double factor) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[i] * factor;
}

// D-layout -> internal layout (IntTransfD.cpp:25-65 Multi_Index / Multi_Index_Inter).
// pairD(i,j) = i(i+1)/2 + j, i>=j (0-based).  pi/pj: (i,j) of each xy-numbered pair (i<=j).
__device__ __forceinline__ int64_t pairD(int64_t i, int64_t j) { return i >= j ? i * (i + 1) / 2 + j : j * (j + 1) / 2 + i; }

__global__ void repack_d_intra_kernel(double *__restrict__ packed /*SRC_SYM_PACKED*/, const double *__restrict__ dpacked,
                                      const int32_t *__restrict__ pi, const int32_t *__restrict__ pj, int64_t M) {
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= M * M) return;
  int64_t a = e / M, b = e % M;
  if (b < a) return;
  int64_t da = pairD(pi[a], pj[a]), db = pairD(pi[b], pj[b]);
  int64_t hi = da >= db ? da : db, lo = da >= db ? db : da;
  packed[a * M - (a * (a + 1)) / 2 + b] = dpacked[hi * (hi + 1) / 2 + lo];
}

__global__ void repack_d_inter_kernel(double *__restrict__ rect /*[Mb][Ma] xy numbering*/, const double *__restrict__ d_rect /*[ij*om+kl]*/,
                                      const int32_t *__restrict__ pia, const int32_t *__restrict__ pja,
                                      const int32_t *__restrict__ pib, const int32_t *__restrict__ pjb, int64_t Ma, int64_t Mb) {
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= Ma * Mb) return;
  int64_t b = e / Ma, a = e % Ma;
  rect[e] = d_rect[pairD(pia[a], pja[a]) * Mb + pairD(pib[b], pjb[b])];
}

// v[ij*M + kl] (dense pair matrix in D numbering) -> ERIS[ij(ij+1)/2 + kl], kl <= ij
__global__ void pack_d_lower_kernel(const double *__restrict__ v, double *__restrict__ dpacked, int64_t M) {
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= M * M) return;
  int64_t ij = e / M, kl = e % M;
  if (kl <= ij) dpacked[ij * (ij + 1) / 2 + kl] = v[e];
}

}  // namespace lowdin
