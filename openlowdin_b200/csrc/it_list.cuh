// it_list.cuh -- list-driven first quarter: the canonical AO list (p,q,r,s,v) stays on the device as it
// was read (16 bytes per integral), no N^4/8 dense tensor is built, and every integral is scattered with its <= 4 permutational
// images straight into the QUARTER-TRANSFORMED slabs
//     T1[slab = (lam sig)][nu][f] += (mu nu|lam sig) * C(mu, f)
// This is the reference's DIRECT first quarter (Libint2Iface.cpp:793-853: GG[bf2][bf3][bf4] += v C(p,bf1) and its three
// images; consumed by TransformIntegralsC.f90:545-558 as auxtempA(:,:,nu)), for a whole batch of MO indices f at once and with the
// integrals coming from the stored list instead of being recomputed per MO index.  Real AO lists are sparse (|v| > 1e-10 only,
// Libint2Iface.cpp:369): the work is 4 n_f FMAs per STORED integral.
#pragma once
#include "it_kernels.cuh"

namespace lowdin {

struct __align__(16) ListEntry {
  uint16_t p, q, r, s;  // 0-based AO indices as stored in the list (intra: p>=q, r>=s, (pq)>=(rs), Iterators.cpp:45-77)
  double v;
};

// Upload: entries of a staged piece before the terminator, index-checked, appended to the resident list
// (order is irrelevant: the loader of the reference is order independent too, C.f90:262-273).
__global__ void __launch_bounds__(256) append_list_kernel(StackView w, int intra, int swapped, int na, int nb, ListEntry *__restrict__ out,
                                                          unsigned long long *__restrict__ count, unsigned long long *__restrict__ state) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  bool keep = false;
  int p = 0, q = 0, r = 0, s = 0;
  double v = 0.0;
  if (k < w.total && (unsigned long long)(w.pos0 + k) < state[0]) {
    stack_entry(w, k, p, q, r, s, v);
    const unsigned lim_pq = (unsigned)(intra ? na : (swapped ? nb : na)), lim_rs = (unsigned)(intra ? na : (swapped ? na : nb));
    if ((unsigned)(p - 1) >= lim_pq || (unsigned)(q - 1) >= lim_pq || (unsigned)(r - 1) >= lim_rs || (unsigned)(s - 1) >= lim_rs)
      atomicMin(state + 1, (unsigned long long)(w.pos0 + k + 1));
    else keep = true;
  }
  // warp-aggregated append: one atomic per warp, positions by ballot prefix
  const unsigned bal = __ballot_sync(0xffffffffu, keep);
  const int lane = threadIdx.x & 31;
  unsigned long long base = 0;
  if (lane == 0 && bal) base = atomicAdd(count, (unsigned long long)__popc(bal));
  base = __shfl_sync(0xffffffffu, base, 0);
  if (keep) {
    ListEntry e;
    e.p = (uint16_t)(p - 1); e.q = (uint16_t)(q - 1); e.r = (uint16_t)(r - 1); e.s = (uint16_t)(s - 1); e.v = v;
    out[base + __popc(bal & ((1u << lane) - 1u))] = e;
  }
}

// Cw[mu][f] = C(mu, f_first + f): the window columns of the pass as rows of nfbp doubles (f contiguous, zero padded)
__global__ void list_window_kernel(const double *__restrict__ C, int64_t ldc, int n, int f_first, int nfb, int nfbp, double *__restrict__ Cw) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (int64_t)n * nfbp) return;
  const int mu = (int)(e / nfbp), f = (int)(e - (int64_t)mu * nfbp);
  Cw[e] = (f < nfb) ? C[(int64_t)(f_first + f) * ldc + mu] : 0.0;
}

// One warp per integral.  The lanes decode the entry once (every lane reads the same 16 bytes: one broadcast transaction) and then
// run over the MO indices f of the pass: each image is ONE coalesced run of nfb reductions (red.global.add.f64) into
// T1[(slab * nc + nu) * nfbp + f].   nc = basis size of the contracted species, n_slab = basis size of the slab species.
// intra (nc == n_slab), entry (i j|k l):  slab (kl): nu=j += v C(i,f);  i!=j: nu=i += v C(j,f)
//                                         (ij)!=(kl): slab (ij): nu=l += v C(k,f);  k!=l: nu=k += v C(l,f)      (Libint2Iface.cpp:803-851)
// inter (first pair on the contracted species A, second on the slab species B; `swapped`: the stacks hold (BB|AA)):
//                                         slab (kl): nu=j += v C(i,f);  i!=j: nu=i += v C(j,f)
// On a communicator every rank holds the whole list and keeps only the images that land in ITS slabs (block-cyclic owner,
// it_kernels.cuh); T1 is then indexed by the rank's local slab number.
__global__ void __launch_bounds__(256) list_first_quarter_kernel(const ListEntry *__restrict__ list, const unsigned long long *__restrict__ count,
                                                                 int intra, int swapped, int nc, int n_slab, const double *__restrict__ Cw,
                                                                 int nfb, int nfbp, double *__restrict__ T1, int logB, int G, int rank) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t n = (int64_t)*count;
  for (int64_t e = warp; e < n; e += nwarps) {
    const ListEntry en = list[e];
    int i = en.p, j = en.q, k = en.r, l = en.s;
    if (!intra && swapped) { int t = i; i = k; k = t; t = j; j = l; l = t; }  // (BB|AA): the contracted pair is the second one
    const double v = en.v;
    const int64_t g_kl = pair0(k, l, n_slab), g_ij = intra ? pair0(i, j, nc) : -1;
    const bool mine_kl = (G <= 1) || slab_owner(g_kl, logB, G) == rank;
    const bool both = intra && (g_ij != g_kl) && ((G <= 1) || slab_owner(g_ij, logB, G) == rank);
    const int64_t slab_kl = (G <= 1) ? g_kl : slab_local(g_kl, logB, G);
    double *row_j = T1 + (slab_kl * nc + j) * (int64_t)nfbp, *row_i = T1 + (slab_kl * nc + i) * (int64_t)nfbp;
    const double *ci = Cw + (int64_t)i * nfbp, *cj = Cw + (int64_t)j * nfbp;
    double *row_l = nullptr, *row_k = nullptr;
    const double *ck = nullptr, *cl = nullptr;
    if (both) {
      const int64_t slab_ij = (G <= 1) ? g_ij : slab_local(g_ij, logB, G);
      row_l = T1 + (slab_ij * nc + l) * (int64_t)nfbp; row_k = T1 + (slab_ij * nc + k) * (int64_t)nfbp;
      ck = Cw + (int64_t)k * nfbp; cl = Cw + (int64_t)l * nfbp;
    }
    for (int f = lane; f < nfb; f += 32) {
      if (mine_kl) {
        atomicAdd(row_j + f, v * __ldg(ci + f));
        if (i != j) atomicAdd(row_i + f, v * __ldg(cj + f));
      }
      if (both) {
        atomicAdd(row_l + f, v * __ldg(ck + f));
        if (k != l) atomicAdd(row_k + f, v * __ldg(cl + f));
      }
    }
  }
}

// T1[slab0 + z][nu][f] -> T1t[f][z][nu]  (the operand layout of the second quarter), 32 x 32 tiles through shared memory.
// grid = (ceil(nc/32), ceil(nfb/32), bc)
__global__ void __launch_bounds__(256) list_transpose_kernel(const double *__restrict__ T1, int64_t slab0, int bc, int nc, int nfb, int nfbp,
                                                             double *__restrict__ T1t, int64_t ldt) {
  __shared__ double tile[32][33];
  const int z = blockIdx.z, nu0 = blockIdx.x * 32, f0 = blockIdx.y * 32;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const double *src = T1 + (slab0 + z) * (int64_t)nc * nfbp;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int nu = nu0 + w + 8 * k, f = f0 + lane;
    tile[w + 8 * k][lane] = (nu < nc && f < nfb) ? src[(int64_t)nu * nfbp + f] : 0.0;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int f = f0 + w + 8 * k, nu = nu0 + lane;
    if (f < nfb && nu < nc) T1t[((int64_t)f * bc + z) * ldt + nu] = tile[lane][w + 8 * k];
  }
}

}  // namespace lowdin
