// TEST INFRASTRUCTURE: the device ERI evaluator of openlowdin_b200/csrc/it_eri.cuh (product source, unchanged; its functions are
// __host__ __device__) run on the HOST, so that the CPU suite can compare the McMurchie-Davidson code with the Rys-form oracle
// before any GPU time is spent.  Not part of the product: the product library has no host evaluation path.
#include "../../openlowdin_b200/csrc/it_kernels.cuh"
#include "../../openlowdin_b200/csrc/it_eri.cuh"
#include "../../include/lowdin_it.h"

using namespace lowdin;

// packed intra tensor (mode 0 of eri_fill_kernel) of one basis, evaluated with the product's eri_raw on the host
extern "C" int mock_eri_packed_intra(int nshells, const lowdin_it_shell *shells, const double *ex, const double *co, double *packed, double *norma) {
  EriHostBasis hb;
  if (!eri_prepare_basis(nshells, shells, ex, co, hb)) return 1;
  EriBasis B{hb.sh.data(), hb.fn.data(), hb.expo.data(), hb.coef.data(), (int)hb.fn.size()};
  const int n = B.nbf;
  const int64_t M = (int64_t)n * (n + 1) / 2;
  for (int f = 0; f < n; ++f) norma[f] = hb.fn[f].norma;
#pragma omp parallel for schedule(dynamic, 4)
  for (int64_t lo = 0; lo < M; ++lo)
    for (int64_t hi = lo; hi < M; ++hi) {
      int i, j, k, l;
      eri_pair_decode(hi, n, i, j);
      eri_pair_decode(lo, n, k, l);
      const double raw = eri_raw(B, hb.fn[i], hb.fn[j], B, hb.fn[k], hb.fn[l]);
      packed[lo * M - lo * (lo + 1) / 2 + hi] = (std::fabs(raw) > 1.0e-10) ? raw * hb.fn[i].norma * hb.fn[j].norma * hb.fn[k].norma * hb.fn[l].norma : 0.0;
    }
  return 0;
}

// The kernel's own element routine (eri_fill_element) driven over the kernel's own grid (eri_fill_grid) on the host: checks the
// block / thread -> element mapping of the three storage modes.  mode 0: packed intra; 1: rectangular [Mb][Ma] (A = first basis);
// 2: the rows rank `rank` of G owns (block-cyclic, 2^logB slabs per block), intra.  dst must hold the mode's element count.
extern "C" int mock_eri_fill(int mode, int nshA, const lowdin_it_shell *shA, const double *exA, const double *coA, int nshB,
                             const lowdin_it_shell *shB, const double *exB, const double *coB, int logB, int G, int rank, double *dst,
                             int64_t *nrows_out) {
  EriHostBasis ha, hb;
  if (!eri_prepare_basis(nshA, shA, exA, coA, ha) || !eri_prepare_basis(nshB, shB, exB, coB, hb)) return 1;
  EriBasis A{ha.sh.data(), ha.fn.data(), ha.expo.data(), ha.coef.data(), (int)ha.fn.size()};
  EriBasis B{hb.sh.data(), hb.fn.data(), hb.expo.data(), hb.coef.data(), (int)hb.fn.size()};
  const int64_t Ma = (int64_t)A.nbf * (A.nbf + 1) / 2, Mb = (int64_t)B.nbf * (B.nbf + 1) / 2;
  int64_t nrows = Mb;
  if (mode == 2) { nrows = 0; for (int64_t s = 0; s < Mb; ++s) nrows += (slab_owner(s, logB, G) == rank); }
  EriFillArgs a{A, B, mode, Ma, nrows, logB, G, rank, mode != 1, dst};
  const dim3 grid = eri_fill_grid(a);
#pragma omp parallel for schedule(dynamic, 1) collapse(2)
  for (int64_t bx = 0; bx < (int64_t)grid.x; ++bx)
    for (int64_t by = 0; by < (int64_t)grid.y; ++by)
      for (int t = 0; t < ERI_FILL_THREADS; ++t) eri_fill_element(a, bx, by, t);
  if (nrows_out) *nrows_out = nrows;
  return 0;
}
