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
