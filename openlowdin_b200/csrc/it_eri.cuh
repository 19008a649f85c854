// it_eri.cuh -- row f4 of SURVEY.md section 8: the AO two-particle integrals evaluated ON THE DEVICE, written straight into the
// transformer's AO storage (packed M(M+1)/2 tensor, rectangular inter-species tensor, or the rows a rank owns on a communicator),
// so that the `.ints` stream between the reference's integrals program and its transformation program never exists.
//
// Replaces LibintInterface::compute_2body_disk (Libint2Iface.cpp:219-416) and ::compute_coupling_disk (:930-1110) with the basis
// handed over as the reference hands it to libint2 (::add_shell, :83-130): Cartesian shells, contraction coefficients rescaled
// like libint2::Shell::renorm() (unit-normalised (L,0,0) primitives), every function divided by the square root of its self
// overlap (`norma`), integrals whose RAW value is <= 1e-10 dropped (:369, :1053).
//
// libint2 itself (Obara-Saika / Head-Gordon-Pople code generated per angular-momentum class) is not restated.  One thread owns one
// element of the tensor -- a pair of function pairs -- and evaluates it by the McMurchie-Davidson scheme: Hermite expansion
// coefficients E_t^{ij} of each primitive pair per Cartesian direction, the Hermite Coulomb integrals R_{tuv} from the Boys function
// by an in-place recursion over one simplex array, and the 6-index contraction.  Stores are coalesced along the packed row.  It is
// an O(K^4 L^6)-per-integral evaluator (no shell-quartet reuse): a correct producer that keeps the whole pipeline on the device,
// not a tuned integral code; s, p, d and f shells.
#pragma once
#include <cmath>
#include <cstdint>
#include <string>
#include <vector>
#include <cuda_runtime.h>

namespace lowdin {

constexpr int ERI_LMAX = 3;                 // highest shell angular momentum (f)
constexpr int ERI_LT = 4 * ERI_LMAX;        // highest Hermite order of a quartet
constexpr int ERI_NR = (ERI_LT + 1) * (ERI_LT + 2) * (ERI_LT + 3) / 6;  // entries of the simplex t + u + v <= ERI_LT

struct EriFunction { int shell, lx, ly, lz; double norma; };
struct EriShell { int l, nprim, first_prim, pad; double x, y, z; };

struct EriBasis {
  const EriShell *shells;      // device
  const EriFunction *fn;       // device, one per basis function
  const double *expo;          // device, primitive exponents
  const double *coef;          // device, contraction coefficient x libint2 primitive normalisation
  int nbf;
};

// Host side of lowdin_it_set_basis: the shells as LibintInterface::add_shell receives them -> device tables.
struct EriHostBasis {
  std::vector<EriShell> sh;
  std::vector<EriFunction> fn;
  std::vector<double> expo, coef;
  std::string err;
};
template <class ShellIn>  // ShellIn: {int l, nprim, first_prim; double origin[3];} (lowdin_it_shell)
inline bool eri_prepare_basis(int nshells, const ShellIn *shells, const double *exponents, const double *coefficients, EriHostBasis &out) {
  out.sh.assign(nshells, EriShell{});
  out.fn.clear();
  int nprim = 0;
  for (int s = 0; s < nshells; ++s) {
    const ShellIn &in = shells[s];
    if (in.l < 0 || in.l > ERI_LMAX) { out.err = "shells up to f (l <= 3) are evaluated on the device"; return false; }
    if (in.nprim < 1 || in.first_prim < 0) { out.err = "bad primitive range"; return false; }
    out.sh[s] = EriShell{in.l, in.nprim, in.first_prim, 0, in.origin[0], in.origin[1], in.origin[2]};
    nprim = nprim > in.first_prim + in.nprim ? nprim : in.first_prim + in.nprim;
  }
  auto dfact = [](int n) { double r = 1.0; for (; n > 1; n -= 2) r *= n; return r; };
  // libint2::Shell::renorm(): the coefficient of a primitive refers to the unit-normalised (l,0,0) Gaussian
  out.expo.assign(exponents, exponents + nprim);
  out.coef.assign(nprim, 0.0);
  for (int s = 0; s < nshells; ++s)
    for (int k = 0; k < out.sh[s].nprim; ++k) {
      const int id = out.sh[s].first_prim + k, l = out.sh[s].l;
      const double two_a = 2.0 * out.expo[id];
      if (!(out.expo[id] > 0.0)) { out.err = "exponents must be positive"; return false; }
      out.coef[id] = coefficients[id] * std::sqrt(std::pow(2.0, l) * std::pow(two_a, l + 1) * std::sqrt(two_a) /
                                                  (5.56832799683170784528481798212 * dfact(2 * l - 1)));
    }
  // functions in libint2's (CCA) order, norma = 1 / sqrt(self overlap) (Libint2Iface.cpp:118-129)
  for (int s = 0; s < nshells; ++s) {
    const EriShell &S = out.sh[s];
    for (int lx = S.l; lx >= 0; --lx)
      for (int ly = S.l - lx; ly >= 0; --ly) {
        const int lz = S.l - lx - ly;
        double ov = 0.0;
        for (int a = 0; a < S.nprim; ++a)
          for (int b = 0; b < S.nprim; ++b) {
            const double p = out.expo[S.first_prim + a] + out.expo[S.first_prim + b];
            ov += out.coef[S.first_prim + a] * out.coef[S.first_prim + b] * std::pow(3.14159265358979323846 / p, 1.5) * dfact(2 * lx - 1) *
                  dfact(2 * ly - 1) * dfact(2 * lz - 1) / std::pow(2.0 * p, S.l);
          }
        out.fn.push_back(EriFunction{s, lx, ly, lz, 1.0 / std::sqrt(ov)});
      }
  }
  return true;
}

// F_n(x) = int_0^1 t^2n exp(-x t^2) dt, n = 0..m
__host__ __device__ inline void eri_boys(int m, double x, double *F) {
  const double ex = exp(-x);
  if (x < 35.0) {
    // F_m(x) = exp(-x) sum_k (2x)^k / ((2m+1)(2m+3)...(2m+2k+1)): all terms positive, then the stable downward recursion
    double term = 1.0 / (2 * m + 1), sum = term;
    for (int k = 1; k < 400; ++k) {
      term *= 2.0 * x / (2 * m + 2 * k + 1);
      sum += term;
      if (term < 1e-17 * sum) break;
    }
    F[m] = ex * sum;
    for (int n = m; n > 0; --n) F[n - 1] = (2.0 * x * F[n] + ex) / (2 * n - 1);
  } else {
    F[0] = 0.5 * sqrt(3.14159265358979323846 / x);  // erf(sqrt(x)) = 1 to 1e-16 from x = 35
    for (int n = 0; n < m; ++n) F[n + 1] = ((2 * n + 1) * F[n] - ex) / (2.0 * x);
  }
}

// index of (t,u,v) in the simplex t + u + v <= ERI_LT
__host__ __device__ __forceinline__ int eri_ridx(int t, int u, int v) {
  const int m = ERI_LT - t;  // u + v <= m inside slice t
  // slices 0..t-1 hold sum_{s<t} (ERI_LT-s+1)(ERI_LT-s+2)/2 entries = NR(ERI_LT) - NR(m) with NR(k) = (k+1)(k+2)(k+3)/6
  return ERI_NR - (m + 1) * (m + 2) * (m + 3) / 6 + u * (m + 1) - (u * (u - 1)) / 2 + v;
}

// Hermite expansion coefficients E_t^{la,lb}, t = 0..la+lb, of one Cartesian direction (without exp(-mu AB^2)):
// E_t^{i+1,j} = E_{t-1}^{ij} / 2p + PA E_t^{ij} + (t+1) E_{t+1}^{ij}, and likewise in j with PB
__host__ __device__ inline void eri_hermite(int la, int lb, double PA, double PB, double inv2p, double *E) {
  double e[ERI_LMAX + 1][ERI_LMAX + 1][2 * ERI_LMAX + 2];
  for (int i = 0; i <= la; ++i)
    for (int j = 0; j <= lb; ++j)
      for (int t = 0; t <= 2 * ERI_LMAX + 1; ++t) e[i][j][t] = 0.0;
  e[0][0][0] = 1.0;
  for (int i = 0; i < la; ++i)
    for (int t = 0; t <= i + 1; ++t)
      e[i + 1][0][t] = (t ? inv2p * e[i][0][t - 1] : 0.0) + PA * e[i][0][t] + (t + 1) * e[i][0][t + 1];
  for (int j = 0; j < lb; ++j)
    for (int i = 0; i <= la; ++i)
      for (int t = 0; t <= i + j + 1; ++t)
        e[i][j + 1][t] = (t ? inv2p * e[i][j][t - 1] : 0.0) + PB * e[i][j][t] + (t + 1) * e[i][j][t + 1];
  for (int t = 0; t <= la + lb; ++t) E[t] = e[la][lb][t];
}

// raw (libint2-convention, before norma) integral (fa fb | fc fd); fa, fb of basis A, fc, fd of basis B
__host__ __device__ inline double eri_raw(const EriBasis &A, const EriFunction &fa, const EriFunction &fb, const EriBasis &B, const EriFunction &fc,
                                 const EriFunction &fd) {
  const EriShell sa = A.shells[fa.shell], sb = A.shells[fb.shell], sc = B.shells[fc.shell], sd = B.shells[fd.shell];
  const double ABx = sa.x - sb.x, ABy = sa.y - sb.y, ABz = sa.z - sb.z, CDx = sc.x - sd.x, CDy = sc.y - sd.y, CDz = sc.z - sd.z;
  const double AB2 = ABx * ABx + ABy * ABy + ABz * ABz, CD2 = CDx * CDx + CDy * CDy + CDz * CDz;
  const int tx = fa.lx + fb.lx, ty = fa.ly + fb.ly, tz = fa.lz + fb.lz;   // bra Hermite orders
  const int kx = fc.lx + fd.lx, ky = fc.ly + fd.ly, kz = fc.lz + fd.lz;   // ket Hermite orders
  const int LT = tx + ty + tz + kx + ky + kz;
  double total = 0.0;
  double R[ERI_NR], F[ERI_LT + 1];
  for (int ia = 0; ia < sa.nprim; ++ia)
    for (int ib = 0; ib < sb.nprim; ++ib) {
      const double a = A.expo[sa.first_prim + ia], b = A.expo[sb.first_prim + ib], p = a + b, inv2p = 0.5 / p;
      const double Px = (a * sa.x + b * sb.x) / p, Py = (a * sa.y + b * sb.y) / p, Pz = (a * sa.z + b * sb.z) / p;
      const double cab = A.coef[sa.first_prim + ia] * A.coef[sb.first_prim + ib] * exp(-a * b / p * AB2);
      double Ex[2 * ERI_LMAX + 1], Ey[2 * ERI_LMAX + 1], Ez[2 * ERI_LMAX + 1];
      eri_hermite(fa.lx, fb.lx, Px - sa.x, Px - sb.x, inv2p, Ex);
      eri_hermite(fa.ly, fb.ly, Py - sa.y, Py - sb.y, inv2p, Ey);
      eri_hermite(fa.lz, fb.lz, Pz - sa.z, Pz - sb.z, inv2p, Ez);
      for (int ic = 0; ic < sc.nprim; ++ic)
        for (int id = 0; id < sd.nprim; ++id) {
          const double c = B.expo[sc.first_prim + ic], d = B.expo[sd.first_prim + id], q = c + d, inv2q = 0.5 / q;
          const double Qx = (c * sc.x + d * sd.x) / q, Qy = (c * sc.y + d * sd.y) / q, Qz = (c * sc.z + d * sd.z) / q;
          const double ccd = B.coef[sc.first_prim + ic] * B.coef[sd.first_prim + id] * exp(-c * d / q * CD2);
          double Kx[2 * ERI_LMAX + 1], Ky[2 * ERI_LMAX + 1], Kz[2 * ERI_LMAX + 1];
          eri_hermite(fc.lx, fd.lx, Qx - sc.x, Qx - sd.x, inv2q, Kx);
          eri_hermite(fc.ly, fd.ly, Qy - sc.y, Qy - sd.y, inv2q, Ky);
          eri_hermite(fc.lz, fd.lz, Qz - sc.z, Qz - sd.z, inv2q, Kz);
          const double X = Px - Qx, Y = Py - Qy, Z = Pz - Qz, alpha = p * q / (p + q);
          eri_boys(LT, alpha * (X * X + Y * Y + Z * Z), F);
          // R^n_{tuv}, n = LT..0, in place: the entries of index sum s at level n need sums s-1 and s-2 of level n+1, so walking s
          // downwards overwrites nothing that is still needed.  R^n_{000} = (-2 alpha)^n F_n.
          double m2a = 1.0;
          for (int n = 0; n < LT; ++n) m2a *= -2.0 * alpha;
          for (int n = LT; n >= 0; --n) {
            for (int s = LT - n; s >= 1; --s)
              for (int t = 0; t <= s; ++t)
                for (int u = 0; u <= s - t; ++u) {
                  const int v = s - t - u;
                  double r;
                  if (t > 0) r = X * R[eri_ridx(t - 1, u, v)] + (t > 1 ? (t - 1) * R[eri_ridx(t - 2, u, v)] : 0.0);
                  else if (u > 0) r = Y * R[eri_ridx(0, u - 1, v)] + (u > 1 ? (u - 1) * R[eri_ridx(0, u - 2, v)] : 0.0);
                  else r = Z * R[eri_ridx(0, 0, v - 1)] + (v > 1 ? (v - 1) * R[eri_ridx(0, 0, v - 2)] : 0.0);
                  R[eri_ridx(t, u, v)] = r;
                }
            R[eri_ridx(0, 0, 0)] = m2a * F[n];
            m2a /= -2.0 * alpha;
          }
          double sum = 0.0;
          for (int t = 0; t <= tx; ++t)
            for (int u = 0; u <= ty; ++u)
              for (int v = 0; v <= tz; ++v) {
                const double eb = Ex[t] * Ey[u] * Ez[v];
                if (eb == 0.0) continue;
                double inner = 0.0;
                for (int tt = 0; tt <= kx; ++tt)
                  for (int uu = 0; uu <= ky; ++uu)
                    for (int vv = 0; vv <= kz; ++vv) {
                      const double ek = Kx[tt] * Ky[uu] * Kz[vv];
                      inner += (((tt + uu + vv) & 1) ? -ek : ek) * R[eri_ridx(t + tt, u + uu, v + vv)];
                    }
                sum += eb * inner;
              }
          total += cab * ccd * 34.98683665524972497 / (p * q * sqrt(p + q)) * sum;  // 2 pi^(5/2)
        }
    }
  return total;
}

__host__ __device__ __forceinline__ void eri_pair_decode(int64_t pair, int n, int &i, int &j) {
  // row-wise upper triangular pair id (pair0 of it_kernels.cuh): row i starts at i n - i (i - 1) / 2
  double nn = (double)n + 0.5;
  int64_t r = (int64_t)(nn - sqrt(nn * nn - 2.0 * (double)pair));
  if (r < 0) r = 0;
  while (r > 0 && r * n - r * (r - 1) / 2 > pair) --r;
  while ((r + 1) * n - (r + 1) * r / 2 <= pair) ++r;
  i = (int)r;
  j = (int)(pair - (r * n - r * (r - 1) / 2)) + i;
}

// One thread per stored element.  mode 0: intra packed (row = slab `lo`, pairs hi = lo..Ma-1); mode 1: rectangular rows
// [row][Ma] with the slab of row r = r (inter-species, A = pair species, B = slab species) ; mode 2: rows a rank owns on a
// communicator (slab = slab_global(row)), A == B allowed.  Grid: x = row, y = tile of 128 columns (eri_fill_grid).
constexpr int ERI_FILL_THREADS = 128;
struct EriFillArgs {
  EriBasis A, B;
  int mode;
  int64_t Ma, nrows;
  int logB, G, rank, strict;
  double *dst;
};
inline dim3 eri_fill_grid(const EriFillArgs &a) { return dim3((unsigned)a.nrows, (unsigned)((a.Ma + ERI_FILL_THREADS - 1) / ERI_FILL_THREADS)); }

__host__ __device__ inline void eri_fill_element(const EriFillArgs &a, int64_t block_x, int64_t block_y, int thread) {
  const int64_t row = block_x;
  if (row >= a.nrows) return;
  const int64_t slab = (a.mode == 2) ? slab_global(row, a.logB, a.G, a.rank) : row;
  const int64_t col0 = (a.mode == 0) ? slab : 0;
  const int64_t pair = col0 + block_y * ERI_FILL_THREADS + thread;
  if (pair >= a.Ma) return;
  int i, j, k, l;
  eri_pair_decode(pair, a.A.nbf, i, j);
  eri_pair_decode(slab, a.B.nbf, k, l);
  const EriFunction fa = a.A.fn[i], fb = a.A.fn[j], fc = a.B.fn[k], fd = a.B.fn[l];
  const double raw = eri_raw(a.A, fa, fb, a.B, fc, fd);
  // the reference keeps |raw| > 1e-10 within one species (Libint2Iface.cpp:369) and |raw| >= 1e-10 between two (:1053)
  const bool keep = a.strict ? (fabs(raw) > 1.0e-10) : (fabs(raw) >= 1.0e-10);
  const double v = keep ? raw * fa.norma * fb.norma * fc.norma * fd.norma : 0.0;
  double *out = (a.mode == 0) ? a.dst + (slab * a.Ma - (slab * (slab + 1)) / 2) : a.dst + row * a.Ma;
  out[pair] = v;
}

__global__ void __launch_bounds__(ERI_FILL_THREADS) eri_fill_kernel(EriFillArgs a) { eri_fill_element(a, blockIdx.x, blockIdx.y, threadIdx.x); }

}  // namespace lowdin
