/* oracle/eri_oracle.c -- TEST INFRASTRUCTURE ONLY (tests/, __graft_entry__.smoke(), bench.py cpu_baseline).
 *
 * CPU checker of row f4: the two-particle AO integrals the reference obtains from libint2
 * (third-party, v2.x, NOT vendored in /root/reference: `#include <libint2.hpp>`, src/ints/Libint2Iface.h) in
 *   LibintInterface::add_shell           Libint2Iface.cpp:83-130   (Cartesian shells, `norma` = 1/sqrt(self overlap))
 *   LibintInterface::compute_2body_disk  Libint2Iface.cpp:219-416  (unique quartets, |raw| > 1e-10 filter, value * norma^4)
 *   LibintInterface::compute_coupling_disk  Libint2Iface.cpp:~930-1110 (two species, p<=q, r<=s)
 *
 * libint2 is absent, so the published definition is restated: (ab|cd) = integral of a(1)b(1) r12^-1 c(2)d(2) over contracted
 * Cartesian Gaussians x^lx y^ly z^lz exp(-alpha r^2), components of a shell in libint2's (CCA) order, primitive coefficients
 * rescaled as libint2::Shell::renorm() does (unit-normalised (L,0,0) primitive), then every function divided by the square root
 * of its self overlap (`norma`).  PARITY UNPINNED against libint2 itself; pinned instead to (i) the textbook H2 / STO-3G values
 * of Szabo & Ostlund (tests/test_eri_oracle.py) and (ii) 50-digit mixed centre-derivatives of the s-type closed form (mpmath), which
 * give every class up to f analytically: agreement 1e-12.
 *
 * The algorithm here is deliberately NOT the device's (McMurchie-Davidson + Boys function, it_eri.cuh): it is the
 * Rys-polynomial form -- the 2-D integrals G_x(n,m; t) of Rys, Dupuis & King (J. Comput. Chem. 4, 154 (1983)) from their
 * three-term recurrences, shifted to (la,lb|lc,ld) by the transfer relation, and the remaining integral over t in [0,1]
 * done with Gauss-Legendre nodes on [0, min(1, sqrt(60/x))] instead of Rys roots: no Boys function, no Hermite expansion.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif
#define ERI_LMAX 3
#define NGL 96

typedef struct { int l, nprim, first_prim; double origin[3]; } orc_shell;

static double gl_x[NGL], gl_w[NGL];
static int gl_ready = 0;

static void gl_init(void) {  /* Gauss-Legendre nodes on [0,1] by Newton iteration on P_n */
  if (gl_ready) return;
  const int n = NGL;
  for (int i = 0; i < n; ++i) {
    double x = cos(M_PI * (i + 0.75) / (n + 0.5)), pp = 0;
    for (int it = 0; it < 100; ++it) {
      double p0 = 1.0, p1 = x;
      for (int k = 2; k <= n; ++k) { double p2 = ((2.0 * k - 1.0) * x * p1 - (k - 1.0) * p0) / k; p0 = p1; p1 = p2; }
      pp = n * (x * p1 - p0) / (x * x - 1.0);
      double dx = p1 / pp;
      x -= dx;
      if (fabs(dx) < 1e-16) break;
    }
    gl_x[i] = 0.5 * (x + 1.0);
    gl_w[i] = 1.0 / ((1.0 - x * x) * pp * pp);  /* = 0.5 * 2 / ((1-x^2) P'^2) on [0,1] */
  }
  gl_ready = 1;
}

static double dfact(int n) { double r = 1.0; for (; n > 1; n -= 2) r *= n; return r; }  /* n!! ; (-1)!! = 0!! = 1 */

/* libint2::Shell::renorm(): coefficient of a primitive of angular momentum l and exponent a */
static double prim_norm(int l, double a) {
  const double two_a = 2.0 * a;
  return sqrt(pow(2.0, l) * pow(two_a, l + 1) * sqrt(two_a) / (5.56832799683170784528481798212 * dfact(2 * l - 1)));
}

/* Cartesian components of a shell in libint2 / CCA order: lx = l..0, ly = l-lx..0 */
static void cart(int l, int k, int *lx, int *ly, int *lz) {
  int c = 0;
  for (int x = l; x >= 0; --x)
    for (int y = l - x; y >= 0; --y, ++c)
      if (c == k) { *lx = x; *ly = y; *lz = l - x - y; return; }
}

int orc_eri_nbf(int nshells, const orc_shell *sh) { int n = 0; for (int s = 0; s < nshells; ++s) n += (sh[s].l + 1) * (sh[s].l + 2) / 2; return n; }

/* 1-D two-centre overlap factor of x^i exp(-a (x-A)^2) x^j exp(-b (x-B)^2) without the exp(-mu AB^2) sqrt(pi/p) prefactor */
static double ov1d(int i, int j, double PA, double PB, double p) {
  double s = 0.0;
  for (int k = 0; k <= i; ++k)
    for (int m = 0; m <= j; ++m) {
      if ((k + m) & 1) continue;
      double binom_i = 1.0, binom_j = 1.0;
      for (int t = 0; t < k; ++t) binom_i = binom_i * (i - t) / (t + 1);
      for (int t = 0; t < m; ++t) binom_j = binom_j * (j - t) / (t + 1);
      s += binom_i * binom_j * pow(PA, i - k) * pow(PB, j - m) * dfact(k + m - 1) / pow(2.0 * p, (k + m) / 2);
    }
  return s;
}

/* norma[i] = 1 / sqrt(<i|i>) with the renormalised primitive coefficients (Libint2Iface.cpp:118-129) */
void orc_eri_norma(int nshells, const orc_shell *sh, const double *ex, const double *co, double *norma) {
  int f = 0;
  for (int s = 0; s < nshells; ++s) {
    const int l = sh[s].l, nc = (l + 1) * (l + 2) / 2;
    for (int k = 0; k < nc; ++k, ++f) {
      int lx, ly, lz;
      cart(l, k, &lx, &ly, &lz);
      double S = 0.0;
      for (int a = 0; a < sh[s].nprim; ++a)
        for (int b = 0; b < sh[s].nprim; ++b) {
          const double ea = ex[sh[s].first_prim + a], eb = ex[sh[s].first_prim + b], p = ea + eb;
          const double ca = co[sh[s].first_prim + a] * prim_norm(l, ea), cb = co[sh[s].first_prim + b] * prim_norm(l, eb);
          S += ca * cb * pow(M_PI / p, 1.5) * ov1d(lx, lx, 0, 0, p) * ov1d(ly, ly, 0, 0, p) * ov1d(lz, lz, 0, 0, p);
        }
      norma[f] = 1.0 / sqrt(S);
    }
  }
}

/* G(n,m) for one Cartesian direction at u = t^2 (Rys-Dupuis-King recurrences), n <= nmax, m <= mmax */
static void g2d(int nmax, int mmax, double C00, double C00p, double B00, double B10, double B01, double G[2 * ERI_LMAX + 2][2 * ERI_LMAX + 2]) {
  G[0][0] = 1.0;
  for (int n = 0; n < nmax; ++n) G[n + 1][0] = C00 * G[n][0] + (n ? n * B10 * G[n - 1][0] : 0.0);
  for (int m = 0; m < mmax; ++m)
    for (int n = 0; n <= nmax; ++n)
      G[n][m + 1] = C00p * G[n][m] + (m ? m * B01 * G[n][m - 1] : 0.0) + (n ? n * B00 * G[n - 1][m] : 0.0);
}

/* transfer (n,0|m,0) -> (la,lb|lc,ld): I(a,b+1) = I(a+1,b) + (A-B) I(a,b), both sides */
static double transfer(double G[2 * ERI_LMAX + 2][2 * ERI_LMAX + 2], int la, int lb, int lc, int ld, double AB, double CD) {
  double s = 0.0;
  for (int i = 0; i <= lb; ++i) {
    double bi = 1.0;
    for (int t = 0; t < i; ++t) bi = bi * (lb - t) / (t + 1);
    const double fi = bi * pow(AB, lb - i);
    for (int j = 0; j <= ld; ++j) {
      double bj = 1.0;
      for (int t = 0; t < j; ++t) bj = bj * (ld - t) / (t + 1);
      s += fi * bj * pow(CD, ld - j) * G[la + i][lc + j];
    }
  }
  return s;
}

typedef struct { int shell, lx, ly, lz; } fn_t;

static void functions(int nshells, const orc_shell *sh, fn_t *fn) {
  int f = 0;
  for (int s = 0; s < nshells; ++s) {
    const int nc = (sh[s].l + 1) * (sh[s].l + 2) / 2;
    for (int k = 0; k < nc; ++k, ++f) { fn[f].shell = s; cart(sh[s].l, k, &fn[f].lx, &fn[f].ly, &fn[f].lz); }
  }
}

/* raw (libint2-convention, before norma) integral (i j | k l); i, j from basis A, k, l from basis B */
static double eri_raw(const orc_shell *shA, const double *exA, const double *coA, const fn_t *fa, const fn_t *fb,
                      const orc_shell *shB, const double *exB, const double *coB, const fn_t *fc, const fn_t *fd) {
  const orc_shell *sa = &shA[fa->shell], *sb = &shA[fb->shell], *sc = &shB[fc->shell], *sd = &shB[fd->shell];
  const double *A = sa->origin, *B = sb->origin, *Cc = sc->origin, *D = sd->origin;
  double AB2 = 0, CD2 = 0;
  for (int d = 0; d < 3; ++d) { AB2 += (A[d] - B[d]) * (A[d] - B[d]); CD2 += (Cc[d] - D[d]) * (Cc[d] - D[d]); }
  const int la[3] = {fa->lx, fa->ly, fa->lz}, lb[3] = {fb->lx, fb->ly, fb->lz}, lc[3] = {fc->lx, fc->ly, fc->lz}, ld[3] = {fd->lx, fd->ly, fd->lz};
  double total = 0.0;
  for (int ia = 0; ia < sa->nprim; ++ia)
    for (int ib = 0; ib < sb->nprim; ++ib) {
      const double a = exA[sa->first_prim + ia], b = exA[sb->first_prim + ib], p = a + b;
      const double cab = coA[sa->first_prim + ia] * prim_norm(sa->l, a) * coA[sb->first_prim + ib] * prim_norm(sb->l, b) * exp(-a * b / p * AB2);
      double P[3];
      for (int d = 0; d < 3; ++d) P[d] = (a * A[d] + b * B[d]) / p;
      for (int ic = 0; ic < sc->nprim; ++ic)
        for (int id = 0; id < sd->nprim; ++id) {
          const double c = exB[sc->first_prim + ic], dd = exB[sd->first_prim + id], q = c + dd;
          const double ccd = coB[sc->first_prim + ic] * prim_norm(sc->l, c) * coB[sd->first_prim + id] * prim_norm(sd->l, dd) * exp(-c * dd / q * CD2);
          double Q[3], PQ2 = 0;
          for (int d = 0; d < 3; ++d) { Q[d] = (c * Cc[d] + dd * D[d]) / q; PQ2 += (P[d] - Q[d]) * (P[d] - Q[d]); }
          const double rho = p * q / (p + q), x = rho * PQ2;
          const double tmax = (x > 60.0) ? sqrt(60.0 / x) : 1.0;  /* exp(-x t^2) < 1e-26 beyond */
          double integral = 0.0;
          for (int g = 0; g < NGL; ++g) {
            const double t = tmax * gl_x[g], u = t * t;
            double prod = exp(-x * u);
            for (int d = 0; d < 3; ++d) {
              const double B00 = u / (2.0 * (p + q)), B10 = (1.0 - q * u / (p + q)) / (2.0 * p), B01 = (1.0 - p * u / (p + q)) / (2.0 * q);
              const double C00 = (P[d] - A[d]) - q * u * (P[d] - Q[d]) / (p + q), C00p = (Q[d] - Cc[d]) + p * u * (P[d] - Q[d]) / (p + q);
              double G[2 * ERI_LMAX + 2][2 * ERI_LMAX + 2];
              g2d(la[d] + lb[d], lc[d] + ld[d], C00, C00p, B00, B10, B01, G);
              prod *= transfer(G, la[d], lb[d], lc[d], ld[d], A[d] - B[d], Cc[d] - D[d]);
            }
            integral += tmax * gl_w[g] * prod;
          }
          total += cab * ccd * 2.0 * pow(M_PI, 2.5) / (p * q * sqrt(p + q)) * integral;
        }
    }
  return total;
}

/* Packed intra-species tensor in the transformer's layout: row lo holds hi = lo..M-1 at lo*M - lo(lo+1)/2 + hi, pair ids
 * row-wise upper triangular (pair0 of it_kernels.cuh; C.f90:214-226, 0-based).  Values with |raw| <= 1e-10 are zero
 * (Libint2Iface.cpp:369), the others raw * norma^4 (:375-377). */
void orc_eri_packed_intra(int nshells, const orc_shell *sh, const double *ex, const double *co, double *packed) {
  gl_init();
  const int n = orc_eri_nbf(nshells, sh);
  const int64_t M = (int64_t)n * (n + 1) / 2;
  fn_t *fn = (fn_t *)malloc(sizeof(fn_t) * n);
  double *norma = (double *)malloc(sizeof(double) * n);
  int *pi = (int *)malloc(sizeof(int) * M), *pj = (int *)malloc(sizeof(int) * M);
  functions(nshells, sh, fn);
  orc_eri_norma(nshells, sh, ex, co, norma);
  int64_t k = 0;
  for (int i = 0; i < n; ++i) for (int j = i; j < n; ++j, ++k) { pi[k] = i; pj[k] = j; }
#pragma omp parallel for schedule(dynamic, 4)
  for (int64_t lo = 0; lo < M; ++lo)
    for (int64_t hi = lo; hi < M; ++hi) {
      const double raw = eri_raw(sh, ex, co, &fn[pi[lo]], &fn[pj[lo]], sh, ex, co, &fn[pi[hi]], &fn[pj[hi]]);
      packed[lo * M - lo * (lo + 1) / 2 + hi] = (fabs(raw) > 1.0e-10) ? raw * norma[pi[lo]] * norma[pj[lo]] * norma[pi[hi]] * norma[pj[hi]] : 0.0;
    }
  free(fn); free(norma); free(pi); free(pj);
}

/* Rectangular inter-species tensor [M_b][M_a] (C.f90:882): (pq of A | rs of B); |raw| < 1e-10 dropped (Libint2Iface.cpp:1053) */
void orc_eri_rect_inter(int nshA, const orc_shell *shA, const double *exA, const double *coA, int nshB, const orc_shell *shB,
                        const double *exB, const double *coB, double *rect) {
  gl_init();
  const int na = orc_eri_nbf(nshA, shA), nb = orc_eri_nbf(nshB, shB);
  const int64_t Ma = (int64_t)na * (na + 1) / 2, Mb = (int64_t)nb * (nb + 1) / 2;
  fn_t *fa = (fn_t *)malloc(sizeof(fn_t) * na), *fb = (fn_t *)malloc(sizeof(fn_t) * nb);
  double *noa = (double *)malloc(sizeof(double) * na), *nob = (double *)malloc(sizeof(double) * nb);
  int *ai = (int *)malloc(sizeof(int) * Ma), *aj = (int *)malloc(sizeof(int) * Ma), *bi = (int *)malloc(sizeof(int) * Mb), *bj = (int *)malloc(sizeof(int) * Mb);
  functions(nshA, shA, fa); functions(nshB, shB, fb);
  orc_eri_norma(nshA, shA, exA, coA, noa); orc_eri_norma(nshB, shB, exB, coB, nob);
  int64_t k = 0;
  for (int i = 0; i < na; ++i) for (int j = i; j < na; ++j, ++k) { ai[k] = i; aj[k] = j; }
  k = 0;
  for (int i = 0; i < nb; ++i) for (int j = i; j < nb; ++j, ++k) { bi[k] = i; bj[k] = j; }
#pragma omp parallel for schedule(dynamic, 4)
  for (int64_t rs = 0; rs < Mb; ++rs)
    for (int64_t pq = 0; pq < Ma; ++pq) {
      const double raw = eri_raw(shA, exA, coA, &fa[ai[pq]], &fa[aj[pq]], shB, exB, coB, &fb[bi[rs]], &fb[bj[rs]]);
      rect[rs * Ma + pq] = (fabs(raw) >= 1.0e-10) ? raw * noa[ai[pq]] * noa[aj[pq]] * nob[bi[rs]] * nob[bj[rs]] : 0.0;
    }
  free(fa); free(fb); free(noa); free(nob); free(ai); free(aj); free(bi); free(bj);
}

/* one normalised integral (i j | k l) of a single basis, 0-based function indices: known-answer tests */
double orc_eri_one(int nshells, const orc_shell *sh, const double *ex, const double *co, int i, int j, int k, int l) {
  gl_init();
  const int n = orc_eri_nbf(nshells, sh);
  fn_t *fn = (fn_t *)malloc(sizeof(fn_t) * n);
  double *norma = (double *)malloc(sizeof(double) * n);
  functions(nshells, sh, fn);
  orc_eri_norma(nshells, sh, ex, co, norma);
  const double v = eri_raw(sh, ex, co, &fn[i], &fn[j], sh, ex, co, &fn[k], &fn[l]) * norma[i] * norma[j] * norma[k] * norma[l];
  free(fn); free(norma);
  return v;
}
