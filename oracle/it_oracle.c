/*
 * oracle/it_oracle.c -- CPU restatement of openLOWDIN's four-index AO->MO
 * integral transformation (src/integralsTransformation).
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is linked, imported or
 * executed by the product (openlowdin_b200/, include/).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may use it, and only as the checker / the CPU baseline.
 *
 * Every routine below follows one loop nest of the reference and cites it
 * (paths relative to /root/reference).  The arithmetic is plain C in the
 * reference's loop order; no BLAS, no reordering "for speed" except an
 * optional OpenMP loop over p in transformer C, which the reference has too
 * (TransformIntegralsC.f90:341-345).
 *
 * Parity pinning: the reference ships no MO-integral golden vectors
 * (SURVEY.md 8c).  This restatement is pinned against the reference's own
 * transformer D (IntTransfD.cpp) compiled in place into oracle/_ref/ (see
 * oracle/Makefile) and against fixtures generated from it
 * (tests/golden/, made by oracle/make_golden.py).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define IDX2(i, j, ld) ((size_t)(i) + (size_t)(j) * (size_t)(ld)) /* column-major, 0-based */

/* ------------------------------------------------------------------ */
/* Index maps                                                          */
/* ------------------------------------------------------------------ */

/* 1-based row-wise upper-triangular pair id.
 * src/core/IndexMap.f90:249-265 (IndexMap_tensorR2ToVectorB); identical to the
 * xy(p,q) table built in TransformIntegralsC.f90:214-221 / E.f90:851-861. */
int64_t orc_pair_id(int64_t i, int64_t j, int64_t n) {
  if (i > j) { int64_t t = i; i = j; j = t; }
  return j - i + ((2 * n * (i - 1) - i * i + 3 * i) / 2);
}

/* ioff(pq): TransformIntegralsC.f90:223-226 recurrence
 *   ioff(1)=0; ioff(pq)=ioff(pq-1)+M-pq+1   (closed form) */
int64_t orc_ioff(int64_t pq, int64_t M) {
  return (pq - 1) * M - (pq * (pq - 1)) / 2;
}

/* 1-based position in the packed intra-species AO array for pair ids pq, rs.
 * TransformIntegralsC.f90:267-271 / E.f90:965-969. */
int64_t orc_packed_index(int64_t pq, int64_t rs, int64_t M) {
  return (pq >= rs) ? orc_ioff(rs, M) + pq : orc_ioff(pq, M) + rs;
}

/* xypair(1:2,m): E.f90:851-861 */
static void build_xypair(int n, int32_t *x1, int32_t *x2) {
  int64_t m = 0;
  for (int p = 1; p <= n; ++p)
    for (int q = p; q <= n; ++q) { x1[m] = p; x2[m] = q; ++m; }
}

/* ------------------------------------------------------------------ */
/* AO list loader (scatter of the .ints stacks)                        */
/* ------------------------------------------------------------------ */

/* Intra: TransformIntegralsC.f90:262-273 (same in E.f90:960-972).
 * Stops at p == -1 like the "last stack" loop (C.f90:279-280). */
int64_t orc_scatter_intra(const int32_t *p, const int32_t *q, const int32_t *r, const int32_t *s,
                          const double *v, int64_t n, int nbf, double *packed) {
  int64_t M = (int64_t)nbf * (nbf + 1) / 2, k;
  for (k = 0; k < n; ++k) {
    if (p[k] == -1) break;
    int64_t pq = orc_pair_id(p[k], q[k], nbf);
    int64_t rs = orc_pair_id(r[k], s[k], nbf);
    packed[orc_packed_index(pq, rs, M) - 1] = v[k];
  }
  return k;
}

/* Inter: index=(rs-1)*M_a+pq, TransformIntegralsC.f90:879-883 / E.f90:1536-1546.
 * swapped!=0 restates the branch C.f90:947-951, taken when the caller passed the
 * species in reversed order: the file is still (p q | r s) = (B B | A A) and is
 * stored at (pq_file-1)*M_a + rs_file, i.e. the same rect[ (B pair) ][ (A pair) ]
 * layout with "a" = the transformer's first species. */
int64_t orc_scatter_inter(const int32_t *p, const int32_t *q, const int32_t *r, const int32_t *s,
                          const double *v, int64_t n, int na, int nb, int swapped, double *rect) {
  int64_t Ma = (int64_t)na * (na + 1) / 2, k;
  for (k = 0; k < n; ++k) {
    if (p[k] == -1) break;
    if (!swapped) {
      int64_t pq = orc_pair_id(p[k], q[k], na);
      int64_t rs = orc_pair_id(r[k], s[k], nb);
      rect[(rs - 1) * Ma + pq - 1] = v[k];
    } else {
      int64_t pq = orc_pair_id(p[k], q[k], nb);
      int64_t rs = orc_pair_id(r[k], s[k], na);
      rect[(pq - 1) * Ma + rs - 1] = v[k];
    }
  }
  return k;
}

/* ------------------------------------------------------------------ */
/* Transformer E ("two-half")                                          */
/* ------------------------------------------------------------------ */

/* ijmap / klmap: E.f90:863-907 (intra), :1429-1471 (inter).
 * pairs (i in [l1,u1], j in [l2, min(i,u2)]) in loop order, value = xy(i,j). */
static int64_t build_pairmap(int l1, int u1, int l2, int u2, int n, int64_t *map) {
  int64_t m = 0;
  for (int i = l1; i <= u1; ++i)
    for (int j = l2; j <= i; ++j) {
      if (j > u2) continue;
      if (map) map[m] = orc_pair_id(i, j, n);
      ++m;
    }
  return m;
}

int64_t orc_pairmap(int l1, int u1, int l2, int u2, int n, int64_t *map) {
  return build_pairmap(l1, u1, l2, u2, n, map);
}

/* One half-transformation of transformer E over a set of slabs.
 *   slab(pq)[rs]  (rs = 1..M of the species being contracted, a plain M-vector
 *   after the packed row has been unpacked)  ->  for each window pair (i,j):
 *   t = sum_mu C(mu,i) * sum_nu X(nu,mu) C(nu,j)
 * E.f90:1045-1110 (first half, intra), :1621-1674 (first half, inter),
 * :1209-1239 (second half, intra), :1772-1802 (second half, inter).
 * w_l..w_u is the window of the FIRST contracted index (the smaller member of
 * the pair, xypair(1,.)). */
static void e_half_slab(int n, const double *C, int ldc, const double *slab /*M*/, const int32_t *x1,
                        const int32_t *x2, int w_l, int w_u, const int64_t *pmap, int64_t npairs,
                        double *tempB /*n*n*/, double *tempBC /*n*n*/, double *out /*npairs*/) {
  int64_t M = (int64_t)n * (n + 1) / 2;
  /* unpack: E.f90:1047-1063 */
  for (int64_t rs = 0; rs < M; ++rs) {
    int r = x1[rs] - 1, s = x2[rs] - 1;
    tempB[IDX2(s, r, n)] = slab[rs];
    tempB[IDX2(r, s, n)] = slab[rs];
  }
  /* tempBC(mu,j) = sum( tempB(:,mu) * C(:,j) )   E.f90:1081-1090 */
  for (int j = w_l; j <= w_u; ++j)
    for (int mu = 0; mu < n; ++mu) {
      double acc = 0.0;
      const double *b = tempB + IDX2(0, mu, n), *c = C + IDX2(0, j - 1, ldc);
      for (int nu = 0; nu < n; ++nu) acc += b[nu] * c[nu];
      tempBC[IDX2(mu, j - 1, n)] = acc;
    }
  /* tempB(i,j) = sum( C(:,i) * tempBC(:,j) )   E.f90:1099-1110 */
  for (int64_t ij = 0; ij < npairs; ++ij) {
    int64_t ij2 = pmap[ij];
    int j = x1[ij2 - 1], i = x2[ij2 - 1];
    double acc = 0.0;
    const double *c = C + IDX2(0, i - 1, ldc), *t = tempBC + IDX2(0, j - 1, n);
    for (int mu = 0; mu < n; ++mu) acc += c[mu] * t[mu];
    out[ij] = acc;
  }
}

/* Intra-species transformer E, in-memory branch: E.f90:821-1277.
 * win = {p_l,p_u,q_l,q_u,r_l,r_u,s_l,s_u}, 1-based inclusive.
 * Half-transformed values with |t| <= 1e-10 are dropped (E.f90:1113) and so
 * are final ones (E.f90:1242).  Output in the reference's loop order:
 * (ij2, kl2, value) with pair ids, E.f90:1244-1247.  Returns the count (the
 * count is returned even when it exceeds cap; only cap entries are written). */
int64_t orc_transform_e_intra(int n, const double *C, int ldc, const double *packed, const int *win,
                              int64_t *oij, int64_t *okl, double *ov, int64_t cap) {
  int64_t M = (int64_t)n * (n + 1) / 2;
  int32_t *x1 = malloc(sizeof(int32_t) * M), *x2 = malloc(sizeof(int32_t) * M);
  build_xypair(n, x1, x2);
  int64_t nij = build_pairmap(win[0], win[1], win[2], win[3], n, NULL);
  int64_t nkl = build_pairmap(win[4], win[5], win[6], win[7], n, NULL);
  int64_t *ijmap = malloc(sizeof(int64_t) * (nij + 1)), *klmap = malloc(sizeof(int64_t) * (nkl + 1));
  build_pairmap(win[0], win[1], win[2], win[3], n, ijmap);
  build_pairmap(win[4], win[5], win[6], win[7], n, klmap);
  double *H = calloc((size_t)(nij > 0 ? nij : 1) * M, sizeof(double)); /* the it2.tmp buckets */
  int64_t *cnt = calloc(nij + 1, sizeof(int64_t));                    /* totalnIJ2 */
  double *slab = malloc(sizeof(double) * M), *tB = malloc(sizeof(double) * n * n),
         *tBC = calloc((size_t)n * n, sizeof(double)), *res = malloc(sizeof(double) * (nij + nkl + 1));
  /* first half: E.f90:1043-1132 */
  for (int64_t pq = 1; pq <= M; ++pq) {
    for (int64_t rs = 1; rs <= M; ++rs) slab[rs - 1] = packed[orc_packed_index(pq, rs, M) - 1];
    e_half_slab(n, C, ldc, slab, x1, x2, win[2], win[3], ijmap, nij, tB, tBC, res);
    for (int64_t ij = 0; ij < nij; ++ij)
      if (fabs(res[ij]) > 1e-10) { H[ij * M + pq - 1] = res[ij]; cnt[ij]++; }
  }
  /* second half: E.f90:1178-1260 */
  int64_t m = 0;
  for (int64_t ij = 0; ij < nij; ++ij) {
    if (cnt[ij] == 0) continue; /* E.f90:1184 */
    e_half_slab(n, C, ldc, H + ij * M, x1, x2, win[6], win[7], klmap, nkl, tB, tBC, res);
    for (int64_t kl = 0; kl < nkl; ++kl)
      if (fabs(res[kl]) > 1e-10) {
        if (m < cap) { oij[m] = ijmap[ij]; okl[m] = klmap[kl]; ov[m] = res[kl]; }
        ++m;
      }
  }
  free(x1); free(x2); free(ijmap); free(klmap); free(H); free(cnt); free(slab); free(tB); free(tBC); free(res);
  return m;
}

/* Inter-species transformer E: E.f90:1285-1837.  rect is the AO array stored as
 * (rs-1)*M_a + pq (E.f90:1545).  First half loops over species-B AO pairs and
 * transforms the species-A indices (E.f90:1619-1695); second half transforms B
 * (E.f90:1750-1822).  win = {p,q of A; r,s of B}. */
int64_t orc_transform_e_inter(int na, int nb, const double *Ca, int ldca, const double *Cb, int ldcb,
                              const double *rect, const int *win, int64_t *oij, int64_t *okl, double *ov,
                              int64_t cap) {
  int64_t Ma = (int64_t)na * (na + 1) / 2, Mb = (int64_t)nb * (nb + 1) / 2;
  int32_t *xa1 = malloc(sizeof(int32_t) * Ma), *xa2 = malloc(sizeof(int32_t) * Ma);
  int32_t *xb1 = malloc(sizeof(int32_t) * Mb), *xb2 = malloc(sizeof(int32_t) * Mb);
  build_xypair(na, xa1, xa2);
  build_xypair(nb, xb1, xb2);
  int64_t nij = build_pairmap(win[0], win[1], win[2], win[3], na, NULL);
  int64_t nkl = build_pairmap(win[4], win[5], win[6], win[7], nb, NULL);
  int64_t *ijmap = malloc(sizeof(int64_t) * (nij + 1)), *klmap = malloc(sizeof(int64_t) * (nkl + 1));
  build_pairmap(win[0], win[1], win[2], win[3], na, ijmap);
  build_pairmap(win[4], win[5], win[6], win[7], nb, klmap);
  int nmax = na > nb ? na : nb;
  double *H = calloc((size_t)(nij > 0 ? nij : 1) * Mb, sizeof(double));
  int64_t *cnt = calloc(nij + 1, sizeof(int64_t));
  double *tB = malloc(sizeof(double) * nmax * nmax), *tBC = calloc((size_t)nmax * nmax, sizeof(double)),
         *res = malloc(sizeof(double) * (nij + nkl + 1));
  for (int64_t pq = 1; pq <= Mb; ++pq) { /* E.f90:1619 */
    e_half_slab(na, Ca, ldca, rect + (pq - 1) * Ma, xa1, xa2, win[2], win[3], ijmap, nij, tB, tBC, res);
    for (int64_t ij = 0; ij < nij; ++ij)
      if (fabs(res[ij]) > 1e-10) { H[ij * Mb + pq - 1] = res[ij]; cnt[ij]++; }
  }
  int64_t m = 0;
  for (int64_t ij = 0; ij < nij; ++ij) { /* E.f90:1750 */
    if (cnt[ij] == 0) continue;
    e_half_slab(nb, Cb, ldcb, H + ij * Mb, xb1, xb2, win[6], win[7], klmap, nkl, tB, tBC, res);
    for (int64_t kl = 0; kl < nkl; ++kl)
      if (fabs(res[kl]) > 1e-10) {
        if (m < cap) { oij[m] = ijmap[ij]; okl[m] = klmap[kl]; ov[m] = res[kl]; }
        ++m;
      }
  }
  free(xa1); free(xa2); free(xb1); free(xb2); free(ijmap); free(klmap); free(H); free(cnt);
  free(tB); free(tBC); free(res);
  return m;
}

/* ------------------------------------------------------------------ */
/* Transformer C ("4N^5")                                              */
/* ------------------------------------------------------------------ */

/* Intra-species transformer C: TransformIntegralsC.f90:341-446.
 * Output (p,q,r,s,value) for |x|>1e-10 in p,q,r,s loop order (the reference's
 * order for one thread).  `symmetric` skips q<p, r<p, s<r (C.f90:380,394,409).
 * OpenMP over p as in the reference (C.f90:341-345); results are gathered per p
 * and concatenated in p order so the output is deterministic. */
int64_t orc_transform_c_intra(int n, const double *C, int ldc, const double *packed, const int *win,
                              int symmetric, int32_t *op, int32_t *oq, int32_t *or_, int32_t *os,
                              double *ov, int64_t cap) {
  int64_t M = (int64_t)n * (n + 1) / 2;
  int P = win[1] - win[0] + 1;
  if (P <= 0) return 0;
  int64_t per_p = (int64_t)(win[3] - win[2] + 1) * (win[5] - win[4] + 1) * (win[7] - win[6] + 1);
  if (per_p < 0) per_p = 0;
  int64_t *cntp = calloc(P, sizeof(int64_t));
  int32_t **bq = calloc(P, sizeof(void *)), **br = calloc(P, sizeof(void *)), **bs = calloc(P, sizeof(void *));
  double **bv = calloc(P, sizeof(void *));
#pragma omp parallel for schedule(dynamic)
  for (int p = win[0]; p <= win[1]; ++p) {
    double *tempA = calloc((size_t)n * n * n, sizeof(double));
    double *tempB = malloc(sizeof(double) * n * n), *tempC = malloc(sizeof(double) * n);
    int32_t *lq = malloc(sizeof(int32_t) * (per_p + 1)), *lr = malloc(sizeof(int32_t) * (per_p + 1)),
            *ls = malloc(sizeof(int32_t) * (per_p + 1));
    double *lv = malloc(sizeof(double) * (per_p + 1));
    int64_t m = 0;
    /* first quarter: C.f90:351-372 */
    for (int mu = 1; mu <= n; ++mu) {
      double cmp = C[IDX2(mu - 1, p - 1, ldc)];
      for (int j = 1; j <= n; ++j) {
        int64_t ij = orc_pair_id(j, mu, n), kl = 0;
        for (int k = 1; k <= n; ++k)
          for (int l = k; l <= n; ++l) {
            ++kl;
            double a = packed[orc_packed_index(ij, kl, M) - 1];
            size_t lk = (size_t)(l - 1) + (size_t)(k - 1) * n + (size_t)(j - 1) * n * n;
            size_t kl_ = (size_t)(k - 1) + (size_t)(l - 1) * n + (size_t)(j - 1) * n * n;
            tempA[lk] += a * cmp;
            tempA[kl_] = tempA[lk];
          }
      }
    }
    for (int q = win[2]; q <= win[3]; ++q) {
      if (q < p && symmetric) continue; /* C.f90:380 */
      /* second quarter: C.f90:382-386 */
      memset(tempB, 0, sizeof(double) * n * n);
      for (int nu = 1; nu <= n; ++nu) {
        double c = C[IDX2(nu - 1, q - 1, ldc)];
        const double *a = tempA + (size_t)(nu - 1) * n * n;
        for (int x = 0; x < n * n; ++x) tempB[x] += c * a[x];
      }
      for (int r = win[4]; r <= win[5]; ++r) {
        if (r < p && symmetric) continue; /* C.f90:394 (n = p) */
        /* third quarter: C.f90:397-402 */
        memset(tempC, 0, sizeof(double) * n);
        for (int lam = 1; lam <= n; ++lam) {
          double c = C[IDX2(lam - 1, r - 1, ldc)];
          const double *b = tempB + (size_t)(lam - 1) * n;
          for (int x = 0; x < n; ++x) tempC[x] += c * b[x];
        }
        for (int s = win[6]; s <= win[7]; ++s) {
          if (s < r && symmetric) continue; /* C.f90:409 */
          double x = 0.0; /* fourth quarter: C.f90:411-416 */
          for (int sg = 1; sg <= n; ++sg) x += C[IDX2(sg - 1, s - 1, ldc)] * tempC[sg - 1];
          if (fabs(x) > 1e-10) { lq[m] = q; lr[m] = r; ls[m] = s; lv[m] = x; ++m; }
        }
      }
    }
    int ip = p - win[0];
    cntp[ip] = m; bq[ip] = lq; br[ip] = lr; bs[ip] = ls; bv[ip] = lv;
    free(tempA); free(tempB); free(tempC);
  }
  int64_t tot = 0;
  for (int ip = 0; ip < P; ++ip) {
    for (int64_t k = 0; k < cntp[ip]; ++k, ++tot)
      if (tot < cap) { op[tot] = win[0] + ip; oq[tot] = bq[ip][k]; or_[tot] = br[ip][k]; os[tot] = bs[ip][k]; ov[tot] = bv[ip][k]; }
    free(bq[ip]); free(br[ip]); free(bs[ip]); free(bv[ip]);
  }
  free(cntp); free(bq); free(br); free(bs); free(bv);
  return tot;
}

/* Inter-species transformer C: TransformIntegralsC.f90:1051-1154 (serial in
 * the reference).  rect stored (kl-1)*M_a + ij (C.f90:1072).  `symmetric`
 * skips q<p and s<r only (C.f90:1090, :1116). */
int64_t orc_transform_c_inter(int na, int nb, const double *Ca, int ldca, const double *Cb, int ldcb,
                              const double *rect, const int *win, int symmetric, int32_t *op, int32_t *oq,
                              int32_t *or_, int32_t *os, double *ov, int64_t cap) {
  int64_t Ma = (int64_t)na * (na + 1) / 2;
  double *tempA = malloc(sizeof(double) * (size_t)nb * nb * na);
  double *tempB = malloc(sizeof(double) * nb * nb), *tempC = malloc(sizeof(double) * nb);
  int64_t m = 0;
  for (int p = win[0]; p <= win[1]; ++p) {
    memset(tempA, 0, sizeof(double) * (size_t)nb * nb * na);
    for (int mu = 1; mu <= na; ++mu) { /* first quarter: C.f90:1057-1083 */
      double cmp = Ca[IDX2(mu - 1, p - 1, ldca)];
      for (int j = 1; j <= na; ++j) {
        int64_t ij = orc_pair_id(j, mu, na), kl = 0;
        for (int k = 1; k <= nb; ++k)
          for (int l = k; l <= nb; ++l) {
            ++kl;
            double a = rect[(kl - 1) * Ma + ij - 1];
            size_t lk = (size_t)(l - 1) + (size_t)(k - 1) * nb + (size_t)(j - 1) * nb * nb;
            size_t kl_ = (size_t)(k - 1) + (size_t)(l - 1) * nb + (size_t)(j - 1) * nb * nb;
            tempA[lk] += a * cmp;
            tempA[kl_] = tempA[lk];
          }
      }
    }
    for (int q = win[2]; q <= win[3]; ++q) {
      if (q < p && symmetric) continue;
      memset(tempB, 0, sizeof(double) * nb * nb);
      for (int nu = 1; nu <= na; ++nu) { /* C.f90:1092-1097 */
        double c = Ca[IDX2(nu - 1, q - 1, ldca)];
        const double *a = tempA + (size_t)(nu - 1) * nb * nb;
        for (int x = 0; x < nb * nb; ++x) tempB[x] += c * a[x];
      }
      for (int r = win[4]; r <= win[5]; ++r) {
        memset(tempC, 0, sizeof(double) * nb);
        for (int lam = 1; lam <= nb; ++lam) { /* C.f90:1106-1111 */
          double c = Cb[IDX2(lam - 1, r - 1, ldcb)];
          const double *b = tempB + (size_t)(lam - 1) * nb;
          for (int x = 0; x < nb; ++x) tempC[x] += c * b[x];
        }
        for (int s = win[6]; s <= win[7]; ++s) {
          if (s < r && symmetric) continue;
          double x = 0.0;
          for (int sg = 1; sg <= nb; ++sg) x += Cb[IDX2(sg - 1, s - 1, ldcb)] * tempC[sg - 1];
          if (fabs(x) > 1e-10) {
            if (m < cap) { op[m] = p; oq[m] = q; or_[m] = r; os[m] = s; ov[m] = x; }
            ++m;
          }
        }
      }
    }
  }
  free(tempA); free(tempB); free(tempC);
  return m;
}

/* ------------------------------------------------------------------ */
/* Transformer D (full, in place, lower-triangular 0-based packing)    */
/* ------------------------------------------------------------------ */

/* IntTransfD.cpp:25-45 with 64-bit arithmetic. */
int64_t orc_d_multi_index(int64_t i, int64_t j, int64_t k, int64_t l) {
  int64_t t;
  if (i < j) { t = i; i = j; j = t; }
  if (k < l) { t = k; k = l; l = t; }
  int64_t ij = i * (i + 1) / 2 + j, kl = k * (k + 1) / 2 + l;
  if (ij < kl) { t = ij; ij = kl; kl = t; }
  return ij * (ij + 1) / 2 + kl;
}

/* Op <- C^T Op C on column-major n x n arrays: IntTransfD.cpp:107-123
 * (dgemm("T","N") then dgemm("N","N")), written as plain loops. */
static void d_similarity(int n, double *Op, const double *C, double *W) {
  for (int j = 0; j < n; ++j)
    for (int i = 0; i < n; ++i) {
      double acc = 0.0;
      for (int k = 0; k < n; ++k) acc += C[IDX2(k, i, n)] * Op[IDX2(k, j, n)];
      W[IDX2(i, j, n)] = acc;
    }
  for (int j = 0; j < n; ++j)
    for (int i = 0; i < n; ++i) {
      double acc = 0.0;
      for (int k = 0; k < n; ++k) acc += W[IDX2(i, k, n)] * C[IDX2(k, j, n)];
      Op[IDX2(i, j, n)] = acc;
    }
}

/* IntTransfD.cpp:125-181 (four_index_trans, array version). */
void orc_transform_d_intra(const double *C, double *eris, int nao) {
  int64_t m = (int64_t)nao * (nao + 1) / 2;
  double *X = malloc(sizeof(double) * nao * nao), *W = malloc(sizeof(double) * nao * nao);
  double *TMP = calloc((size_t)m * m, sizeof(double));
  int64_t ij = 0;
  for (int i = 0; i < nao; ++i)
    for (int j = 0; j <= i; ++j, ++ij) {
      for (int k = 0; k < nao; ++k)
        for (int l = 0; l <= k; ++l) X[k * nao + l] = X[l * nao + k] = eris[orc_d_multi_index(i, j, k, l)];
      d_similarity(nao, X, C, W);
      int64_t kl = 0;
      for (int k = 0; k < nao; ++k)
        for (int l = 0; l <= k; ++l, ++kl) TMP[kl * m + ij] = X[k * nao + l];
    }
  int64_t kl = 0;
  for (int k = 0; k < nao; ++k)
    for (int l = 0; l <= k; ++l, ++kl) {
      ij = 0;
      for (int i = 0; i < nao; ++i)
        for (int j = 0; j <= i; ++j, ++ij) X[i * nao + j] = X[j * nao + i] = TMP[kl * m + ij];
      d_similarity(nao, X, C, W);
      for (int i = 0; i < nao; ++i)
        for (int j = 0; j <= i; ++j) eris[orc_d_multi_index(k, l, i, j)] = X[i * nao + j];
    }
  free(X); free(W); free(TMP);
}

/* IntTransfD.cpp:245-323 (four_index_trans_inter): ERIS[ij*om + kl]. */
void orc_transform_d_inter(const double *C, const double *OC, double *eris, int nao, int onao) {
  int64_t m = (int64_t)nao * (nao + 1) / 2, om = (int64_t)onao * (onao + 1) / 2;
  int nmax = nao > onao ? nao : onao;
  double *X = malloc(sizeof(double) * nmax * nmax), *W = malloc(sizeof(double) * nmax * nmax);
  double *TMP = calloc((size_t)m * om, sizeof(double));
  int64_t ij = 0, kl;
  for (int i = 0; i < nao; ++i)
    for (int j = 0; j <= i; ++j, ++ij) {
      kl = 0;
      for (int k = 0; k < onao; ++k)
        for (int l = 0; l <= k; ++l, ++kl) X[k * onao + l] = X[l * onao + k] = eris[ij * om + kl];
      d_similarity(onao, X, OC, W);
      kl = 0;
      for (int k = 0; k < onao; ++k)
        for (int l = 0; l <= k; ++l, ++kl) TMP[ij * om + kl] = X[k * onao + l];
    }
  kl = 0;
  for (int k = 0; k < onao; ++k)
    for (int l = 0; l <= k; ++l, ++kl) {
      ij = 0;
      for (int i = 0; i < nao; ++i)
        for (int j = 0; j <= i; ++j, ++ij) X[i * nao + j] = X[j * nao + i] = TMP[ij * om + kl];
      d_similarity(nao, X, C, W);
      ij = 0;
      for (int i = 0; i < nao; ++i)
        for (int j = 0; j <= i; ++j, ++ij) eris[ij * om + kl] = X[i * nao + j];
    }
  free(X); free(W); free(TMP);
}

/* ------------------------------------------------------------------ */
/* Synthetic AO generator (benchmark definition)          */
/* ------------------------------------------------------------------ */

static inline uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

/* kind H: value at canonical 0-based pair ids PQ>=RS (intra) = 2u-1,
 * u = (splitmix64(seed ^ (PQ*M+RS)) >> 11) * 2^-53; inter: key PQ*M_b+RS. */
double orc_hash_value(uint64_t seed, uint64_t key) {
  return 2.0 * ((double)(splitmix64(seed ^ key) >> 11) * (1.0 / 9007199254740992.0)) - 1.0;
}

/* kind F: the same value map over mulfold64(x) = ((x*K1) ^ ((x*K1) >> 32)) * K2 (two odd multiplies around a
 * 32-bit fold: a bijection of the key, a third of splitmix64's integer work on the GPU). */
static inline uint64_t mulfold64(uint64_t x) {
  x *= 0x9E3779B97F4A7C15ull;
  x ^= x >> 32;
  return x * 0xD6E8FEB86659FD93ull;
}
double orc_gen_value(int kind, uint64_t seed, uint64_t key) {
  if (kind == 2) return 2.0 * ((double)(mulfold64(seed ^ key) >> 11) * (1.0 / 9007199254740992.0)) - 1.0;
  return orc_hash_value(seed, key);
}

/* Fill the packed intra array (C/E layout, 1-based pair ids pq<=rs stored at
 * ioff(pq)+rs) with kind-H values.  Canonical key uses 0-based ids, PQ>=RS. */
void orc_fill_hash_intra(uint64_t seed, int nbf, double *packed) {
  int64_t M = (int64_t)nbf * (nbf + 1) / 2;
  for (int64_t lo = 1; lo <= M; ++lo)
    for (int64_t hi = lo; hi <= M; ++hi)
      packed[orc_ioff(lo, M) + hi - 1] = orc_hash_value(seed, (uint64_t)((hi - 1) * M + (lo - 1)));
}

/* rect[(rs-1)*Ma + pq] = hash(seed, PQ*Mb + RS), PQ of species A, RS of species B (0-based). */
void orc_fill_gen_intra(int kind, uint64_t seed, int nbf, double *packed) {
  int64_t M = (int64_t)nbf * (nbf + 1) / 2;
  for (int64_t lo = 1; lo <= M; ++lo)
    for (int64_t hi = lo; hi <= M; ++hi)
      packed[orc_ioff(lo, M) + hi - 1] = orc_gen_value(kind, seed, (uint64_t)((hi - 1) * M + (lo - 1)));
}
void orc_fill_gen_inter(int kind, uint64_t seed, int na, int nb, double *rect) {
  int64_t Ma = (int64_t)na * (na + 1) / 2, Mb = (int64_t)nb * (nb + 1) / 2;
  for (int64_t rs = 0; rs < Mb; ++rs)
    for (int64_t pq = 0; pq < Ma; ++pq) rect[rs * Ma + pq] = orc_gen_value(kind, seed, (uint64_t)(pq * Mb + rs));
}
void orc_fill_hash_inter(uint64_t seed, int na, int nb, double *rect) {
  int64_t Ma = (int64_t)na * (na + 1) / 2, Mb = (int64_t)nb * (nb + 1) / 2;
  for (int64_t rs = 0; rs < Mb; ++rs)
    for (int64_t pq = 0; pq < Ma; ++pq) rect[rs * Ma + pq] = orc_hash_value(seed, (uint64_t)(pq * Mb + rs));
}

/* ------------------------------------------------------------------ */
/* Timed windowed half-transform sample for the CPU baseline            */
/* ------------------------------------------------------------------ */

/* First half of transformer E (E.f90:1043-1132) on slabs [pq0, pq0+npq) of a
 * kind-H synthetic intra tensor generated on the fly (so N=500..1500 samples
 * need no N^4/8 array).  Returns a checksum so the work cannot be elided.
 * The reference's compute loops are serial (its OMP directives are commented
 * out: E.f90:1078-1080); nthreads > 1 spreads the independent slabs over
 * threads, which is the most the reference's structure would allow. */
double orc_e_first_half_sample(uint64_t seed, int n, const double *C, int ldc, const int *win, int64_t pq0,
                               int64_t npq, int nthreads) {
  int64_t M = (int64_t)n * (n + 1) / 2;
  int32_t *x1 = malloc(sizeof(int32_t) * M), *x2 = malloc(sizeof(int32_t) * M);
  build_xypair(n, x1, x2);
  int64_t nij = build_pairmap(win[0], win[1], win[2], win[3], n, NULL);
  int64_t *ijmap = malloc(sizeof(int64_t) * (nij + 1));
  build_pairmap(win[0], win[1], win[2], win[3], n, ijmap);
  double chk = 0.0;
  if (nthreads < 1) nthreads = 1;
#pragma omp parallel num_threads(nthreads) reduction(+ : chk)
  {
    double *slab = malloc(sizeof(double) * M), *tB = malloc(sizeof(double) * n * n),
           *tBC = calloc((size_t)n * n, sizeof(double)), *res = malloc(sizeof(double) * (nij + 1));
#pragma omp for schedule(dynamic)
    for (int64_t pq = pq0 + 1; pq <= pq0 + npq; ++pq) {
      if (pq > M) continue;
      for (int64_t rs = 1; rs <= M; ++rs) {
        int64_t hi = pq >= rs ? pq : rs, lo = pq >= rs ? rs : pq;
        slab[rs - 1] = orc_hash_value(seed, (uint64_t)((hi - 1) * M + (lo - 1)));
      }
      e_half_slab(n, C, ldc, slab, x1, x2, win[2], win[3], ijmap, nij, tB, tBC, res);
      for (int64_t ij = 0; ij < nij; ++ij)
        if (fabs(res[ij]) > 1e-10) chk += res[ij];
    }
    free(slab); free(tB); free(tBC); free(res);
  }
  free(x1); free(x2); free(ijmap);
  return chk;
}

/* The same first half, returning the values: out[ij * npq + z] for window pair ij (ijmap order, E.f90:878-884) and slab
 * pq0 + z, with |t| <= 1e-10 -> 0 (the values E would not write to its bucket file, E.f90:1113).  gen_kind 1 = kind H,
 * 2 = kind F.  Checked against the device's first half at N_bf = 1500 (bench.py) and in the GPU tests. */
int64_t orc_e_first_half_values(int gen_kind, uint64_t seed, int n, const double *C, int ldc, const int *win, int64_t pq0,
                                int64_t npq, int nthreads, double *out) {
  int64_t M = (int64_t)n * (n + 1) / 2;
  int32_t *x1 = malloc(sizeof(int32_t) * M), *x2 = malloc(sizeof(int32_t) * M);
  build_xypair(n, x1, x2);
  int64_t nij = build_pairmap(win[0], win[1], win[2], win[3], n, NULL);
  int64_t *ijmap = malloc(sizeof(int64_t) * (nij + 1));
  build_pairmap(win[0], win[1], win[2], win[3], n, ijmap);
  if (nthreads < 1) nthreads = 1;
  if (out) {
#pragma omp parallel num_threads(nthreads)
    {
      double *slab = malloc(sizeof(double) * M), *tB = malloc(sizeof(double) * n * n),
             *tBC = calloc((size_t)n * n, sizeof(double)), *res = malloc(sizeof(double) * (nij + 1));
#pragma omp for schedule(dynamic)
      for (int64_t z = 0; z < npq; ++z) {
        int64_t pq = pq0 + z + 1;
        if (pq > M) continue;
        for (int64_t rs = 1; rs <= M; ++rs) {
          int64_t hi = pq >= rs ? pq : rs, lo = pq >= rs ? rs : pq;
          slab[rs - 1] = orc_gen_value(gen_kind, seed, (uint64_t)((hi - 1) * M + (lo - 1)));
        }
        e_half_slab(n, C, ldc, slab, x1, x2, win[2], win[3], ijmap, nij, tB, tBC, res);
        for (int64_t ij = 0; ij < nij; ++ij) out[ij * npq + z] = fabs(res[ij]) > 1e-10 ? res[ij] : 0.0;
      }
      free(slab); free(tB); free(tBC); free(res);
    }
  }
  free(x1); free(x2); free(ijmap);
  return nij;
}

/* ------------------------------------------------------------------ */
/* DIRECT first quarter (the list-driven, un-densified form)           */
/* ------------------------------------------------------------------ */

/* Libint2Iface.cpp:793-868 with the integrals taken from a canonical list instead of being recomputed: for the MO index
 * whose AO coefficients are coef[mu] (the reference's C(p, bf)), every stored integral (bf1 bf2|bf3 bf4) -- 1-based in the
 * list, bf1>=bf2, bf3>=bf4, (12)>=(34), Iterators.cpp:45-77 -- adds v*coef to up to four elements of GG[n][n][n]
 * (:803-851), the last two indices ordered; then GG is symmetrised in its last two indices (:858-868).
 * GG[nu][lam][sig] = sum_mu (mu nu|lam sig) coef[mu], i.e. auxtempA(lam, sig, nu) of TransformIntegralsC.f90:545-558. */
void orc_direct_first_quarter(int n, const double *coef, const int32_t *p, const int32_t *q, const int32_t *r, const int32_t *s,
                              const double *v, int64_t count, double *GG) {
#define GG3(a, b, c) GG[((int64_t)(a) * n + (b)) * n + (c)]
  for (int64_t e = 0; e < (int64_t)n * n * n; ++e) GG[e] = 0.0;
  for (int64_t e = 0; e < count; ++e) {
    if (p[e] == -1) break;
    const int bf1 = p[e] - 1, bf2 = q[e] - 1, bf3 = r[e] - 1, bf4 = s[e] - 1;
    const double val = v[e];
    if (bf3 < bf4) GG3(bf2, bf3, bf4) += val * coef[bf1]; else GG3(bf2, bf4, bf3) += val * coef[bf1];
    if (bf1 != bf2) {
      if (bf3 < bf4) GG3(bf1, bf3, bf4) += val * coef[bf2]; else GG3(bf1, bf4, bf3) += val * coef[bf2];
    }
    if (bf1 != bf3 || bf2 != bf4) {
      if (bf1 < bf2) GG3(bf4, bf1, bf2) += val * coef[bf3]; else GG3(bf4, bf2, bf1) += val * coef[bf3];
      if (bf3 != bf4) {
        if (bf1 < bf2) GG3(bf3, bf1, bf2) += val * coef[bf4]; else GG3(bf3, bf2, bf1) += val * coef[bf4];
      }
    }
  }
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j)
      for (int k = j; k < n; ++k) GG3(i, k, j) = GG3(i, j, k);
#undef GG3
}
