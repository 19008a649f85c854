/*
 * tests/mock/mock_it.c -- TEST INFRASTRUCTURE ONLY.  A CPU stand-in for the subset of include/lowdin_it.h that the host
 * mirror (openlowdin_b200/csrc/host_mirror.cpp) calls, answered by the oracle's restatement of the reference
 * (oracle/it_oracle.c).  Linked with host_mirror.cpp into tests/mock/_build/libmock_host.so so that the `-m "not gpu"`
 * suite can run the mirror's file-to-file calls (stream readers, record writers, the program loop, method D) without a
 * GPU.  Nothing under openlowdin_b200/ links or loads this; the product library has no CPU path.
 */
#include "../../include/lowdin_it.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

int64_t orc_scatter_intra(const int32_t *, const int32_t *, const int32_t *, const int32_t *, const double *, int64_t, int, double *);
int64_t orc_scatter_inter(const int32_t *, const int32_t *, const int32_t *, const int32_t *, const double *, int64_t, int, int, int,
                          double *);
int64_t orc_transform_e_intra(int, const double *, int, const double *, const int *, int64_t *, int64_t *, double *, int64_t);
int64_t orc_transform_e_inter(int, int, const double *, int, const double *, int, const double *, const int *, int64_t *, int64_t *,
                              double *, int64_t);
int64_t orc_transform_c_intra(int, const double *, int, const double *, const int *, int, int32_t *, int32_t *, int32_t *, int32_t *,
                              double *, int64_t);
int64_t orc_transform_c_inter(int, int, const double *, int, const double *, int, const double *, const int *, int, int32_t *, int32_t *,
                              int32_t *, int32_t *, double *, int64_t);
void orc_transform_d_intra(const double *, double *, int);
void orc_transform_d_inter(const double *, const double *, double *, int, int);

struct lowdin_it_ctx {
  int n[8], ncols[8];
  double *C[8];           /* column-major, ld = n */
  double *ao[8][8];
  int up_a, up_b, up_sw;
  int conv;
  int64_t count;
  int64_t *ij, *kl;
  int32_t *q4[4];
  double *v;
  char err[256];
};

static char g_err[256];

static int mfail(lowdin_it_handle h, const char *m) {
  snprintf(h ? h->err : g_err, 256, "%s", m);
  return 1;
}

int lowdin_it_create(int device, lowdin_it_handle *out) {
  (void)device;
  *out = (lowdin_it_handle)calloc(1, sizeof(struct lowdin_it_ctx));
  (*out)->up_a = -1;
  return 0;
}

int lowdin_it_destroy(lowdin_it_handle h) {
  if (!h) return 0;
  for (int a = 0; a < 8; ++a) { free(h->C[a]); for (int b = 0; b < 8; ++b) free(h->ao[a][b]); }
  free(h->ij); free(h->kl); free(h->v);
  for (int k = 0; k < 4; ++k) free(h->q4[k]);
  free(h);
  return 0;
}

const char *lowdin_it_last_error(lowdin_it_handle h) { return h ? h->err : g_err; }

int lowdin_it_set_species(lowdin_it_handle h, int slot, int nao, const double *C, int ldc, int ncols) {
  if (!h || slot < 0 || slot > 7 || nao <= 0 || ncols <= 0 || ldc < nao || !C) return mfail(h, "bad coefficient matrix arguments");
  free(h->C[slot]);
  h->C[slot] = (double *)malloc(sizeof(double) * (size_t)nao * ncols);
  for (int k = 0; k < ncols; ++k) memcpy(h->C[slot] + (size_t)k * nao, C + (size_t)k * ldc, sizeof(double) * nao);
  h->n[slot] = nao; h->ncols[slot] = ncols;
  return 0;
}

static int64_t npairs(int64_t n) { return n * (n + 1) / 2; }

int lowdin_it_ao_begin(lowdin_it_handle h, int a, int b, int swapped) {
  if (!h || !h->n[a] || !h->n[b]) return mfail(h, "ao_begin: species not set");
  const int64_t Ma = npairs(h->n[a]), Mb = npairs(h->n[b]);
  free(h->ao[a][b]);
  h->ao[a][b] = (double *)calloc((size_t)(a == b ? Ma * (Ma + 1) / 2 : Ma * Mb), sizeof(double));
  h->up_a = a; h->up_b = b; h->up_sw = swapped;
  return 0;
}

int lowdin_it_ao_push_stacks(lowdin_it_handle h, const int32_t *p, const int32_t *q, const int32_t *r, const int32_t *s,
                             const double *v, int64_t n) {
  if (!h || h->up_a < 0) return mfail(h, "ao_push_stacks without ao_begin");
  const int a = h->up_a, b = h->up_b;
  const int lim_pq = (a == b) ? h->n[a] : (h->up_sw ? h->n[b] : h->n[a]), lim_rs = (a == b) ? h->n[a] : (h->up_sw ? h->n[a] : h->n[b]);
  for (int64_t k = 0; k < n && p[k] != -1; ++k)
    if (p[k] < 1 || q[k] < 1 || r[k] < 1 || s[k] < 1 || p[k] > lim_pq || q[k] > lim_pq || r[k] > lim_rs || s[k] > lim_rs)
      return mfail(h, "AO stack entry has an index outside the basis");
  if (a == b) orc_scatter_intra(p, q, r, s, v, n, h->n[a], h->ao[a][b]);
  else orc_scatter_inter(p, q, r, s, v, n, h->n[a], h->n[b], h->up_sw, h->ao[a][b]);
  return 0;
}

int lowdin_it_ao_push_blocks(lowdin_it_handle h, const void *blocks, int64_t nblocks, int S) {
  const unsigned char *raw = (const unsigned char *)blocks;
  for (int64_t t = 0; t < nblocks; ++t) {
    const int32_t *p = (const int32_t *)(raw + (size_t)t * 24 * S);
    int rc = lowdin_it_ao_push_stacks(h, p, p + S, p + 2 * S, p + 3 * S, (const double *)(p + 4 * S), S);
    if (rc) return rc;
    for (int i = 0; i < S; ++i) if (p[i] == -1) return 0; /* the terminator ends this call's stream */
  }
  return 0;
}

int lowdin_it_ao_end(lowdin_it_handle h) {
  if (!h || h->up_a < 0) return mfail(h, "ao_end without ao_begin");
  h->up_a = h->up_b = -1;
  return 0;
}

int lowdin_it_transform(lowdin_it_handle h, int a, int b, const int win[8], int conv, int symmetric, double drop_tol) {
  (void)drop_tol; /* the oracle applies the reference's 1e-10 */
  if (!h || !h->ao[a][b]) return mfail(h, "AO integrals for this species pair were not uploaded");
  const int lim[4] = {h->ncols[a], h->ncols[a], h->ncols[b], h->ncols[b]};
  int64_t cap = 1;
  for (int w = 0; w < 4; ++w) {
    if (win[2 * w] < 1 || win[2 * w + 1] > lim[w]) return mfail(h, "window out of range");
    const int c = win[2 * w + 1] - win[2 * w] + 1;
    cap *= (c > 0 ? c : 0);
  }
  if (cap < 1) cap = 1;
  free(h->ij); free(h->kl); free(h->v);
  for (int k = 0; k < 4; ++k) { free(h->q4[k]); h->q4[k] = NULL; }
  h->ij = h->kl = NULL;
  h->v = (double *)malloc(sizeof(double) * cap);
  h->conv = conv;
  if (conv == LOWDIN_IT_CONV_E) {
    h->ij = (int64_t *)malloc(sizeof(int64_t) * cap); h->kl = (int64_t *)malloc(sizeof(int64_t) * cap);
    h->count = (a == b) ? orc_transform_e_intra(h->n[a], h->C[a], h->n[a], h->ao[a][b], win, h->ij, h->kl, h->v, cap)
                        : orc_transform_e_inter(h->n[a], h->n[b], h->C[a], h->n[a], h->C[b], h->n[b], h->ao[a][b], win, h->ij, h->kl, h->v, cap);
  } else {
    for (int k = 0; k < 4; ++k) h->q4[k] = (int32_t *)malloc(sizeof(int32_t) * cap);
    h->count = (a == b) ? orc_transform_c_intra(h->n[a], h->C[a], h->n[a], h->ao[a][b], win, symmetric, h->q4[0], h->q4[1], h->q4[2], h->q4[3], h->v, cap)
                        : orc_transform_c_inter(h->n[a], h->n[b], h->C[a], h->n[a], h->C[b], h->n[b], h->ao[a][b], win, symmetric, h->q4[0], h->q4[1],
                                                h->q4[2], h->q4[3], h->v, cap);
  }
  return 0;
}

int lowdin_it_result_count(lowdin_it_handle h, int64_t *count) { *count = h->count; return 0; }

int lowdin_it_download_pairs(lowdin_it_handle h, int64_t *ij, int64_t *kl, double *v) {
  if (h->conv != LOWDIN_IT_CONV_E) return mfail(h, "last transform did not use the E (pair id) convention");
  memcpy(ij, h->ij, sizeof(int64_t) * h->count); memcpy(kl, h->kl, sizeof(int64_t) * h->count); memcpy(v, h->v, sizeof(double) * h->count);
  return 0;
}

int lowdin_it_download_quads(lowdin_it_handle h, int32_t *p, int32_t *q, int32_t *r, int32_t *s, double *v) {
  if (h->conv != LOWDIN_IT_CONV_C) return mfail(h, "last transform did not use the C (quad) convention");
  int32_t *o[4] = {p, q, r, s};
  for (int k = 0; k < 4; ++k) memcpy(o[k], h->q4[k], sizeof(int32_t) * h->count);
  memcpy(v, h->v, sizeof(double) * h->count);
  return 0;
}

/* row f4 is a device computation: not in the mock */
int lowdin_it_ao_compute(lowdin_it_handle h, int a, int b) { (void)h; (void)a; (void)b; return 1; }
int lowdin_it_ao_download(lowdin_it_handle h, int a, int b, double *out, int64_t cap) { (void)h; (void)a; (void)b; (void)out; (void)cap; return 1; }

/* group calls: the mock is one handle */
int lowdin_it_group_transform(lowdin_it_handle *hs, int n, int a, int b, const int win[8], int conv, int symmetric, double tol) {
  (void)n; return lowdin_it_transform(hs[0], a, b, win, conv, symmetric, tol);
}
int lowdin_it_group_result_count(lowdin_it_handle *hs, int n, int64_t *count) { (void)n; return lowdin_it_result_count(hs[0], count); }
int lowdin_it_group_download_pairs(lowdin_it_handle *hs, int n, int64_t *ij, int64_t *kl, double *v) { (void)n; return lowdin_it_download_pairs(hs[0], ij, kl, v); }
int lowdin_it_group_download_quads(lowdin_it_handle *hs, int n, int32_t *p, int32_t *q, int32_t *r, int32_t *s, double *v) {
  (void)n; return lowdin_it_download_quads(hs[0], p, q, r, s, v);
}

int lowdin_it_transform_all(const double *coeff, double *ints, int nao) { orc_transform_d_intra(coeff, ints, nao); return 0; }
int lowdin_it_transform_inter_all(const double *coeff, const double *ocoeff, double *ints, int nao, int onao) {
  orc_transform_d_inter(coeff, ocoeff, ints, nao, onao);
  return 0;
}
