/*
 * oracle/blas_shim.c -- dgemm_ provider for the reference's transformer D build
 * (oracle/_ref/libref_d.so).  TEST INFRASTRUCTURE ONLY.
 *
 * The reference links any system BLAS (configure:52, "-lblas -llapack"); this
 * image has none on the link path.  orc_blas_bind(path) dlopen()s a BLAS that
 * exports a plain dgemm_ (the OpenBLAS bundled in site-packages' opencv libs);
 * if that is not bound, a straightforward OpenMP column-major dgemm is used.
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <stddef.h>

typedef void (*dgemm_fn)(char *, char *, int *, int *, int *, double *, double *, int *, double *, int *,
                         double *, double *, int *);
static dgemm_fn g_ext = NULL;

int orc_blas_bind(const char *libgfortran_path, const char *blas_path) {
  if (libgfortran_path && *libgfortran_path) dlopen(libgfortran_path, RTLD_NOW | RTLD_GLOBAL);
  void *h = dlopen(blas_path, RTLD_NOW | RTLD_LOCAL);
  if (!h) return 1;
  dgemm_fn f = (dgemm_fn)dlsym(h, "dgemm_");
  if (!f) return 2;
  g_ext = f;
  return 0;
}
int orc_blas_is_external(void) { return g_ext != NULL; }

void dgemm_(char *ta, char *tb, int *pm, int *pn, int *pk, double *palpha, double *A, int *plda, double *B,
            int *pldb, double *pbeta, double *C, int *pldc) {
  if (g_ext) { g_ext(ta, tb, pm, pn, pk, palpha, A, plda, B, pldb, pbeta, C, pldc); return; }
  int m = *pm, n = *pn, k = *pk, lda = *plda, ldb = *pldb, ldc = *pldc;
  int tA = (*ta == 'T' || *ta == 't'), tB = (*tb == 'T' || *tb == 't');
  double alpha = *palpha, beta = *pbeta;
#pragma omp parallel for schedule(static)
  for (int j = 0; j < n; ++j) {
    for (int i = 0; i < m; ++i) C[i + (size_t)j * ldc] = (beta == 0.0) ? 0.0 : beta * C[i + (size_t)j * ldc];
    if (!tA) {
      for (int l = 0; l < k; ++l) {
        double b = alpha * (tB ? B[j + (size_t)l * ldb] : B[l + (size_t)j * ldb]);
        const double *a = A + (size_t)l * lda;
        double *c = C + (size_t)j * ldc;
        for (int i = 0; i < m; ++i) c[i] += a[i] * b;
      }
    } else {
      for (int i = 0; i < m; ++i) {
        const double *a = A + (size_t)i * lda;
        double acc = 0.0;
        if (!tB) { const double *b = B + (size_t)j * ldb; for (int l = 0; l < k; ++l) acc += a[l] * b[l]; }
        else     { for (int l = 0; l < k; ++l) acc += a[l] * B[j + (size_t)l * ldb]; }
        C[i + (size_t)j * ldc] += alpha * acc;
      }
    }
  }
}
