/*
 * oracle/energy_oracle.c -- CPU restatement of the consumers of the MO
 * integrals that the downstream-energy clause is checked with: the MO-integral
 * reader addressing (src/core/ReadTransformedIntegrals.f90) and the APMO-MP2
 * second-order formula (src/MBPT/MPFunctions.f90).
 *
 * TEST INFRASTRUCTURE ONLY (see it_oracle.c header).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

int64_t orc_pair_id(int64_t i, int64_t j, int64_t n);

/* IndexMap_tensorR4ToVectorB, intra: src/core/IndexMap.f90:224-230 */
static int64_t r4b(int64_t i, int64_t j, int64_t k, int64_t l, int64_t n) {
  int64_t M = n * (n + 1) / 2;
  return orc_pair_id(orc_pair_id(i, j, n), orc_pair_id(k, l, n), M);
}

/* Reader, method E layout: ReadTransformedIntegrals.f90:302-311.
 * packed has M(M+1)/2 entries (1-based index -> packed[idx-1]); later entries
 * overwrite earlier ones at the same address, as in the reference. */
void orc_reader_pairs_intra(const int64_t *ij, const int64_t *kl, const double *v, int64_t cnt, int n,
                            double *packed) {
  int64_t M = (int64_t)n * (n + 1) / 2;
  for (int64_t k = 0; k < cnt; ++k) packed[orc_pair_id(ij[k], kl[k], M) - 1] = v[k];
}

/* Reader, method C layout: ReadTransformedIntegrals.f90:225-236. */
void orc_reader_quads_intra(const int32_t *p, const int32_t *q, const int32_t *r, const int32_t *s,
                            const double *v, int64_t cnt, int n, double *packed) {
  for (int64_t k = 0; k < cnt; ++k) packed[r4b(p[k], q[k], r[k], s[k], n) - 1] = v[k];
}

/* Reader, inter, method E: auxIndex = M_b*(pq-1)+rs, ReadTransformedIntegrals.f90:853-863. */
void orc_reader_pairs_inter(const int64_t *ij, const int64_t *kl, const double *v, int64_t cnt, int na,
                            int nb, double *rect) {
  int64_t Mb = (int64_t)nb * (nb + 1) / 2;
  for (int64_t k = 0; k < cnt; ++k) rect[Mb * (ij[k] - 1) + kl[k] - 1] = v[k];
}

/* Reader, inter, method C ("AB" order): IndexMap_tensorR4ToVectorB with two
 * basis sizes, IndexMap.f90:232-239; ReadTransformedIntegrals.f90:640-660. */
void orc_reader_quads_inter(const int32_t *p, const int32_t *q, const int32_t *r, const int32_t *s,
                            const double *v, int64_t cnt, int na, int nb, double *rect) {
  int64_t Mb = (int64_t)nb * (nb + 1) / 2;
  for (int64_t k = 0; k < cnt; ++k)
    rect[Mb * (orc_pair_id(p[k], q[k], na) - 1) + orc_pair_id(r[k], s[k], nb) - 1] = v[k];
}

/* Intra-species second-order correction before the charge / particle scaling:
 * src/MBPT/MPFunctions.f90:450-512 ("independentEnergyCorrection").
 * eps 1-based semantics (eps[a-1]); lambda = particles per orbital. */
double orc_mp2_intra(const double *packed, int n, int occ, int frozen, int active, double lambda,
                     const double *eps) {
  double e = 0.0;
  for (int a = frozen + 1; a <= occ; ++a)
    for (int b = frozen + 1; b <= occ; ++b)
      for (int r = occ + 1; r <= active; ++r)
        for (int s = r; s <= active; ++s) {
          double A = packed[r4b(a, r, b, s, n) - 1];
          if (!(fabs(A) > 1.0e-10)) continue;
          double den = eps[a - 1] + eps[b - 1] - eps[r - 1] - eps[s - 1];
          if (s > r) {
            if (a == b) {
              if (fabs(lambda - 1.0) > 1e-12) e += 2.0 * A * A * (lambda - 1.0) / den;
            } else {
              double B = packed[r4b(r, b, s, a, n) - 1];
              e += 2.0 * A * (lambda * A - B) / den;
            }
          } else if (a == b && r == s) {
            if (fabs(lambda - 1.0) > 1e-12) e += A * A * (lambda - 1.0) / (2.0 * (eps[a - 1] - eps[r - 1]));
          } else {
            if (fabs(lambda - 1.0) > 1e-12) e += A * A * (lambda - 1.0) / den;
          }
        }
  return e;
}

/* Scaling of the intra term: MPFunctions.f90:649-662. */
double orc_mp2_intra_scale(double e, double charge, int is_alpha_or_beta, double particles_fraction) {
  e *= pow(charge, 4.0);
  return is_alpha_or_beta ? e / 2.0 : e / (particles_fraction * 2.0);
}

/* Inter-species coupling term: MPFunctions.f90:744-765 and :878-879.
 * rect indexed M_b*(pair_a(a,r)-1)+pair_b(p,t) (1-based). */
double orc_mp2_inter(const double *rect, int na, int nb, int occ_a, int occ_b, int frozen_a, int frozen_b,
                     int active_a, int active_b, double charge_a, double charge_b, double lambda_a,
                     double lambda_b, const double *eps_a, const double *eps_b) {
  int64_t Mb = (int64_t)nb * (nb + 1) / 2;
  double e = 0.0;
  for (int a = frozen_a + 1; a <= occ_a; ++a)
    for (int p = frozen_b + 1; p <= occ_b; ++p)
      for (int r = occ_a + 1; r <= active_a; ++r)
        for (int t = occ_b + 1; t <= active_b; ++t) {
          double x = rect[Mb * (orc_pair_id(a, r, na) - 1) + orc_pair_id(p, t, nb) - 1] * charge_a * charge_b;
          e += x * x / (eps_a[a - 1] + eps_b[p - 1] - eps_a[r - 1] - eps_b[t - 1]);
        }
  return lambda_a * lambda_b * e;
}
