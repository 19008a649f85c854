// host_mirror.cpp -- C++ mirror of the reference's Fortran host side for transformers C and E
// (include/lowdin_it_host.h): window tables, partialTransform choice, .ints stream readers, moint.dat record
// writers and the per-species / per-pair transformer calls on top of the C ABI of lowdin_it.h.
// No arithmetic of the transformation lives here (that is csrc/it_kernels.cuh); this file is integer / file logic.
#include "../../include/lowdin_it_host.h"

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace {

thread_local std::string g_err;
int hfail(const std::string &m) { g_err = m; return 1; }

struct Win { int pl, pu, ql, qu, rl, ru, sl, su; };
void put(const Win &w, int out[8]) {
  const int v[8] = {w.pl, w.pu, w.ql, w.qu, w.rl, w.ru, w.sl, w.su};
  memcpy(out, v, sizeof v);
}
bool is(const lowdin_host_control *c, const char *s) { return strncmp(c->partial_transform, s, sizeof c->partial_transform) == 0; }

std::string trimmed(const char *s, size_t cap) {
  std::string t(s, strnlen(s, cap));
  while (!t.empty() && t.back() == ' ') t.pop_back();
  return t;
}

void ionize_flags(const lowdin_host_control *c, const lowdin_host_species *a, const lowdin_host_species *b, bool &ia, bool &ib) {
  ia = ib = false;  // C.f90:1703-1716 / E.f90:2156-2168: species names compared with every IONIZE_SPECIES entry
  const std::string na = trimmed(a->name, sizeof a->name), nb = trimmed(b->name, sizeof b->name);
  for (int s = 0; s < c->n_ionize_species && s < 4; ++s) {
    const std::string t = trimmed(c->ionize_species[s], 32);
    if (na == t) ia = true;
    if (nb == t) ib = true;
  }
}

// ---- transformer C, intra: TransformIntegralsC.f90:1436-1622 ------------------------------------
Win win_c_intra(const lowdin_host_control *c, const lowdin_host_species *a, int *symmetric) {
  const int occ = a->occupation, core = a->core_orbitals, act = a->active_orbitals ? a->active_orbitals : a->nao, mo = c->ionize_mo;
  *symmetric = 1;
  Win w{core + 1, act, core + 1, act, core + 1, act, core + 1, act};
  if (is(c, "ALL")) w = Win{1, a->nao, 1, a->nao, 1, a->nao, 1, a->nao};
  if (is(c, "ALLACTIVE")) w = Win{1, act, 1, act, 1, act, 1, act};
  if (is(c, "MP2")) w = Win{core + 1, occ, occ + 1, act, core + 1, occ, occ + 1, act};
  if (is(c, "PT2") || is(c, "MP2-PT2")) {
    *symmetric = 0;
    const bool both = is(c, "MP2-PT2");
    int pl, pu;
    if (mo == 0) { pl = both ? core + 1 : occ; pu = occ + 1; }          // HOMO..LUMO (PT2) / core+1..LUMO (MP2-PT2)
    else if (both) { pl = core + 1; pu = std::max(mo, occ); }
    else { pl = mo; pu = mo; }
    const int sl = (mo != 0 && c->pt_transition_operator) ? core + 1 : occ + 1;
    w = Win{pl, pu, core + 1, act, core + 1, occ, sl, act};
  }
  return w;
}

// ---- transformer E, intra: TransformIntegralsE.f90:1899-2073 (roles of the pair members swapped; no ALL case) ----
Win win_e_intra(const lowdin_host_control *c, const lowdin_host_species *a) {
  const int occ = a->occupation, core = a->core_orbitals, act = a->active_orbitals ? a->active_orbitals : a->nao, mo = c->ionize_mo;
  Win w{core + 1, act, core + 1, act, core + 1, act, core + 1, act};
  if (is(c, "ALLACTIVE")) w = Win{1, act, 1, act, 1, act, 1, act};
  if (is(c, "MP2")) w = Win{occ + 1, act, core + 1, occ, occ + 1, act, core + 1, occ};
  if (is(c, "PT2") || is(c, "MP2-PT2")) {
    int ql, qu;
    if (mo == 0) { ql = core + 1; qu = occ + 1; }
    else if (is(c, "MP2-PT2")) { ql = core + 1; qu = std::max(mo, occ); }
    else { ql = mo; qu = mo; }
    const int rl = (mo != 0 && c->pt_transition_operator) ? core + 1 : occ + 1;
    w = Win{core + 1, act, ql, qu, rl, act, core + 1, occ};
  }
  return w;
}

// ---- transformer C, inter: TransformIntegralsC.f90:1625-1963 ------------------------------------
Win win_c_inter(const lowdin_host_control *c, const lowdin_host_species *a, const lowdin_host_species *b, int *symmetric) {
  const int oa = a->occupation, ob = b->occupation, ca = a->core_orbitals, cb = b->core_orbitals;
  const int aa = a->active_orbitals ? a->active_orbitals : a->nao, ab = b->active_orbitals ? b->active_orbitals : b->nao;
  const int mo = c->ionize_mo;
  *symmetric = 1;
  Win w{ca + 1, aa, ca + 1, aa, cb + 1, ab, cb + 1, ab};
  if (is(c, "ALL")) w = Win{1, a->nao, 1, a->nao, 1, b->nao, 1, b->nao};
  if (is(c, "ALLACTIVE")) w = Win{1, aa, 1, aa, 1, ab, 1, ab};
  if (is(c, "MP2")) w = Win{ca + 1, oa, oa + 1, aa, cb + 1, ob, ob + 1, ab};
  if (is(c, "PT2") || is(c, "MP2-PT2")) {
    const bool both = is(c, "MP2-PT2");
    w = Win{ca + 1, oa + 1, ca + 1, aa, cb + 1, ob + 1, cb + 1, ab};
    if (c->n_ionize_species > 0) {
      *symmetric = 0;
      bool ia, ib;
      ionize_flags(c, a, b, ia, ib);
      if (mo == 0) {
        if (ia && ib) w = Win{ca + 1, oa + 1, ca + 1, aa, cb + 1, ob + 1, cb + 1, ab};
        else if (ia) w = Win{ca + 1, oa + 1, ca + 1, aa, cb + 1, ob, ob + 1, ab};
        else if (ib) w = Win{ca + 1, oa, oa + 1, aa, cb + 1, ob + 1, cb + 1, ab};
      } else {
        if (ia && ib) {
          if (mo <= oa && mo <= ob) w = Win{ca + 1, oa, ca + 1, aa, cb + 1, ob, cb + 1, ab};
          else if (mo > oa && mo > ob) w = Win{ca + 1, mo, ca + 1, aa, cb + 1, mo, cb + 1, ab};
        } else if (ia) {
          w = both ? Win{ca + 1, std::max(mo, oa), ca + 1, aa, cb + 1, ob, ob + 1, ab} : Win{mo, mo, ca + 1, aa, cb + 1, ob, ob + 1, ab};
        } else if (ib) {
          w = both ? Win{ca + 1, oa, oa + 1, aa, cb + 1, std::max(mo, ob), cb + 1, ab} : Win{ca + 1, oa, oa + 1, aa, mo, mo, cb + 1, ab};
        }
      }
    }
  }
  return w;
}

// ---- transformer E, inter: TransformIntegralsE.f90:2076-2418 ------------------------------------
// Quirks kept as they are in the reference: the PT2 default takes s_l and r_l from the FIRST species' core
// orbitals (E.f90:2142-2145) and the MP2-PT2 default sets r_u to the first species' active count (E.f90:2283).
Win win_e_inter(const lowdin_host_control *c, const lowdin_host_species *a, const lowdin_host_species *b) {
  const int oa = a->occupation, ob = b->occupation, ca = a->core_orbitals, cb = b->core_orbitals;
  const int aa = a->active_orbitals ? a->active_orbitals : a->nao, ab = b->active_orbitals ? b->active_orbitals : b->nao;
  const int mo = c->ionize_mo;
  Win w{ca + 1, aa, ca + 1, aa, cb + 1, ab, cb + 1, ab};
  if (is(c, "ALLACTIVE")) w = Win{1, aa, 1, aa, 1, ab, 1, ab};
  if (is(c, "MP2")) w = Win{oa + 1, aa, ca + 1, oa, ob + 1, ab, cb + 1, ob};
  if (is(c, "PT2") || is(c, "MP2-PT2")) {
    const bool both = is(c, "MP2-PT2");
    w = both ? Win{ca + 1, aa, ca + 1, oa + 1, cb + 1, aa, cb + 1, ob + 1} : Win{ca + 1, aa, ca + 1, oa + 1, ca + 1, ab, ca + 1, ob + 1};
    if (c->n_ionize_species > 0) {
      bool ia, ib;
      ionize_flags(c, a, b, ia, ib);
      if (mo == 0) {
        if (ia && ib) w = Win{ca + 1, aa, ca + 1, oa + 1, cb + 1, ab, cb + 1, ob + 1};
        else if (ia) w = Win{ca + 1, aa, ca + 1, oa + 1, ob + 1, ab, cb + 1, ob};
        else if (ib) w = Win{oa + 1, aa, ca + 1, oa, cb + 1, ab, cb + 1, ob + 1};
      } else {
        if (ia && ib) {
          if (mo <= oa && mo <= ob) w = Win{ca + 1, aa, ca + 1, oa, cb + 1, ab, cb + 1, ob};
          else if (mo > oa && mo > ob) w = Win{ca + 1, aa, ca + 1, aa, cb + 1, ab, cb + 1, ab};
        } else if (ia) {
          int ql = both ? ca + 1 : mo, qu = both ? std::max(mo, oa) : mo;
          if (c->pt_transition_operator) { ql = ca + 1; qu = aa; }
          w = Win{ca + 1, aa, ql, qu, ob + 1, ab, cb + 1, ob};
        } else if (ib) {
          w = Win{oa + 1, aa, ca + 1, oa, cb + 1, ab, cb + 1, ab};
        }
      }
    }
  }
  return w;
}

// ---- transformer D: TransformIntegralsD.f90:595-700.  Printed, never used (D always transforms everything); the intra
// table is 0-based (:606-626) and the inter table 1-based (:668-690), as in the reference. ----
Win win_d(const lowdin_host_control *c, const lowdin_host_species *a, const lowdin_host_species *b) {
  if (!b) {
    const int n = a->nao, occ = a->occupation;
    if (is(c, "MP2")) return Win{0, occ - 1, occ, n - 1, 0, occ - 1, occ, n - 1};
    return Win{0, n - 1, 0, n - 1, 0, n - 1, 0, n - 1};
  }
  if (is(c, "MP2")) return Win{1, a->occupation, a->occupation + 1, a->nao, 1, b->occupation, b->occupation + 1, b->nao};
  return Win{1, a->nao, 1, a->nao, 1, b->nao, 1, b->nao};
}

std::string join(const char *dir, const std::string &file) {
  std::string d = trimmed(dir, 512);
  if (d.empty()) return file;
  if (d.back() != '/') d += '/';
  return d + file;
}

int check_ctl(const lowdin_host_control *c) {
  if (!c) return hfail("null control block");
  if (c->method != 'C' && c->method != 'E' && c->method != 'D') return hfail("method must be 'C', 'D' or 'E'");
  if (c->integral_stack_size < 1) return hfail("integral_stack_size < 1");
  if (c->nfiles < 1) return hfail("nfiles < 1");
  return 0;
}

// Reads every stack of one stream file and hands it to `sink` (the reader loops of C.f90:251-295).
template <class Sink>
int for_each_stack(const std::string &path, int S, Sink sink) {
  FILE *f = fopen(path.c_str(), "rb");
  if (!f) return hfail("cannot open " + path);
  std::vector<int32_t> p(S), q(S), r(S), s(S);
  std::vector<double> v(S);
  int rc = 0;
  for (;;) {
    size_t a = fread(p.data(), 4, S, f);
    if (a == 0) break;  // the reference trusts filesize/24/S; a file without terminator simply ends
    if (a != (size_t)S || fread(q.data(), 4, S, f) != (size_t)S || fread(r.data(), 4, S, f) != (size_t)S ||
        fread(s.data(), 4, S, f) != (size_t)S || fread(v.data(), 8, S, f) != (size_t)S) { rc = hfail("truncated stack in " + path); break; }
    bool last = false;
    for (int i = 0; i < S; ++i) if (p[i] == -1) { last = true; break; }
    if ((rc = sink(p.data(), q.data(), r.data(), s.data(), v.data(), S)) != 0) break;
    if (last) break;
  }
  fclose(f);
  return rc;
}

int write_record(FILE *f, const void *a, size_t na, const void *b, size_t nb, const void *c, size_t nc, const void *d, size_t nd,
                 const void *e, size_t ne) {
  const uint32_t len = (uint32_t)(na + nb + nc + nd + ne);  // gfortran sequential unformatted: 4-byte length before and after
  if (fwrite(&len, 4, 1, f) != 1) return 1;
  if (na && fwrite(a, 1, na, f) != na) return 1;
  if (nb && fwrite(b, 1, nb, f) != nb) return 1;
  if (nc && fwrite(c, 1, nc, f) != nc) return 1;
  if (nd && fwrite(d, 1, nd, f) != nd) return 1;
  if (ne && fwrite(e, 1, ne, f) != ne) return 1;
  return fwrite(&len, 4, 1, f) != 1;
}

// hs[0..nh): one handle, or the handles of an in-process group (lowdin_it_comm_init_local): every handle is pushed every byte and
// keeps the rows of the AO tensor it owns
int load_ints(lowdin_it_handle *hs, int nh, const lowdin_host_control *ctl, const lowdin_host_species *a, const lowdin_host_species *b) {
  const int slot_b = b ? 1 : 0;
  int swapped = 0;
  char name[256];
  if (lowdin_host_ints_filename(0, a, b, name, &swapped)) return 1;
  for (int d = 0; d < nh; ++d)
    if (lowdin_it_ao_begin(hs[d], 0, slot_b, swapped)) return hfail(lowdin_it_last_error(hs[d]));
  // every per-thread stream file as RAW bytes, up to 64 MiB of whole stacks per call: the device decodes the blocks
  // `pp, qq, rr, ss, shellIntegrals` (terminator, index check, scatter: TransformIntegralsC.f90:258-296)
  const size_t S = (size_t)ctl->integral_stack_size, block = 24 * S;
  const size_t per_read = std::max<size_t>(1, ((size_t)64 << 20) / block);
  std::vector<unsigned char> raw(per_read * block);
  for (int tid = 0; tid < ctl->nfiles; ++tid) {
    if (lowdin_host_ints_filename(tid, a, b, name, &swapped)) return 1;
    const std::string path = join(ctl->scratch_dir, name);
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) return hfail("cannot open " + path);
    for (;;) {
      const size_t got = fread(raw.data(), 1, raw.size(), f);
      if (got == 0) break;
      if (got % block) { fclose(f); return hfail("truncated stack in " + path); }  // the reference trusts filesize/24/S
      for (int d = 0; d < nh; ++d)
        if (lowdin_it_ao_push_blocks(hs[d], raw.data(), (int64_t)(got / block), (int)S)) { fclose(f); return hfail(lowdin_it_last_error(hs[d])); }
      if (got < raw.size()) break;
    }
    fclose(f);
  }
  for (int d = 0; d < nh; ++d)
    if (lowdin_it_ao_end(hs[d])) return hfail(lowdin_it_last_error(hs[d]));
  return 0;
}

// Method D, file to file (TransformIntegralsD.f90:141-273 intra, :370-490 inter): ReadIntegrals_* scatter the .ints streams into
// D's packed array on the HOST (ReadIntegrals.f90:23-99, :102-172), the transform runs in place behind the
// c_integrals_transform_all signature, one record per integral is written.
inline int64_t d_index2(int64_t i, int64_t j) { return i > j ? i * (i + 1) / 2 + j : j * (j + 1) / 2 + i; }

int run_and_write_d(const lowdin_host_control *ctl, const lowdin_host_species *a, const lowdin_host_species *b, int64_t *nonzero) {
  if (!a || !a->coeff || (b && !b->coeff)) return hfail("null species / coefficients");
  const int nao = a->nao, onao = b ? b->nao : 0;
  if (a->ncols < nao || (b && b->ncols < onao)) return hfail("method D needs the square coefficient matrix");
  const int64_t sze = (int64_t)nao * (nao + 1) / 2, osze = (int64_t)onao * (onao + 1) / 2;
  std::vector<double> ints((size_t)(b ? sze * osze : sze * (sze + 1) / 2), 0.0);
  const std::string na = trimmed(a->name, 32), nb = b ? trimmed(b->name, 32) : std::string();
  const std::string stem = b ? na + "." + nb : (na == "E-BETA" ? std::string("E-ALPHA") : na);  // ReadIntegrals.f90:64-68; no alias for pairs (:136)
  for (int tid = 0; tid < ctl->nfiles; ++tid) {
    bool bad = false;
    int rc = for_each_stack(join(ctl->scratch_dir, std::to_string(tid) + stem + ".ints"), ctl->integral_stack_size,
                            [&](const int32_t *p, const int32_t *q, const int32_t *r, const int32_t *s, const double *v, int n) {
                              for (int i = 0; i < n; ++i) {
                                if (p[i] == -1) break;
                                const int lim2 = b ? onao : nao;
                                if (p[i] < 1 || q[i] < 1 || r[i] < 1 || s[i] < 1 || p[i] > nao || q[i] > nao || r[i] > lim2 || s[i] > lim2) { bad = true; return 1; }
                                const int64_t ij = d_index2(p[i] - 1, q[i] - 1), kl = d_index2(r[i] - 1, s[i] - 1);
                                ints[(size_t)(b ? ij * osze + kl : d_index2(ij, kl))] = v[i];  // ReadIntegrals_index4Inter / index4Intra, 0-based
                              }
                              return 0;
                            });
    if (bad) return hfail("AO stack entry with an index outside the basis");
    if (rc) return rc;
  }
  // coeff(j,k) = coefficients%values(j,k), j,k <= nao (TransformIntegralsD.f90:189-196)
  std::vector<double> ca((size_t)nao * nao), cb((size_t)onao * onao);
  for (int k = 0; k < nao; ++k) memcpy(&ca[(size_t)k * nao], a->coeff + (size_t)k * a->ldc, sizeof(double) * nao);
  for (int k = 0; k < onao; ++k) memcpy(&cb[(size_t)k * onao], b->coeff + (size_t)k * b->ldc, sizeof(double) * onao);
  if (b ? lowdin_it_transform_inter_all(ca.data(), cb.data(), ints.data(), nao, onao) : lowdin_it_transform_all(ca.data(), ints.data(), nao))
    return hfail(std::string("transformer-D entry point failed: ") + lowdin_it_last_error(nullptr));
  const std::string path = join(ctl->scratch_dir, (b ? na + "." + nb : na) + "moint.dat");
  int64_t n = 0;
  if (b ? lowdin_host_write_moint_d_inter(path.c_str(), nao, onao, ints.data(), &n) : lowdin_host_write_moint_d_intra(path.c_str(), nao, ints.data(), &n))
    return 1;
  if (ctl->verbose) printf(" Non zero transformed repulsion integrals: %lld\n", (long long)n);  // D.f90:269
  if (nonzero) *nonzero = n;
  return 0;
}

int run_and_write(lowdin_it_handle *hs, int nh, const lowdin_host_control *ctl, const lowdin_host_species *a, const lowdin_host_species *b,
                  int64_t *nonzero) {
  if (check_ctl(ctl)) return 1;
  if (ctl->method == 'D') return run_and_write_d(ctl, a, b, nonzero);
  if (!hs || nh < 1 || !hs[0] || !a || !a->coeff) return hfail("null handle / species");
  lowdin_it_handle h = hs[0];
  int win[8], symmetric = 0;
  if (lowdin_host_windows(ctl, a, b, win, &symmetric)) return 1;
  if (ctl->verbose) {  // the "Transformation boundaries" table of C.f90:1614-1620
    printf("              Transformation boundaries \n              orbital   lower upper\n");
    const char *nm = "pqrs";
    for (int w = 0; w < 4; ++w) printf("                   %c%6d%6d\n", nm[w], win[2 * w], win[2 * w + 1]);
  }
  for (int d = 0; d < nh; ++d) {
    if (lowdin_it_set_species(hs[d], 0, a->nao, a->coeff, a->ldc, a->ncols)) return hfail(lowdin_it_last_error(hs[d]));
    if (b && lowdin_it_set_species(hs[d], 1, b->nao, b->coeff, b->ldc, b->ncols)) return hfail(lowdin_it_last_error(hs[d]));
  }
  if (load_ints(hs, nh, ctl, a, b)) return 1;
  const int conv = (ctl->method == 'E') ? LOWDIN_IT_CONV_E : LOWDIN_IT_CONV_C;
  // one handle: lowdin_it_transform; a group: the collective transform on every handle at once, merged downloads below
  if (lowdin_it_group_transform(hs, nh, 0, b ? 1 : 0, win, conv, symmetric, 1e-10)) return hfail(lowdin_it_last_error(h));
  int64_t n = 0;
  if (lowdin_it_group_result_count(hs, nh, &n)) return hfail(lowdin_it_last_error(h));
  const std::string prefix = b ? trimmed(a->name, 32) + "." + trimmed(b->name, 32) : trimmed(a->name, 32);  // C.f90:192, :788
  const std::string path = join(ctl->scratch_dir, prefix + "moint.dat");
  const size_t m = (size_t)std::max<int64_t>(n, 1);
  std::vector<double> v(m);
  int rc;
  if (conv == LOWDIN_IT_CONV_E) {
    std::vector<int64_t> ij(m), kl(m);
    if (lowdin_it_group_download_pairs(hs, nh, ij.data(), kl.data(), v.data())) return hfail(lowdin_it_last_error(h));
    rc = lowdin_host_write_moint_pairs(path.c_str(), ctl->integral_stack_size, ij.data(), kl.data(), v.data(), n);
  } else {
    std::vector<int32_t> p(m), q(m), r(m), s(m);
    if (lowdin_it_group_download_quads(hs, nh, p.data(), q.data(), r.data(), s.data(), v.data())) return hfail(lowdin_it_last_error(h));
    rc = lowdin_host_write_moint_quads(path.c_str(), ctl->integral_stack_size, p.data(), q.data(), r.data(), s.data(), v.data(), n);
  }
  if (rc) return rc;
  if (ctl->verbose) printf("   %36s%12lld\n", "Non-zero transformed integrals: ", (long long)n);  // C.f90:469
  if (nonzero) *nonzero = n;
  return 0;
}

// Algorithmic flops of one call, with the library's rule "smaller window contracted first".
double call_flops(const lowdin_host_control *ctl, const lowdin_host_species *a, const lowdin_host_species *b, const int win[8],
                  int symmetric) {
  const lowdin_host_species *sb = b ? b : a;
  const double Na = a->nao, Nb = sb->nao, Mb = Nb * (Nb + 1) / 2;
  if (ctl->method == 'D')  // two similarity transforms per pair vector, whatever the printed windows say: 8 M N^3 intra (IntTransfD.cpp:145-178)
    return 4.0 * Na * Na * Na * Mb + 4.0 * Nb * Nb * Nb * (Na * (Na + 1) / 2);
  auto cnt = [&](int w) { return (double)std::max(0, win[2 * w + 1] - win[2 * w] + 1); };
  const double nf1 = std::min(cnt(0), cnt(1)), ns1 = std::max(cnt(0), cnt(1));
  const double nf2 = std::min(cnt(2), cnt(3)), ns2 = std::max(cnt(2), cnt(3));
  double npairs = 0;  // first pairs entering the second half: E keeps q<=p (E.f90:878-884), C skips q<p when symmetric (C.f90:380)
  for (int x = win[0]; x <= win[1]; ++x)
    for (int y = win[2]; y <= win[3]; ++y)
      if ((ctl->method == 'E') ? (y <= x) : !(symmetric && y < x)) npairs += 1;
  return 2.0 * Na * nf1 * (Na + ns1) * Mb + 2.0 * Nb * nf2 * (Nb + ns2) * npairs;
}

bool pt2_filter_on(const lowdin_host_control *c) { return is(c, "PT2") && c->n_ionize_species > 0; }
bool named_for_ionization(const lowdin_host_control *c, const lowdin_host_species *a) {
  const std::string na = trimmed(a->name, sizeof a->name);
  for (int s = 0; s < c->n_ionize_species && s < 4; ++s)
    if (na == trimmed(c->ionize_species[s], 32)) return true;
  return false;
}

int plan_program(const lowdin_host_control *ctl, const lowdin_host_species *sp, int n, int nranks, std::vector<lowdin_host_task> &out) {
  if (check_ctl(ctl)) return 1;
  if (!sp || n < 1) return hfail("no species");
  if (nranks < 1) return hfail("nranks < 1");
  out.clear();
  auto add = [&](int first, int second) -> int {
    lowdin_host_task t{};
    t.first = first; t.second = second; t.rank = 0;
    const lowdin_host_species *a = &sp[first], *b = second >= 0 ? &sp[second] : nullptr;
    if (lowdin_host_windows(ctl, a, b, t.win, &t.symmetric)) return 1;
    t.flops = call_flops(ctl, a, b, t.win, t.symmetric);
    out.push_back(t);
    return 0;
  };
  for (int i = 0; i < n; ++i) {
    // IntegralTransformation.f90:176-185
    if (!pt2_filter_on(ctl) || named_for_ionization(ctl, &sp[i])) { if (add(i, -1)) return 1; }
    for (int j = i + 1; j < n; ++j) {
      // :259-269
      if (pt2_filter_on(ctl) && !named_for_ionization(ctl, &sp[i]) && !named_for_ionization(ctl, &sp[j])) continue;
      // :322-334 -- method C calls with the species of fewer occupied orbitals first
      const bool keep = (ctl->method != 'C') || (sp[i].occupation <= sp[j].occupation);
      if (keep ? add(i, j) : add(j, i)) return 1;
    }
  }
  // longest-processing-time-first over the ranks; ties broken by program order and by the lowest rank
  std::vector<int> order(out.size());
  for (size_t k = 0; k < order.size(); ++k) order[k] = (int)k;
  std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return out[x].flops > out[y].flops; });
  std::vector<double> load(nranks, 0.0);
  for (int k : order) {
    int best = 0;
    for (int r = 1; r < nranks; ++r) if (load[r] < load[best]) best = r;
    out[k].rank = best;
    load[best] += out[k].flops;
  }
  return 0;
}

}  // namespace

extern "C" {

const char *lowdin_host_last_error(void) { return g_err.c_str(); }

int lowdin_host_partial_transform(int mp, int pt, int en, int ci_none, char out[16]) {
  const char *r = "BOUNDS";  // IntegralTransformation.f90:106-126
  if (mp == 2 && pt == 0 && en == 0 && ci_none) r = "MP2";
  else if (pt == 2 && mp == 0 && en == 0 && ci_none) r = "PT2";
  else if (pt == 2 && mp == 2 && en == 0 && ci_none) r = "MP2-PT2";
  else if (!ci_none) r = "ALL";
  memset(out, 0, 16);
  strncpy(out, r, 15);
  return 0;
}

int lowdin_host_windows(const lowdin_host_control *ctl, const lowdin_host_species *a, const lowdin_host_species *b, int win[8],
                        int *symmetric) {
  if (check_ctl(ctl)) return 1;
  if (!a || !win) return hfail("null species / window output");
  int sym = 1;
  Win w;
  if (ctl->method == 'C') w = b ? win_c_inter(ctl, a, b, &sym) : win_c_intra(ctl, a, &sym);
  else if (ctl->method == 'D') { w = win_d(ctl, a, b); sym = 0; }
  else { w = b ? win_e_inter(ctl, a, b) : win_e_intra(ctl, a); sym = 0; }
  put(w, win);
  if (symmetric) *symmetric = sym;
  return 0;
}

int lowdin_host_ints_filename(int tid, const lowdin_host_species *a, const lowdin_host_species *b, char out[256], int *swapped) {
  if (!a || !out) return hfail("null argument");
  const std::string na = trimmed(a->name, 32);
  std::string file;
  int sw = 0;
  if (!b) {
    file = (na == "E-BETA") ? "E-ALPHA" : na;  // C.f90:241-245
  } else {
    const std::string nb = trimmed(b->name, 32);
    // the stream on disk was written for the pair in molecular-system order (lower id first); E-BETA reads E-ALPHA's
    const bool forward = a->id < b->id;  // C.f90:838 / :906
    const std::string &first = forward ? na : nb, &second = forward ? nb : na;
    sw = forward ? 0 : 1;
    if (first == "E-ALPHA" && second == "E-BETA") file = "E-ALPHA.E-BETA";
    else if (second == "E-BETA") file = first + ".E-ALPHA";
    else if (first == "E-BETA") file = "E-ALPHA." + second;
    else file = first + "." + second;
  }
  snprintf(out, 256, "%d%s.ints", tid, file.c_str());
  if (swapped) *swapped = sw;
  return 0;
}

int lowdin_host_write_ints_file(const char *path, int S, const int32_t *p, const int32_t *q, const int32_t *r, const int32_t *s,
                                const double *v, int64_t n) {
  if (S < 1) return hfail("stack size < 1");
  FILE *f = fopen(path, "wb");
  if (!f) return hfail(std::string("cannot create ") + path);
  std::vector<int32_t> bp(S), bq(S), br(S), bs(S);
  std::vector<double> bv(S);
  int64_t done = 0;
  bool terminated = false;
  while (!terminated) {  // Libint2Iface.cpp:3414-3426: full blocks, the last one carrying p = -1 after its last entry
    const int64_t m = std::min<int64_t>(S, n - done);
    std::fill(bp.begin(), bp.end(), 0); std::fill(bq.begin(), bq.end(), 0); std::fill(br.begin(), br.end(), 0);
    std::fill(bs.begin(), bs.end(), 0); std::fill(bv.begin(), bv.end(), 0.0);
    for (int64_t i = 0; i < m; ++i) { bp[i] = p[done + i]; bq[i] = q[done + i]; br[i] = r[done + i]; bs[i] = s[done + i]; bv[i] = v[done + i]; }
    if (m < S) { bp[m] = -1; terminated = true; }
    done += m;
    if (fwrite(bp.data(), 4, S, f) != (size_t)S || fwrite(bq.data(), 4, S, f) != (size_t)S || fwrite(br.data(), 4, S, f) != (size_t)S ||
        fwrite(bs.data(), 4, S, f) != (size_t)S || fwrite(bv.data(), 8, S, f) != (size_t)S) { fclose(f); return hfail("short write"); }
  }
  fclose(f);
  return 0;
}

int lowdin_host_read_ints_file(const char *path, int S, int32_t *p, int32_t *q, int32_t *r, int32_t *s, double *v, int64_t cap,
                               int64_t *n) {
  if (S < 1) return hfail("stack size < 1");
  int64_t cnt = 0;
  int rc = for_each_stack(path, S, [&](const int32_t *bp, const int32_t *bq, const int32_t *br, const int32_t *bs, const double *bv, int m) {
    for (int i = 0; i < m; ++i) {
      if (bp[i] == -1) break;
      if (cnt < cap) { p[cnt] = bp[i]; q[cnt] = bq[i]; r[cnt] = br[i]; s[cnt] = bs[i]; v[cnt] = bv[i]; }
      ++cnt;
    }
    return 0;
  });
  if (n) *n = cnt;
  return rc;
}

int lowdin_host_write_moint_quads(const char *path, int S, const int32_t *p, const int32_t *q, const int32_t *r, const int32_t *s,
                                  const double *v, int64_t n) {
  if (S < 1) return hfail("stack size < 1");
  FILE *f = fopen(path, "wb");
  if (!f) return hfail(std::string("cannot create ") + path);
  std::vector<int32_t> bp(S, 0), bq(S, 0), br(S, 0), bs(S, 0);
  std::vector<double> bv(S, 0.0);
  int m = 0, rc = 0;
  for (int64_t k = 0; k < n && !rc; ++k) {  // C.f90:419-440
    bp[m] = p[k]; bq[m] = q[k]; br[m] = r[k]; bs[m] = s[k]; bv[m] = v[k];
    if (++m == S) {
      rc = write_record(f, bp.data(), 4 * (size_t)S, bq.data(), 4 * (size_t)S, br.data(), 4 * (size_t)S, bs.data(), 4 * (size_t)S, bv.data(), 8 * (size_t)S);
      m = 0;
      std::fill(bp.begin(), bp.end(), 0); std::fill(bq.begin(), bq.end(), 0); std::fill(br.begin(), br.end(), 0);
      std::fill(bs.begin(), bs.end(), 0); std::fill(bv.begin(), bv.end(), 0.0);
    }
  }
  bp[m] = -1;  // C.f90:450-456: the terminator record is always written
  if (!rc) rc = write_record(f, bp.data(), 4 * (size_t)S, bq.data(), 4 * (size_t)S, br.data(), 4 * (size_t)S, bs.data(), 4 * (size_t)S, bv.data(), 8 * (size_t)S);
  fclose(f);
  return rc ? hfail("short write") : 0;
}

int lowdin_host_write_moint_pairs(const char *path, int S, const int64_t *ij, const int64_t *kl, const double *v, int64_t n) {
  if (S < 1) return hfail("stack size < 1");
  FILE *f = fopen(path, "wb");
  if (!f) return hfail(std::string("cannot create ") + path);
  std::vector<int64_t> bi(S, 0), bk(S, 0);
  std::vector<double> bv(S, 0.0);
  int m = 0, rc = 0;
  for (int64_t k = 0; k < n && !rc; ++k) {  // E.f90:1244-1257
    bi[m] = ij[k]; bk[m] = kl[k]; bv[m] = v[k];
    if (++m == S) {
      rc = write_record(f, bi.data(), 8 * (size_t)S, bk.data(), 8 * (size_t)S, bv.data(), 8 * (size_t)S, nullptr, 0, nullptr, 0);
      m = 0;
      std::fill(bi.begin(), bi.end(), 0); std::fill(bk.begin(), bk.end(), 0); std::fill(bv.begin(), bv.end(), 0.0);
    }
  }
  bi[m] = -1;  // E.f90:1263-1268
  if (!rc) rc = write_record(f, bi.data(), 8 * (size_t)S, bk.data(), 8 * (size_t)S, bv.data(), 8 * (size_t)S, nullptr, 0, nullptr, 0);
  fclose(f);
  return rc ? hfail("short write") : 0;
}

int lowdin_host_atomic_to_molecular_one_species(lowdin_it_handle h, const lowdin_host_control *ctl, const lowdin_host_species *a,
                                                int64_t *nonzero) {
  return run_and_write(&h, 1, ctl, a, nullptr, nonzero);
}

int lowdin_host_atomic_to_molecular_two_species(lowdin_it_handle h, const lowdin_host_control *ctl, const lowdin_host_species *a,
                                                const lowdin_host_species *b, int64_t *nonzero) {
  if (!b) return hfail("second species missing");
  return run_and_write(&h, 1, ctl, a, b, nonzero);
}

// The same calls for a ONE-PROCESS host driving several GPUs: `handles` are the members of an in-process group
// (lowdin_it_comm_init_local).  The .ints streams are read once and pushed to every handle (each keeps the rows of the AO tensor
// it owns), the transform is collective, and ONE moint.dat is written with the entries in the single-GPU order.
int lowdin_host_group_atomic_to_molecular(lowdin_it_handle *handles, int nhandles, const lowdin_host_control *ctl,
                                          const lowdin_host_species *a, const lowdin_host_species *b, int64_t *nonzero) {
  return run_and_write(handles, nhandles, ctl, a, b, nonzero);
}

// ---- row f4: the integrals program's stream files from integrals evaluated on the device ------------------------
// The entries LibintInterface::compute_2body_disk (Libint2Iface.cpp:219-416; p >= q, r >= s, (p,q) >= (r,s), 1-based) or
// ::compute_coupling_disk (:930-1110; p <= q, r <= s) write -- the same set, in tensor order instead of shell-quartet order --
// spread over ctl->nfiles stream files as the reference's threads spread theirs.
int lowdin_host_write_computed_ints(lowdin_it_handle h, const lowdin_host_control *ctl, const lowdin_host_species *a,
                                    const lowdin_host_species *b, int slot_a, int slot_b, int64_t *nonzero) {
  if (check_ctl(ctl)) return 1;
  if (!h || !a) return hfail("null handle / species");
  const bool intra = (b == nullptr);
  if (intra) slot_b = slot_a;
  if (lowdin_it_ao_compute(h, slot_a, slot_b)) return hfail(lowdin_it_last_error(h));
  const int64_t na = a->nao, nb = intra ? a->nao : b->nao;
  const int64_t Ma = na * (na + 1) / 2, Mb = nb * (nb + 1) / 2;
  const int64_t count = intra ? Ma * (Ma + 1) / 2 : Ma * Mb;
  std::vector<double> t((size_t)count);
  if (lowdin_it_ao_download(h, slot_a, slot_b, t.data(), count)) return hfail(lowdin_it_last_error(h));
  auto pair_table = [](int64_t n, std::vector<int32_t> &lo, std::vector<int32_t> &hi) {
    for (int64_t i = 0; i < n; ++i) for (int64_t j = i; j < n; ++j) { lo.push_back((int32_t)i + 1); hi.push_back((int32_t)j + 1); }
  };
  std::vector<int32_t> alo, ahi, blo, bhi;
  pair_table(na, alo, ahi);
  if (!intra) pair_table(nb, blo, bhi);
  std::vector<int32_t> p, q, r, s;
  std::vector<double> v;
  if (intra) {
    for (int64_t lo = 0; lo < Ma; ++lo)
      for (int64_t hi = lo; hi < Ma; ++hi) {
        const double x = t[(size_t)(lo * Ma - lo * (lo + 1) / 2 + hi)];
        if (x == 0.0) continue;  // dropped on the device by the reference's raw-value filter
        int32_t pp = ahi[hi], qq = alo[hi], rr = ahi[lo], ss = alo[lo];  // p >= q, r >= s
        if (pp < rr || (pp == rr && qq < ss)) { std::swap(pp, rr); std::swap(qq, ss); }
        p.push_back(pp); q.push_back(qq); r.push_back(rr); s.push_back(ss); v.push_back(x);
      }
  } else {
    for (int64_t rs = 0; rs < Mb; ++rs)
      for (int64_t pq = 0; pq < Ma; ++pq) {
        const double x = t[(size_t)(rs * Ma + pq)];
        if (x == 0.0) continue;
        p.push_back(alo[pq]); q.push_back(ahi[pq]); r.push_back(blo[rs]); s.push_back(bhi[rs]); v.push_back(x);
      }
  }
  const int64_t n = (int64_t)v.size();
  for (int tid = 0; tid < ctl->nfiles; ++tid) {
    char name[256];
    int swapped = 0;
    if (lowdin_host_ints_filename(tid, a, b, name, &swapped)) return 1;
    if (swapped) return hfail("write_computed_ints: give the pair in file order (lower species id first)");
    const int64_t lo = n * tid / ctl->nfiles, hi = n * (tid + 1) / ctl->nfiles;
    if (lowdin_host_write_ints_file(join(ctl->scratch_dir, name).c_str(), ctl->integral_stack_size, p.data() + lo, q.data() + lo,
                                    r.data() + lo, s.data() + lo, v.data() + lo, hi - lo))
      return 1;
  }
  if (nonzero) *nonzero = n;
  return 0;
}

// ---- transformer-D record layout ------------------------------------------------------------------------------
namespace {
inline int64_t index2(int64_t i, int64_t j) { return i > j ? i * (i + 1) / 2 + j : j * (j + 1) / 2 + i; }  // ReadIntegrals.f90:175-186
int write_d_record(FILE *f, int32_t p, int32_t q, int32_t r, int32_t s, double v) {
  unsigned char rec[4 + 24 + 4];
  const uint32_t len = 24;
  memcpy(rec, &len, 4); memcpy(rec + 4, &p, 4); memcpy(rec + 8, &q, 4); memcpy(rec + 12, &r, 4); memcpy(rec + 16, &s, 4);
  memcpy(rec + 20, &v, 8); memcpy(rec + 28, &len, 4);
  return fwrite(rec, 1, sizeof rec, f) != sizeof rec;
}
}  // namespace

int lowdin_host_write_moint_d_intra(const char *path, int nao, const double *ints, int64_t *nrecords) {
  if (!path || !ints || nao < 1) return hfail("bad arguments");
  FILE *f = fopen(path, "wb");
  if (!f) return hfail(std::string("cannot create ") + path);
  int64_t cnt = 0;
  int rc = 0;
  for (int p = 1; p <= nao && !rc; ++p)  // TransformIntegralsD.f90:221-236
    for (int q = 1; q <= p && !rc; ++q)
      for (int r = 1; r <= p && !rc; ++r) {
        const int smax = (p == r) ? q : r;
        for (int s_ = 1; s_ <= smax && !rc; ++s_) {
          rc = write_d_record(f, p, q, r, s_, ints[index2(index2(p - 1, q - 1), index2(r - 1, s_ - 1))]);  // ReadIntegrals_index4Intra - 1
          ++cnt;
        }
      }
  if (!rc) rc = write_d_record(f, -1, 0, 0, 0, 0.0);  // :266
  fclose(f);
  if (nrecords) *nrecords = cnt;
  return rc ? hfail("short write") : 0;
}

int lowdin_host_write_moint_d_inter(const char *path, int nao, int onao, const double *ints, int64_t *nrecords) {
  if (!path || !ints || nao < 1 || onao < 1) return hfail("bad arguments");
  FILE *f = fopen(path, "wb");
  if (!f) return hfail(std::string("cannot create ") + path);
  const int64_t osze = (int64_t)onao * (onao + 1) / 2;
  int64_t cnt = 0;
  int rc = 0;
  for (int p = 1; p <= nao && !rc; ++p)  // TransformIntegralsD.f90:447-457
    for (int q = p; q <= nao && !rc; ++q)
      for (int r = 1; r <= onao && !rc; ++r)
        for (int s_ = r; s_ <= onao && !rc; ++s_) {
          rc = write_d_record(f, p, q, r, s_, ints[index2(p - 1, q - 1) * osze + index2(r - 1, s_ - 1)]);  // ReadIntegrals_index4Inter - 1
          ++cnt;
        }
  if (!rc) rc = write_d_record(f, -1, 0, 0, 0, 0.0);  // :484
  fclose(f);
  if (nrecords) *nrecords = cnt;
  return rc ? hfail("short write") : 0;
}

// ---- lowdin.wfn labelled records ------------------------------------------------------------------------------
namespace {
// One sequential unformatted record (4-byte gfortran markers).  Returns 0 ok, -1 end of file, 1 malformed.
int read_record(FILE *f, std::vector<unsigned char> &buf) {
  uint32_t len = 0, tail = 0;
  if (fread(&len, 4, 1, f) != 1) return -1;
  if (len > (1u << 31)) return 1;  // split (negative-marker) records: not produced below 2 GiB
  buf.resize(len);
  if (len && fread(buf.data(), 1, len, f) != len) return 1;
  if (fread(&tail, 4, 1, f) != 1 || tail != len) return 1;
  return 0;
}
std::string rtrim(const unsigned char *p, size_t n) {
  while (n > 0 && (p[n - 1] == ' ' || p[n - 1] == 0)) --n;
  return std::string((const char *)p, n);
}
}  // namespace

int lowdin_host_wfn_read(const char *path, const char *label, const char *species, double *out, int64_t cap, int64_t *n) {
  if (!path || !label || !species) return hfail("bad arguments");
  FILE *f = fopen(path, "rb");
  if (!f) return hfail(std::string("cannot open ") + path);
  const std::string a1 = rtrim((const unsigned char *)label, strlen(label)), a2 = rtrim((const unsigned char *)species, strlen(species));
  std::vector<unsigned char> rec;
  int rc = 0;
  bool found = false;
  while (!found) {  // Matrix.f90:712-748
    long at = ftell(f);
    rc = read_record(f, rec);
    if (rc == -1) { fclose(f); return hfail("End of file! " + a1 + " " + a2 + " not in " + path); }
    if (rc) { fclose(f); return hfail(std::string("malformed record in ") + path); }
    // read(unit) line(1:len_trim(arguments(1))): the first characters of the record
    if (rtrim(rec.data(), std::min(rec.size(), a1.size())) != a1) continue;
    fseek(f, at, SEEK_SET);  // backspace
    for (int i = 0; i < 2; ++i) {  // every argument is read again in full; the LAST comparison decides (Matrix.f90:730-742)
      rc = read_record(f, rec);
      if (rc) { fclose(f); return hfail(rc == -1 ? "End of file! " + a1 + " " + a2 + " not in " + path : std::string("malformed record in ") + path); }
      found = (rtrim(rec.data(), rec.size()) == (i == 0 ? a1 : a2));
    }
  }
  int64_t total = 0;
  rc = read_record(f, rec);
  if (rc || rec.size() != 8) { fclose(f); return hfail("missing element count after " + a1 + " " + a2); }
  memcpy(&total, rec.data(), 8);
  rc = read_record(f, rec);
  fclose(f);
  if (rc || total < 0 || rec.size() != (size_t)total * 8) return hfail("value record of " + a1 + " " + a2 + " does not hold the stored count");
  if (n) *n = total;
  if (out && cap > 0) memcpy(out, rec.data(), (size_t)std::min<int64_t>(cap, total) * 8);
  return 0;
}

int lowdin_host_wfn_append(const char *path, const char *label, const char *species, int label_len, const double *values, int64_t n,
                           int truncate) {
  if (!path || !label || !species || label_len < 1 || n < 0 || (n && !values)) return hfail("bad arguments");
  if ((int)strlen(label) > label_len || (int)strlen(species) > label_len) return hfail("label longer than label_len");
  FILE *f = fopen(path, truncate ? "wb" : "ab");
  if (!f) return hfail(std::string("cannot create ") + path);
  auto put_rec = [&](const void *p, size_t len) {
    const uint32_t m = (uint32_t)len;
    return fwrite(&m, 4, 1, f) != 1 || (len && fwrite(p, 1, len, f) != len) || fwrite(&m, 4, 1, f) != 1;
  };
  std::string l1(label), l2(species);
  l1.resize(label_len, ' '); l2.resize(label_len, ' ');  // write(unit) arguments(m): the full declared length, blank padded
  int rc = put_rec(l1.data(), l1.size()) || put_rec(l2.data(), l2.size()) || put_rec(&n, 8) || put_rec(values, (size_t)n * 8);
  fclose(f);
  return rc ? hfail("short write") : 0;
}

int lowdin_host_wfn_load_species(const char *path, lowdin_host_species *sp, double *coeff, double *eps) {
  if (!sp || !coeff || sp->nao < 1) return hfail("bad arguments");
  const std::string name = trimmed(sp->name, sizeof sp->name);
  const int rows = sp->nao, cols = std::max(sp->nao, sp->occupation);  // IntegralTransformation.f90:200-202
  int64_t n = 0;
  if (lowdin_host_wfn_read(path, "COEFFICIENTS", name.c_str(), coeff, (int64_t)rows * cols, &n)) return 1;
  if (n != (int64_t)rows * cols) return hfail("The dimensions of the matrix COEFFICIENTS " + name + " are wrong");  // Matrix.f90:753, :786
  if (eps) {
    if (lowdin_host_wfn_read(path, "ORBITALS", name.c_str(), eps, rows, &n)) return 1;
    if (n != rows) return hfail("The dimensions of the vector ORBITALS " + name + " are wrong");  // Vector.f90:802, :816
  }
  sp->coeff = coeff; sp->ldc = rows; sp->ncols = cols;  // values(m) fills column by column (Matrix.f90:773-778): already column-major
  return 0;
}

int lowdin_host_plan_program(const lowdin_host_control *ctl, const lowdin_host_species *species, int nspecies, int nranks,
                             lowdin_host_task *tasks, int cap, int *ntasks) {
  std::vector<lowdin_host_task> plan;
  if (plan_program(ctl, species, nspecies, nranks, plan)) return 1;
  if (ntasks) *ntasks = (int)plan.size();
  if (tasks) {
    if ((int)plan.size() > cap) return hfail("task buffer too small");
    std::copy(plan.begin(), plan.end(), tasks);
  }
  return 0;
}

namespace {
// the calls of the plan assigned to `rank` (program order) on one handle, or on the handles of an in-process group (every call collective)
int run_program(lowdin_it_handle *hs, int nh, const lowdin_host_control *ctl, const lowdin_host_species *species, int nspecies, int rank,
                int nranks, int64_t *nonzero, int *ncalls);
}  // namespace

int lowdin_host_run_program(lowdin_it_handle h, const lowdin_host_control *ctl, const lowdin_host_species *species, int nspecies,
                            int rank, int nranks, int64_t *nonzero, int *ncalls) {
  return run_program(&h, 1, ctl, species, nspecies, rank, nranks, nonzero, ncalls);
}

// The program's loop for a ONE-PROCESS host driving several GPUs: every call of the loop, in program order, runs on the whole group
// (the stored AO tensor sharded by rows, the transform collective, one moint.dat per call).
int lowdin_host_group_run_program(lowdin_it_handle *handles, int nhandles, const lowdin_host_control *ctl, const lowdin_host_species *species,
                                  int nspecies, int64_t *nonzero, int *ncalls) {
  if (!handles || nhandles < 1) return hfail("null handles");
  return run_program(handles, nhandles, ctl, species, nspecies, 0, 1, nonzero, ncalls);
}

namespace {
int run_program(lowdin_it_handle *hs, int nh, const lowdin_host_control *ctl, const lowdin_host_species *species, int nspecies, int rank,
                int nranks, int64_t *nonzero, int *ncalls) {
  if (rank < 0 || rank >= nranks) return hfail("bad rank");
  std::vector<lowdin_host_task> plan;
  if (plan_program(ctl, species, nspecies, nranks, plan)) return 1;
  int64_t total = 0;
  int calls = 0;
  for (const lowdin_host_task &t : plan) {
    if (t.rank != rank) continue;
    const lowdin_host_species *a = &species[t.first], *b = t.second >= 0 ? &species[t.second] : nullptr;
    if (ctl->verbose) {  // IntegralTransformation.f90:188-191, :272-276
      const lowdin_host_species *lo = (b && t.second < t.first) ? b : a, *hi = (b && t.second < t.first) ? a : b;  // named in (i, j>i) order
      if (b) printf("\n Inter-species integrals transformation for: %s/%s\n\n", trimmed(lo->name, 32).c_str(), trimmed(hi->name, 32).c_str());
      else printf("\n Integrals transformation for: %s\n\n", trimmed(a->name, 32).c_str());
    }
    int64_t n = 0;
    if (run_and_write(hs, nh, ctl, a, b, &n)) return 1;
    total += n;
    ++calls;
  }
  if (nonzero) *nonzero = total;
  if (ncalls) *ncalls = calls;
  return 0;
}
}  // namespace

}  // extern "C"
