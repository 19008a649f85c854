/*
 * lowdin_it_host.h -- host-side mirror of the reference's transformer interface (methods C and E) above the
 * C ABI of lowdin_it.h.  The reference host is Fortran (src/integralsTransformation/IntegralTransformation.f90);
 * there is no Fortran compiler in this build environment, so the same per-species calls, window tables, file names
 * and record layouts are provided in C++ (openlowdin_b200/csrc/host_mirror.cpp) with a C interface:
 *
 *   TransformIntegralsC_atomicToMolecularOfOneSpecie / OfTwoSpecies  (TransformIntegralsC.f90:141, :728)
 *   TransformIntegralsE_atomicToMolecularOfOneSpecie / OfTwoSpecies  (TransformIntegralsE.f90:153, :1285)
 *        -> lowdin_host_atomic_to_molecular_one_species / _two_species
 *   TransformIntegralsC_checkMOIntegralType / checkInterMOIntegralType   (C.f90:1436-1963)
 *   TransformIntegralsE_checkMOIntegralType / checkInterMOIntegralType   (E.f90:1899-2418)
 *        -> lowdin_host_windows
 *   partialTransform choice (IntegralTransformation.f90:106-126)            -> lowdin_host_partial_transform
 *   <tid><name>.ints stream files (Libint2Iface.cpp:3414-3426; readers C.f90:231-298, :852-972)
 *        -> lowdin_host_ints_filename, lowdin_host_read_ints_file, lowdin_host_write_ints_file
 *   <prefix>moint.dat sequential unformatted records (C.f90:419-456, E.f90:1244-1268)
 *        -> lowdin_host_write_moint_quads / _pairs
 *   transformer D's one-integral-per-record moint.dat (TransformIntegralsD.f90:221-236, :447-457)
 *        -> lowdin_host_write_moint_d_intra / _inter
 *   lowdin.wfn labelled records (Matrix.f90:697-792, Vector.f90:738-821; IntegralTransformation.f90:193-215)
 *        -> lowdin_host_wfn_read, lowdin_host_wfn_append, lowdin_host_wfn_load_species
 *   program IntegralsTransformation, species / species-pair loop (IntegralTransformation.f90:171-355)
 *        -> lowdin_host_plan_program, lowdin_host_run_program
 *
 * Only the functions that take a lowdin_it_handle need a GPU.  Every function returns 0 on success; on error the
 * message is available through lowdin_host_last_error().
 */
#ifndef LOWDIN_IT_HOST_H
#define LOWDIN_IT_HOST_H

#include <stdint.h>
#include "lowdin_it.h"

#ifdef __cplusplus
extern "C" {
#endif

/* The CONTROL_instance / InputCI fields this path reads. */
typedef struct lowdin_host_control {
  char method;                  /* 'C' or 'E': whose window roles, skip rules and record layout (CONTROL.f90:354) */
  char partial_transform[16];   /* "MP2", "PT2", "MP2-PT2", "ALL", "ALLACTIVE", "BOUNDS" (IntegralTransformation.f90:106-126) */
  int integral_stack_size;      /* INTEGRAL_STACK_SIZE, default 30000 (CONTROL.f90:1106) */
  int nfiles;                   /* number of <tid>*.ints files = OMP threads of lowdin-ints.x (Libint2Iface.cpp:282-286) */
  int ionize_mo;                /* IONIZE_MO(1); 0 = not set */
  int pt_transition_operator;   /* PT_TRANSITION_OPERATOR */
  int n_ionize_species;         /* 0 = IONIZE_SPECIES(1) is "NONE" */
  char ionize_species[4][32];
  char scratch_dir[512];        /* directory of the .ints / moint.dat files ("" = cwd) */
  int verbose;                  /* print the reference's phase lines */
} lowdin_host_control;

/* What the transformer reads of one quantum species (MolecularSystem_*, InputCI_Instance). */
typedef struct lowdin_host_species {
  char name[32];          /* "E-", "E-ALPHA", "E-BETA", "H_1", "POSITRON", ... */
  int id;                 /* species index in the molecular system (1-based, decides the A.B file order: C.f90:838) */
  int nao;                /* MolecularSystem_getTotalNumberOfContractions */
  int occupation;         /* MolecularSystem_getOcupationNumber */
  int core_orbitals;      /* InputCI_Instance%coreOrbitals (0 = none) */
  int active_orbitals;    /* InputCI_Instance%activeOrbitals (0 = all) */
  const double *coeff;    /* column-major C(mu,p), leading dimension ldc, ncols columns */
  int ldc, ncols;
} lowdin_host_species;

const char *lowdin_host_last_error(void);

/* IntegralTransformation.f90:106-126. ci_level_is_none != 0 when CONFIGURATION_INTERACTION_LEVEL == "NONE". */
int lowdin_host_partial_transform(int moller_plesset_correction, int pt_order, int epstein_nesbet_correction,
                                  int ci_level_is_none, char out[16]);

/* Window table of the selected method.  b == NULL: intra-species.  win = {p_l,p_u,q_l,q_u,r_l,r_u,s_l,s_u}. */
int lowdin_host_windows(const lowdin_host_control *ctl, const lowdin_host_species *a, const lowdin_host_species *b,
                        int win[8], int *symmetric);

/* File name (without directory) of thread `tid`'s AO stream for species a (b == NULL) or the pair (a,b) in CALL order,
 * including the E-BETA -> E-ALPHA aliasing (C.f90:241-245, :838-848, :906-915).  *swapped = 1 when the file holds (b b|a a). */
int lowdin_host_ints_filename(int tid, const lowdin_host_species *a, const lowdin_host_species *b, char out[256], int *swapped);

/* One .ints stream file: blocks of int32 p[S],q[S],r[S],s[S], double v[S]; terminator p=-1 in the last block. */
int lowdin_host_write_ints_file(const char *path, int stack_size, const int32_t *p, const int32_t *q, const int32_t *r,
                                const int32_t *s, const double *v, int64_t n);
/* Reads up to cap entries (stops at the terminator); *n = number of entries in the file. */
int lowdin_host_read_ints_file(const char *path, int stack_size, int32_t *p, int32_t *q, int32_t *r, int32_t *s, double *v,
                               int64_t cap, int64_t *n);

/* moint.dat writers: Fortran sequential unformatted records with 4-byte gfortran markers.
 * quads: record = int32 pp[S],qq[S],rr[S],ss[S], real64 v[S]; last record has pp(m+1) = -1   (C.f90:419-456)
 * pairs: record = int64 ij[S],kl[S], real64 v[S];            last record has ij(m+1) = -1   (E.f90:1244-1268) */
int lowdin_host_write_moint_quads(const char *path, int stack_size, const int32_t *p, const int32_t *q, const int32_t *r,
                                  const int32_t *s, const double *v, int64_t n);
int lowdin_host_write_moint_pairs(const char *path, int stack_size, const int64_t *ij, const int64_t *kl, const double *v, int64_t n);

/* Transformer-D record layout (TransformIntegralsD.f90:221-236 intra, :447-457 inter, terminator :266 / :484; reader
 * ReadTransformedIntegrals.f90:255-263): ONE Fortran record per integral, int32 p,q,r,s + real64 value, last record
 * (-1,0,0,0,0.0).  `ints` is the in-place result of lowdin_it_transform_all / _inter_all (D packing).
 * intra order: p, q<=p, r<=p, s<=(r, or q when r==p); inter order: p, q>=p, r, s>=r. */
int lowdin_host_write_moint_d_intra(const char *path, int nao, const double *ints, int64_t *nrecords);
int lowdin_host_write_moint_d_inter(const char *path, int nao, int onao, const double *ints, int64_t *nrecords);

/* lowdin.wfn: the labelled records the transformation program reads its coefficients from (IntegralTransformation.f90:193-215;
 * Matrix_getFromFile, Matrix.f90:697-792; Vector_getFromFile, Vector.f90:738-821; written by MultiSCF_saveWfn,
 * MultiSCF.f90:1346-1389 with character(30) labels).  A block is four sequential unformatted records:
 * label, species name, int64 element count, real64 values.  lowdin_host_wfn_read scans from the start of the file exactly as the
 * reference does (first record that STARTS with `label`, accepted when the record after it equals `species`) and copies up
 * to cap values; *n = the count stored in the file.  lowdin_host_wfn_append writes one block (label_len = declared length of the
 * writer's labels, 30 in MultiSCF_saveWfn); truncate != 0 starts a new file. */
int lowdin_host_wfn_read(const char *path, const char *label, const char *species, double *out, int64_t cap, int64_t *n);
int lowdin_host_wfn_append(const char *path, const char *label, const char *species, int label_len, const double *values, int64_t n,
                           int truncate);
/* What the program loads per species (IntegralTransformation.f90:193-215): COEFFICIENTS as rows = nao, columns =
 * max(nao, occupation) -- a stored count that differs is the reference's "dimensions of the matrix ... are wrong" error --
 * and ORBITALS (nao eigenvalues; eps may be NULL).  sp->name, nao, occupation must be set; sets sp->coeff/ldc/ncols. */
int lowdin_host_wfn_load_species(const char *path, lowdin_host_species *sp, double *coeff, double *eps);

/* The transformer calls.  Read <tid><name>.ints (or <tid><A>.<B>.ints) from ctl->scratch_dir, upload, transform on the
 * GPU with the method's semantics and write <name>moint.dat (or <A>.<B>moint.dat).  For method C the caller applies the
 * reference's pair ordering (species with fewer occupied orbitals first, IntegralTransformation.f90:322-334) by the order
 * in which it passes a and b.  *nonzero = number of integrals written ("Non-zero transformed integrals"). */
int lowdin_host_atomic_to_molecular_one_species(lowdin_it_handle h, const lowdin_host_control *ctl, const lowdin_host_species *a,
                                                int64_t *nonzero);
int lowdin_host_atomic_to_molecular_two_species(lowdin_it_handle h, const lowdin_host_control *ctl, const lowdin_host_species *a,
                                                const lowdin_host_species *b, int64_t *nonzero);
/* The same for a ONE-PROCESS host driving several GPUs (the reference's transformation program is one process): `handles` are
 * the members of an in-process group (lowdin_it_comm_init_local).  The .ints streams are read once and pushed to every handle
 * (each keeps the rows of the AO tensor it owns), the transform is collective, ONE moint.dat is written with the entries in the
 * single-GPU order.  b == NULL: one species. */
int lowdin_host_group_atomic_to_molecular(lowdin_it_handle *handles, int nhandles, const lowdin_host_control *ctl,
                                          const lowdin_host_species *a, const lowdin_host_species *b, int64_t *nonzero);

/* Row f4: the integrals program's stream files (<tid><name>.ints, ctl->nfiles of them, in ctl->scratch_dir) from integrals
 * EVALUATED ON THE DEVICE for the basis given to lowdin_it_set_basis(h, slot_a / slot_b): the entries compute_2body_disk
 * (Libint2Iface.cpp:219-416; b == NULL) or compute_coupling_disk (:930-1110) write, in tensor order.  The AO set of the pair stays
 * resident, so a transform can follow without reading the files back.  *nonzero = entries written. */
int lowdin_host_write_computed_ints(lowdin_it_handle h, const lowdin_host_control *ctl, const lowdin_host_species *a,
                                    const lowdin_host_species *b, int slot_a, int slot_b, int64_t *nonzero);

/* ---- the program's species loop (IntegralTransformation.f90:171-355) and its division among devices ------------
 * One entry per transformer call the reference program would make, in program order: species i (skipped under PT2 when
 * IONIZE_SPECIES is set and does not name it, :176-185), then every pair (i, j>i) (skipped under PT2 unless one of the
 * two is named, :259-269).  Method C passes the species with FEWER occupied orbitals first (:322-334); E keeps (i,j).
 * Calls are independent (each reads its own .ints streams and writes its own moint.dat), so with several processes,
 * one per GPU, whole calls are assigned to ranks -- no collective (SURVEY.md 8e "species pairs are scheduled across
 * devices").  The assignment is longest-processing-time-first on the algorithmic flop count of each call and is a pure
 * function of the inputs: every rank computes the same plan without communicating. */
typedef struct lowdin_host_task {
  int first, second;  /* indices into the species array, in CALL order; second = -1 for a one-species call */
  int win[8];         /* window table of the call (lowdin_host_windows) */
  int symmetric;
  double flops;       /* algorithmic flops: 2 N nf (N+ns) n_slabs + 2 N' nf' (N'+ns') n_pairs */
  int rank;           /* process / device that runs the call */
} lowdin_host_task;

int lowdin_host_plan_program(const lowdin_host_control *ctl, const lowdin_host_species *species, int nspecies, int nranks,
                             lowdin_host_task *tasks, int cap, int *ntasks);
/* Runs the calls of the plan assigned to `rank` (program order) on handle h.  *nonzero = integrals written by this rank,
 * *ncalls = calls it made. */
int lowdin_host_run_program(lowdin_it_handle h, const lowdin_host_control *ctl, const lowdin_host_species *species, int nspecies,
                            int rank, int nranks, int64_t *nonzero, int *ncalls);
/* The same loop for a ONE-PROCESS host driving several GPUs (the reference's program is one process): `handles` are the members of
 * an in-process group (lowdin_it_comm_init_local); every call of the loop, in program order, runs on the whole group and writes
 * its one moint.dat. */
int lowdin_host_group_run_program(lowdin_it_handle *handles, int nhandles, const lowdin_host_control *ctl,
                                  const lowdin_host_species *species, int nspecies, int64_t *nonzero, int *ncalls);

#ifdef __cplusplus
}
#endif
#endif /* LOWDIN_IT_HOST_H */
