/*
 * pmc_oracle.h -- CPU ORACLE for the ParELAGMC per-sample hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This is a plain-C restatement of the reference algorithm; it is never linked into, imported by or
 * called from the product path (parelagmc_b200/).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference leg may load it.
 *
 * PARITY STATUS: "parity unpinned" for the third-party pieces.  The arithmetic of this path lives in
 * TRNG 4.19, ParELAG 2.0, MFEM and hypre, none of which is vendored under /root/reference and none of
 * which can be built here (no MPI/MFEM/hypre/TRNG).  What IS pinned, in tests/: the reference's own
 * known answers that need no third party -- Darcy Q = 2 and dof counts 17152/2240/304
 * (/root/reference/examples/CMakeLists.txt:62-66), the Matern scaling values of
 * /root/reference/src/Utilities.hpp:188-200, the exact (log-)normal moments of
 * /root/reference/examples/PDESamplerTest.cpp:207-209, the manager formulas of
 * /root/reference/src/MLMC_Manager.cpp:300-401 -- plus an independent sparse direct solve (scipy) of
 * every linear system.  The yarn5 integer stream follows the published algorithm (MRG of order 5 mod
 * 2^31-1 with the LEcuyer1 parameter set and the YARN power map) and is self-checked (jump vs. step,
 * split vs. leapfrog); no known-answer vector of the stream exists in the reference tree.
 */
#ifndef PMC_ORACLE_H
#define PMC_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ----------------------------------------------------------------------------------------------
 * RNG: trng::yarn5 + trng::normal_dist<double>, as wrapped by
 * /root/reference/src/NormalDistributionSampler.cpp:17-37  [TRNG 4.19, restated from the published
 * algorithm; UPSTREAM-UNVERIFIED]
 * -------------------------------------------------------------------------------------------- */
typedef struct {
    int32_t a[5]; /* recurrence coefficients (LEcuyer1 by default) */
    int32_t r[5]; /* state, r[0] most recent */
} po_yarn5;

void po_yarn5_init(po_yarn5 *g);                             /* trng::yarn5()                       */
int32_t po_yarn5_next(po_yarn5 *g);                          /* operator()(): step + power map       */
void po_yarn5_jump(po_yarn5 *g, uint64_t n);                 /* jump(n): advance n steps             */
void po_yarn5_split(po_yarn5 *g, unsigned s, unsigned n);    /* split(s, n): leapfrog sub-stream     */
void po_yarn5_fill_int(po_yarn5 *g, int64_t n, int32_t *out);
double po_uniformoo(int32_t x);                              /* utility::uniformoo<double>           */
double po_inv_Phi(double u);                                 /* math::inv_Phi (Acklam + Halley)      */
double po_inv_Phi_libm(double u);                            /* same, with glibc's erf/erfc/exp/log   */
/* which: 0 exp, 1 log, 2 erf, 3 erfc (parelagmc_b200/csrc/detmath.h), 4 inv_Phi, 5 inv_Phi over libm */
void po_det_eval(int which, int64_t n, const double *x, double *y);
void po_normal_map(int64_t n, const int32_t *engine, double mu, double sigma, double *out);
/* NormalDistributionSampler::operator()(Vector&): out[i] = mu + sigma * inv_Phi(uniformoo(rng())) */
void po_normal_fill(po_yarn5 *g, double mu, double sigma, int64_t n, double *out);

/* ----------------------------------------------------------------------------------------------
 * Problem handle: sampler levels + Darcy levels (the data ParELAG leaves on the host, see
 * parelagmc_b200/hierarchy.py for the schema) and the per-sample path on top of it.
 * -------------------------------------------------------------------------------------------- */
typedef struct po_problem po_problem;

po_problem *po_create(int nlevels);
void po_destroy(po_problem *p);
/* Krylov parameters of "MINRES-BJ-GS" (/root/reference/examples/example_helpers/CreateMLMCParameterList.hpp:58-70) */
void po_set_tolerances(po_problem *p, double rel, double abs_, int maxit);

/* PDESampler::BuildHierarchy output for one level (/root/reference/src/PDESampler.cpp:218-284).
 * M, B already eliminated; Wdiag positive; P = Ps[level] (Ne x Ne_coarse) or NULL on the coarsest. */
int po_set_sampler_level(po_problem *p, int level, int Ne, int Nf,
                         const int *M_rowptr, const int *M_col, const double *M_val,
                         const int *B_rowptr, const int *B_col, const double *B_val,
                         const double *Wdiag,
                         int P_cols, const int *P_rowptr, const int *P_col, const double *P_val,
                         double alpha, double matern_coeff, int lognormal);

/* Optional transfer of the sampled field to the forward problem's mesh, applied before exp:
 * s = row_scale .* (T field).  EmbeddedPDESampler (/root/reference/src/EmbeddedPDESampler.cpp:426-435): T = meshP (0/1
 * selection), row_scale = NULL.  L2ProjectionPDESampler (/root/reference/src/L2ProjectionPDESampler.cpp:595-611):
 * T = G^T, row_scale = 1/diag(W_orig).  T is n_out x Ne(level). */
int po_set_field_transfer(po_problem *p, int level, int n_out, const int *T_rowptr, const int *T_col,
                          const double *T_val, const double *row_scale);

/* DarcySolver level data (/root/reference/src/DarcySolver.cpp:60-414). B un-eliminated. */
int po_set_darcy_level(po_problem *p, int level, int Ne, int Nf,
                       const int *elem_ptr, const int *elem_dofs, const double *elem_mat,
                       const int *B_rowptr, const int *B_col, const double *B_val,
                       const int *ess_u, const double *ess_data, const double *rhs, const double *obs,
                       int Pp_cols, const int *Pp_rowptr, const int *Pp_col, const double *Pp_val);

/* PDESampler::Eval (5-argument form, /root/reference/src/PDESampler.cpp:411-535; use_init < 0 selects
 * the 3-argument form :342-409).  xi has the size of level `xi_level` <= level.  s_out: the (log)normal
 * field at `level`; embed_s: in (coarser Gaussian field of size Ne[init_level]) / out (Gaussian field). */
int po_sampler_eval(po_problem *p, int level, int xi_level, const double *xi, double *s_out,
                    double *embed_s, int init_level, int use_init, int *iters);

/* DarcySolver::SolveFwd (/root/reference/src/DarcySolver.cpp:416-437): assemble M(k), eliminate,
 * solve from zero, Q = obs . sol, C = N.  sol_out (size N) may be NULL. */
int po_darcy_solve(po_problem *p, int level, const double *k, double *Q, double *C, double *sol_out,
                   int *iters);

/* One level of MLMC_Manager::InitRun (/root/reference/src/MLMC_Manager.cpp:110-175): nsamples
 * realisations starting at absolute stream position pos0 (stride 1).  sums[9] in the order of the
 * enum at /root/reference/src/MLMC_Manager.hpp:65 {Y2,Y,ABSY,Q2,Q,ABSQ,C,Y3,Y4} is ACCUMULATED into.
 * rows (nsamples x 4: Y,Q,Qc,C) may be NULL.  nthreads > 1 spreads samples over OpenMP threads (each
 * jumps to its sample's stream position); accumulation stays in sample order. */
int po_mlmc_level(po_problem *p, int level, int nlevels, int nsamples, uint64_t pos0,
                  double mu, double sigma, double *sums, double *rows, int nthreads,
                  int64_t *total_iters);

/* BayesianInverseProblem (/root/reference/src/BayesianInverseProblem.cpp:26-218): m pressure functionals g [m][Ne]
 * (un-normalised), observed data, noise variance; and one level of ML_BayesRatio_Manager::InitRun
 * (/root/reference/src/ML_BayesRatio_Manager.hpp:313-424): sums[20] in the manager's enum layout is ACCUMULATED into,
 * rows (nsamples x 5: R, Y_R, Z, Y_Z, c) may be NULL. */
int po_set_observations(po_problem *p, int level, int m, const double *g, const double *G_obs, double noise);
int po_bayes_level(po_problem *p, int level, int nlevels, int nsamples, uint64_t pos0, double mu, double sigma,
                   double *sums, double *rows, int nthreads);

/* ----------------------------------------------------------------------------------------------
 * Manager statistics
 * -------------------------------------------------------------------------------------------- */
/* expWRegression (/root/reference/src/Utilities.cpp:257-283) */
double po_exp_w_regression(const double *y, const double *x, int n, int skip_n_last);

typedef struct {
    double estimate, ml_estimator_variance, bias2, actual_mse, eps2;
    double alpha, alpha_abs, beta, gamma;
} po_mlmc_stats;
/* MLMC_Manager::computeNSamplesMSE (/root/reference/src/MLMC_Manager.cpp:300-401).
 * sums: nlevels x 9 row-major; M: dofs per level; cost: per-level cost (wall time or eC) or NULL for eC.
 * Outputs per level arrays (each nlevels): eY,eABSY,eQ,eABSQ,eC,varY,varQ,consistency,kurtosis,VC,
 * and missing[] = level_nsamples_missing. */
void po_mlmc_compute(int nlevels, const double *sums, const int *nsamples, const double *M,
                     const double *cost, double eps2_in, double ratio,
                     double *eY, double *eABSY, double *eQ, double *eABSQ, double *eC, double *varY,
                     double *varQ, double *consistency, double *kurtosis, double *VC, int *missing,
                     po_mlmc_stats *out);
/* MC_Manager::computeNSamplesMSE (/root/reference/src/MC_Manager.cpp:194-239); sums[4] = {Q2,Q,ABSQ,C} */
void po_mc_compute(const double *sums, int nsamples, const double *cost, double eps2_in, double ratio,
                   double *eQ, double *eABSQ, double *eC, double *varQ, int *missing, po_mlmc_stats *out);

#ifdef __cplusplus
}
#endif
#endif
