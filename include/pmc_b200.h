/*
 * pmc_b200.h -- C ABI of the B200-native per-sample hot path of ParELAGMC
 * (SPDE white-noise sampling + mixed-FE saddle-point solves + QoI + per-level moment sums, batched over
 * independent Monte-Carlo realisations).  Implemented in parelagmc_b200/csrc/ as hand-written CUDA for
 * sm_100a; built into parelagmc_b200/lib/libpmc_b200.so.
 *
 * This is the drop-in boundary: these are the entry points the reference's L3 classes would bind to
 * (see INTEGRATION.md for the C++ stubs).  Each entry point cites the reference interface it replaces;
 * paths are relative to the reference tree (LLNL/parelagmc).
 *
 * Conventions
 *   - extern "C", opaque handle, plain pointers and sizes.  Return value: 0 = ok, negative = error
 *     (pmc_last_error() gives the message).  Nothing throws across this boundary.  There is NO CPU
 *     fallback: every entry point that computes fails with PMC_ERR_CUDA when no sm_100 device is usable.
 *   - One host thread per handle; one handle per GPU/process.  All kernels are launched on the handle's
 *     stream (its own, or the one given to pmc_set_stream).  Host pointers are borrowed for the duration
 *     of the call; device memory is owned by the handle.
 *   - Host-side batched vectors use the reference's natural layout: sample-major, one contiguous
 *     mfem::Vector-like array per realisation, i.e. v[sample * n + i].
 *   - CSR arrays are `int rowptr[rows+1]`, `int col[nnz]`, `double val[nnz]` exactly as the host library
 *     (ParELAG/MFEM) produced them.
 *   - Levels: 0 = finest ... nlevels-1 = coarsest, as in the reference.
 *   - The random stream is addressed by ABSOLUTE position (number of engine draws since the default seed),
 *     so any GPU can generate any realisation's noise: realisation j of a level whose noise vectors have
 *     n entries starts at pos0 + j*n (one engine draw per normal, src/NormalDistributionSampler.cpp:31-37).
 */
#ifndef PMC_B200_H
#define PMC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pmc_context_s *pmc_handle;

enum {
    PMC_OK = 0,
    PMC_ERR_ARG = -1,     /* bad argument / level not uploaded                                   */
    PMC_ERR_CUDA = -2,    /* CUDA runtime error or no usable device                              */
    PMC_ERR_STATE = -3,   /* call order violated (e.g. solve before upload)                      */
    PMC_ERR_NOMEM = -4    /* device or host allocation failed                                    */
};

/* ---- lifetime ---------------------------------------------------------------------------------- */
/* Constructed where the reference constructs PDESampler / DarcySolver (src/PDESampler.hpp:52-55,
 * src/DarcySolver.hpp:40-42); one context serves both, as both share one hierarchy. */
int pmc_create(int device, int nlevels, pmc_handle *out);
void pmc_destroy(pmc_handle h);
/* Message of the last error on this handle (h may be NULL: last error of pmc_create). */
const char *pmc_last_error(pmc_handle h);
/* Optional: run on a caller-owned cudaStream_t instead of the handle's own stream. */
int pmc_set_stream(pmc_handle h, void *cuda_stream);
/* Synchronise the handle's stream. */
int pmc_synchronize(pmc_handle h);

/* Page-locked host buffers (cudaMallocHost) for the batched host-side vectors: with them the host<->device copies of the
 * *_batch entry points run at full PCIe rate and asynchronously; any host pointer is accepted, pinned or not. */
int pmc_host_alloc(size_t bytes, void **out);
void pmc_host_free(void *p);

/* ---- solver configuration ---------------------------------------------------------------------- */
/* Krylov stopping rule of the reference's "MINRES-BJ-GS" entry: relative/absolute tolerance on the
 * preconditioned residual norm and the iteration cap (examples/example_helpers/
 * CreateMLMCParameterList.hpp:58-70; defaults here are the same: 1e-6, 1e-12, 300). */
int pmc_set_tolerances(pmc_handle h, double rel_tol, double abs_tol, int max_iter);
/* Block-diagonal preconditioner shape (replaces hypre L1-GS x3 on M and BoomerAMG on B diag(M)^-1 B^T,
 * CreateMLMCParameterList.hpp:85-118): Chebyshev-Jacobi degree on the RT mass block, Chebyshev degree and
 * eigenvalue ratio of the V-cycle smoother on the pressure Schur complement, and of its coarsest level.
 * Values <= 0 keep the current setting. */
int pmc_set_preconditioner(pmc_handle h, int mass_degree, int schur_degree, double schur_ratio,
                           int coarse_degree, double coarse_ratio);
/* Fine-grained options, to be set before the first solve / pmc_prepare.  Keys "sampler.<k>" / "darcy.<k>" with
 * <k> in {mass_degree, schur_degree, schur_ratio, coarse_degree, coarse_ratio, omega (over-correction factor of the
 * coarse-grid correction), method (sampler only: 0 = MINRES on the saddle system [M B^T; B -alpha W], 1 = Jacobi-preconditioned CG
 * on its SPD form (M + alpha^-1 B^T W^-1 B) u = alpha^-1 B^T W^-1 f, s = (B u - f) / (alpha W) -- the elimination of
 * src/PDESampler_Legacy.cpp:172-176 -- , 2 = Chebyshev semi-iteration on the same SPD form: the operator does not depend
 * on the realisation, so its spectrum is computed once at set-up and the number of steps for the requested residual
 * reduction is known a priori (no dot products; the true residual norm is still checked per realisation after the steps),
 * -1 = the SPD form (Chebyshev) when alpha W dominates the Schur complement, i.e. for short correlation lengths,
 * MINRES otherwise), cheb_lo_scale / cheb_hi_scale (sampler: safety margins on the Ritz estimates of the extreme
 * eigenvalues the Chebyshev interval is built from; defaults 0.96 / 1.01), schur_degree_coarse (smoother degree on the
 * V-levels between the finest and the coarsest; 0 = as schur_degree, -1 = choose), mass_scale (relative scaling of the Jacobi mass block against the Schur block of the MINRES
 * preconditioner; default 1), max_vlevels (depth of the Schur V-cycle; 0 = full hierarchy, -1 = decide from the mass
 * term, sampler only), amg (coarse spaces of the Schur V-cycle: 0 = the hierarchy's L2 prolongators, 1 = strength-aware
 * pairwise aggregation built at set-up, -1 = aggregation only when the couplings are anisotropic), amg_passes (pairwise
 * matching passes per level of that aggregation: aggregates of up to 2^passes rows; default 3), amg_smooth (damping of the
 * Jacobi step that smooths its tentative prolongators at set-up -- smoothed aggregation on the strength-filtered operator at
 * k = 1; 0 = plain aggregation; default 0.9)}; and the launch
 * shape of the solver kernel (all optional, the library chooses): "max_batch", "cta_threads" (64/128/256/512),
 * "cluster_size" (1/2/4/8 CTAs of a thread-block cluster per tile of 4 realisations), "group_size" (G > 1: G co-resident
 * CTAs of a cooperative launch per tile, for one or two tiles of very large levels; -1 = never), "solo_rows" (in a group,
 * operations with at most this many rows run on its first CTA), "stage_operators" (0 = read operator entries from L2
 * instead of staging them through shared memory with TMA bulk copies), "defer_x" (0 = update the MINRES solution
 * every iteration instead of once per iteration pair), "qoi_only" (0 = Darcy solves form the solution vector even when the
 * caller reads Q only; by default such solves carry Q = obs . x by scalar recurrences of the MINRES coefficients),
 * "cheb_three_term" (0 = keep a separate update vector in the sampler's
 * Chebyshev steps instead of reading the previous iterate from the buffer a step overwrites), "single_wave" (1 = prefer one wave of smaller CTAs), and
 * "renumber" (0 = keep the caller's numbering of the RT dofs inside the library; to be set before the uploads), and
 * "cache_results" (see pmc_sampler_eval_batch). */
int pmc_set_option(pmc_handle h, const char *key, double value);
/* Largest number of realisations processed per kernel launch (0 = choose from free device memory), and
 * how many MINRES iterations are queued between convergence checks. */
int pmc_set_batch(pmc_handle h, int max_batch, int check_every);

/* ---- host-once uploads ------------------------------------------------------------------------- */
/* End of PDESampler::BuildHierarchy (src/PDESampler.cpp:218-284): the operators of
 * [M B^T; B -alpha W] for one level.  M (Nf x Nf) and B (Ne x Nf) are the ELIMINATED matrices (:236-246),
 * Wdiag the positive diagonal of W_s before the -alpha scaling (:248-258), P = Ps[level] (Ne x P_cols,
 * :189-193) or NULL on the coarsest level, matern_coeff from ComputeScalingCoefficientForSPDE
 * (src/Utilities.hpp:188-200), alpha = 1/corlen^2 (src/PDESampler.cpp:42). */
int pmc_upload_sampler_level(pmc_handle h, int level, int Ne, int Nf,
                             const int *M_rowptr, const int *M_col, const double *M_val,
                             const int *B_rowptr, const int *B_col, const double *B_val,
                             const double *Wdiag,
                             int P_cols, const int *P_rowptr, const int *P_col, const double *P_val,
                             double alpha, double matern_coeff, int lognormal);

/* DarcySolver after BuildHierachySpaces / Build*ObservationFunctional / SetEssBdrConditions /
 * BuildForcingTerms (src/DarcySolver.cpp:60-414): per element (agglomerate) e the RT dofs
 * elem_dofs[elem_ptr[e]..elem_ptr[e+1]) and the dense row-major local mass block (n_e x n_e, blocks
 * concatenated in elem_mat) that ComputeMassOperator(uform, k) scales by k_e (:479); the un-eliminated
 * B = W D (:203-207); the 0/1 mask of essential RT dofs and ess_data (:360-384), rhs (:386-414) and the
 * observation functional (:246-358), all of size Nf+Ne; Pp = L2-form prolongator to the next coarser
 * level (Ne x Pp_cols) or NULL on the coarsest. */
int pmc_upload_darcy_level(pmc_handle h, int level, int Ne, int Nf,
                           const int *elem_ptr, const int *elem_dofs, const double *elem_mat,
                           const int *B_rowptr, const int *B_col, const double *B_val,
                           const int *ess_u, const double *ess_data, const double *rhs, const double *obs,
                           int Pp_cols, const int *Pp_rowptr, const int *Pp_col, const double *Pp_val);

/* Enlarged-domain samplers: the SPDE is solved on an enlarged mesh and the field is brought to the forward problem's
 * mesh before exp, s = row_scale .* (T field).  EmbeddedPDESampler (src/EmbeddedPDESampler.cpp:63-89, applied at
 * :426-435): T = meshP[level] (0/1 selection), row_scale = NULL.  L2ProjectionPDESampler
 * (src/L2ProjectionPDESampler.cpp:488-514, applied at :595-611): T = Gt[level], row_scale = 1/diag(W_orig).
 * T is n_out x Ne(level) in CSR; after this call the level's sampler outputs have n_out entries (= the Darcy level's
 * Ne) while noise vectors and embed_s keep the enlarged size. */
int pmc_upload_field_transfer(pmc_handle h, int level, int n_out, const int *T_rowptr, const int *T_col,
                              const double *T_val, const double *row_scale);

/* A second handle on the same device with the same levels, options, tolerances and random stream but its own CUDA
 * stream and workspace.  The uploaded operators are SHARED with the source (one copy per device, freed with the last
 * handle that uses them); pmc_clone prepares the source first.  The level loops of one InitRun are independent of one
 * another, so a manager keeps one handle per level and issues the level batches from one host thread each: their
 * kernels then share the GPU (a single fine-level batch of ~1000 realisations does not fill it).  Create the clones
 * before the threads that use them start. */
int pmc_clone(pmc_handle src, pmc_handle *out);

/* ---- ranks that share one sample budget (replaces the communicator of MLMC_Manager / MC_Manager,
 * src/MLMC_Manager.hpp:34, src/MC_Manager.hpp:32) -------------------------------------------------------------------
 * Every rank (one handle per GPU: processes under MPI / torchrun, or host threads of one process) owns a slice of every
 * level's realisations and the per-level sums are combined once per InitRun with ncclAllReduce(ncclDouble, ncclSum)
 * over NVLink.  Rank 0 obtains PMC_COMM_ID_BYTES opaque bytes from pmc_comm_unique_id and ships them to the other
 * ranks by whatever the host program uses (MPI_Bcast in the reference's drivers); every rank then calls pmc_comm_init
 * (collective).  pmc_allreduce_sums reduces `count` doubles in place (host array) on the handle's stream and returns
 * when the result is in `sums`; with one rank both calls are no-ops.  libnccl.so.2 is loaded on first use (dlopen;
 * PMC_NCCL_LIB overrides the name), so the library has no link-time dependency on NCCL. */
#define PMC_COMM_ID_BYTES 128
int pmc_comm_unique_id(void *id_out);
int pmc_comm_init(pmc_handle h, int nranks, int rank, const void *id);
int pmc_allreduce_sums(pmc_handle h, double *sums, int count);
int pmc_comm_destroy(pmc_handle h);

/* Build every derived device structure now (Schur-complement hierarchies, block operators) instead of on
 * first use, so that set-up time stays out of timed regions. */
int pmc_prepare(pmc_handle h);

/* ---- NormalDistributionSampler (src/NormalDistributionSampler.hpp:27-66) ------------------------- */
/* ctor (:31) + Split (:46; trng::yarn5::split leapfrog, no-op for nparts <= 1). */
int pmc_rng_init(pmc_handle h, double mu, double sigma, int nparts, int mypart);
/* Raw engine output (trng::yarn5::operator()) at stream positions pos .. pos+n-1; host out. */
int pmc_rng_fill_int(pmc_handle h, uint64_t pos, int64_t n, int32_t *out);
/* operator()(mfem::Vector&) (src/NormalDistributionSampler.cpp:31-37): out[i] = mu + sigma *
 * inv_Phi(uniformoo(engine draw pos+i)); host out. */
int pmc_rng_fill(pmc_handle h, uint64_t pos, int64_t n, double *out);
/* The map of trng::normal_dist<double> alone (src/NormalDistributionSampler.hpp:64): out[i] = mu + sigma *
 * inv_Phi(uniformoo(engine[i])) for caller-chosen engine outputs in [0, 2^31 - 2] (both extreme tails are reachable
 * this way; the stream visits them once in 2^31 draws); host in, host out. */
int pmc_rng_map(pmc_handle h, int64_t n, const int32_t *engine, double *out);

/* ---- PDESampler (src/PDESampler.hpp:68-206) ----------------------------------------------------- */
/* Sample(level, xi) (src/PDESampler.cpp:336-340) for nsamples consecutive realisations: xi_out is
 * [nsamples][Ne(level)], realisation j drawn from stream positions pos0 + j*Ne(level) ... */
int pmc_sampler_sample_batch(pmc_handle h, int level, int nsamples, uint64_t pos0, double *xi_out);
/* Eval(level, xi, s) (:342-409) and Eval(level, xi, s, embed_s, use_init) (:411-535), batched.
 * xi: [nsamples][Ne(xi_level)] (xi_level <= level is passed explicitly instead of being inferred from the
 * vector length, :349,:419).  use_init < 0: 3-argument form.  use_init > 0: init_s [nsamples][Ne(init_level)]
 * is the coarser Gaussian field, prolongated to `level` as the initial guess (:496-511).  s_out
 * [nsamples][Ne(level)] receives exp(field) if lognormal else the field (:529-533); embed_s_out (may be
 * NULL) receives the Gaussian field (:523-527); iters_out (may be NULL) the MINRES iterations per sample. */
/* Chained calls (the managers' sequence Sample -> Eval -> SolveFwd on the same realisations): with the option
 * "cache_results" = 1 the handle keeps the results of pmc_sampler_sample_batch (noise) and pmc_sampler_eval_batch (s_out,
 * embed_s_out) in device memory as well, and a following call may pass NULL for xi, for init_s (with use_init > 0) and,
 * in pmc_darcy_solve_batch, for k: the input is then taken from device memory instead of being copied from the host again
 * (same level and nsamples as the call that produced it, else PMC_ERR_STATE).  Results are still written to the host. */
int pmc_sampler_eval_batch(pmc_handle h, int level, int xi_level, int nsamples, const double *xi,
                           const double *init_s, int init_level, int use_init,
                           double *s_out, double *embed_s_out, int *iters_out);

/* ---- DarcySolver (src/DarcySolver.hpp:55-169) --------------------------------------------------- */
/* SolveFwd(ilevel, k_over_k_ref, Q, C) (src/DarcySolver.cpp:416-437) and SolveFwd_RtnPressure (:439-470),
 * batched: k [nsamples][Ne]; Q_out [nsamples]; C_out [nsamples] (= Nf+Ne, :429) may be NULL; sol_out
 * [nsamples][Nf+Ne] (flux block then pressure block) may be NULL; iters_out may be NULL. */
int pmc_darcy_solve_batch(pmc_handle h, int level, int nsamples, const double *k,
                          double *Q_out, double *C_out, double *sol_out, int *iters_out);
/* y = A_bc(k) x with A_bc the eliminated block operator of DarcySolver::assemble (:472-520):
 * [[M(k) B^T],[B 0]] after EliminateRowCol on the essential dofs.  x, y: [nsamples][Nf+Ne].
 * (Exposes the batched element kernel on its own, for parity tests of the reassembly.) */
int pmc_darcy_apply_batch(pmc_handle h, int level, int nsamples, const double *k, const double *x,
                          double *y);

/* ---- MLMC_Manager / MC_Manager inner loops ------------------------------------------------------ */
/* One level of MLMC_Manager::InitRun (src/MLMC_Manager.cpp:110-138 for level == nlevels-1, :140-175
 * otherwise), fully on the device: noise -> Eval (both levels of the pair, the fine one warm-started) ->
 * SolveFwd (both) -> Y = Q - Qc.  sums[9], in the order of the enum at src/MLMC_Manager.hpp:65
 * {Y2, Y, ABSY, Q2, Q, ABSQ, C, Y3, Y4}, is ACCUMULATED into.  rows (may be NULL): [nsamples][4] =
 * (Y, Q, Qc, C), the MLMC.dat log row (:134-135,:171-172).  total_iters (may be NULL): MINRES iterations
 * summed over all solves. */
int pmc_mlmc_level_batch(pmc_handle h, int level, int nlevels, int nsamples, uint64_t pos0,
                         double *sums, double *rows, int64_t *total_iters);
/* MC_Manager::InitRun (src/MC_Manager.cpp:82-116): single-level loop; sums[4] in the order of the enum at
 * src/MC_Manager.hpp:61 {Q2, Q, ABSQ, C} is ACCUMULATED into; rows (may be NULL): [nsamples][2] = (Q, C). */
int pmc_mc_level_batch(pmc_handle h, int level, int nsamples, uint64_t pos0, double *sums, double *rows,
                       int64_t *total_iters);

/* ---- BayesianInverseProblem + ratio-estimator managers ------------------------------------------------------ */
/* BayesianInverseProblem set-up data for one level (src/BayesianInverseProblem.cpp:26-128): the m pressure functionals
 * g_obs_func[i][level] (g: [m][Ne(level)], un-normalised as the reference holds them; ComputeG divides by their sums,
 * :186-187), the observed data G_obs[m] (:130-175) and the noise variance ("Noise"). */
int pmc_upload_observations(pmc_handle h, int level, int m, const double *g, const double *G_obs, double noise);
/* One level of ML_BayesRatio_Manager::InitRun (src/ML_BayesRatio_Manager.hpp:313-424; the coarsest level is also the
 * loop of SL_BayesRatio_Manager): per realisation two independent prior draws (zxi, then xi; stream positions
 * pos0 + (2j) Ne and pos0 + (2j+1) Ne), Z = likelihood(zxi), R = Q * likelihood(xi)
 * (BayesianInverseProblem::ComputeLikelihood / ComputeR, src/BayesianInverseProblem.cpp:190-218), on `level` and, unless
 * it is the coarsest, on level + 1 with the 3-argument Eval.  sums[20] in the layout of the manager's enum
 * (hpp:67-70) {YZ2, YZ, ABS_YZ, Z2, Z, ABS_Z, YR2, YR, ABS_YR, R2, R, ABS_R, -, -, -, -, -, -, C, -} is ACCUMULATED
 * into; rows (may be NULL): [nsamples][5] = (R, Y_R, Z, Y_Z, c), the log row (:355-358). */
int pmc_bayes_level_batch(pmc_handle h, int level, int nlevels, int nsamples, uint64_t pos0, double *sums,
                          double *rows, int64_t *total_iters);

/* ---- instrumentation ---------------------------------------------------------------------------- */
/* Kernel classes of pmc_kernel_stats. */
enum {
    PMC_K_SADDLE_APPLY = 0, /* block operator apply  q = A u (+ fused u.q)                          */
    PMC_K_LANCZOS_UPDATE,   /* v0 = cq q + cv1 v1 + cv0 v0                                           */
    PMC_K_SOLUTION_UPDATE,  /* w0 = ..., x += cx w0                                                  */
    PMC_K_MASS_SMOOTH,      /* Chebyshev-Jacobi steps on the RT mass block                           */
    PMC_K_SCHUR_SMOOTH,     /* Chebyshev steps / residuals on Schur-complement levels                */
    PMC_K_TRANSFER,         /* prolongation / restriction SpMM                                       */
    PMC_K_SETUP,            /* per-solve value set-up (diag M(k), Schur values, l1 norms)            */
    PMC_K_SCALAR,           /* per-sample scalar recurrences                                         */
    PMC_K_RNG,              /* yarn5 + inverse CDF + W^{1/2} scaling                                 */
    PMC_K_MISC,             /* transposes, exp, QoI, moment sums, fills                              */
    PMC_K_COUNT
};
typedef struct {
    int64_t launches[PMC_K_COUNT];        /* kernel launches outside the persistent solver kernel, per class          */
    double algo_bytes[PMC_K_COUNT];       /* algorithmic bytes per class (DESIGN.md formulas), in-kernel ops included   */
    double ms[PMC_K_COUNT];               /* kernel_ms split by the in-kernel cycle share of each operation class       */
    int64_t timed_launches[PMC_K_COUNT];  /* operations of each class executed inside the persistent kernel (x tiles)   */
    double class_cycle_share[PMC_K_COUNT];/* fraction of the persistent kernel's CTA cycles spent in each class         */
    int64_t kernel_launches;              /* launches of the persistent solver kernel (one per level batch)             */
    int64_t other_launches;               /* all other kernel launches (layout conversion, RNG fills, moment sums)      */
    double kernel_ms;                     /* CUDA-event time of the persistent kernel launches (always measured)        */
    double kernel_algo_bytes;             /* algorithmic bytes moved by the operations executed inside them             */
    int64_t ops_executed;                 /* operations executed inside them, summed over tiles                         */
    int64_t minres_iterations;            /* MINRES iterations, summed over realisations and solves                     */
} pmc_kernel_stats_t;
int pmc_reset_stats(pmc_handle h);
int pmc_kernel_stats(pmc_handle h, pmc_kernel_stats_t *out);

#ifdef __cplusplus
}
#endif
#endif /* PMC_B200_H */
