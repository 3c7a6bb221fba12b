/*
 * sfhcuda.h -- C-ABI of libsfhcuda.so: the B200 (sm_100a) fitting hot path of
 * StarFormationHistories.jl (composite -> Poisson log-likelihood -> gradient, plus the MZR/AMR
 * chain rules and the batched-walker log-likelihood).
 *
 * The reference (pure Julia, /root/reference) has NO FFI for this path; it is reached by multiple
 * dispatch.  Each entry point below therefore names the Julia METHOD whose body it replaces
 * (file:line under /root/reference/src).  INTEGRATION.md shows the `ccall` method bodies a
 * maintainer would add, and the ctypes binding used by this repo's tests.
 *
 * Conventions
 *   - plain C: pointers, sizes, ints.  No C++/torch types.  Every call returns an sfh_status;
 *     results are written only on SFH_OK; nothing throws or calls back into the host runtime.
 *   - host pointers are borrowed for the duration of the call only.
 *   - scalars in/out are always double, whatever the storage dtype of the stack.
 *   - one sfh_ctx == one concurrent caller (mirrors TaskLocalValue, fitting/hmc_sample.jl:127);
 *     the sfh_stack is immutable after creation and may be shared by any number of contexts
 *     on any number of host threads.
 *   - there is NO CPU fallback: without a CUDA device every compute entry point returns
 *     SFH_ERR_NO_DEVICE.
 */
#ifndef SFHCUDA_H
#define SFHCUDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SFH_VERSION_MAJOR 0
#define SFH_VERSION_MINOR 1

typedef struct sfh_stack sfh_stack; /* device-resident template stack + data (one row shard) */
typedef struct sfh_ctx sfh_ctx;     /* per-caller stream, scratch, pinned result buffers     */

typedef enum sfh_status {
    SFH_OK = 0,
    SFH_ERR_INVALID_ARG = 1, /* NULL pointer, bad enum, bad size                                        */
    SFH_ERR_SHAPE = 2,       /* the reference's @argcheck ArgumentError (fitting_base.jl:58-59,271-272;  */
                             /* solvers.jl:9-12; mzr.jl:55-57)                                          */
    SFH_ERR_NO_DEVICE = 3,   /* no usable CUDA device / driver                                          */
    SFH_ERR_CUDA = 4,        /* a CUDA runtime/driver call failed (message in sfh_last_error)           */
    SFH_ERR_OOM = 5,         /* device or pinned-host allocation failed                                 */
    SFH_ERR_NCCL = 6,        /* NCCL missing or a collective failed                                     */
    SFH_ERR_UNSUPPORTED = 7, /* valid request this build cannot serve (e.g. not sm_100)                 */
    SFH_ERR_NOT_BOUND = 8,   /* hierarchical call before sfh_hier_bind                                  */
    SFH_ERR_IO = 9           /* file missing, unreadable, truncated or failing its checksums            */
} sfh_status;

/* SFH_U8 is valid for file arrays only (free-form metadata), never for a stack or its data */
typedef enum sfh_dtype { SFH_F32 = 0, SFH_F64 = 1, SFH_I64 = 2, SFH_U8 = 3 } sfh_dtype;

/* metallicity models with a device-side chain rule (hierarchical/mzr.jl:263-284,
 * hierarchical/amr.jl:181-213, :250-298) and the generic escape hatch */
typedef enum sfh_mh_kind {
    SFH_MH_POWERLAW_MZR = 0, /* fixed = {logMstar0}                    */
    SFH_MH_LINEAR_AMR = 1,   /* fixed = {T_max}                        */
    SFH_MH_LOG_AMR = 2       /* fixed = {T_max, solZ, Y_p, gamma}      */
} sfh_mh_kind;
typedef enum sfh_disp_kind { SFH_DISP_GAUSSIAN = 0 } sfh_disp_kind; /* dispersion_models.jl:80-110 */

/* Options for stack creation.  Zero-initialise, set struct_size, override what you need. */
typedef struct sfh_opts {
    int32_t struct_size; /* = sizeof(sfh_opts)                                                      */
    int32_t device;      /* CUDA device ordinal (default 0)                                         */
    int64_t row_begin;   /* bin-row shard [row_begin, row_end) of the host matrix this stack holds; */
    int64_t row_end;     /* 0,0 = all rows.  (one process per GPU)            */
    double clamp_eps;    /* max(m, eps) clamp; 0 = eps(storage dtype) like the reference            */
                         /* (fitting_base.jl:90,277)                                                */
    int32_t tile_bins;   /* 0 = auto; else bins per tile of the fused kernel (a power of two)        */
    int32_t cluster;     /* 0 = auto; else thread-block-cluster size (1,2,4,8,16)                   */
    int32_t force_unfused; /* 1 = always use the two-pass kernels (debug / A-B measurements)        */
    int32_t consumer_warps; /* 0 = auto; 8 or 16 consumer warps per CTA (cluster-tile kernel only)    */
    int32_t variant;     /* 0 = auto; 1 = cluster-tile kernel (sfh_fused.cuh); 4 = warp-specialised   */
                         /* stream kernel (sfh_fused2.cuh)                                           */
    int32_t reserved;
} sfh_opts;

typedef struct sfh_info {
    int64_t nbins_total, ntemplates, row_begin, row_end, ld; /* ld = padded leading dimension (elements) */
    int32_t dtype, device;
    int32_t fused;     /* 1 if the single-pass fused kernel is in use                  */
    int32_t tile_bins, cluster, chunks_per_tile, ring_slots, n_clusters, consumer_warps;
    int32_t sm_count, cc_major, cc_minor;
    int32_t variant;      /* fused kernel in use: 0 none (two-pass), 1 cluster-tile kernel, 4 warp-specialised stream kernel */
    int32_t panel_layout; /* 1: device copy stored as bin-major panels of tile_bins bins (host layout unchanged) */
    int32_t l2_resident_mb; /* MB of the stack the fused kernel keeps L2-resident between evaluations (0: none / stack fits L2) */
    int64_t stack_bytes; /* device bytes held by the stack (padded)                    */
    double clamp_eps;
} sfh_info;

typedef struct sfh_stats {
    int64_t evals;          /* evaluations issued through this context                               */
    int64_t kernel_launches;/* kernels launched by this context                                      */
    double last_device_ms;  /* device time of the last evaluation (CUDA events on the ctx stream)    */
} sfh_stats;

/* ---- library ------------------------------------------------------------------------------ */
int sfh_version(void);              /* major*100 + minor                                         */
const char *sfh_last_error(void);   /* thread-local message for the last non-OK status           */
int sfh_device_count(int *count);   /* SFH_OK with *count = 0 when there is no driver            */

/* ---- stack: the device mirror of stack_models (src/fitting/utilities.jl:12-13) ------------- */
/* models: host, column-major nbins x ntemplates, leading dimension nbins (exactly the Matrix
 * stack_models returns).  data: host vector of nbins (F32/F64/I64: fitting_core_test.jl:38 passes
 * Int64).  Replaces the per-call host arrays captured by every driver closure
 * (solvers.jl:88,181-196; hmc_sample.jl:109; mcmc_sample.jl:102; generic_fitting.jl:306,317).  */
int sfh_stack_create(sfh_stack **out, const void *models, int64_t nbins, int64_t ntemplates,
                     int dtype, const void *data, int data_dtype, const sfh_opts *opts);
/* Synthetic stack generated ON DEVICE: M_ij = scale*U(0,1) from
 * Philox4x32-10 keyed by seed with counter = global linear index i + nbins*j (so any sharding sees
 * the same matrix); data_i ~ Poisson((M x_true)_i) (x_true host, ntemplates doubles).          */
int sfh_stack_create_synthetic(sfh_stack **out, int64_t nbins, int64_t ntemplates, int dtype,
                               uint64_t seed, double scale, const double *x_true,
                               const sfh_opts *opts);
int sfh_stack_destroy(sfh_stack *s); /* idempotent on NULL; safe from a finalizer thread        */
int sfh_stack_info(const sfh_stack *s, sfh_info *info);
int sfh_stack_set_data(sfh_stack *s, const void *data, int data_dtype); /* MCMCModelDistance-style rebinding */
/* copy the shard back to host (column-major rows x ntemplates, ld = rows) -- used by tests      */
int sfh_stack_download(const sfh_stack *s, void *models_out, double *data_out);

/* ---- context ----------------------------------------------------------------------------- */
/* stream: an existing cudaStream_t to enqueue on (e.g. torch's current stream), or NULL for a
 * private non-blocking stream.                                                                 */
int sfh_ctx_create(sfh_stack *s, void *stream, sfh_ctx **out);
/* LIFETIME: the library does not track the contexts of a stack.  Destroy every context before its stack where you control the
 * order; sfh_ctx_destroy itself never touches the stack, so finalizers that run in the other order are harmless -- USING a
 * context whose stack is gone is not.  Keep the number of live contexts bounded (one per CONCURRENT caller, e.g. a pool): each
 * holds a stream, O(nbins + clusters * ntemplates) doubles of device scratch and pinned staging buffers.  Idempotent on NULL. */
int sfh_ctx_destroy(sfh_ctx *c);
int sfh_ctx_stats(const sfh_ctx *c, sfh_stats *out);

/* ---- core path --------------------------------------------------------------------------- */
/* fg!(F, G, coeffs, models, data, composite)  src/fitting/solvers.jl:20-38.
 *   neg_logL != NULL  <=>  F !== nothing ;  G != NULL  <=>  G !== nothing.
 *   *neg_logL = -logL (logL == 0 -> +Inf, fitting_base.jl:95);  G[j] = sum_i M_ij (1 - n_i/m_i).
 *   composite_out (nullable, nbins doubles): what the reference leaves in `composite`:
 *   the residual 1 - n/m when G != NULL (fitting_base.jl:219), else M*coeffs.
 * Host-synchronous: returns when neg_logL / G hold the answer.  On the fused path the call does not synchronise the stream:
 * the results arrive in a pinned buffer as self-validating packets which the calling thread polls (DESIGN.md section 3,
 * "Completion by packets"); the caller's arrays may be pageable.  Environment switches for A/B measurements only:
 * SFH_HOST_PACKETS=0 (synchronise instead), SFH_PDL_EARLY=<bit mask>, SFH_NO_GRAPH=1.                                    */
int sfh_eval_fg(sfh_ctx *c, const double *coeffs, double *neg_logL, double *G, double *composite_out);
/* composite!(C, coeffs, models)  fitting_base.jl:55-65 */
int sfh_composite(sfh_ctx *c, const double *coeffs, double *composite_out);
/* loglikelihood(composite, data)  fitting_base.jl:84-96 (data = the vector bound to the stack) */
int sfh_loglikelihood(sfh_ctx *c, const double *composite, double *logL);
/* loglikelihood(coeffs, models, data)  fitting_base.jl:117-125 */
int sfh_loglikelihood_coeffs(sfh_ctx *c, const double *coeffs, double *logL);
/* grad-loglikelihood!(G, composite, models, data)  fitting_base.jl:265-285:
 * composite (in/out, nbins) is overwritten with 1 - n/max(composite,eps); G = -M' * that.     */
int sfh_grad_loglikelihood(sfh_ctx *c, double *composite_inout, double *G);

/* colsum_j = sum_i M_ij (all ranks' shards when sharded).  Helper for the post-fit summaries that are composite-
 * shaped products against the resident stack, e.g. mdf_amr(coeffs, logAge, MH, models)  src/fitting/mdf.jl:54-74. */
int sfh_column_sums(sfh_ctx *c, double *colsums_out);

/* ---- hierarchical path -------------------------------------------------------------------- */
/* Precompute the age grouping of mzr.jl:131-140 / amr.jl:131 from value-equality of logAge
 * entries in first-appearance order.  *n_ages_out = length(unique(logAge)).                   */
int sfh_hier_bind(sfh_ctx *c, const double *logAge, const double *metallicities, int64_t *n_ages_out);
/* calculate_coeffs(MHmodel, dispmodel, R, logAge, MH)  mzr.jl:50-79 / amr.jl:50-73.
 * variables = [R_1..R_Nj, alpha, beta, sigma] in natural units.                               */
int sfh_calculate_coeffs(sfh_ctx *c, int mh_kind, const double *mh_fixed, int disp_kind,
                         const double *variables, double *coeffs_out);
/* fg!(F, G, MHmodel0, dispmodel0, variables, models, data, composite, logAge, MH)
 * mzr.jl:84-215 ("fg_mzr!") and amr.jl:78-173 ("fg_amr!").  free_mask[3] = free_params of
 * (alpha, beta, sigma); fixed ones receive 0 (mzr.jl:175,196,201).  G has Nj+3 entries.       */
int sfh_eval_fg_hier(sfh_ctx *c, int mh_kind, const double *mh_fixed, int disp_kind,
                     const double *variables, const uint8_t *free_mask,
                     double *neg_logL, double *G);

/* ---- template construction on the device --------------------------------------
 * Builds the stack from ragged per-template point lists instead of uploading host-built Hess diagrams: template t
 * is bin_cmd_smooth (src/StarFormationHistories.jl:574-621) of the points [offsets[t], offsets[t+1]) -- one addstar!
 * (:348-408) per point with the GaussianPSFAsymmetric (cov_mult[t] == 0, :223-266) or GaussianPSFCovariant
 * (cov_mult[t] == +-1, :272-338) kernel -- on a Hess diagram of nx x ny uniform bins whose first left edges are
 * xfirst / yfirst and widths xstep / ystep (> 0).  Bin (ix, iy) is stack row ix + nx*iy (vec of the column-major
 * weights matrix, fitting/utilities.jl:12-13).  colors/mags/color_err/mag_err/weights hold offsets[ntemplates]
 * entries.  data (nullable: zeros) has nx*ny entries.                                                         */
int sfh_stack_create_from_points(sfh_stack **out, int64_t nx, int64_t ny, double xfirst, double xstep,
                                 double yfirst, double ystep, int64_t ntemplates, const int64_t *offsets,
                                 const double *colors, const double *mags, const double *color_err,
                                 const double *mag_err, const double *weights, const int32_t *cov_mult,
                                 int dtype, const void *data, int data_dtype, const sfh_opts *opts);

/* ---- on-disk container for stacks and results ------------------------------------
 * The reference defines no file format (users Serialization.serialize their templates by hand, examples/fitting1.ipynb
 * cell 96).  This one is a little-endian, memory-mappable table of named arrays: a 128-byte header, one 128-byte entry
 * per array, then the arrays, each on a 4096-byte boundary and stored as they sit in host memory (column-major), each
 * with an order-independent 64-bit checksum (csrc/sfh_file.h documents the bytes; tests/file_ref.py restates them in
 * numpy).  Files are written to a temporary name and renamed, so a reader never sees a partial file.              */
typedef struct sfh_file sfh_file; /* a read-only mapping of one container file */
typedef enum sfh_file_kind { SFH_FILE_GENERIC = 0, SFH_FILE_STACK = 1, SFH_FILE_RESULT = 2 } sfh_file_kind;
typedef struct sfh_array_desc {
    char name[48];      /* NUL-terminated, unique within the file                        */
    int32_t dtype;      /* sfh_dtype                                                     */
    int32_t ndim;       /* 1..4; dims[ndim..3] are 1                                     */
    int64_t dims[4];    /* column-major: dims[0] varies fastest                          */
    int64_t nbytes;     /* filled in by the library                                      */
    uint64_t checksum;  /* filled in by the library                                      */
} sfh_array_desc;
/* sum over the 64-bit little-endian words w_i (tail zero-padded) of mix64(w_i xor (i+1)*0x9E3779B97F4A7C15), mod 2^64 */
int sfh_checksum64(const void *data, int64_t nbytes, uint64_t *out);
/* attrs8 (nullable): 8 free int64 attributes stored in the header.  ptrs[i] holds descs[i]'s dims in column-major order. */
int sfh_file_write(const char *path, int kind, const int64_t *attrs8, int narrays, const sfh_array_desc *descs,
                   const void *const *ptrs);
int sfh_file_open(const char *path, sfh_file **out); /* validates header and array table, not the payloads */
int sfh_file_close(sfh_file *f);                     /* idempotent on NULL; invalidates pointers from sfh_file_array */
int sfh_file_info(const sfh_file *f, int *kind, int *narrays, int64_t *attrs8); /* each output nullable */
int sfh_file_find(const sfh_file *f, const char *name, int *index);             /* *index = -1 when absent (still SFH_OK) */
int sfh_file_array(const sfh_file *f, int index, sfh_array_desc *desc, const void **data); /* data points into the mapping */
int sfh_file_verify(const sfh_file *f, int index);   /* recompute the checksum of one array (index < 0: all); SFH_ERR_IO on mismatch */

/* A stack file (kind SFH_FILE_STACK) holds "models" (rows x ntemplates, F32/F64: the stack_models matrix of the rows it
 * covers), "data" (rows, F64) and optionally "logAge" / "MH" (ntemplates, F64); attrs = {nbins_total, row_begin, row_end,
 * nx, ny, stack dtype, 0, 0}.  sfh_stack_save copies the device stack (one bin-row shard, if sharded) straight into the
 * mapped file in column blocks -- no host copy of the stack is made.  nx, ny (0 = unknown) record the Hess-diagram shape. */
int sfh_stack_save(const sfh_stack *s, const char *path, int64_t nx, int64_t ny, const double *logAge, const double *MH);
/* Upload a stack from a file.  opts->row_begin/row_end select the bin-row shard this process holds (it must lie inside
 * the rows the file covers; 0,0 = all rows of the file): only those rows of the mapping are touched, so 8 ranks loading
 * one 40 GB file each read an eighth of it.  verify != 0 checks the payload checksums first (reads the whole file).   */
int sfh_stack_create_from_file(sfh_stack **out, const char *path, int verify, const sfh_opts *opts);

/* ---- batched walkers (new; per-walker semantics of MCMCModel, fitting/mcmc_sample.jl:12-23) -- */
/* X: ntemplates x W column-major.  logL[w] = -Inf if any X[:,w] < 0 (:15-19) else
 * loglikelihood(M*X[:,w], data).                                                               */
int sfh_eval_logl_batched(sfh_ctx *c, const double *X, int64_t W, double *logL);

/* fg! for C coefficient vectors at once (multi-chain HMC: hmc_sample.jl:123-141, generic_fitting.jl:617-626 run one
 * chain per thread; here one device pass serves all chains).  X: ntemplates x C column-major; neg_logL[C];
 * G: ntemplates x C column-major (nullable).  Per-vector semantics are exactly those of sfh_eval_fg.            */
int sfh_eval_fg_batched(sfh_ctx *c, const double *X, int64_t C, double *neg_logL, double *G);

/* Hierarchical fg! for C variable vectors at once (the chains of sample_sfh / tsample_sfh,
 * fitting/hierarchical/generic_fitting.jl:564-665): V is (Nj+3) x C column-major in natural units, neg_logL[C],
 * G (nullable) (Nj+3) x C.  Per-vector semantics are exactly those of sfh_eval_fg_hier.                       */
int sfh_eval_fg_hier_batched(sfh_ctx *c, int mh_kind, const double *mh_fixed, int disp_kind, const double *V,
                             int64_t C, const uint8_t *free_mask, double *neg_logL, double *G);

/* Device-resident affine-invariant ensemble sampler (Goodman & Weare stretch move, the algorithm of KissMCMC.emcee
 * that mcmc_sample drives, fitting/mcmc_sample.jl:97-108): proposal, the log-likelihood of each half-ensemble
 * (sfh_eval_logl_batched's kernel) and accept/reject all stay on the device.  X: ntemplates x W column-major, in = the
 * starting walkers, out = the final ones; W even.  Every nthin-th step is stored: chain (nullable) is
 * (nsteps/nthin) x [ntemplates x W], logl_chain (nullable) (nsteps/nthin) x W; logl_final[W], accept_frac nullable.
 * Random numbers are Philox4x32-10 keyed by `seed` with counter (step, half, walker): a run is reproducible and
 * independent of the GPU count.                                                                              */
int sfh_mcmc_run(sfh_ctx *c, double *X, int64_t W, int64_t nsteps, int64_t nthin, double a_scale, uint64_t seed,
                 double *chain, double *logl_chain, double *logl_final, double *accept_frac);

/* ---- native driver loops (csrc/sfh_drivers.h) ---------------------------------------
 * The reference's fit_templates (src/fitting/solvers.jl:163-221), fit_templates_fast (:238-275), fixed_amr
 * (fitting/hierarchical/fixed_amr.jl:41-180) and fit_sfh (fitting/hierarchical/generic_fitting.jl:242-411) all run
 * Optim.optimize(only_fg!(...), x0, BFGS(alphaguess = InitialStatic(1.0, true), linesearch = HagerZhang())) and read the
 * final inverse Hessian off the trace.  These entry points run that loop (dense BFGS, strong-Wolfe line search) natively
 * around the device evaluations, so one call = one whole optimisation and only the answer crosses the boundary.  The
 * engine is third-party in the reference: iterates differ, converged answers and the role of invH do not.        */
typedef int (*sfh_objective_fn)(void *user, const double *x, int64_t n, double *f, double *g); /* 0 = OK, else aborts */
typedef struct sfh_bfgs_opts {
    int32_t struct_size;   /* = sizeof(sfh_bfgs_opts)                                                            */
    int32_t alphaguess;    /* 0 = default (1); 1: InitialStatic(1.0, scaled = true) like the reference; 2: from the */
                           /* previous decrease (Nocedal & Wright p. 59)                                         */
    double g_abstol;       /* stop when max|g_i| <= g_abstol; 0 = 1e-8 (solvers.jl:206)                          */
    int64_t maxiter;       /* 0 = 5000 (solvers.jl:203)                                                          */
    int32_t device_hessian;/* 1: keep the n x n inverse Hessian in device memory (matrix-vector product and       */
                           /* rank-two update as kernels; pays off from ~500 variables).  Only for the entry       */
                           /* points that take an sfh_ctx.  Experimental, opt-in.                                  */
    int32_t reserved;
} sfh_bfgs_opts;
typedef struct sfh_bfgs_report {
    double f, g_norm;             /* objective and max|g_i| at the returned point                               */
    int64_t iterations, f_calls;
    int32_t converged;            /* 1: g_norm <= g_abstol                                                      */
    int32_t status;               /* 0 converged; 1 iteration limit; 2 line search failed; 3 start not finite   */
} sfh_bfgs_report;
/* x: n doubles, in = start, out = minimiser.  invH (nullable): n x n column-major final inverse-Hessian estimate.   */
int sfh_minimize_bfgs(sfh_objective_fn fn, void *user, int64_t n, double *x, const sfh_bfgs_opts *opts,
                      sfh_bfgs_report *report, double *invH);
/* fit_templates / fit_templates_fast on the resident stack.  theta (ntemplates, in/out) lives in the fitting space:
 *   SFH_FIT_LOG_MAP   theta = log coeffs, objective -logL - sum(theta), gradient G*x - 1      (solvers.jl:178-186)
 *   SFH_FIT_LOG_MLE   theta = log coeffs, objective -logL,              gradient G*x          (solvers.jl:187-195)
 *   SFH_FIT_SQRT_MLE  theta = sqrt coeffs, objective -logL,             gradient 2 G theta    (solvers.jl:254-261)  */
typedef enum sfh_fit_transform { SFH_FIT_LOG_MAP = 0, SFH_FIT_LOG_MLE = 1, SFH_FIT_SQRT_MLE = 2 } sfh_fit_transform;
int sfh_fit_templates_bfgs(sfh_ctx *c, int transform, double *theta, const sfh_bfgs_opts *opts,
                           sfh_bfgs_report *report, double *invH);
/* fixed_amr (fixed_amr.jl:96-167): coeffs_k = relweights_k * exp(theta[age_index_k]), one theta per unique logAge
 * (age_index: 0-based position of template k's age in unique(logAge), n_ages of them); gradient contracted per age
 * (:119-121, :147-151); jacobian != 0 adds the -sum(theta) / -1 terms of the MAP objective (:112, :120).           */
int sfh_fit_fixed_amr_bfgs(sfh_ctx *c, const double *relweights, const int32_t *age_index, int64_t n_ages, int jacobian,
                           double *theta, const sfh_bfgs_opts *opts, sfh_bfgs_report *report, double *invH);
/* fit_sfh: BFGS over xvec = [log R_1..R_Nj, transformed FREE parameters] on the objective of
 * LogDensityProblems.logdensity_and_gradient(::HierarchicalOptimizer, xvec) (generic_fitting.jl:90-199), negated as
 * fg_map! / fg_mle! do (:306-325).  params0[3] = (alpha, beta, sigma) in natural units (fixed ones are used as they
 * are, :134-136); transforms[3] in {1, 0} for free parameters (transformations.jl:18,44; the reference's -1 branch is
 * unvalidated, :155-159, and is refused); free_mask[3]; jacobian_corrections: 1 = MAP,
 * 0 = MLE.  xvec holds Nj + (number of free parameters) doubles.  Requires sfh_hier_bind.                          */
int sfh_fit_sfh_bfgs(sfh_ctx *c, int mh_kind, const double *mh_fixed, int disp_kind, const double *params0,
                     const int32_t *transforms, const uint8_t *free_mask, int jacobian_corrections, double *xvec,
                     const sfh_bfgs_opts *opts, sfh_bfgs_report *report, double *invH);
/* The same loop and the same transformed objective around a CALLER-SUPPLIED hierarchical fg! over the natural variables
 * [R_1..R_n_ages, n_params model parameters] -> (-logL, gradient): the path for user-defined AbstractMetallicityModel /
 * AbstractDispersionModel subtypes (SURVEY.md section 8b "GENERIC"), whose chain rule lives on the host.              */
int sfh_fit_sfh_bfgs_generic(sfh_objective_fn inner_fg, void *user, int64_t n_ages, int32_t n_params, const double *params0,
                             const int32_t *transforms, const uint8_t *free_mask, int jacobian_corrections, double *xvec,
                             const sfh_bfgs_opts *opts, sfh_bfgs_report *report, double *invH);

/* fit_templates_lbfgsb (src/fitting/solvers.jl:70-90) hands fg! to LBFGSB.jl with lb = 0, ub = Inf, m = 10, factr = 1e-12,
 * pgtol = 1e-5.  The same box-constrained limited-memory method (Byrd, Lu, Nocedal & Zhu 1995: generalized Cauchy point,
 * subspace minimisation, projected-gradient stopping rule) runs natively here (csrc/sfh_drivers.h).                  */
typedef struct sfh_lbfgsb_opts {
    int32_t struct_size;   /* = sizeof(sfh_lbfgsb_opts)                                                          */
    int32_t m;             /* stored correction pairs; 0 = 10                                                    */
    double factr;          /* relative-reduction tolerance in units of machine epsilon, as in the Fortran code;  */
                           /* 0 = 1e-12 (solvers.jl:82)                                                          */
    double pgtol;          /* projected-gradient tolerance; 0 = 1e-5                                             */
    int64_t maxiter;       /* 0 = 100000                                                                         */
    int64_t maxfun;        /* 0 = 100000                                                                         */
} sfh_lbfgsb_opts;
typedef struct sfh_lbfgsb_report {
    double f, pg_norm;     /* objective and infinity norm of the projected gradient at the returned point        */
    int64_t iterations, f_calls;
    int32_t status;        /* 0: pg_norm <= pgtol; 1: relative reduction <= factr*eps; 2: iteration / evaluation  */
                           /* limit; 3: line search failed (abnormal termination); 4: start not finite           */
    int32_t reserved;
} sfh_lbfgsb_report;
/* lb / ub: n entries each or NULL for unbounded below / above (use +-HUGE_VAL entries for partially bounded problems). */
int sfh_minimize_lbfgsb(sfh_objective_fn fn, void *user, int64_t n, double *x, const double *lb, const double *ub,
                        const sfh_lbfgsb_opts *opts, sfh_lbfgsb_report *report);
/* fit_templates_lbfgsb on the resident stack: coeffs (ntemplates, in = x0 after renormalize_x0, out = best fit), bounds
 * 0 <= coeffs (solvers.jl:82-90); report->f = -logL at the solution.                                                 */
int sfh_fit_templates_lbfgsb(sfh_ctx *c, double *coeffs, const sfh_lbfgsb_opts *opts, sfh_lbfgsb_report *report);

/* Native multi-chain NUTS (csrc/sfh_nuts.h): every chain is a host thread running the No-U-Turn recursion (Hoffman & Gelman
 * 2014, alg. 6, dual-averaging step size, Gaussian kinetic energy) -- the role DynamicHMC plays in hmc_sample
 * (src/fitting/hmc_sample.jl:105-143: one chain per thread) and sample_sfh / tsample_sfh (generic_fitting.jl:456-665: one
 * task per short chain) -- but the gradient requests of all live chains are served by ONE batched device pass per round
 * (sfh_eval_fg_batched / sfh_eval_fg_hier_batched).  Random numbers: Philox4x32-10 keyed by seed, counter (chain, draw).   */
typedef int (*sfh_batch_logdensity_fn)(void *user, const double *Theta /* n x C */, int64_t n, int64_t C, double *logp /* C */,
                                       double *grad /* n x C */);   /* 0 = OK, else aborts the run with that status */
typedef struct sfh_nuts_opts {
    int32_t struct_size;  /* = sizeof(sfh_nuts_opts)                                                             */
    int32_t max_depth;    /* 0 = 8                                                                               */
    int64_t nwarmup;      /* dual-averaging warm-up draws per chain (not stored); used as given                  */
    double delta;         /* target mean acceptance; 0 = 0.8                                                     */
    double eps0;          /* > 0: fixed initial step size (the reference's epsilon, generic_fitting.jl:479-482); */
                          /* 0: doubling heuristic                                                               */
    uint64_t seed;
    int32_t mass_kind;    /* M^-1 of the kinetic energy: 0 identity, 1 diagonal (inv_mass[n]), 2 dense            */
                          /* (inv_mass[n x n], e.g. MAP.invH as in GaussianKineticEnergy(MAP.invH))               */
    int32_t reserved;
} sfh_nuts_opts;
/* theta0: n x nchains starts; nsteps[nchains] draws per chain; samples: n x sum(nsteps) column-major, chains concatenated
 * in order; logps[sum(nsteps)]; step_sizes[nchains] (nullable); n_batches / n_evals (nullable): batched passes made and
 * chain-evaluations served.                                                                                      */
int sfh_nuts_run(sfh_batch_logdensity_fn fn, void *user, int64_t n, int64_t nchains, const double *theta0, const int64_t *nsteps,
                 const double *inv_mass, const sfh_nuts_opts *opts, double *samples, double *logps, double *step_sizes,
                 int64_t *n_batches, int64_t *n_evals);
/* hmc_sample: the chains of HMCModel (hmc_sample.jl:24-37: theta = log coeffs, logp = -fg! + sum(theta), grad = -G x + 1)
 * on the resident stack; theta0 is ntemplates x nchains; samples come back in the theta (log) space.              */
int sfh_hmc_sample_nuts(sfh_ctx *c, int64_t nchains, const double *theta0, const int64_t *nsteps, const double *inv_mass,
                        const sfh_nuts_opts *opts, double *samples, double *logps, double *step_sizes, int64_t *n_batches,
                        int64_t *n_evals);
/* sample_sfh / tsample_sfh: the chains of the HierarchicalOptimizer log-density with Jacobian corrections
 * (generic_fitting.jl:90-199, :477) over x = [log R_j, transformed free parameters]; arguments as sfh_fit_sfh_bfgs.
 * theta0 is (Nj + nfree) x nchains; samples come back in the transformed space.  Requires sfh_hier_bind.          */
int sfh_sample_sfh_nuts(sfh_ctx *c, int mh_kind, const double *mh_fixed, int disp_kind, const double *params0,
                        const int32_t *transforms, const uint8_t *free_mask, int64_t nchains, const double *theta0,
                        const int64_t *nsteps, const double *inv_mass, const sfh_nuts_opts *opts, double *samples, double *logps,
                        double *step_sizes, int64_t *n_batches, int64_t *n_evals);
/* The same chains around a CALLER-SUPPLIED batched hierarchical fg! over the natural variables (V: (n_ages + n_params) x C ->
 * -logL[C] in `logp`, gradient in `grad`): user-defined metallicity / dispersion models (SURVEY.md section 8b "GENERIC").   */
int sfh_sample_sfh_nuts_generic(sfh_batch_logdensity_fn inner_fg, void *user, int64_t n_ages, int32_t n_params, const double *params0,
                                const int32_t *transforms, const uint8_t *free_mask, int64_t nchains, const double *theta0,
                                const int64_t *nsteps, const double *inv_mass, const sfh_nuts_opts *opts, double *samples,
                                double *logps, double *step_sizes, int64_t *n_batches, int64_t *n_evals);

/* ---- multi-GPU: bin-row shards, one process per GPU -------------------- */
/* 128-byte NCCL unique id (rank 0 creates, the host runtime broadcasts it).                    */
int sfh_comm_unique_id(void *id128);
/* After this, every evaluation all-reduces [logL, G_1..G_T] (sum, FP64) over the ranks on the
 * context stream before results are returned, so each rank receives the full-stack answer.     */
int sfh_comm_init(sfh_ctx *c, int nranks, int rank, const void *id128);
/* Optional upgrade of the fused path's reduction to a ONE-SHOT all-reduce over NVLink peer memory, fused into the
 * finalize kernel (no NCCL call on the critical path).  Each rank obtains the 64-byte CUDA-IPC handle of its inbox,
 * the host runtime all-gathers them (nranks x 64 bytes, rank order), every rank calls sfh_comm_p2p_init.  Requires
 * sfh_comm_init first (NCCL remains the fallback for the two-pass and batched-walker paths).                    */
int sfh_comm_p2p_handle(sfh_ctx *c, int nranks, void *handle64_out);
int sfh_comm_p2p_init(sfh_ctx *c, int nranks, int rank, const void *handles);
/* on = 0: back to the NCCL all-reduce (what a rank does when some OTHER rank's sfh_comm_p2p_init failed: the switch must be
 * taken by all ranks or none, else the ranks that switched wait for flags that never come); on = 1 re-enables after a
 * successful sfh_comm_p2p_init.  Must be called by every rank between evaluations, never concurrently with one.          */
int sfh_comm_p2p_enable(sfh_ctx *c, int on);
/* How this context reduces: *mode = 0 single GPU, 1 NCCL all-reduce on the context stream, 2 one-shot NVLink exchange
 * inside the finalize kernel (fused path) with NCCL for the other paths.  Every output is nullable.                */
int sfh_ctx_comm_info(const sfh_ctx *c, int *nranks, int *rank, int *mode);

/* ---- device-side plumbing (no host round trip; used by bench.py and torch interop) ---------- */
/* ---- multi-GPU from ONE process (`devices, ndev` of sfh_stack_create) ------------------------------
 * Every reference caller is a single Julia process (fit_sfh, generic_fitting.jl:242-411; fit_templates, solvers.jl:82-90), so a
 * stack too large for one GPU must be shardable without a process launcher.  sfh_group_create takes the arguments of
 * sfh_stack_create plus the device list (devices == NULL: ordinals 0..ndev-1), splits the bin rows over the GPUs
 * (sfh_shard_rows, 128-bin boundaries), enables peer access between all pairs and wires the one-shot NVLink exchange of the
 * finalize kernel with plain peer pointers -- no NCCL, no CUDA-IPC, no launcher.  sfh_group_ctx returns the group's PRIMARY
 * context: sfh_hier_bind, sfh_eval_fg, sfh_eval_fg_hier, sfh_composite-free drivers built on them (sfh_fit_templates_lbfgsb,
 * sfh_fit_templates_bfgs, sfh_fit_fixed_amr_bfgs, sfh_fit_sfh_bfgs) accept it unchanged and return the all-reduced answer; the
 * other GPUs are driven by worker threads inside the library.  Entry points that reduce with NCCL (two-pass direct calls,
 * batched walkers, the samplers built on them) answer SFH_ERR_UNSUPPORTED on a group context.  Every shard must get a fused
 * tiling (SFH_ERR_UNSUPPORTED otherwise).  The primary context is owned by the group: sfh_group_destroy releases everything. */
typedef struct sfh_group sfh_group;
int sfh_shard_rows(int64_t nbins, int nshards, int i, int64_t align, int64_t *row_begin, int64_t *row_end); /* pure arithmetic */
int sfh_group_create(sfh_group **out, const void *models, int64_t nbins, int64_t ntemplates, int dtype, const void *data,
                     int data_dtype, const int *devices, int ndev, const sfh_opts *opts);
int sfh_group_create_synthetic(sfh_group **out, int64_t nbins, int64_t ntemplates, int dtype, uint64_t seed, double scale,
                               const double *x_true, const int *devices, int ndev, const sfh_opts *opts);
int sfh_group_destroy(sfh_group *g);            /* idempotent on NULL */
int sfh_group_ctx(sfh_group *g, sfh_ctx **primary);
int sfh_group_info(const sfh_group *g, int *ndev, sfh_info *infos /* nullable, ndev entries */);
/* bench plumbing: mean device time of `reps` back-to-back all-reduced evaluations (CUDA events per GPU, max over the GPUs) */
int sfh_group_time_fg(sfh_group *g, const double *coeffs, int reps, int want_G, double *ms_per_eval_out);

/* Enqueue one fused evaluation on the ctx stream.  d_coeffs: device, ntemplates doubles.
 * d_out: device, 1+ntemplates doubles = [logL (raw sum, un-guarded), G...].  Asynchronous.
 * want_G = 0 skips the gradient pass.                                                          */
int sfh_enqueue_fg(sfh_ctx *c, const double *d_coeffs, double *d_out, int want_G);
int sfh_enqueue_logl_batched(sfh_ctx *c, const double *d_X, int64_t W, double *d_logL);
int sfh_ctx_synchronize(sfh_ctx *c);
/* Time `reps` back-to-back evaluations with CUDA events on the ctx stream (inputs resident).
 * flush_l2 != 0 writes a >L2-sized buffer between evaluations (outside the per-kernel events).
 * ms_kernel_out (nullable): mean duration of the fused kernel alone.                           */
int sfh_time_fg(sfh_ctx *c, const double *coeffs, int reps, int want_G, int flush_l2,
                double *ms_per_eval_out, double *ms_kernel_out);

#ifdef __cplusplus
}
#endif
#endif /* SFHCUDA_H */
