/*
 * desc_b200.h -- C ABI of the B200-native DESC solver hot path.
 *
 * This is the drop-in boundary.  The reference (ColeWyeth/DESC) has no FFI of its own: the
 * boundary there is the MATLAB function signature
 *
 *     [R_est, R_init, S_vec] = DESC     (Ind, RijMat, params)   Algorithms/DESC.m:14
 *     [S_vec]                = DESC_PGD (Ind, RijMat, params)   Algorithms/DESC_PGD.m:14
 *     [R_est, S_vec]         = DESC_init(Ind, RijMat, params)   Algorithms/DESC_init.m:14
 *     R_est                  = GCW(Ind, AdjMat, RijMat, S_vec)  Utils/GCW.m:1
 *
 * called from Demo/compare_algorithms.m:72.  The entry points below are what a MEX gateway
 * (mex/desc_b200_mex.c) or any other FFI (ctypes: desc_b200/_lib.py) binds; every buffer in a
 * signature is a plain pointer in MATLAB's own memory layout, so an mxArray's mxGetPr()
 * pointer can be passed straight through:
 *
 *   Ind    : m x 2 double, column-major (all i, then all j), 1-based, i<j, rows sorted by
 *            (i,j)                              (Models/Uniform_Topology.m:33-34, DESC.m:19-22)
 *   RijMat : 3 x 3 x m double, column-major: element (r,c) of edge e at 9*e + r + 3*c
 *                                               (Uniform_Topology.m:47-51, DESC.m:65)
 *   S_vec  : 1 x m double, 1.0 for edges without 3-cycles   (DESC.m:148)
 *   R_est  : 3 x 3 x n double, column-major                 (GCW.m:29)
 *
 * Every function returns DESC_B200_OK (0) or a negative error code; the message is available
 * from desc_b200_last_error() (thread-local).  There is NO CPU fallback: without a CUDA
 * device every call fails with DESC_B200_ERR_CUDA.
 *
 * All device memory is owned by the handle.  Host buffers are owned by the caller and are
 * not referenced after the call returns (exception: create with DESC_B200_INPUTS_ON_DEVICE
 * borrows the RijMat device buffer until destroy).
 */
#ifndef DESC_B200_H
#define DESC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DESC_B200_VERSION 102

#define DESC_B200_OK 0
#define DESC_B200_ERR_ARG -1      /* bad argument / input violates the layout contract */
#define DESC_B200_ERR_CUDA -2     /* CUDA runtime error, or no device */
#define DESC_B200_ERR_STATE -3    /* stages called out of order */
#define DESC_B200_ERR_LIMIT -4    /* problem exceeds a compiled-in limit */
#define DESC_B200_ERR_NCCL -5     /* NCCL error */
#define DESC_B200_ERR_NOCONV -6   /* spectral iteration did not converge */

typedef struct desc_b200_handle desc_b200_handle;

/* flags for desc_b200_opts.flags */
#define DESC_B200_INPUTS_ON_DEVICE 1u /* Ind / RijMat passed to create are device pointers */

typedef struct desc_b200_opts {
    int32_t device;       /* CUDA device ordinal; -1 = current device                        */
    uint32_t flags;
    void* stream;         /* cudaStream_t to run on; NULL = the library creates its own      */
    /* multi-GPU: one process per GPU.  world<=1 means single GPU.  nccl_id is the 128-byte
       ncclUniqueId made by desc_b200_nccl_unique_id() on rank 0 and broadcast by the host.
       Every rank must pass the SAME Ind / RijMat: with host inputs each rank uploads 1/world of
       them and the slices are all-gathered over NVLink.                                        */
    int32_t rank;
    int32_t world;
    const void* nccl_id;
} desc_b200_opts;

/* Step rules: params.Gradient of the reference (DESC.m:207).
   kind 0: Utils/ConstantStepSize.m:9-11    step = -lr*grad
   kind 1: Utils/PiecewiseStepSize.m:13-18  t++; step = -lr/(fix(t/decay_interval)+1)*grad
   kind 2: Utils/HybridGradient.m:23-41     strategy 0: Adam(beta_1,beta_2,eps 1e-8, bias corrected)
                                            strategy 1: step = -100*lr/(fix(t/decay_interval)+1)*grad
   `t` is the object's call counter on entry; on return it has advanced by iters_run.
   Adam's moments m_t / v_t (m_cycle doubles each) live on the device inside the handle and are zeroed
   when t == 0 (HybridGradient.m:24-27).  A rule with t > 0 therefore continues only on the handle that
   ran its earlier steps; on any other handle (or after build_incidence, or after a run that the early
   stop ended: its moments are one step ahead of the reported state) pgd returns DESC_B200_ERR_STATE.  */
typedef struct desc_b200_step_rule {
    int32_t kind;
    int32_t strategy;
    double lr;
    double decay_interval;
    double beta_1;
    double beta_2;
    int64_t t;
} desc_b200_step_rule;

/* Stage timings in milliseconds (CUDA events on the library's stream), desc_b200_get_timings */
typedef struct desc_b200_timings {
    double h2d_ms;        /* create: host->device copies (0 with INPUTS_ON_DEVICE)           */
    double graph_ms;      /* create: Ind conversion/validation + adjacency                   */
    double build_ms;      /* build_incidence: co-degree, sampling, CSR incidence, flags      */
    double cycle_ms;      /* cycle_inconsistency                                             */
    double pgd_ms;        /* pgd: all iterations actually run                                */
    double gcw_ms;        /* gcw: weights, power iteration, projection                       */
    double d2h_ms;        /* device->host copies of results of the last pgd/gcw call         */
    double pgd_iter_ms;   /* mean duration of the kernels of one PGD iteration (pass 1 + pass 2) */
    int32_t pgd_launches; /* kernels launched by the last pgd call                           */
    int32_t gcw_iters;    /* power iterations of the last gcw call                           */
    int32_t total_launches; /* kernels launched by this handle so far                        */
    int32_t reserved;
    double pgd_pass1_ms;  /* mean duration of the update kernel (pass over smaller endpoints) */
    double pgd_pass2_ms;  /* mean duration of the pass over larger endpoints (0: single-kernel paths) */
    double pgd_comm_ms;   /* mean per-iteration time in collectives (all-gather S + all-reduce sums) */
    double laa_ms;        /* refine: all IRLS iterations                                      */
    int32_t laa_iters;    /* IRLS iterations of the last refine call                          */
    int32_t laa_cg_iters; /* CG iterations (all IRLS iterations together) of the last refine  */
    double cemp_ms;       /* cemp: initial mean + all reweighting iterations                  */
    int32_t cemp_iters;   /* reweighting iterations of the last cemp call                     */
    int32_t reserved2;
    double mst_ms;        /* mst_init: spanning tree + propagation                            */
    double gcw_spmv_ms;   /* gcw: mean device time of one block-sparse SpMV kernel            */
} desc_b200_timings;

const char* desc_b200_last_error(void);
int desc_b200_version(void);
/* number of visible CUDA devices, or a negative error code */
int desc_b200_device_count(void);
/* fills 128 bytes with a fresh ncclUniqueId (call on rank 0, broadcast, pass via opts) */
int desc_b200_nccl_unique_id(void* out128);
/* NCCL communicators are cached per (id, rank, world, device) and shared by all handles created
   with the same id; this destroys them (call once at shutdown, after destroying the handles). */
int desc_b200_comm_finalize(void);
/* The library keeps freed device buffers in a per-device pool and reuses them in later solves
   (cudaMalloc/cudaFree of GB-sized buffers cost more than some kernels); this returns the cached
   buffers to the driver.  No reference counterpart (MATLAB manages its own heap). */
int desc_b200_trim(void);

/* A1 (DESC.m:19-24): take the graph.  n may be 0 (= max(Ind(:)), as the reference does) or
   an explicit node count >= max(Ind(:)).  Validates the layout contract (1<=i<j<=n,
   strictly sorted rows, every node has an edge) and returns DESC_B200_ERR_ARG otherwise.   */
int desc_b200_create(desc_b200_handle** out, int32_t n, int64_t m, const double* Ind,
                     const double* RijMat, const desc_b200_opts* opts);
void desc_b200_destroy(desc_b200_handle* h);

/* A2-A4 (DESC.m:29-127): co-degrees, sampling budget, CSR edge->3-cycle incidence with
   reciprocal-slot flags, all on device.
   n_sample: 0 = reference rule max(ceil(median(codeg_pos)/4),30) (DESC.m:43); >0 = that
   budget; <0 = keep every triangle.  An edge whose co-degree exceeds n_sample keeps the
   n_sample common neighbours with the smallest desc_b200 sampler key (seed, edge, apex) --
   the deterministic stand-in for datasample (DESC.m:84).
   Explicit lists: cyc_ptr (m+1, host) / cyc_apex (cyc_ptr[m], host, 0-based) replace the
   sampler (e.g. the lists a MATLAB run of the reference drew); pass NULL to sample.        */
int desc_b200_build_incidence(desc_b200_handle* h, int32_t n_sample, uint64_t seed,
                              const int64_t* cyc_ptr, const int32_t* cyc_apex);

/* A5 (DESC.m:129-147): S0_long = abs(acos((trace(Rij*Rjk*Rki)-1)/2))/pi per slot, with the
   reference's unfused operation order.                                                      */
int desc_b200_cycle_inconsistency(desc_b200_handle* h);

/* A6-A12 (DESC.m:148-261): initial weights + up to `iters` fused projected-gradient
   iterations with the reference's early stop (30 consecutive objective decreases < 1e-5).
   S_vec_out: m doubles (may be NULL).  hist_out: 2*iters doubles, row t = [average_change,
   objective] of iteration t+1 (may be NULL).  iters_run_out: iterations executed.          */
int desc_b200_pgd(desc_b200_handle* h, int32_t iters, desc_b200_step_rule* rule,
                  double* S_vec_out, double* hist_out, int32_t* iters_run_out);

/* A13 (Utils/GCW.m:1-38): weighted spectral recovery.  S_vec: m doubles on the host, or NULL
   to use the result of the last desc_b200_pgd on this handle.  R_out: 9*n doubles.         */
int desc_b200_gcw(desc_b200_handle* h, const double* S_vec, double* R_out);

/* DESC_init.m:14 in one call: build + cycle + pgd + gcw (R_out may be NULL: DESC_PGD.m:14).*/
int desc_b200_solve(desc_b200_handle* h, int32_t n_sample, uint64_t seed, int32_t iters,
                    desc_b200_step_rule* rule, double* S_vec_out, double* R_out,
                    double* hist_out, int32_t* iters_run_out);

/* ---- getters (parity tests, MEX diagnostics) ---- */
/* info[0]=n, [1]=m, [2]=m_pos, [3]=m_cycle (global), [4]=n_sample, [5]=max slots per edge,
   [6]=first local edge, [7]=one-past-last local edge, [8]=local slot count, [9]=max co-degree */
int desc_b200_get_info(desc_b200_handle* h, int64_t info[10]);
/* co-degree of every edge (m int32) */
int desc_b200_get_codeg(desc_b200_handle* h, int32_t* codeg);
/* CSR incidence over ALL edges: rowptr (m+1 int64), apex (m_cycle int32, 0-based).  Any NULL skipped. */
int desc_b200_get_incidence(desc_b200_handle* h, int64_t* rowptr, int32_t* apex);
/* per LOCAL slot (all slots when world==1): e_jk / e_ki (0-based edge ids), appears flags
   (IKJ_appears / JKI_appears, DESC.m:113,124) as bytes.  Any pointer may be NULL.           */
int desc_b200_get_slots(desc_b200_handle* h, int32_t* e_jk, int32_t* e_ki, uint8_t* ikj_appears,
                        uint8_t* jki_appears);
/* S0_long and the final wijk of the LOCAL slots */
int desc_b200_get_s0(desc_b200_handle* h, double* S0_long);
int desc_b200_get_w(desc_b200_handle* h, double* wijk);
/* last gcw call: info[0]=power iterations, [1]=final residual ||(N+I)/2 X - X H||_F,
   [2..4]=the three Ritz values of D^-1/2 (W o R) D^-1/2 (= eigenvalues of GCW.m:27), descending */
int desc_b200_get_gcw_info(desc_b200_handle* h, double info[8]);
/* DESC step 5 (DESC.m:265-312 + Utils/Weighted_LAA.m, Build_Amatrix.m, R2Q.m, q2R.m): iteratively
   re-weighted Lie-algebraic averaging started from R_init.  S_vec = NULL: the S_vec of the last pgd
   on this handle; R_init = NULL: the rotations of the last gcw (DESC.m:267).  R_out: 3x3xn.  scores
   (may be NULL): the `score` the reference prints per iteration (DESC.m:305), at most 99 values.
   Stops like the reference: score <= 1e-3 or 99 iterations (DESC.m:272,287).  On a multi-GPU handle every
   rank runs the (small) stage on the replicated graph and returns the same result. */
int desc_b200_refine(desc_b200_handle* h, const double* S_vec, const double* R_init, double* R_out,
                     int32_t* iters_run, double* scores);

/* ---- the comparators that share DESC's incidence ----
   CEMP (Algorithms/CEMP.m:25-131): call after build_incidence (n_sample = CEMP_parameters.nsample, or
   explicit cycle lists -- which may repeat an apex, as the reference's with-replacement draw CEMP.m:63
   does) and cycle_inconsistency.  SVec_0 = mean of the edge's d_ijk (:101), then max_iter reweightings
   SVec(l) = sum_s w_s d_s / sum_s w_s, w_s = exp(-beta_t (SVec(e_ki)+SVec(e_jk))) (:106-126); edges
   without a 3-cycle keep 1 (:102,125).  reweighting: n_reweighting values beta_t, padded with the last
   one as CEMP.m:31-35 does.  SVec_out: m doubles (may be NULL).                                      */
int desc_b200_cemp(desc_b200_handle* h, int32_t max_iter, const double* reweighting, int32_t n_reweighting,
                   double* SVec_out);
/* CEMP_GCW.m:127-159: GCW with the weights 1./(SVec+1e-8) (:141).  SVec = NULL: the last cemp.       */
int desc_b200_cemp_gcw(desc_b200_handle* h, const double* SVec, double* R_out);
/* Algorithms/Spectral.m:15-47: top-3 eigenvectors of the unweighted, un-normalised block matrix of the Rij,
   projected to SO(3) like GCW.m:28-36 (the `Spectral` row of the demo's table).                          */
int desc_b200_spectral(desc_b200_handle* h, double* R_out);
/* One cycle reweighting of an arbitrary edge vector x (m doubles): out(l) = sum_s w_s d_s / sum_s w_s,
   w_s = exp(-beta (x(e_ki)+x(e_jk))); edges without cycles get empty_value.  This is CEMP.m:109-125
   (empty_value 1) and the HVec step of MPLS.m:219-233 with x = ResVec.                                */
int desc_b200_cycle_reweight(desc_b200_handle* h, const double* x, double beta, double empty_value, double* out);
/* MPLS (Algorithms/MPLS.m:28) = cemp + mst_init + mpls_refine.
   mst_init (MPLS.m:152-195): minimum spanning tree of the graph weighted by SVec+1 (ties broken by the edge index),
   R_1 = I, rotations multiplied along the tree -- the CEMP+MST estimate the demo reports (compare_algorithms.m:77).
   SVec = NULL: the last cemp.  DESC_B200_ERR_ARG if the graph is not connected.  Replicated on every rank of a
   multi-GPU handle.                                                                                        */
int desc_b200_mst_init(desc_b200_handle* h, const double* SVec, double* R_out);
typedef struct desc_b200_mpls_params {
    double stop_threshold;            /* MPLS_parameters.stop_threshold                               */
    int32_t max_iter;                 /* MPLS_parameters.max_iter                                     */
    int32_t n_reweighting;            /* lengths of the three vectors below; short vectors are padded */
    int32_t n_thresholding;           /* with their last element (MPLS.m:43-63)                       */
    int32_t n_cycle_info_ratio;
    const double* reweighting;        /* beta_t                                                       */
    const double* thresholding;       /* tau_t: quantile above which an edge weight drops to 1e-4     */
    const double* cycle_info_ratio;   /* alpha_t: weight of the cycle information hij vs the residual */
} desc_b200_mpls_params;
/* mpls_refine (MPLS.m:198-256): Weighted_LAA step, residuals r_ij, h_ij = cycle reweighting of the residuals
   (needs build_incidence + cycle_inconsistency), weights (alpha h + (1-alpha) r)^-0.75 capped at 1e4, edges above the
   tau-quantile set to 1e-4; stops at score <= stop_threshold or max_iter-1 iterations.  SVec = NULL: the last
   cemp (initial weights); R_init = NULL: the last mst_init.  scores (may be NULL): max_iter doubles.  On a multi-GPU
   handle the LAA step is replicated, the cycle reweighting is edge-sharded (all ranks must call it together).    */
int desc_b200_mpls_refine(desc_b200_handle* h, const double* SVec, const double* R_init,
                          const desc_b200_mpls_params* params, double* R_out, int32_t* iters_run, double* scores);

/* ---- evaluation / diagnostics ----
   Utils/Rotation_Alignment.m:13-38 (== GlobalSOdCorrectRight.m): R_est, R_gt 3x3xn (n of the handle);
   R_out = R_est*R_align (9n, may be NULL), R_align 9 doubles (may be NULL), errors in degrees.        */
int desc_b200_rotation_alignment(desc_b200_handle* h, const double* R_est, const double* R_gt, double* R_out,
                                 double* R_align, double* mean_error, double* median_error);
/* desc_b200_pgd with params.make_plots = true (DESC.m:235-239): additionally, after every iteration t,
   diag_out[3(t-1)..] = { mean(abs(ErrVec - S_vec)), MSE_mean, MSE_median } where the last two come from
   GCW(S_vec) aligned to R_orig.  ErrVec: m doubles, R_orig: 3x3xn, diag_out: 3*iters doubles (rows past
   iters_run are zero).  The figures themselves (DESC.m:315-344) are the caller's business.             */
int desc_b200_pgd_diag(desc_b200_handle* h, int32_t iters, desc_b200_step_rule* rule, const double* ErrVec,
                       const double* R_orig, double* S_vec_out, double* hist_out, double* diag_out,
                       int32_t* iters_run_out);

/* ---- the reference's data generators on the device ----
   Models/Uniform_Topology.m:24 `Uniform_Topology(n,p,q,sigma,model)` (kind 0: model 'uniform', kind 1: any other
   model string = self-consistent corruption) and Models/Nonuniform_Topology.m:26
   `Nonuniform_Topology(n,p,p_node_crpt,p_edge_crpt,sigma_in,sigma_out,crpt_type)` (kind 2 'uniform',
   3 'self-consistent', 4 'adv'; sigma = sigma_in).  topology 1 restricts the Erdos-Renyi draw to pairs whose
   circular distance is <= window (SfM-shaped graphs, BASELINE.json configs[4]; kinds 0/1 only).  MATLAB's RNG
   stream is replaced by counter-based draws of `seed` (csrc/gen.cu).  The model owns device buffers in exactly the
   layout desc_b200_create takes with DESC_B200_INPUTS_ON_DEVICE.                                             */
typedef struct desc_b200_model desc_b200_model;
typedef struct desc_b200_gen_opts {
    int32_t device;       /* CUDA device ordinal; -1 = current device */
    int32_t topology;     /* 0 Erdos-Renyi G(n,p); 1 ring window       */
    int32_t n;
    int32_t window;
    int32_t kind;
    int32_t reserved;
    double p, q, sigma, sigma_out, p_node_crpt, p_edge_crpt;
    uint64_t seed;
} desc_b200_gen_opts;
int desc_b200_generate(const desc_b200_gen_opts* opts, desc_b200_model** out);
void desc_b200_model_destroy(desc_b200_model* mo);
/* info[0]=n, [1]=m, [2]=kernels launched, [3]=device; gen_ms (may be NULL): device time of the generation */
int desc_b200_model_info(desc_b200_model* mo, int64_t info[4], double* gen_ms);
/* copies to host buffers (any may be NULL): Ind 2m, RijMat 9m, R_orig 9n, ErrVec m, Rij_orig 9m, corrupted m bytes
   (the reference's corrIndLog / crptInd) */
int desc_b200_model_fetch(desc_b200_model* mo, double* Ind, double* RijMat, double* R_orig, double* ErrVec,
                          double* Rij_orig, uint8_t* corrupted);
/* device pointers of the model's buffers (valid until desc_b200_model_destroy) */
int desc_b200_model_device(desc_b200_model* mo, const double** Ind, const double** RijMat, const double** R_orig,
                           const double** ErrVec);
/* Host-only (no device needed): the vertex-aligned shard boundaries a multi-GPU handle of `world` ranks would use.
   cum_slots[v] / cum_adj[v] (n+1 values each) = slots / adjacency entries of the vertex blocks before v; v_bounds gets
   world+1 boundaries.  slots_only != 0: equal slot counts (round 1); 0: the cost model of csrc/build.cu (pass 1 pays
   its slots + the tables of its own vertex blocks, pass 2 its slots + the tables of every vertex above its first).  */
int desc_b200_plan_shards(int32_t n, int32_t world, const int64_t* cum_slots, const int32_t* cum_adj,
                          int32_t slots_only, int32_t* v_bounds);
int desc_b200_get_timings(desc_b200_handle* h, desc_b200_timings* t);
/* synchronise the handle's stream (for callers timing from outside) */
int desc_b200_sync(desc_b200_handle* h);

#ifdef __cplusplus
}
#endif
#endif /* DESC_B200_H */
