/*
 * desc_b200_mex.c -- thin MEX gateway: MATLAB host code -> C ABI (include/desc_b200.h) -> CUDA.
 *
 *   out = desc_b200_mex('solve', Ind, RijMat, iters, rule, n_sample, seed, want_R)     want_R: 0 | 1 (GCW) | 2 (GCW + refinement)
 *   out = desc_b200_mex('solve', Ind, RijMat, iters, rule, n_sample, seed, want_R, ErrVec, R_orig)
 *         (params.make_plots = true, DESC.m:235-239: adds out.diag = iters_run x 3
 *          [svec_errors, MSE_means, MSE_medians])
 *   out = desc_b200_mex('cemp',  Ind, RijMat, max_iter, reweighting, nsample, seed, want_R)
 *         (Algorithms/CEMP.m / CEMP_GCW.m: out.SVec 1 x m, out.R_est)
 *   out = desc_b200_mex('mpls',  Ind, RijMat, cemp_max_iter, cemp_reweighting, nsample, seed, stop_threshold,
 *                       max_iter, reweighting, thresholding, cycle_info_ratio)
 *         (Algorithms/MPLS.m: out.R_est, out.R_init (CEMP+MST), out.scores)
 *   R   = desc_b200_mex('spectral', Ind, RijMat)                 (Algorithms/Spectral.m)
 *   out = desc_b200_mex('align', R_est, R_gt)   (Utils/Rotation_Alignment.m: out.R_out, out.R_align,
 *          out.mean_error, out.median_error)
 *   R   = desc_b200_mex('gcw',   Ind, RijMat, S_vec)
 *   out = desc_b200_mex('refine', Ind, RijMat, S_vec, R_init)    (DESC.m:265-312; out.R_est, out.scores)
 *   mo  = desc_b200_mex('generate', kind, topology, n, window, p, q, sigma, sigma_out, p_node_crpt, p_edge_crpt, seed)
 *         (Models/Uniform_Topology.m / Nonuniform_Topology.m on the device: mo.Ind, mo.RijMat, mo.Rij_orig,
 *          mo.R_orig, mo.ErrVec, mo.AdjMat)
 *   n   = desc_b200_mex('device_count')
 *
 * 'solve' runs Algorithms/DESC.m:14-263 (== DESC_PGD.m:14-261 / DESC_init.m:14-253) on the GPU and
 * returns a struct with fields S_vec (1 x m), R_est (3 x 3 x n, [] if want_R is false), hist
 * (iters_run x 2: [average_change, objective] per iteration -- what DESC.m:241 prints),
 * iters_run, t (the step rule's advanced call counter), n_sample, m_cycle.
 * `rule` is a struct made by matlab/desc_b200_rule.m from params.Gradient.
 *
 * Build (on a machine with MATLAB):  mex -I../include desc_b200_mex.c -L../desc_b200 -ldesc_b200
 * The gateway only moves pointers: mxGetPr() of Ind / RijMat already has the layout the C ABI wants.
 * Every library error becomes mexErrMsgIdAndTxt('DESC:b200', ...); there is no CPU fallback.
 */
#include <string.h>

#include "mex.h"
#include "desc_b200.h"

static void fail_if(int rc, desc_b200_handle* h) {
    if (rc != DESC_B200_OK) {
        char msg[1100];
        strncpy(msg, desc_b200_last_error(), sizeof(msg) - 1);
        msg[sizeof(msg) - 1] = 0;
        if (h) desc_b200_destroy(h);
        mexErrMsgIdAndTxt("DESC:b200", "desc_b200 error %d: %s", rc, msg);
    }
}

static double field_or(const mxArray* s, const char* name, double dflt) {
    const mxArray* f = mxGetField(s, 0, name);
    return (f && !mxIsEmpty(f)) ? mxGetScalar(f) : dflt;
}

static void check_inputs(const mxArray* Ind, const mxArray* Rij, mwSize* m_out) {
    if (!mxIsDouble(Ind) || mxIsComplex(Ind) || mxGetN(Ind) != 2)
        mexErrMsgIdAndTxt("DESC:b200", "Ind must be a real double m x 2 matrix");
    const mwSize m = mxGetM(Ind);
    const mwSize* d = mxGetDimensions(Rij);
    const mwSize nd = mxGetNumberOfDimensions(Rij);
    const mwSize m3 = nd == 3 ? d[2] : (nd == 2 ? 1 : 0);
    if (!mxIsDouble(Rij) || mxIsComplex(Rij) || d[0] != 3 || d[1] != 3 || m3 != m)
        mexErrMsgIdAndTxt("DESC:b200", "RijMat must be a real double 3 x 3 x m array");
    *m_out = m;
}

void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
    char cmd[32];
    (void)nlhs;
    if (nrhs < 1 || mxGetString(prhs[0], cmd, sizeof(cmd)) != 0)
        mexErrMsgIdAndTxt("DESC:b200", "first argument must be a command string");

    if (strcmp(cmd, "device_count") == 0) {
        plhs[0] = mxCreateDoubleScalar((double)desc_b200_device_count());
        return;
    }
    if (strcmp(cmd, "solve") == 0) {
        if (nrhs != 8 && nrhs != 10) mexErrMsgIdAndTxt("DESC:b200", "solve: 7 or 9 arguments expected");
        const int diag = nrhs == 10;
        mwSize m;
        check_inputs(prhs[1], prhs[2], &m);
        const int iters = (int)mxGetScalar(prhs[3]);
        const mxArray* rs = prhs[4];
        if (!mxIsStruct(rs)) mexErrMsgIdAndTxt("DESC:b200", "rule must be a struct (desc_b200_rule.m)");
        desc_b200_step_rule rule;
        rule.kind = (int32_t)field_or(rs, "kind", 0);
        rule.strategy = (int32_t)field_or(rs, "strategy", 0);
        rule.lr = field_or(rs, "lr", 0.01);
        rule.decay_interval = field_or(rs, "decay_interval", 1);
        rule.beta_1 = field_or(rs, "beta_1", 0.9);
        rule.beta_2 = field_or(rs, "beta_2", 0.999);
        rule.t = (int64_t)field_or(rs, "t", 0);
        const int n_sample = (int)mxGetScalar(prhs[5]);
        const uint64_t seed = (uint64_t)mxGetScalar(prhs[6]);
        /* want_R: false / 0 = S_vec only (DESC_PGD.m), true / 1 = + GCW rotations (DESC_init.m), 2 = + the refinement
           stage DESC.m:265-312 on the SAME handle (DESC.m: no second upload of Ind / RijMat, no second graph build) */
        const int want_R = mxIsLogicalScalarTrue(prhs[7]) ? 1 : (mxIsNumeric(prhs[7]) ? (int)mxGetScalar(prhs[7]) : 0);

        desc_b200_handle* h = NULL;
        fail_if(desc_b200_create(&h, 0, (int64_t)m, mxGetPr(prhs[1]), mxGetPr(prhs[2]), NULL), NULL);
        int64_t info[10];
        fail_if(desc_b200_get_info(h, info), h);
        const mwSize n = (mwSize)info[0];

        mxArray* S = mxCreateDoubleMatrix(1, m, mxREAL);               /* DESC.m:148: 1 x m row */
        mxArray* hist = mxCreateDoubleMatrix(2, iters > 0 ? iters : 1, mxREAL);
        mxArray* R = NULL;
        if (want_R) {
            mwSize dims[3] = {3, 3, 0};
            dims[2] = n;
            R = mxCreateNumericArray(3, dims, mxDOUBLE_CLASS, mxREAL);   /* GCW.m:29 */
        } else {
            R = mxCreateDoubleMatrix(0, 0, mxREAL);
        }
        int32_t iters_run = 0;
        mxArray* dg = NULL;
        if (!diag) {
            fail_if(desc_b200_solve(h, n_sample, seed, iters, &rule, mxGetPr(S), want_R ? mxGetPr(R) : NULL,
                                    mxGetPr(hist), &iters_run), h);
        } else {   /* params.make_plots: the diagnostics branch DESC.m:235-239 runs on the device too */
            if (!mxIsDouble(prhs[8]) || mxGetNumberOfElements(prhs[8]) != m || !mxIsDouble(prhs[9]) ||
                mxGetNumberOfElements(prhs[9]) != 9 * n) {
                desc_b200_destroy(h);
                mexErrMsgIdAndTxt("DESC:b200", "make_plots: params.ErrVec must have m entries and params.R_orig must be 3 x 3 x n");
            }
            dg = mxCreateDoubleMatrix(3, iters > 0 ? iters : 1, mxREAL);
            fail_if(desc_b200_build_incidence(h, n_sample, seed, NULL, NULL), h);
            fail_if(desc_b200_cycle_inconsistency(h), h);
            fail_if(desc_b200_pgd_diag(h, iters, &rule, mxGetPr(prhs[8]), mxGetPr(prhs[9]), mxGetPr(S), mxGetPr(hist),
                                       mxGetPr(dg), &iters_run), h);
            if (want_R) fail_if(desc_b200_gcw(h, NULL, mxGetPr(R)), h);
        }
        mxArray *Rref = NULL, *scores = NULL;
        if (want_R >= 2) {   /* DESC.m:265-312 from the GCW rotations and S_vec that are still on the device */
            mwSize dims[3] = {3, 3, 0};
            dims[2] = n;
            Rref = mxCreateNumericArray(3, dims, mxDOUBLE_CLASS, mxREAL);
            mxArray* sc = mxCreateDoubleMatrix(1, 100, mxREAL);
            int32_t lrun = 0;
            fail_if(desc_b200_refine(h, NULL, NULL, mxGetPr(Rref), &lrun, mxGetPr(sc)), h);
            scores = mxCreateDoubleMatrix(1, lrun, mxREAL);
            for (int t = 0; t < lrun; t++) mxGetPr(scores)[t] = mxGetPr(sc)[t];
            mxDestroyArray(sc);
        } else {
            Rref = mxCreateDoubleMatrix(0, 0, mxREAL);
            scores = mxCreateDoubleMatrix(1, 0, mxREAL);
        }
        fail_if(desc_b200_get_info(h, info), h);
        desc_b200_destroy(h);

        /* hist comes back as rows [change, objective] per iteration = 2 x iters column-major */
        mxArray* histT = mxCreateDoubleMatrix(iters_run, 2, mxREAL);
        for (int t = 0; t < iters_run; t++) {
            mxGetPr(histT)[t] = mxGetPr(hist)[2 * t];
            mxGetPr(histT)[t + iters_run] = mxGetPr(hist)[2 * t + 1];
        }
        mxDestroyArray(hist);
        mxArray* dgT = mxCreateDoubleMatrix(dg ? iters_run : 0, 3, mxREAL);
        if (dg) {
            for (int t = 0; t < iters_run; t++)
                for (int c = 0; c < 3; c++) mxGetPr(dgT)[t + c * iters_run] = mxGetPr(dg)[3 * t + c];
            mxDestroyArray(dg);
        }
        const char* fields[] = {"S_vec", "R_est", "hist", "iters_run", "t", "n_sample", "m_cycle", "diag", "R_refined", "scores"};
        plhs[0] = mxCreateStructMatrix(1, 1, 10, fields);
        mxSetField(plhs[0], 0, "R_refined", Rref);
        mxSetField(plhs[0], 0, "scores", scores);
        mxSetField(plhs[0], 0, "diag", dgT);
        mxSetField(plhs[0], 0, "S_vec", S);
        mxSetField(plhs[0], 0, "R_est", R);
        mxSetField(plhs[0], 0, "hist", histT);
        mxSetField(plhs[0], 0, "iters_run", mxCreateDoubleScalar((double)iters_run));
        mxSetField(plhs[0], 0, "t", mxCreateDoubleScalar((double)rule.t));
        mxSetField(plhs[0], 0, "n_sample", mxCreateDoubleScalar((double)info[4]));
        mxSetField(plhs[0], 0, "m_cycle", mxCreateDoubleScalar((double)info[3]));
        return;
    }
    if (strcmp(cmd, "gcw") == 0) {
        if (nrhs != 4) mexErrMsgIdAndTxt("DESC:b200", "gcw: 3 arguments expected");
        mwSize m;
        check_inputs(prhs[1], prhs[2], &m);
        if (!mxIsDouble(prhs[3]) || mxGetNumberOfElements(prhs[3]) != m)
            mexErrMsgIdAndTxt("DESC:b200", "S_vec must have one entry per edge");
        desc_b200_handle* h = NULL;
        fail_if(desc_b200_create(&h, 0, (int64_t)m, mxGetPr(prhs[1]), mxGetPr(prhs[2]), NULL), NULL);
        int64_t info[10];
        fail_if(desc_b200_get_info(h, info), h);
        mwSize dims[3] = {3, 3, 0};
        dims[2] = (mwSize)info[0];
        plhs[0] = mxCreateNumericArray(3, dims, mxDOUBLE_CLASS, mxREAL);
        fail_if(desc_b200_gcw(h, mxGetPr(prhs[3]), mxGetPr(plhs[0])), h);
        desc_b200_destroy(h);
        return;
    }
    if (strcmp(cmd, "refine") == 0) {
        if (nrhs != 5) mexErrMsgIdAndTxt("DESC:b200", "refine: 4 arguments expected");
        mwSize m;
        check_inputs(prhs[1], prhs[2], &m);
        if (!mxIsDouble(prhs[3]) || mxGetNumberOfElements(prhs[3]) != m)
            mexErrMsgIdAndTxt("DESC:b200", "S_vec must have one entry per edge");
        desc_b200_handle* h = NULL;
        fail_if(desc_b200_create(&h, 0, (int64_t)m, mxGetPr(prhs[1]), mxGetPr(prhs[2]), NULL), NULL);
        int64_t info[10];
        fail_if(desc_b200_get_info(h, info), h);
        if (!mxIsDouble(prhs[4]) || mxGetNumberOfElements(prhs[4]) != (mwSize)(9 * info[0])) {
            desc_b200_destroy(h);
            mexErrMsgIdAndTxt("DESC:b200", "R_init must be 3 x 3 x n");
        }
        mwSize dims[3] = {3, 3, 0};
        dims[2] = (mwSize)info[0];
        mxArray* R = mxCreateNumericArray(3, dims, mxDOUBLE_CLASS, mxREAL);
        mxArray* sc = mxCreateDoubleMatrix(100, 1, mxREAL);
        int32_t run = 0;
        fail_if(desc_b200_refine(h, mxGetPr(prhs[3]), mxGetPr(prhs[4]), mxGetPr(R), &run, mxGetPr(sc)), h);
        desc_b200_destroy(h);
        mxArray* scores = mxCreateDoubleMatrix(run, 1, mxREAL);
        for (int t = 0; t < run; t++) mxGetPr(scores)[t] = mxGetPr(sc)[t];
        mxDestroyArray(sc);
        const char* fields[] = {"R_est", "scores"};
        plhs[0] = mxCreateStructMatrix(1, 1, 2, fields);
        mxSetField(plhs[0], 0, "R_est", R);
        mxSetField(plhs[0], 0, "scores", scores);
        return;
    }
    if (strcmp(cmd, "generate") == 0) {
        if (nrhs != 12) mexErrMsgIdAndTxt("DESC:b200", "generate: 11 arguments expected");
        desc_b200_gen_opts o;
        memset(&o, 0, sizeof(o));
        o.device = -1;
        o.kind = (int32_t)mxGetScalar(prhs[1]);
        o.topology = (int32_t)mxGetScalar(prhs[2]);
        o.n = (int32_t)mxGetScalar(prhs[3]);
        o.window = (int32_t)mxGetScalar(prhs[4]);
        o.p = mxGetScalar(prhs[5]);
        o.q = mxGetScalar(prhs[6]);
        o.sigma = mxGetScalar(prhs[7]);
        o.sigma_out = mxGetScalar(prhs[8]);
        o.p_node_crpt = mxGetScalar(prhs[9]);
        o.p_edge_crpt = mxGetScalar(prhs[10]);
        o.seed = (uint64_t)mxGetScalar(prhs[11]);
        desc_b200_model* mo = NULL;
        fail_if(desc_b200_generate(&o, &mo), NULL);
        int64_t info[4];
        int rc = desc_b200_model_info(mo, info, NULL);
        if (rc != DESC_B200_OK) {
            desc_b200_model_destroy(mo);
            fail_if(rc, NULL);
        }
        const mwSize n = (mwSize)info[0], m = (mwSize)info[1];
        mwSize dm[3] = {3, 3, 0}, dn[3] = {3, 3, 0};
        dm[2] = m;
        dn[2] = n;
        mxArray* Ind = mxCreateDoubleMatrix(m, 2, mxREAL);
        mxArray* Rij = mxCreateNumericArray(3, dm, mxDOUBLE_CLASS, mxREAL);
        mxArray* Rij0 = mxCreateNumericArray(3, dm, mxDOUBLE_CLASS, mxREAL);
        mxArray* Ro = mxCreateNumericArray(3, dn, mxDOUBLE_CLASS, mxREAL);
        mxArray* Err = mxCreateDoubleMatrix(1, m, mxREAL);
        rc = desc_b200_model_fetch(mo, mxGetPr(Ind), mxGetPr(Rij), mxGetPr(Ro), mxGetPr(Err), mxGetPr(Rij0), NULL);
        desc_b200_model_destroy(mo);
        fail_if(rc, NULL);
        mxArray* Adj = mxCreateDoubleMatrix(n, n, mxREAL);            /* Uniform_Topology.m:32 */
        for (mwSize e = 0; e < m; e++) {
            const mwSize i = (mwSize)mxGetPr(Ind)[e] - 1, j = (mwSize)mxGetPr(Ind)[e + m] - 1;
            mxGetPr(Adj)[i + n * j] = 1.0;
            mxGetPr(Adj)[j + n * i] = 1.0;
        }
        const char* fields[] = {"AdjMat", "Ind", "RijMat", "Rij_orig", "R_orig", "ErrVec"};
        plhs[0] = mxCreateStructMatrix(1, 1, 6, fields);
        mxSetField(plhs[0], 0, "AdjMat", Adj);
        mxSetField(plhs[0], 0, "Ind", Ind);
        mxSetField(plhs[0], 0, "RijMat", Rij);
        mxSetField(plhs[0], 0, "Rij_orig", Rij0);
        mxSetField(plhs[0], 0, "R_orig", Ro);
        mxSetField(plhs[0], 0, "ErrVec", Err);
        return;
    }
    if (strcmp(cmd, "cemp") == 0) {
        if (nrhs != 8) mexErrMsgIdAndTxt("DESC:b200", "cemp: 7 arguments expected");
        mwSize m;
        check_inputs(prhs[1], prhs[2], &m);
        const int max_iter = (int)mxGetScalar(prhs[3]);
        if (!mxIsDouble(prhs[4]) || mxIsEmpty(prhs[4]))
            mexErrMsgIdAndTxt("DESC:b200", "CEMP_parameters.reweighting must be a non-empty double vector");
        const int nsample = (int)mxGetScalar(prhs[5]);
        const uint64_t seed = (uint64_t)mxGetScalar(prhs[6]);
        /* want_R: false / 0 = S_vec only (DESC_PGD.m), true / 1 = + GCW rotations (DESC_init.m), 2 = + the refinement
           stage DESC.m:265-312 on the SAME handle (DESC.m: no second upload of Ind / RijMat, no second graph build) */
        const int want_R = mxIsLogicalScalarTrue(prhs[7]) ? 1 : (mxIsNumeric(prhs[7]) ? (int)mxGetScalar(prhs[7]) : 0);
        if (nsample <= 0) mexErrMsgIdAndTxt("DESC:b200", "CEMP_parameters.nsample must be positive");
        desc_b200_handle* h = NULL;
        fail_if(desc_b200_create(&h, 0, (int64_t)m, mxGetPr(prhs[1]), mxGetPr(prhs[2]), NULL), NULL);
        int64_t info[10];
        fail_if(desc_b200_get_info(h, info), h);
        mxArray* S = mxCreateDoubleMatrix(1, m, mxREAL);               /* CEMP.m:101: 1 x m row */
        mxArray* R = NULL;
        if (want_R) {
            mwSize dims[3] = {3, 3, 0};
            dims[2] = (mwSize)info[0];
            R = mxCreateNumericArray(3, dims, mxDOUBLE_CLASS, mxREAL);
        } else {
            R = mxCreateDoubleMatrix(0, 0, mxREAL);
        }
        fail_if(desc_b200_build_incidence(h, nsample, seed, NULL, NULL), h);
        fail_if(desc_b200_cycle_inconsistency(h), h);
        fail_if(desc_b200_cemp(h, max_iter, mxGetPr(prhs[4]), (int32_t)mxGetNumberOfElements(prhs[4]), mxGetPr(S)), h);
        if (want_R) fail_if(desc_b200_cemp_gcw(h, NULL, mxGetPr(R)), h);
        desc_b200_destroy(h);
        const char* fields[] = {"SVec", "R_est"};
        plhs[0] = mxCreateStructMatrix(1, 1, 2, fields);
        mxSetField(plhs[0], 0, "SVec", S);
        mxSetField(plhs[0], 0, "R_est", R);
        return;
    }
    if (strcmp(cmd, "mpls") == 0) {
        if (nrhs != 12) mexErrMsgIdAndTxt("DESC:b200", "mpls: 11 arguments expected");
        mwSize m;
        check_inputs(prhs[1], prhs[2], &m);
        for (int a = 4; a <= 11; a += (a == 4 ? 5 : 1))
            if (!mxIsDouble(prhs[a]) || mxIsEmpty(prhs[a]))
                mexErrMsgIdAndTxt("DESC:b200", "reweighting / thresholding / cycle_info_ratio must be non-empty double vectors");
        const int cemp_iter = (int)mxGetScalar(prhs[3]);
        const int nsample = (int)mxGetScalar(prhs[5]);
        const uint64_t seed = (uint64_t)mxGetScalar(prhs[6]);
        desc_b200_mpls_params P;
        P.stop_threshold = mxGetScalar(prhs[7]);
        P.max_iter = (int32_t)mxGetScalar(prhs[8]);
        P.reweighting = mxGetPr(prhs[9]);
        P.n_reweighting = (int32_t)mxGetNumberOfElements(prhs[9]);
        P.thresholding = mxGetPr(prhs[10]);
        P.n_thresholding = (int32_t)mxGetNumberOfElements(prhs[10]);
        P.cycle_info_ratio = mxGetPr(prhs[11]);
        P.n_cycle_info_ratio = (int32_t)mxGetNumberOfElements(prhs[11]);
        if (nsample <= 0 || P.max_iter < 1) mexErrMsgIdAndTxt("DESC:b200", "nsample and MPLS max_iter must be positive");
        desc_b200_handle* h = NULL;
        fail_if(desc_b200_create(&h, 0, (int64_t)m, mxGetPr(prhs[1]), mxGetPr(prhs[2]), NULL), NULL);
        int64_t info[10];
        fail_if(desc_b200_get_info(h, info), h);
        mwSize dims[3] = {3, 3, 0};
        dims[2] = (mwSize)info[0];
        mxArray* R = mxCreateNumericArray(3, dims, mxDOUBLE_CLASS, mxREAL);
        mxArray* R0 = mxCreateNumericArray(3, dims, mxDOUBLE_CLASS, mxREAL);
        mxArray* sc = mxCreateDoubleMatrix(P.max_iter, 1, mxREAL);
        int32_t run = 0;
        fail_if(desc_b200_build_incidence(h, nsample, seed, NULL, NULL), h);
        fail_if(desc_b200_cycle_inconsistency(h), h);
        fail_if(desc_b200_cemp(h, cemp_iter, mxGetPr(prhs[4]), (int32_t)mxGetNumberOfElements(prhs[4]), NULL), h);
        fail_if(desc_b200_mst_init(h, NULL, mxGetPr(R0)), h);
        fail_if(desc_b200_mpls_refine(h, NULL, NULL, &P, mxGetPr(R), &run, mxGetPr(sc)), h);
        desc_b200_destroy(h);
        mxArray* scores = mxCreateDoubleMatrix(run, 1, mxREAL);
        for (int t = 0; t < run; t++) mxGetPr(scores)[t] = mxGetPr(sc)[t];
        mxDestroyArray(sc);
        const char* fields[] = {"R_est", "R_init", "scores"};
        plhs[0] = mxCreateStructMatrix(1, 1, 3, fields);
        mxSetField(plhs[0], 0, "R_est", R);
        mxSetField(plhs[0], 0, "R_init", R0);
        mxSetField(plhs[0], 0, "scores", scores);
        return;
    }
    if (strcmp(cmd, "spectral") == 0) {
        if (nrhs != 3) mexErrMsgIdAndTxt("DESC:b200", "spectral: 2 arguments expected");
        mwSize m;
        check_inputs(prhs[1], prhs[2], &m);
        desc_b200_handle* h = NULL;
        fail_if(desc_b200_create(&h, 0, (int64_t)m, mxGetPr(prhs[1]), mxGetPr(prhs[2]), NULL), NULL);
        int64_t info[10];
        fail_if(desc_b200_get_info(h, info), h);
        mwSize dims[3] = {3, 3, 0};
        dims[2] = (mwSize)info[0];
        plhs[0] = mxCreateNumericArray(3, dims, mxDOUBLE_CLASS, mxREAL);
        fail_if(desc_b200_spectral(h, mxGetPr(plhs[0])), h);
        desc_b200_destroy(h);
        return;
    }
    if (strcmp(cmd, "align") == 0) {
        if (nrhs != 3) mexErrMsgIdAndTxt("DESC:b200", "align: 2 arguments expected");
        const mwSize ne = mxGetNumberOfElements(prhs[1]);
        if (!mxIsDouble(prhs[1]) || !mxIsDouble(prhs[2]) || ne != mxGetNumberOfElements(prhs[2]) || ne % 9 != 0 || ne < 18)
            mexErrMsgIdAndTxt("DESC:b200", "R_est and R_gt must both be 3 x 3 x n, n >= 2");
        const mwSize n = ne / 9;
        /* the metric needs no graph: a path on n nodes with identity edges carries the device context */
        mxArray* Ind = mxCreateDoubleMatrix(n - 1, 2, mxREAL);
        mwSize dims[3] = {3, 3, 0};
        dims[2] = n - 1;
        mxArray* Rij = mxCreateNumericArray(3, dims, mxDOUBLE_CLASS, mxREAL);
        for (mwSize e = 0; e + 1 < n; e++) {
            mxGetPr(Ind)[e] = (double)(e + 1);
            mxGetPr(Ind)[e + (n - 1)] = (double)(e + 2);
            mxGetPr(Rij)[9 * e] = mxGetPr(Rij)[9 * e + 4] = mxGetPr(Rij)[9 * e + 8] = 1.0;
        }
        desc_b200_handle* h = NULL;
        fail_if(desc_b200_create(&h, 0, (int64_t)(n - 1), mxGetPr(Ind), mxGetPr(Rij), NULL), NULL);
        dims[2] = n;
        mxArray* R_out = mxCreateNumericArray(3, dims, mxDOUBLE_CLASS, mxREAL);
        mxArray* R_align = mxCreateDoubleMatrix(3, 3, mxREAL);
        double mean_error = 0.0, median_error = 0.0;
        fail_if(desc_b200_rotation_alignment(h, mxGetPr(prhs[1]), mxGetPr(prhs[2]), mxGetPr(R_out), mxGetPr(R_align),
                                             &mean_error, &median_error), h);
        desc_b200_destroy(h);
        mxDestroyArray(Ind);
        mxDestroyArray(Rij);
        const char* fields[] = {"R_out", "R_align", "mean_error", "median_error"};
        plhs[0] = mxCreateStructMatrix(1, 1, 4, fields);
        mxSetField(plhs[0], 0, "R_out", R_out);
        mxSetField(plhs[0], 0, "R_align", R_align);
        mxSetField(plhs[0], 0, "mean_error", mxCreateDoubleScalar(mean_error));
        mxSetField(plhs[0], 0, "median_error", mxCreateDoubleScalar(median_error));
        return;
    }
    mexErrMsgIdAndTxt("DESC:b200", "unknown command '%s'", cmd);
}
