/*
 * desc_b200_mex.c -- thin MEX gateway: MATLAB host code -> C ABI (include/desc_b200.h) -> CUDA.
 *
 *   out = desc_b200_mex('solve', Ind, RijMat, iters, rule, n_sample, seed, want_R)
 *   R   = desc_b200_mex('gcw',   Ind, RijMat, S_vec)
 *   out = desc_b200_mex('refine', Ind, RijMat, S_vec, R_init)    (DESC.m:265-312; out.R_est, out.scores)
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
        if (nrhs != 8) mexErrMsgIdAndTxt("DESC:b200", "solve: 7 arguments expected");
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
        const int want_R = mxIsLogicalScalarTrue(prhs[7]) || (mxIsNumeric(prhs[7]) && mxGetScalar(prhs[7]) != 0);

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
        fail_if(desc_b200_solve(h, n_sample, seed, iters, &rule, mxGetPr(S), want_R ? mxGetPr(R) : NULL,
                                mxGetPr(hist), &iters_run), h);
        fail_if(desc_b200_get_info(h, info), h);
        desc_b200_destroy(h);

        /* hist comes back as rows [change, objective] per iteration = 2 x iters column-major */
        mxArray* histT = mxCreateDoubleMatrix(iters_run, 2, mxREAL);
        for (int t = 0; t < iters_run; t++) {
            mxGetPr(histT)[t] = mxGetPr(hist)[2 * t];
            mxGetPr(histT)[t + iters_run] = mxGetPr(hist)[2 * t + 1];
        }
        mxDestroyArray(hist);
        const char* fields[] = {"S_vec", "R_est", "hist", "iters_run", "t", "n_sample", "m_cycle"};
        plhs[0] = mxCreateStructMatrix(1, 1, 7, fields);
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
    mexErrMsgIdAndTxt("DESC:b200", "unknown command '%s'", cmd);
}
