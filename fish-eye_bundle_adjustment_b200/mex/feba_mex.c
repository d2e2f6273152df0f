/*
 * feba_mex.c -- MEX gateway: MATLAB <-> the C ABI of include/feba.h (libfeba.so).
 *
 * Build (where MATLAB exists):  mex -I<repo>/include feba_mex.c -L<repo>/fish-eye_bundle_adjustment_b200 -lfeba
 * Usage from the reference's main.m (see INTEGRATION.md and feba_main_loop.m):
 *     h        = feba_mex('create', S);          S: packed numeric struct built once after main.m:384
 *                feba_mex('set_xhat', h, xhat);  after Buildxhat (main.m:388)
 *     deltasum = feba_mex('iterate', h);         replaces the while-body main.m:416-487
 *     xhat     = feba_mex('get_xhat', h);
 *     [v, rsd, stats] = feba_mex('residuals', h);   replaces main.m:569-601 (rsd: n_obs x 5 = r vx vy vr vt)
 *     q = feba_mex('cov_diag', h);  B = feba_mex('cov_block', h, idx);   diag(Cx)/sigma02, Cx/sigma02 blocks
 *                feba_mex('destroy', h);
 *     [S, err] = feba_mex('pack', pho, ext, cnt, int, tie, NK, allgcp);   native problem build (feba_pack.h):
 *                the numeric arrays of S straight from the five text files, instead of main.m:196-384
 *                (tie = '' when there is no .tie file); add the settings fields and pass S to 'create'
 *   BatchRun sweep (BatchRun.m:57-65: many independent blocks), all blocks advanced by ONE graph launch per step:
 *     b = feba_mex('batch_create', hs);          hs: uint64 vector of handles (one per block)
 *         feba_mex('batch_iterate', b);          one Gauss-Newton step of every block, asynchronous
 *     deltasum = feba_mex('sync', h);            per block: wait + main.m:487 of the step just issued
 *         feba_mex('batch_destroy', b);          (handles stay valid; drop converged blocks by making a new batch)
 *     feba_mex('iterate_async', h) / 'sync' do the same for one handle
 *     [info, flop] = feba_mex('plan_info', h);   row order / pattern chosen for the reduced system (feba_plan_info)
 * Recoverable errors are NOT raised with mexErrMsgIdAndTxt: like the reference's 0/1 `error` flags
 * (main.m:417-421) the gateway returns the status as the LAST output and prints feba_last_error.
 * The handle travels as a uint64 scalar; all arrays stay owned by MATLAB.
 */
#include <string.h>

#include "feba.h"
#include "feba_pack.h"
#include "mex.h"

static feba_handle* get_handle(const mxArray* a) {
    if (!mxIsClass(a, "uint64") || mxGetNumberOfElements(a) != 1)
        mexErrMsgIdAndTxt("feba:handle", "handle must be a uint64 scalar");
    return (feba_handle*)(uintptr_t)(*(uint64_t*)mxGetData(a));
}

static const mxArray* field(const mxArray* s, const char* name) {
    const mxArray* f = mxGetField(s, 0, name);
    if (!f) mexErrMsgIdAndTxt("feba:field", "missing field %s", name);
    return f;
}

static const double* dfield(const mxArray* s, const char* name, size_t count) {
    const mxArray* f = field(s, name);
    if (!mxIsDouble(f) || mxGetNumberOfElements(f) != count)
        mexErrMsgIdAndTxt("feba:field", "field %s must be double with %d elements", name, (int)count);
    return mxGetPr(f);
}

static const int32_t* ifield(const mxArray* s, const char* name, size_t count) {
    const mxArray* f = field(s, name);
    if (!mxIsClass(f, "int32") || mxGetNumberOfElements(f) != count)
        mexErrMsgIdAndTxt("feba:field", "field %s must be int32 with %d elements", name, (int)count);
    return (const int32_t*)mxGetData(f);
}

static double sfield(const mxArray* s, const char* name) { return mxGetScalar(field(s, name)); }

static void status_out(int nlhs, mxArray* plhs[], int slot, int rc, const feba_handle* h) {
    if (rc) mexPrintf("%s\n", feba_last_error(h));      /* main.m:418 disp('Error building A and w') follows */
    if (nlhs > slot) plhs[slot] = mxCreateDoubleScalar((double)(rc != 0));
}

void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
    char cmd[32];
    if (nrhs < 1 || mxGetString(prhs[0], cmd, sizeof(cmd))) mexErrMsgIdAndTxt("feba:cmd", "first argument: command");
    if (!strcmp(cmd, "create")) {
        /* S: n_obs n_img n_cam n_pts n_tie | obs_x obs_y (double) obs_img obs_pt (int32, 0-based) img_cam
         * pt_tie (int32) | eop0 (6 x n_img) iop0 ((3+NK+2) x n_cam) cam_box (5 x n_cam) xyz0 (3 x n_pts),
         * column-major = the row-major layout of feba.h | settings fields */
        if (nrhs < 2 || !mxIsStruct(prhs[1])) mexErrMsgIdAndTxt("feba:create", "create needs the packed problem struct");
        const mxArray* S = prhs[1];
        feba_problem p;
        memset(&p, 0, sizeof(p));
        p.n_obs = (int64_t)sfield(S, "n_obs");
        p.n_img = (int32_t)sfield(S, "n_img");
        p.n_cam = (int32_t)sfield(S, "n_cam");
        p.n_pts = (int32_t)sfield(S, "n_pts");
        p.n_tie = (int32_t)sfield(S, "n_tie");
        feba_settings* s = &p.settings;
        const double* ee = dfield(S, "estimate_eop", 6);
        for (int q = 0; q < 6; ++q) s->estimate_eop[q] = (int32_t)ee[q];
        s->estimate_xp = (int32_t)sfield(S, "Estimate_xp");
        s->estimate_yp = (int32_t)sfield(S, "Estimate_yp");
        s->estimate_c = (int32_t)sfield(S, "Estimate_c");
        s->estimate_radial = (int32_t)sfield(S, "Estimate_radial");
        s->num_radial = (int32_t)sfield(S, "Num_Radial_Distortions");
        if (s->num_radial < 1) s->num_radial = 1;                           /* BuildAwG.m:18-20 */
        s->estimate_decent = (int32_t)sfield(S, "Estimate_decent");
        s->inner_constraints = (int32_t)sfield(S, "Inner_Constraints");
        s->type = (int32_t)sfield(S, "typeint");
        s->iteration_cap = (int32_t)sfield(S, "Iteration_Cap");
        s->sigma_x = sfield(S, "Meas_std");
        s->sigma_y = sfield(S, "Meas_std_y");
        s->threshold = sfield(S, "threshold");
        const size_t n = (size_t)p.n_obs, NC = (size_t)s->num_radial + 5;
        p.obs_x = dfield(S, "obs_x", n);
        p.obs_y = dfield(S, "obs_y", n);
        p.obs_img = ifield(S, "obs_img", n);
        p.obs_pt = ifield(S, "obs_pt", n);
        p.img_cam = ifield(S, "img_cam", (size_t)p.n_img);
        p.pt_tie = ifield(S, "pt_tie", (size_t)p.n_pts);
        p.eop0 = dfield(S, "eop0", 6 * (size_t)p.n_img);
        p.iop0 = dfield(S, "iop0", NC * (size_t)p.n_cam);
        p.cam_box = dfield(S, "cam_box", 5 * (size_t)p.n_cam);
        p.xyz0 = dfield(S, "xyz0", 3 * (size_t)p.n_pts);
        feba_handle* h = NULL;
        const int rc = feba_create(&p, &h);
        plhs[0] = mxCreateNumericMatrix(1, 1, mxUINT64_CLASS, mxREAL);
        *(uint64_t*)mxGetData(plhs[0]) = (uint64_t)(uintptr_t)h;
        status_out(nlhs, plhs, 1, rc, NULL);
        return;
    }
    if (!strcmp(cmd, "pack")) {
        /* main.m:196-384 in one call: tokenise the files, str2double, resolve image / camera / target /
         * tie IDs (first match wins).  Index fields are 0-based int32 as 'create' expects them. */
        char path[5][1024];
        if (nrhs < 8) mexErrMsgIdAndTxt("feba:pack", "pack needs pho, ext, cnt, int, tie, NK, allgcp");
        for (int k = 0; k < 5; ++k)
            if (mxGetString(prhs[1 + k], path[k], sizeof(path[k]))) mexErrMsgIdAndTxt("feba:pack", "file names must be char");
        feba_pack* pk = NULL;
        const int rc = feba_pack_read(path[0], path[1], path[2], path[3], path[4][0] ? path[4] : NULL,
                                      (int32_t)mxGetScalar(prhs[6]), (int32_t)mxGetScalar(prhs[7]), 0, &pk);
        if (rc) {
            mexPrintf("%s\n", feba_pack_last_error());       /* the text of main.m's errordlg (main.m:293,316,352) */
            plhs[0] = mxCreateDoubleMatrix(0, 0, mxREAL);
            if (nlhs > 1) plhs[1] = mxCreateDoubleScalar(1.0);
            return;
        }
        feba_pack_view v;
        feba_pack_get(pk, &v);
        static const char* names[] = {"n_obs", "n_img", "n_cam", "n_pts", "n_tie", "obs_x", "obs_y", "obs_img", "obs_pt",
                                      "img_cam", "pt_tie", "tie_pt", "eop0", "iop0", "cam_box", "xyz0"};
        mxArray* S = mxCreateStructMatrix(1, 1, 16, names);
        const double counts[5] = {(double)v.n_obs, v.n_img, v.n_cam, v.n_pts, v.n_tie};
        for (int k = 0; k < 5; ++k) mxSetField(S, 0, names[k], mxCreateDoubleScalar(counts[k]));
        const struct { const char* name; const double* src; size_t rows, cols; } dd[] = {
            {"obs_x", v.obs_x, (size_t)v.n_obs, 1}, {"obs_y", v.obs_y, (size_t)v.n_obs, 1},
            {"eop0", v.eop0, 6, (size_t)v.n_img}, {"iop0", v.iop0, (size_t)v.n_iop_cols, (size_t)v.n_cam},
            {"cam_box", v.cam_box, 5, (size_t)v.n_cam}, {"xyz0", v.xyz0, 3, (size_t)v.n_pts}};
        for (size_t k = 0; k < sizeof(dd) / sizeof(dd[0]); ++k) {    /* row-major C = column-major MATLAB transposed */
            mxArray* a = mxCreateDoubleMatrix((mwSize)dd[k].rows, (mwSize)dd[k].cols, mxREAL);
            if (dd[k].rows && dd[k].cols) memcpy(mxGetPr(a), dd[k].src, dd[k].rows * dd[k].cols * sizeof(double));
            mxSetField(S, 0, dd[k].name, a);
        }
        const struct { const char* name; const int32_t* src; size_t count; } ii[] = {
            {"obs_img", v.obs_img, (size_t)v.n_obs}, {"obs_pt", v.obs_pt, (size_t)v.n_obs},
            {"img_cam", v.img_cam, (size_t)v.n_img}, {"pt_tie", v.pt_tie, (size_t)v.n_pts},
            {"tie_pt", v.tie_pt, (size_t)v.n_tie}};
        for (size_t k = 0; k < sizeof(ii) / sizeof(ii[0]); ++k) {
            mxArray* a = mxCreateNumericMatrix((mwSize)ii[k].count, 1, mxINT32_CLASS, mxREAL);
            if (ii[k].count) memcpy(mxGetData(a), ii[k].src, ii[k].count * sizeof(int32_t));
            mxSetField(S, 0, ii[k].name, a);
        }
        feba_pack_free(pk);
        plhs[0] = S;
        if (nlhs > 1) plhs[1] = mxCreateDoubleScalar(0.0);
        return;
    }
    if (!strcmp(cmd, "batch_create")) {
        if (nrhs < 2 || !mxIsClass(prhs[1], "uint64") || mxGetNumberOfElements(prhs[1]) < 1)
            mexErrMsgIdAndTxt("feba:batch", "batch_create needs a uint64 vector of handles");
        const size_t n = mxGetNumberOfElements(prhs[1]);
        const uint64_t* raw = (const uint64_t*)mxGetData(prhs[1]);
        feba_handle** hs = (feba_handle**)mxMalloc(sizeof(feba_handle*) * n);
        for (size_t i = 0; i < n; ++i) hs[i] = (feba_handle*)(uintptr_t)raw[i];
        feba_batch* b = NULL;
        const int rc = feba_batch_create(hs, (int32_t)n, &b);
        mxFree(hs);
        plhs[0] = mxCreateNumericMatrix(1, 1, mxUINT64_CLASS, mxREAL);
        *(uint64_t*)mxGetData(plhs[0]) = (uint64_t)(uintptr_t)b;
        if (rc != 0) mexPrintf("feba: batch_create failed (%d): blocks must be idle handles of one device\n", rc);
        if (nlhs > 1) plhs[1] = mxCreateDoubleScalar((double)(rc != 0));
        return;
    }
    if (!strcmp(cmd, "batch_iterate") || !strcmp(cmd, "batch_destroy")) {
        if (nrhs < 2 || !mxIsClass(prhs[1], "uint64") || mxGetNumberOfElements(prhs[1]) != 1)
            mexErrMsgIdAndTxt("feba:batch", "%s needs the batch handle", cmd);
        feba_batch* b = (feba_batch*)(uintptr_t)(*(uint64_t*)mxGetData(prhs[1]));
        if (cmd[6] == 'd') {
            feba_batch_destroy(b);
        } else {
            const int rc = feba_batch_iterate_async(b);
            if (rc != 0) mexPrintf("feba: batch_iterate failed (%d)\n", rc);
            if (nlhs > 0) plhs[0] = mxCreateDoubleScalar((double)(rc != 0));
        }
        return;
    }
    if (nrhs < 2) mexErrMsgIdAndTxt("feba:cmd", "%s needs a handle", cmd);
    feba_handle* h = get_handle(prhs[1]);
    int64_t u = 0, uc = 0;
    feba_num_unknowns(h, &u, &uc);
    if (!strcmp(cmd, "destroy")) {
        feba_destroy(h);
    } else if (!strcmp(cmd, "set_xhat")) {
        if (nrhs < 3 || !mxIsDouble(prhs[2])) mexErrMsgIdAndTxt("feba:xhat", "set_xhat needs a double vector");
        status_out(nlhs, plhs, 0, feba_set_xhat(h, mxGetPr(prhs[2]), mxGetNumberOfElements(prhs[2])), h);
    } else if (!strcmp(cmd, "get_xhat") || !strcmp(cmd, "get_delta")) {
        plhs[0] = mxCreateDoubleMatrix((mwSize)u, 1, mxREAL);
        const int rc = cmd[4] == 'x' ? feba_get_xhat(h, mxGetPr(plhs[0]), (size_t)u)
                                     : feba_get_delta(h, mxGetPr(plhs[0]), (size_t)u);
        status_out(nlhs, plhs, 1, rc, h);
    } else if (!strcmp(cmd, "iterate")) {
        double deltasum = 0.0;
        const int rc = feba_iterate(h, &deltasum);
        plhs[0] = mxCreateDoubleScalar(deltasum);                          /* main.m:487 */
        status_out(nlhs, plhs, 1, rc, h);
    } else if (!strcmp(cmd, "iterate_async")) {
        status_out(nlhs, plhs, 0, feba_iterate_async(h), h);
    } else if (!strcmp(cmd, "sync")) {
        double deltasum = 0.0;
        const int rc = feba_sync(h, &deltasum);
        plhs[0] = mxCreateDoubleScalar(deltasum);                          /* main.m:487 */
        status_out(nlhs, plhs, 1, rc, h);
    } else if (!strcmp(cmd, "plan_info")) {
        int32_t info[8];
        double flop[2];
        const int rc = feba_plan_info(h, info, flop);
        plhs[0] = mxCreateDoubleMatrix(1, 8, mxREAL);
        for (int i = 0; i < 8; ++i) mxGetPr(plhs[0])[i] = rc == 0 ? (double)info[i] : 0.0;
        if (nlhs > 1) {
            plhs[1] = mxCreateDoubleMatrix(1, 2, mxREAL);
            mxGetPr(plhs[1])[0] = rc == 0 ? flop[0] : 0.0;
            mxGetPr(plhs[1])[1] = rc == 0 ? flop[1] : 0.0;
        }
        status_out(nlhs, plhs, 2, rc, h);
    } else if (!strcmp(cmd, "solve")) {
        int32_t it = 0;
        const int cap = 4096;
        double* trace = (double*)mxMalloc(sizeof(double) * cap);
        const int rc = feba_solve(h, &it, trace, cap);
        plhs[0] = mxCreateDoubleScalar((double)it);
        if (nlhs > 1) {
            plhs[1] = mxCreateDoubleMatrix(1, (mwSize)(it < cap ? it : cap), mxREAL);   /* deltasumarr, main.m:488 */
            memcpy(mxGetPr(plhs[1]), trace, sizeof(double) * (size_t)(it < cap ? it : cap));
        }
        mxFree(trace);
        status_out(nlhs, plhs, 2, rc, h);
    } else if (!strcmp(cmd, "residuals")) {
        const int64_t n_obs = feba_num_obs(h);                             /* the handle's own count, never the caller's */
        if (n_obs < 0) mexErrMsgIdAndTxt("feba:handle", "residuals: invalid handle");
        const size_t n = (size_t)n_obs;
        plhs[0] = mxCreateDoubleMatrix((mwSize)(2 * n), 1, mxREAL);        /* v, main.m:569 */
        mxArray* rsd = mxCreateDoubleMatrix(5, (mwSize)n, mxREAL);         /* 5 x n_obs column-major = n_obs x 5 row-major */
        mxArray* st = mxCreateDoubleMatrix(1, 6, mxREAL);
        const int rc = feba_residuals(h, mxGetPr(plhs[0]), mxGetPr(rsd), mxGetPr(st));
        if (nlhs > 1) plhs[1] = rsd;
        if (nlhs > 2) plhs[2] = st;                                        /* RMSx RMSy RMS sigma02 sum_vx2 sum_vy2 */
        status_out(nlhs, plhs, 3, rc, h);
    } else if (!strcmp(cmd, "cov_diag")) {                                 /* diag(Cx)/sigma02, main.m:432-482 */
        plhs[0] = mxCreateDoubleMatrix((mwSize)u, 1, mxREAL);
        status_out(nlhs, plhs, 1, feba_cov_diag(h, mxGetPr(plhs[0]), (size_t)u), h);
    } else if (!strcmp(cmd, "cov_block")) {                                /* Cx/sigma02 block, 1-based indices */
        if (nrhs < 3 || !mxIsDouble(prhs[2])) mexErrMsgIdAndTxt("feba:cov", "cov_block needs an index vector");
        const int32_t k = (int32_t)mxGetNumberOfElements(prhs[2]);
        int64_t* idx = (int64_t*)mxMalloc(sizeof(int64_t) * (size_t)k);
        for (int32_t i = 0; i < k; ++i) idx[i] = (int64_t)mxGetPr(prhs[2])[i] - 1;
        plhs[0] = mxCreateDoubleMatrix((mwSize)k, (mwSize)k, mxREAL);      /* symmetric: layout-agnostic */
        const int rc = feba_cov_block(h, idx, k, mxGetPr(plhs[0]));
        mxFree(idx);
        status_out(nlhs, plhs, 1, rc, h);
    } else {
        mexErrMsgIdAndTxt("feba:cmd", "unknown command %s", cmd);
    }
}
