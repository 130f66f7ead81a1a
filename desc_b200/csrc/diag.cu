// the evaluation / diagnostics path on the device.
//
//  * Utils/Rotation_Alignment.m:13-38 (== Utils/GlobalSOdCorrectRight.m:27-50, the function DESC.m:238 calls):
//      A = sum_k R_est_k' R_gt_k;  [U1,~,V1] = svd(A);  R_align = U1 diag(1,1,det(U1 V1')) V1'
//      R_out_k = R_est_k R_align;  MSE_k = abs(acos((trace(R_gt_k R_out_k') - 1)/2))/pi*180;  mean, median
//  * the make_plots branch of the PGD loop (DESC.m:235-239): per iteration
//      svec_errors = mean(abs(ErrVec - S_vec));  R_est = GCW(...S_vec);  [~, MSE_mean, MSE_median] = align(R_est, R_orig)
//
// All reductions use a fixed grid and a fixed summation order (bit-reproducible).  Bytes: 144*n per
// alignment (two reads of both rotation sets) -- negligible next to the GCW it follows.
#include "internal.cuh"
#include "so3.cuh"

#include <algorithm>
#include <cmath>

namespace {
constexpr int DG_BLOCKS = DESC_SMS;
constexpr int DG_TB = 256;
// diag_work layout (doubles)
constexpr int DW_ALIGN = 0;      // 9: R_align
constexpr int DW_SUM = 9;        // 1: result of the last sum
constexpr int DW_PART = 16;      // DG_BLOCKS*9 partials
constexpr int DW_ERRS = DW_PART + DG_BLOCKS * 9;   // n per-node errors

template <int K>
__device__ __forceinline__ void dg_block_reduce(double (&v)[K], double* __restrict__ out) {
    __shared__ double sh[K][DG_TB / 32];
#pragma unroll
    for (int q = 0; q < K; q++) {
        v[q] = group_sum<32>(v[q]);
        if ((threadIdx.x & 31) == 0) sh[q][threadIdx.x >> 5] = v[q];
    }
    __syncthreads();
    if (threadIdx.x < K) {
        double t = 0.0;
        for (int w = 0; w < DG_TB / 32; w++) t += sh[threadIdx.x][w];
        out[threadIdx.x] = t;
    }
}

// partial[b*9 + (a+3c)] = sum over the block's nodes of (R_est_k' R_gt_k)(a,c)     Rotation_Alignment.m:17-21
__global__ void __launch_bounds__(DG_TB)
k_align_partial(const double* __restrict__ Re, const double* __restrict__ Rg, int n, double* __restrict__ partial) {
    double acc[9];
#pragma unroll
    for (int q = 0; q < 9; q++) acc[q] = 0.0;
    for (int k = blockIdx.x * DG_TB + threadIdx.x; k < n; k += DG_BLOCKS * DG_TB) {
        double u[9], v[9];
#pragma unroll
        for (int q = 0; q < 9; q++) {
            u[q] = Re[9 * (int64_t)k + q];
            v[q] = Rg[9 * (int64_t)k + q];
        }
#pragma unroll
        for (int a = 0; a < 3; a++)
#pragma unroll
            for (int c = 0; c < 3; c++)
                acc[a + 3 * c] += u[0 + 3 * a] * v[0 + 3 * c] + u[1 + 3 * a] * v[1 + 3 * c] + u[2 + 3 * a] * v[2 + 3 * c];
    }
    dg_block_reduce<9>(acc, partial + 9 * blockIdx.x);
}

__global__ void k_align_finish(const double* __restrict__ partial, double* __restrict__ R_align) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double A[9];
    for (int q = 0; q < 9; q++) {
        double t = 0.0;
        for (int b = 0; b < DG_BLOCKS; b++) t += partial[9 * b + q];
        A[q] = t;
    }
    proj_so3_dev(A, R_align);   // :23-24
}

// errs[k] in degrees (:28-33); Rout (may be null) = R_est_k R_align
__global__ void k_align_errors(const double* __restrict__ Re, const double* __restrict__ Rg,
                               const double* __restrict__ R_align, int n, double* __restrict__ errs,
                               double* __restrict__ Rout) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    double u[9], q[9];
#pragma unroll
    for (int x = 0; x < 9; x++) {
        u[x] = Re[9 * (int64_t)k + x];
        q[x] = R_align[x];
    }
    double tr = 0.0;
#pragma unroll
    for (int c = 0; c < 3; c++)
#pragma unroll
        for (int r = 0; r < 3; r++) {
            const double o = u[r + 0] * q[0 + 3 * c] + u[r + 3] * q[1 + 3 * c] + u[r + 6] * q[2 + 3 * c];
            if (Rout) Rout[9 * (int64_t)k + r + 3 * c] = o;
            tr += Rg[9 * (int64_t)k + r + 3 * c] * o;     // trace(R_gt * R_out')
        }
    errs[k] = abs_acos_dev((tr - 1.0) / 2.0) / 3.14159265358979323846 * 180.0;
}

// partial[b] = sum of a[e] (b == null) or |a[e] - b[e]| over the block's elements
__global__ void __launch_bounds__(DG_TB)
k_absdiff_partial(const double* __restrict__ a, const double* __restrict__ b, int64_t count, double* __restrict__ partial) {
    double acc[1] = {0.0};
    for (int64_t e = blockIdx.x * (int64_t)DG_TB + threadIdx.x; e < count; e += (int64_t)DG_BLOCKS * DG_TB)
        acc[0] += b ? fabs(a[e] - b[e]) : a[e];
    dg_block_reduce<1>(acc, partial + blockIdx.x);
}
__global__ void k_sum_finish(const double* __restrict__ partial, double scale, double* __restrict__ out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double t = 0.0;
    for (int b = 0; b < DG_BLOCKS; b++) t += partial[b];
    out[0] = t * scale;
}

int diag_scratch(desc_b200_handle* h) {
    if (!h->diag_work) CUDA_TRY(cudaMalloc(&h->diag_work, (size_t)(DW_ERRS + h->n + 8) * sizeof(double)));
    if (!h->diag_hist) CUDA_TRY(cudaMalloc(&h->diag_hist, 2048 * sizeof(unsigned)));
    return DESC_B200_OK;
}

// mean of a (b == null) or of |a - b| over `count` elements -> host
int diag_mean(desc_b200_handle* h, const double* a, const double* b, int64_t count, double* out) {
    double* W = h->diag_work;
    k_absdiff_partial<<<DG_BLOCKS, DG_TB, 0, h->stream>>>(a, b, count, W + DW_PART);
    KERNEL_CHECK(h);
    k_sum_finish<<<1, 32, 0, h->stream>>>(W + DW_PART, 1.0 / (double)count, W + DW_SUM);
    KERNEL_CHECK(h);
    CUDA_TRY(cudaMemcpyAsync(out, W + DW_SUM, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return DESC_B200_OK;
}
}  // namespace

int desc_align_impl(desc_b200_handle* h, const double* d_Rest, const double* d_Rgt, double* d_Rout, double out[11]) {
    const int n = h->n;
    DESC_TRY(diag_scratch(h));
    double* W = h->diag_work;
    cudaStream_t st = h->stream;
    k_align_partial<<<DG_BLOCKS, DG_TB, 0, st>>>(d_Rest, d_Rgt, n, W + DW_PART);
    KERNEL_CHECK(h);
    k_align_finish<<<1, 32, 0, st>>>(W + DW_PART, W + DW_ALIGN);
    KERNEL_CHECK(h);
    k_align_errors<<<(n + 127) / 128, 128, 0, st>>>(d_Rest, d_Rgt, W + DW_ALIGN, n, W + DW_ERRS, d_Rout);
    KERNEL_CHECK(h);
    CUDA_TRY(cudaMemcpyAsync(out + 2, W + DW_ALIGN, 9 * sizeof(double), cudaMemcpyDeviceToHost, st));
    DESC_TRY(diag_mean(h, W + DW_ERRS, nullptr, n, &out[0]));                  // :35
    // MATLAB median (:36): middle order statistic, or the mean of the two middle ones
    double lo = 0.0, hi = 0.0;
    DESC_TRY(desc_select_kth(h, W + DW_ERRS, n, (n + 1) / 2, h->diag_hist, &lo));
    hi = lo;
    if (n % 2 == 0) DESC_TRY(desc_select_kth(h, W + DW_ERRS, n, n / 2 + 1, h->diag_hist, &hi));
    out[1] = (lo + hi) / 2.0;
    return DESC_B200_OK;
}

int desc_diag_record(desc_b200_handle* h, int t, const double* d_S) {
    if (t < 1 || t > h->diag_cap || !h->diag_out) return DESC_B200_OK;
    DESC_TRY(diag_scratch(h));
    double* row = h->diag_out + 3 * (size_t)(t - 1);
    DESC_TRY(diag_mean(h, h->diag_err, d_S, h->m, &row[0]));                    // DESC.m:236
    const int rule = h->gcw_weight_rule;
    h->gcw_weight_rule = 0;
    const int rc = desc_gcw_impl(h, d_S);                                       // DESC.m:237
    h->gcw_weight_rule = rule;
    if (rc == DESC_B200_ERR_NOCONV) {   // keep iterating: the diagnostics of this iteration are undefined
        row[1] = row[2] = NAN;
        return DESC_B200_OK;
    }
    DESC_TRY(rc);
    double a[11];
    DESC_TRY(desc_align_impl(h, h->R_est, h->diag_Rgt, nullptr, a));            // DESC.m:238
    row[1] = a[0];
    row[2] = a[1];
    return DESC_B200_OK;
}
