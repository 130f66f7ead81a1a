// A13: GCW weighted spectral rotation recovery (reference: Utils/GCW.m:1-38).
//
//   omega_e = 1/(s_e^1.5 + 1e-8)                      GCW.m:20
//   d_i     = sum_{e incident to i} omega_e           GCW.m:21 (row normalisation)
//   M       = D^-1 (W o R)   (3n x 3n, block (i,j) = omega_e/d_i * R_ij, (j,i) = omega_e/d_j * R_ij')
//   V       = top-3 eigenvectors of M ('la'), unit 2-norm columns      GCW.m:27
//   flip column 1 if det(V(1:3,:)) < 0                                  GCW.m:28
//   R_i     = U diag(1,1,det(U V')) V' with [U,~,V] = svd(V_i)          GCW.m:30-36
//
// The reference forms the dense 3n x 3n matrix and calls eigs.  Here M is never formed: it is
// similar to the symmetric N = D^-1/2 (W o R) D^-1/2, eigenvectors V = D^-1/2 U, and
// the top-3 invariant subspace of N is found by block power (subspace) iteration on the shifted
// operator (N + I)/2 -- spectrum in [0,1], order preserved, so "largest algebraic" is "largest
// magnitude" -- with a block-sparse SpMV over the symmetric CSR adjacency built in build.cu:
//
//   y_i = sum_{p in row i}  c_e * (i<j ? R_e : R_e') * x_j ,   c_e = omega_e / sqrt(d_i d_j)
//
// One warp per node, lanes over the node's edges; node-centric so no atomics and a fixed
// summation order (deterministic).  Each R_e is read twice per SpMV (once per endpoint).
// The 3n x 3 block is re-orthonormalised by Cholesky-QR every step; all 3x3 algebra runs in
// one-thread kernels so the loop needs no host round trip except the convergence poll.
#include "internal.cuh"

#include <algorithm>
#include <cmath>

#define GCW_RED_BLOCKS 148
#define GCW_RED_TB 256
#define GCW_NRED 16  // 9 (H = X'Y) + 6 (G = Y'Y upper) + 1 (residual)

// small[] layout (doubles)
#define SM_H 0      // 9: H = X' Y
#define SM_T 9      // 9: T with X_new = Y T (upper triangular inverse Cholesky factor transposed)
#define SM_Z 18     // 9: Ritz rotation
#define SM_RES 27   // 1: residual^2 of the previous apply
#define SM_NRM 28   // 3: column scale factors for the final V
#define SM_SGN 31   // 1: sign of column 1
#define SM_THETA 32 // 3: Ritz values of N
#define SM_FLAG 35  // 1: Cholesky breakdown flag
#define SM_SIZE 40

__global__ void k_gcw_weights(const double* __restrict__ S, int64_t m, double* __restrict__ omega) {
    int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e < m) omega[e] = 1.0 / (pow(S[e], 1.5) + 1e-8);  // GCW.m:20
}

// one warp per node: d_i, isd_i = 1/sqrt(d_i)
__global__ void k_gcw_degree(const int* __restrict__ rowstart, const int* __restrict__ adj_eid,
                             const double* __restrict__ omega, int n, double* __restrict__ isd) {
    const int node = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (node >= n) return;
    double d = 0.0;
    for (int p = rowstart[node] + lane; p < rowstart[node + 1]; p += 32) d += omega[adj_eid[p]];
    d = group_sum<32>(d);
    if (lane == 0) isd[node] = 1.0 / sqrt(d);
}

__global__ void k_gcw_coef(const int* __restrict__ ei, const int* __restrict__ ej,
                           const double* __restrict__ omega, const double* __restrict__ isd,
                           int64_t m, double* __restrict__ coef) {
    int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e < m) coef[e] = omega[e] * isd[ei[e]] * isd[ej[e]];
}

__global__ void k_gcw_init(double* __restrict__ X, int64_t n9) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= n9) return;
    uint64_t z = desc_key(0x6a09e667f3bcc908ull, (uint64_t)t, 7ull);
    X[t] = (double)(z >> 11) * (1.0 / 9007199254740992.0) - 0.5;
}

// Y[node] = 0.5 * (sum_p coef * op(R_e) * X[nbr] + X[node]) for node in [n0, n1)
__global__ void __launch_bounds__(256)
k_gcw_spmv(const int* __restrict__ rowstart, const int* __restrict__ adj_nbr,
           const int* __restrict__ adj_eid, const double* __restrict__ Rij,
           const double* __restrict__ coef, const double* __restrict__ X, double* __restrict__ Y,
           int n0, int n1) {
    const int node = n0 + ((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    if (node >= n1) return;
    double acc[9];
#pragma unroll
    for (int x = 0; x < 9; x++) acc[x] = 0.0;
    const int p1 = rowstart[node + 1];
    for (int p = rowstart[node] + lane; p < p1; p += 32) {
        const int nb = adj_nbr[p];
        const int e = adj_eid[p];
        const double c = coef[e];
        const double* pr = Rij + 9 * (int64_t)e;
        const double* px = X + 9 * (int64_t)nb;
        double r[9], x[9];
#pragma unroll
        for (int q = 0; q < 9; q++) {
            r[q] = __ldg(pr + q);
            x[q] = px[q];
        }
        if (node < nb) {
            // block (node, nb) = c * R_e : y(a,col) += c * sum_b R(a,b) x(b,col)
#pragma unroll
            for (int col = 0; col < 3; col++)
#pragma unroll
                for (int a = 0; a < 3; a++)
                    acc[a + 3 * col] += c * (r[a] * x[3 * col] + r[a + 3] * x[1 + 3 * col] + r[a + 6] * x[2 + 3 * col]);
        } else {
            // block (node, nb) = c * R_e'
#pragma unroll
            for (int col = 0; col < 3; col++)
#pragma unroll
                for (int a = 0; a < 3; a++)
                    acc[a + 3 * col] += c * (r[3 * a] * x[3 * col] + r[3 * a + 1] * x[1 + 3 * col] + r[3 * a + 2] * x[2 + 3 * col]);
        }
    }
#pragma unroll
    for (int x = 0; x < 9; x++) acc[x] = group_sum<32>(acc[x]);
    if (lane < 9) {
        double v = 0.0;
#pragma unroll
        for (int x = 0; x < 9; x++)
            if (lane == x) v = acc[x];
        Y[9 * (int64_t)node + lane] = 0.5 * (v + X[9 * (int64_t)node + lane]);
    }
}

template <int K>
__device__ __forceinline__ void block_reduce_store(double (&v)[K], double* __restrict__ out) {
    __shared__ double sh[GCW_RED_TB / 32][K];
#pragma unroll
    for (int x = 0; x < K; x++) v[x] = group_sum<32>(v[x]);
    if ((threadIdx.x & 31) == 0)
#pragma unroll
        for (int x = 0; x < K; x++) sh[threadIdx.x >> 5][x] = v[x];
    __syncthreads();
    if (threadIdx.x < K) {
        double s = 0.0;
        for (int w = 0; w < GCW_RED_TB / 32; w++) s += sh[w][threadIdx.x];
        out[threadIdx.x] = s;
    }
}

// partial[b][0..8] = X'Y, [9..14] = upper(Y'Y) over the nodes of block b
__global__ void __launch_bounds__(GCW_RED_TB)
k_gcw_reduce(const double* __restrict__ X, const double* __restrict__ Y, int n,
             double* __restrict__ partial) {
    double v[15];
#pragma unroll
    for (int x = 0; x < 15; x++) v[x] = 0.0;
    for (int node = blockIdx.x * blockDim.x + threadIdx.x; node < n; node += gridDim.x * blockDim.x) {
        double xb[9], yb[9];
#pragma unroll
        for (int q = 0; q < 9; q++) {
            xb[q] = X[9 * (int64_t)node + q];
            yb[q] = Y[9 * (int64_t)node + q];
        }
#pragma unroll
        for (int a = 0; a < 3; a++)
#pragma unroll
            for (int b = 0; b < 3; b++)
                v[a + 3 * b] += xb[3 * a] * yb[3 * b] + xb[3 * a + 1] * yb[3 * b + 1] + xb[3 * a + 2] * yb[3 * b + 2];
        int idx = 9;
#pragma unroll
        for (int a = 0; a < 3; a++)
#pragma unroll
            for (int b = a; b < 3; b++) {
                v[idx] += yb[3 * a] * yb[3 * b] + yb[3 * a + 1] * yb[3 * b + 1] + yb[3 * a + 2] * yb[3 * b + 2];
                idx++;
            }
    }
    block_reduce_store<15>(v, partial + (size_t)blockIdx.x * GCW_NRED);
}

// one thread: H, G from partials; Cholesky of G; T = inv(L)'; also folds the residual partials
// of the previous apply into small[SM_RES] and the history.
__global__ void k_gcw_small_orth(const double* __restrict__ partial, int nblocks,
                                 double* __restrict__ small, double* __restrict__ res_hist, int it) {
    double s[GCW_NRED];
    for (int x = 0; x < GCW_NRED; x++) s[x] = 0.0;
    for (int b = 0; b < nblocks; b++)
        for (int x = 0; x < GCW_NRED; x++) s[x] += partial[(size_t)b * GCW_NRED + x];
    for (int x = 0; x < 9; x++) small[SM_H + x] = s[x];
    const double g00 = s[9], g01 = s[10], g02 = s[11], g11 = s[12], g12 = s[13], g22 = s[14];
    // G = L L'
    double l00 = sqrt(g00);
    double l10 = g01 / l00, l20 = g02 / l00;
    double d1 = g11 - l10 * l10;
    double l11 = sqrt(d1);
    double l21 = (g12 - l20 * l10) / l11;
    double d2 = g22 - l20 * l20 - l21 * l21;
    double l22 = sqrt(d2);
    if (!(g00 > 0.0) || !(d1 > 0.0) || !(d2 > 0.0)) small[SM_FLAG] = 1.0;
    // Linv (lower)
    double i00 = 1.0 / l00, i11 = 1.0 / l11, i22 = 1.0 / l22;
    double i10 = -l10 * i00 * i11;
    double i21 = -l21 * i11 * i22;
    double i20 = -(l20 * i00 + l21 * i10) * i22;
    // T = Linv' (upper), column-major T[r + 3c]
    small[SM_T + 0] = i00; small[SM_T + 1] = 0.0; small[SM_T + 2] = 0.0;
    small[SM_T + 3] = i10; small[SM_T + 4] = i11; small[SM_T + 5] = 0.0;
    small[SM_T + 6] = i20; small[SM_T + 7] = i21; small[SM_T + 8] = i22;
    if (it >= 0) res_hist[it] = 0.0;  // filled by k_gcw_small_res
}

// X_new[node] = Y[node] * T ; residual partial = sum || Y - X H ||^2
__global__ void __launch_bounds__(GCW_RED_TB)
k_gcw_apply(double* __restrict__ X, const double* __restrict__ Y, int n,
            const double* __restrict__ small, double* __restrict__ partial) {
    double H[9], T[9];
#pragma unroll
    for (int q = 0; q < 9; q++) {
        H[q] = small[SM_H + q];
        T[q] = small[SM_T + q];
    }
    double v[1] = {0.0};
    for (int node = blockIdx.x * blockDim.x + threadIdx.x; node < n; node += gridDim.x * blockDim.x) {
        double xb[9], yb[9];
#pragma unroll
        for (int q = 0; q < 9; q++) {
            xb[q] = X[9 * (int64_t)node + q];
            yb[q] = Y[9 * (int64_t)node + q];
        }
#pragma unroll
        for (int c = 0; c < 3; c++)
#pragma unroll
            for (int r = 0; r < 3; r++) {
                const double xh = xb[r] * H[3 * c] + xb[r + 3] * H[1 + 3 * c] + xb[r + 6] * H[2 + 3 * c];
                const double d = yb[r + 3 * c] - xh;
                v[0] += d * d;
            }
#pragma unroll
        for (int c = 0; c < 3; c++)
#pragma unroll
            for (int r = 0; r < 3; r++)
                X[9 * (int64_t)node + r + 3 * c] =
                    yb[r] * T[3 * c] + yb[r + 3] * T[1 + 3 * c] + yb[r + 6] * T[2 + 3 * c];
    }
    block_reduce_store<1>(v, partial + (size_t)blockIdx.x * GCW_NRED + 15);
}

__global__ void k_gcw_small_res(const double* __restrict__ partial, int nblocks,
                                double* __restrict__ small, double* __restrict__ res_hist, int it) {
    double s = 0.0;
    for (int b = 0; b < nblocks; b++) s += partial[(size_t)b * GCW_NRED + 15];
    small[SM_RES] = s;
    res_hist[it] = sqrt(s);
}

// symmetric 3x3 eigen-decomposition by cyclic Jacobi; eigenvalues descending, Z columns
__device__ void jacobi_eig3(const double* A, double* evals, double* Z) {
    double a[3][3], z[3][3];
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) {
            a[r][c] = 0.5 * (A[r + 3 * c] + A[c + 3 * r]);
            z[r][c] = r == c ? 1.0 : 0.0;
        }
    for (int sweep = 0; sweep < 60; sweep++) {
        double off = fabs(a[0][1]) + fabs(a[0][2]) + fabs(a[1][2]);
        if (off == 0.0) break;
        for (int p = 0; p < 2; p++)
            for (int q = p + 1; q < 3; q++) {
                if (a[p][q] == 0.0) continue;
                double theta = (a[q][q] - a[p][p]) / (2.0 * a[p][q]);
                double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < 3; k++) {  // A <- A J
                    double akp = a[k][p], akq = a[k][q];
                    a[k][p] = c * akp - s * akq;
                    a[k][q] = s * akp + c * akq;
                }
                for (int k = 0; k < 3; k++) {  // A <- J' A
                    double apk = a[p][k], aqk = a[q][k];
                    a[p][k] = c * apk - s * aqk;
                    a[q][k] = s * apk + c * aqk;
                }
                for (int k = 0; k < 3; k++) {
                    double zkp = z[k][p], zkq = z[k][q];
                    z[k][p] = c * zkp - s * zkq;
                    z[k][q] = s * zkp + c * zkq;
                }
            }
    }
    int ord[3] = {0, 1, 2};
    for (int x = 0; x < 2; x++)
        for (int y = x + 1; y < 3; y++)
            if (a[ord[y]][ord[y]] > a[ord[x]][ord[x]]) {
                int t = ord[x];
                ord[x] = ord[y];
                ord[y] = t;
            }
    for (int c = 0; c < 3; c++) {
        evals[c] = a[ord[c]][ord[c]];
        for (int r = 0; r < 3; r++) Z[r + 3 * c] = z[r][ord[c]];
    }
}

// Rayleigh-Ritz on H = X'(N+I)/2 X (from partials): Z, theta (of N)
__global__ void k_gcw_small_ritz(const double* __restrict__ partial, int nblocks,
                                 double* __restrict__ small) {
    double s[9];
    for (int x = 0; x < 9; x++) s[x] = 0.0;
    for (int b = 0; b < nblocks; b++)
        for (int x = 0; x < 9; x++) s[x] += partial[(size_t)b * GCW_NRED + x];
    double ev[3], Z[9];
    jacobi_eig3(s, ev, Z);
    for (int x = 0; x < 9; x++) small[SM_Z + x] = Z[x];
    for (int x = 0; x < 3; x++) small[SM_THETA + x] = 2.0 * ev[x] - 1.0;
}

// V[node] = isd[node] * X[node] * Z ; partial column sums of squares
__global__ void __launch_bounds__(GCW_RED_TB)
k_gcw_ritz_apply(const double* __restrict__ X, const double* __restrict__ isd, int n,
                 const double* __restrict__ small, double* __restrict__ V,
                 double* __restrict__ partial) {
    double Z[9];
#pragma unroll
    for (int q = 0; q < 9; q++) Z[q] = small[SM_Z + q];
    double v[3] = {0.0, 0.0, 0.0};
    for (int node = blockIdx.x * blockDim.x + threadIdx.x; node < n; node += gridDim.x * blockDim.x) {
        double xb[9];
#pragma unroll
        for (int q = 0; q < 9; q++) xb[q] = X[9 * (int64_t)node + q];
        const double sc = isd[node];
#pragma unroll
        for (int c = 0; c < 3; c++)
#pragma unroll
            for (int r = 0; r < 3; r++) {
                const double o = sc * (xb[r] * Z[3 * c] + xb[r + 3] * Z[1 + 3 * c] + xb[r + 6] * Z[2 + 3 * c]);
                V[9 * (int64_t)node + r + 3 * c] = o;
                v[c] += o * o;
            }
    }
    block_reduce_store<3>(v, partial + (size_t)blockIdx.x * GCW_NRED);
}

// column norms + sign rule of GCW.m:28 (det of the first node's block after normalisation)
__global__ void k_gcw_small_final(const double* __restrict__ partial, int nblocks,
                                  const double* __restrict__ V, double* __restrict__ small) {
    double s[3] = {0.0, 0.0, 0.0};
    for (int b = 0; b < nblocks; b++)
        for (int x = 0; x < 3; x++) s[x] += partial[(size_t)b * GCW_NRED + x];
    double sc[3];
    for (int x = 0; x < 3; x++) {
        sc[x] = 1.0 / sqrt(s[x]);
        small[SM_NRM + x] = sc[x];
    }
    double a[9];
    for (int c = 0; c < 3; c++)
        for (int r = 0; r < 3; r++) a[r + 3 * c] = V[r + 3 * c] * sc[c];
    const double det = a[0] * (a[4] * a[8] - a[7] * a[5]) - a[3] * (a[1] * a[8] - a[7] * a[2]) +
                       a[6] * (a[1] * a[5] - a[4] * a[2]);
    small[SM_SGN] = det < 0.0 ? -1.0 : 1.0;
}

// nearest rotation in the reference's sense: [U,~,V]=svd(A); U*diag(1,1,det(U*V'))*V'
// one-sided (Hestenes) Jacobi SVD of a 3x3, singular values sorted descending like LAPACK.
__device__ void proj_so3_dev(const double* Ain, double* Rout) {
    double A[3][3], V[3][3];
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) {
            A[r][c] = Ain[r + 3 * c];
            V[r][c] = r == c ? 1.0 : 0.0;
        }
    for (int sweep = 0; sweep < 60; sweep++) {
        bool rotated = false;
        for (int p = 0; p < 2; p++)
            for (int q = p + 1; q < 3; q++) {
                double alpha = 0.0, beta = 0.0, gamma = 0.0;
                for (int k = 0; k < 3; k++) {
                    alpha += A[k][p] * A[k][p];
                    beta += A[k][q] * A[k][q];
                    gamma += A[k][p] * A[k][q];
                }
                if (gamma == 0.0 || fabs(gamma) <= 2e-16 * sqrt(alpha * beta)) continue;
                rotated = true;
                const double zeta = (beta - alpha) / (2.0 * gamma);
                const double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
                for (int k = 0; k < 3; k++) {
                    const double akp = A[k][p], akq = A[k][q];
                    A[k][p] = c * akp - s * akq;
                    A[k][q] = s * akp + c * akq;
                    const double vkp = V[k][p], vkq = V[k][q];
                    V[k][p] = c * vkp - s * vkq;
                    V[k][q] = s * vkp + c * vkq;
                }
            }
        if (!rotated) break;
    }
    double sig[3];
    int ord[3] = {0, 1, 2};
    for (int c = 0; c < 3; c++) sig[c] = sqrt(A[0][c] * A[0][c] + A[1][c] * A[1][c] + A[2][c] * A[2][c]);
    for (int x = 0; x < 2; x++)
        for (int y = x + 1; y < 3; y++)
            if (sig[ord[y]] > sig[ord[x]]) {
                int t = ord[x];
                ord[x] = ord[y];
                ord[y] = t;
            }
    double U[3][3], W[3][3];
    for (int c = 0; c < 3; c++) {
        const int o = ord[c];
        const double inv = sig[o] > 0.0 ? 1.0 / sig[o] : 0.0;
        for (int r = 0; r < 3; r++) {
            U[r][c] = A[r][o] * inv;
            W[r][c] = V[r][o];
        }
    }
    // rank-deficient block: complete U with a cross product so that it stays orthogonal
    if (!(sig[ord[2]] > 1e-300 * (sig[ord[0]] + 1e-300))) {
        U[0][2] = U[1][0] * U[2][1] - U[2][0] * U[1][1];
        U[1][2] = U[2][0] * U[0][1] - U[0][0] * U[2][1];
        U[2][2] = U[0][0] * U[1][1] - U[1][0] * U[0][1];
    }
    auto det3 = [](double M[3][3]) {
        return M[0][0] * (M[1][1] * M[2][2] - M[1][2] * M[2][1]) - M[0][1] * (M[1][0] * M[2][2] - M[1][2] * M[2][0]) +
               M[0][2] * (M[1][0] * M[2][1] - M[1][1] * M[2][0]);
    };
    // det(U*V') evaluated the way the reference does (a value close to +-1, used as a factor)
    double UVt[3][3];
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) UVt[r][c] = U[r][0] * W[c][0] + U[r][1] * W[c][1] + U[r][2] * W[c][2];
    const double d = det3(UVt);
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++)
            Rout[r + 3 * c] = U[r][0] * W[c][0] + U[r][1] * W[c][1] + d * U[r][2] * W[c][2];
}

__global__ void k_gcw_project(const double* __restrict__ V, const double* __restrict__ small, int n,
                              double* __restrict__ R) {
    const int node = blockIdx.x * blockDim.x + threadIdx.x;
    if (node >= n) return;
    double a[9], out[9];
    for (int c = 0; c < 3; c++) {
        const double sc = small[SM_NRM + c] * (c == 0 ? small[SM_SGN] : 1.0);
        for (int r = 0; r < 3; r++) a[r + 3 * c] = V[9 * (int64_t)node + r + 3 * c] * sc;
    }
    proj_so3_dev(a, out);
    for (int q = 0; q < 9; q++) R[9 * (int64_t)node + q] = out[q];
}

// node ranges balanced by adjacency entries
static void node_bounds(desc_b200_handle* h, const std::vector<int>& rowstart, std::vector<int64_t>& nb) {
    const int n = h->n;
    nb.assign(h->world + 1, 0);
    nb[h->world] = n;
    const int64_t tot = rowstart[n];
    for (int r = 1; r < h->world; r++) {
        const int64_t target = tot * r / h->world;
        nb[r] = std::lower_bound(rowstart.begin(), rowstart.end(), (int)target) - rowstart.begin();
        if (nb[r] > n) nb[r] = n;
        if (nb[r] < nb[r - 1]) nb[r] = nb[r - 1];
    }
}

int desc_gcw_impl(desc_b200_handle* h, const double* d_S) {
    const int n = h->n;
    const int64_t m = h->m;
    const int64_t n9 = 9 * (int64_t)n;
    cudaStream_t st = h->stream;
    if (!h->omega) {
        CUDA_TRY(cudaMalloc(&h->omega, m * sizeof(double)));
        CUDA_TRY(cudaMalloc(&h->gcw_coef, m * sizeof(double)));
        CUDA_TRY(cudaMalloc(&h->isd, (size_t)n * sizeof(double)));
        CUDA_TRY(cudaMalloc(&h->X[0], n9 * sizeof(double)));
        CUDA_TRY(cudaMalloc(&h->X[1], n9 * sizeof(double)));
        CUDA_TRY(cudaMalloc(&h->R_est, n9 * sizeof(double)));
        CUDA_TRY(cudaMalloc(&h->gcw_red, (size_t)GCW_RED_BLOCKS * GCW_NRED * sizeof(double)));
        CUDA_TRY(cudaMalloc(&h->gcw_small, SM_SIZE * sizeof(double)));
        CUDA_TRY(cudaMalloc(&h->gcw_res, (size_t)(DESC_GCW_MAXIT + 8) * sizeof(double)));
        CUDA_TRY(cudaMallocHost(&h->gcw_res_host, (size_t)(DESC_GCW_MAXIT + 8) * sizeof(double)));
    }
    std::vector<int64_t> nb(h->world + 1, 0);
    nb[h->world] = n;
    if (h->world > 1) {
        std::vector<int> rs(n + 1);
        CUDA_TRY(cudaMemcpyAsync(rs.data(), h->rowstart, (size_t)(n + 1) * sizeof(int), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        node_bounds(h, rs, nb);
    }
    std::vector<int64_t> nb9(nb);
    for (auto& v : nb9) v *= 9;
    const int n0 = (int)nb[h->rank], n1 = (int)nb[h->rank + 1];

    const unsigned gbm = (unsigned)((m + 255) / 256);
    CUDA_TRY(cudaMemsetAsync(h->gcw_small, 0, SM_SIZE * sizeof(double), st));
    CUDA_TRY(cudaMemsetAsync(h->gcw_red, 0, (size_t)GCW_RED_BLOCKS * GCW_NRED * sizeof(double), st));
    k_gcw_weights<<<gbm, 256, 0, st>>>(d_S, m, h->omega);
    KERNEL_CHECK(h);
    k_gcw_degree<<<(unsigned)(((int64_t)n * 32 + 255) / 256), 256, 0, st>>>(h->rowstart, h->adj_eid, h->omega, n, h->isd);
    KERNEL_CHECK(h);
    k_gcw_coef<<<gbm, 256, 0, st>>>(h->ei, h->ej, h->omega, h->isd, m, h->gcw_coef);
    KERNEL_CHECK(h);
    double* X = h->X[0];
    double* Y = h->X[1];
    k_gcw_init<<<(unsigned)((n9 + 255) / 256), 256, 0, st>>>(X, n9);
    KERNEL_CHECK(h);
    // orthonormalise the start block (Cholesky-QR twice)
    for (int rep = 0; rep < 2; rep++) {
        CUDA_TRY(cudaMemcpyAsync(Y, X, n9 * sizeof(double), cudaMemcpyDeviceToDevice, st));
        k_gcw_reduce<<<GCW_RED_BLOCKS, GCW_RED_TB, 0, st>>>(X, Y, n, h->gcw_red);
        KERNEL_CHECK(h);
        k_gcw_small_orth<<<1, 1, 0, st>>>(h->gcw_red, GCW_RED_BLOCKS, h->gcw_small, h->gcw_res, -1);
        KERNEL_CHECK(h);
        k_gcw_apply<<<GCW_RED_BLOCKS, GCW_RED_TB, 0, st>>>(X, Y, n, h->gcw_small, h->gcw_red);
        KERNEL_CHECK(h);
    }
    const unsigned spmv_grid = (unsigned)((((int64_t)(n1 - n0)) * 32 + 255) / 256);
    const double tol = 1e-13;
    const int poll = 4;
    int it = 0;
    bool converged = false;
    double last_res = INFINITY;
    while (it < DESC_GCW_MAXIT && !converged) {
        if (spmv_grid > 0) {
            k_gcw_spmv<<<spmv_grid, 256, 0, st>>>(h->rowstart, h->adj_nbr, h->adj_eid, h->Rij, h->gcw_coef, X, Y, n0, n1);
            KERNEL_CHECK(h);
        }
        DESC_TRY(desc_allgather_ranges(h, Y, sizeof(double), nb9));
        k_gcw_reduce<<<GCW_RED_BLOCKS, GCW_RED_TB, 0, st>>>(X, Y, n, h->gcw_red);
        KERNEL_CHECK(h);
        k_gcw_small_orth<<<1, 1, 0, st>>>(h->gcw_red, GCW_RED_BLOCKS, h->gcw_small, h->gcw_res, it);
        KERNEL_CHECK(h);
        k_gcw_apply<<<GCW_RED_BLOCKS, GCW_RED_TB, 0, st>>>(X, Y, n, h->gcw_small, h->gcw_red);
        KERNEL_CHECK(h);
        k_gcw_small_res<<<1, 1, 0, st>>>(h->gcw_red, GCW_RED_BLOCKS, h->gcw_small, h->gcw_res, it);
        KERNEL_CHECK(h);
        it++;
        if (it % poll == 0 || it == DESC_GCW_MAXIT) {
            CUDA_TRY(cudaMemcpyAsync(h->gcw_res_host, h->gcw_res, (size_t)it * sizeof(double), cudaMemcpyDeviceToHost, st));
            CUDA_TRY(cudaStreamSynchronize(st));
            last_res = h->gcw_res_host[it - 1];
            if (!(last_res == last_res)) {
                desc_set_error("GCW: non-finite residual at iteration %d (non-finite S_vec or RijMat?)", it);
                return DESC_B200_ERR_NOCONV;
            }
            // converged, or stagnated at the rounding floor
            if (last_res <= tol) converged = true;
            if (it >= 3 * poll && last_res <= 1e-11 && last_res >= 0.5 * h->gcw_res_host[it - 1 - poll]) converged = true;
        }
    }
    h->tm.gcw_iters = it;
    h->gcw_last_res = last_res;
    if (!converged && !(last_res <= 1e-9)) {
        desc_set_error("GCW subspace iteration did not converge: residual %.3e after %d iterations", last_res, it);
        return DESC_B200_ERR_NOCONV;
    }
    // Rayleigh-Ritz in the converged subspace, back-transform, normalise, sign, project
    if (spmv_grid > 0) {
        k_gcw_spmv<<<spmv_grid, 256, 0, st>>>(h->rowstart, h->adj_nbr, h->adj_eid, h->Rij, h->gcw_coef, X, Y, n0, n1);
        KERNEL_CHECK(h);
    }
    DESC_TRY(desc_allgather_ranges(h, Y, sizeof(double), nb9));
    k_gcw_reduce<<<GCW_RED_BLOCKS, GCW_RED_TB, 0, st>>>(X, Y, n, h->gcw_red);
    KERNEL_CHECK(h);
    k_gcw_small_ritz<<<1, 1, 0, st>>>(h->gcw_red, GCW_RED_BLOCKS, h->gcw_small);
    KERNEL_CHECK(h);
    k_gcw_ritz_apply<<<GCW_RED_BLOCKS, GCW_RED_TB, 0, st>>>(X, h->isd, n, h->gcw_small, Y, h->gcw_red);
    KERNEL_CHECK(h);
    k_gcw_small_final<<<1, 1, 0, st>>>(h->gcw_red, GCW_RED_BLOCKS, Y, h->gcw_small);
    KERNEL_CHECK(h);
    k_gcw_project<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(Y, h->gcw_small, n, h->R_est);
    KERNEL_CHECK(h);
    double flag = 0.0;
    CUDA_TRY(cudaMemcpyAsync(&flag, h->gcw_small + SM_FLAG, sizeof(double), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(h->gcw_theta, h->gcw_small + SM_THETA, 3 * sizeof(double), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    if (flag != 0.0) {
        desc_set_error("GCW: Cholesky-QR breakdown (rank-deficient iterate)");
        return DESC_B200_ERR_NOCONV;
    }
    return DESC_B200_OK;
}
