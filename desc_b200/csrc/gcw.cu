// A13: GCW weighted spectral rotation recovery (reference: Utils/GCW.m:1-38).
//
//   omega_e = 1/(s_e^1.5 + 1e-8)                      GCW.m:20
//   d_i     = sum_{e incident to i} omega_e           GCW.m:21 (row normalisation)
//   M       = D^-1 (W o R)   (3n x 3n, block (i,j) = omega_e/d_i * R_ij, (j,i) = omega_e/d_j * R_ij')
//   V       = top-3 eigenvectors of M ('la'), unit 2-norm columns      GCW.m:27
//   flip column 1 if det(V(1:3,:)) < 0                                  GCW.m:28
//   R_i     = U diag(1,1,det(U V')) V' with [U,~,V] = svd(V_i)          GCW.m:30-36
//
// The reference forms the dense 3n x 3n matrix and calls eigs (ARPACK).  Here M is never formed:
// it is similar to the symmetric N = D^-1/2 (W o R) D^-1/2, eigenvectors
// V = D^-1/2 U, and the three algebraically largest eigenpairs of N come from a thick-restarted
// block Lanczos iteration (block size 3 = the multiplicity of the wanted cluster, full
// re-orthogonalisation, Rayleigh-Ritz on the host over a <= 75-dimensional projected matrix)
// driven by a block-sparse SpMV over the symmetric CSR adjacency built in build.cu:
//
//   y_i = sum_{p in row i}  c_e * (i<j ? R_e : R_e') * x_j ,   c_e = omega_e / sqrt(d_i d_j)
//
// One warp per node, lanes over the node's edges; node-centric so no atomics and a fixed
// summation order (deterministic).  Each R_e is read twice per SpMV (once per endpoint).
// All long-vector work (SpMV, block inner products, block updates, Cholesky-QR) runs on the
// device with two-stage fixed-order reductions; only the tiny projected eigenproblem is solved
// on the host.  Convergence is decided on the Lanczos residual estimate and then verified with
// an explicitly computed residual || N X - X (X'NX) ||_F.
#include "internal.cuh"
#include "so3.cuh"

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
#define SM_R 36     // 9: R factor of the last Cholesky-QR (block = Q R)
#define SM_SIZE 48

// rule 0: GCW.m:20 `1./(S.^(3/2)+1e-8)`;  rule 1: CEMP_GCW.m:141 `1./(S+1e-8)`;
// rule 2: Spectral.m:36-40 -- the unweighted, un-normalised block matrix (S unused).  Its eigenvectors are those of
// Rij_blk/c for any c > 0; c = max degree keeps the spectrum inside [-1,1] like the normalised rules.
__global__ void k_gcw_weights(const double* __restrict__ S, int64_t m, int rule, double scale, double* __restrict__ omega) {
    int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= m) return;
    omega[e] = rule == 2 ? scale : 1.0 / ((rule == 1 ? S[e] : pow(S[e], 1.5)) + 1e-8);
}
__global__ void k_gcw_unit(double* __restrict__ isd, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) isd[i] = 1.0;
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
// the same coefficients in adjacency order (what the SpMV streams)
__global__ void k_gcw_coef_adj(const int* __restrict__ adj_eid, const double* __restrict__ coef, int64_t n2m,
                               double* __restrict__ coef_adj) {
    int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (p < n2m) coef_adj[p] = coef[adj_eid[p]];
}

__global__ void k_gcw_init(double* __restrict__ X, int64_t n9) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= n9) return;
    uint64_t z = desc_key(0x6a09e667f3bcc908ull, (uint64_t)t, 7ull);
    X[t] = (double)(z >> 11) * (1.0 / 9007199254740992.0) - 0.5;
}

// Y[node] = sum_p coef * op(R_e) * X[nbr] for node in [n0, n1).
// One warp per node, 32 adjacency entries at a time.  The two 72-byte records of an entry (R_e and the neighbour's X
// block; 8-byte aligned, so no vector loads) are gathered COOPERATIVELY like in cycle.cu: in 9 rounds lane L loads word
// 32 t + L of the 32 concatenated records -- 9 consecutive lanes read one record as one contiguous piece (1-2 L1
// wavefronts per record instead of 9) -- and a per-warp staging buffer (stride 9 doubles, conflict-free) hands the
// record to the lane that owns the entry.  coef is kept in adjacency order (coalesced).
#define SPMV_WARPS 8
__global__ void __launch_bounds__(SPMV_WARPS * 32, 3)
k_gcw_spmv(const int* __restrict__ rowstart, const int* __restrict__ adj_nbr,
           const int* __restrict__ adj_eid, const double* __restrict__ Rij,
           const double* __restrict__ coef_adj, const double* __restrict__ X, double* __restrict__ Y,
           int n0, int n1) {
    __shared__ double stage[SPMV_WARPS][2][288];
    const int wib = threadIdx.x >> 5;
    const int node = n0 + blockIdx.x * SPMV_WARPS + wib;
    const int lane = threadIdx.x & 31;
    if (node >= n1) return;
    double* sr = stage[wib][0];
    double* sx = stage[wib][1];
    double acc[9];
#pragma unroll
    for (int x = 0; x < 9; x++) acc[x] = 0.0;
    const int p0 = rowstart[node], p1 = rowstart[node + 1];
    for (int pb = p0; pb < p1; pb += 32) {
        const int p = pb + lane;
        const bool ok = p < p1;
        int nb = 0, e = 0;
        double c = 0.0;
        if (ok) {
            nb = adj_nbr[p];
            e = adj_eid[p];
            c = coef_adj[p];
        }
        const int cnt = min(32, p1 - pb);
#pragma unroll
        for (int t = 0; t < 9; t++) {
            const int w = t * 32 + lane;
            const int rec = w / 9, el = w - 9 * rec;
            const int qe = __shfl_sync(0xffffffffu, e, rec), qn = __shfl_sync(0xffffffffu, nb, rec);
            double vr = 0.0, vx = 0.0;
            if (rec < cnt) {
                vr = __ldg(Rij + 9 * (int64_t)qe + el);
                vx = X[9 * (int64_t)qn + el];
            }
            sr[w] = vr;
            sx[w] = vx;
        }
        __syncwarp();
        double r[9], x[9];
#pragma unroll
        for (int q = 0; q < 9; q++) {
            r[q] = sr[9 * lane + q];
            x[q] = sx[9 * lane + q];
        }
        __syncwarp();
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
        Y[9 * (int64_t)node + lane] = v;
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

// one thread: H, G from partials; Cholesky of G = L L'; T = inv(L)' (so Q = Y T), R = L' (Y = Q R).
// accumulate != 0: R <- R_new * R_old (second Cholesky-QR pass).  A pivot that collapses relative
// to the block's scale raises the breakdown flag (the block is numerically rank deficient).
// (launched with one warp: lane l sums the partials of blocks l, l+32, ... and a fixed-shape shuffle tree adds the
// lanes -- deterministic, and ~20x shorter than the serial sum of 148 x 15 dependent loads it replaces)
template <int K>
__device__ __forceinline__ void warp_sum_partials(const double* __restrict__ partial, int nblocks, int offset, double (&s)[K]) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int x = 0; x < K; x++) s[x] = 0.0;
    for (int b = lane; b < nblocks; b += 32)
#pragma unroll
        for (int x = 0; x < K; x++) s[x] += partial[(size_t)b * GCW_NRED + offset + x];
#pragma unroll
    for (int x = 0; x < K; x++) s[x] = group_sum<32>(s[x]);
}

__global__ void k_gcw_small_orth(const double* __restrict__ partial, int nblocks,
                                 double* __restrict__ small, int accumulate) {
    double s[15];
    warp_sum_partials<15>(partial, nblocks, 0, s);
    if (threadIdx.x != 0) return;
    for (int x = 0; x < 9; x++) small[SM_H + x] = s[x];
    const double g00 = s[9], g01 = s[10], g02 = s[11], g11 = s[12], g12 = s[13], g22 = s[14];
    const double scale = fmax(g00, fmax(g11, g22));
    double l00 = sqrt(g00);
    double l10 = g01 / l00, l20 = g02 / l00;
    double d1 = g11 - l10 * l10;
    double l11 = sqrt(d1);
    double l21 = (g12 - l20 * l10) / l11;
    double d2 = g22 - l20 * l20 - l21 * l21;
    double l22 = sqrt(d2);
    const double thr = 1e-13 * scale;
    if (!(scale > 1e-280) || !(g00 > thr) || !(d1 > thr) || !(d2 > thr)) small[SM_FLAG] = 1.0;
    double i00 = 1.0 / l00, i11 = 1.0 / l11, i22 = 1.0 / l22;
    double i10 = -l10 * i00 * i11;
    double i21 = -l21 * i11 * i22;
    double i20 = -(l20 * i00 + l21 * i10) * i22;
    small[SM_T + 0] = i00; small[SM_T + 1] = 0.0; small[SM_T + 2] = 0.0;
    small[SM_T + 3] = i10; small[SM_T + 4] = i11; small[SM_T + 5] = 0.0;
    small[SM_T + 6] = i20; small[SM_T + 7] = i21; small[SM_T + 8] = i22;
    // R = L' (upper), column-major
    double Rn[9] = {l00, 0.0, 0.0, l10, l11, 0.0, l20, l21, l22};
    if (accumulate) {
        double Ro[9], Rt[9];
        for (int x = 0; x < 9; x++) Ro[x] = small[SM_R + x];
        for (int c = 0; c < 3; c++)
            for (int r = 0; r < 3; r++)
                Rt[r + 3 * c] = Rn[r] * Ro[3 * c] + Rn[r + 3] * Ro[1 + 3 * c] + Rn[r + 6] * Ro[2 + 3 * c];
        for (int x = 0; x < 9; x++) small[SM_R + x] = Rt[x];
    } else {
        for (int x = 0; x < 9; x++) small[SM_R + x] = Rn[x];
    }
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

// residual partial only: sum || Y - X H ||^2 with H = small[SM_H] (explicit verification)
__global__ void __launch_bounds__(GCW_RED_TB)
k_gcw_residual(const double* __restrict__ X, const double* __restrict__ Y, int n,
               const double* __restrict__ small, double* __restrict__ partial) {
    double H[9];
#pragma unroll
    for (int q = 0; q < 9; q++) H[q] = small[SM_H + q];
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
    }
    block_reduce_store<1>(v, partial + (size_t)blockIdx.x * GCW_NRED + 15);
}

__global__ void k_gcw_small_res(const double* __restrict__ partial, int nblocks,
                                double* __restrict__ small, double* __restrict__ res_hist, int it) {
    double sv[1];
    warp_sum_partials<1>(partial, nblocks, 15, sv);
    if (threadIdx.x != 0) return;
    const double s = sv[0];
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
    double s[3];
    warp_sum_partials<3>(partial, nblocks, 0, s);
    if (threadIdx.x != 0) return;
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


// ------------------------------------------------------------------------------------------
// block Lanczos building blocks.  Basis blocks are stored back to back: block b at V + b*9n.
// ------------------------------------------------------------------------------------------
#define GCW_PROJ_BLOCKS 37   // x nb blocks of the basis; 4 resident CTAs per SM at nb >= 16
#define GCW_DMAX 24          // processed basis blocks before a thick restart (72 columns)
#define GCW_KEEP 2           // Ritz blocks kept at a restart (6 vectors)

// partial[(b*gridDim.x + bx)*9 + a + 3c] = sum over nodes of V_b(:,a)' W(:,c)
__global__ void __launch_bounds__(GCW_RED_TB)
k_gcw_proj(const double* __restrict__ V, const double* __restrict__ W, int n, int64_t n9,
           double* __restrict__ partial) {
    const double* Vb = V + (size_t)blockIdx.y * n9;
    double v[9];
#pragma unroll
    for (int x = 0; x < 9; x++) v[x] = 0.0;
    for (int node = blockIdx.x * blockDim.x + threadIdx.x; node < n; node += gridDim.x * blockDim.x) {
        double xb[9], yb[9];
#pragma unroll
        for (int q = 0; q < 9; q++) {
            xb[q] = Vb[9 * (int64_t)node + q];
            yb[q] = W[9 * (int64_t)node + q];
        }
#pragma unroll
        for (int a = 0; a < 3; a++)
#pragma unroll
            for (int b = 0; b < 3; b++)
                v[a + 3 * b] += xb[3 * a] * yb[3 * b] + xb[3 * a + 1] * yb[3 * b + 1] + xb[3 * a + 2] * yb[3 * b + 2];
    }
    block_reduce_store<9>(v, partial + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 9);
}

// c[b*9 + q] = sum over bx of partial (fixed order); csum (+)= c
__global__ void k_gcw_proj_finish(const double* __restrict__ partial, int nb, int nbx,
                                  double* __restrict__ c, double* __restrict__ csum, int accumulate) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nb * 9) return;
    const int b = t / 9, q = t % 9;
    double s = 0.0;
    for (int bx = 0; bx < nbx; bx++) s += partial[((size_t)b * nbx + bx) * 9 + q];
    c[t] = s;
    csum[t] = accumulate ? csum[t] + s : s;
}

// W -= sum_b V_b c_b
__global__ void __launch_bounds__(256)
k_gcw_subtract(const double* __restrict__ V, double* __restrict__ W, int n, int64_t n9, int nb,
               const double* __restrict__ c) {
    extern __shared__ double sc[];
    for (int t = threadIdx.x; t < nb * 9; t += blockDim.x) sc[t] = c[t];
    __syncthreads();
    const int node = blockIdx.x * blockDim.x + threadIdx.x;
    if (node >= n) return;
    double w[9];
#pragma unroll
    for (int q = 0; q < 9; q++) w[q] = W[9 * (int64_t)node + q];
    for (int b = 0; b < nb; b++) {
        const double* vb = V + (size_t)b * n9 + 9 * (int64_t)node;
        const double* cb = sc + 9 * b;
        double x[9];
#pragma unroll
        for (int q = 0; q < 9; q++) x[q] = vb[q];
#pragma unroll
        for (int col = 0; col < 3; col++)
#pragma unroll
            for (int r = 0; r < 3; r++)
                w[r + 3 * col] -= x[r] * cb[3 * col] + x[r + 3] * cb[1 + 3 * col] + x[r + 6] * cb[2 + 3 * col];
    }
#pragma unroll
    for (int q = 0; q < 9; q++) W[9 * (int64_t)node + q] = w[q];
}

// Out block o (blockIdx.y) = sum_b V_b * Z[b][o]   (Z: nb x nout blocks of 3x3, column-major 3x3)
__global__ void __launch_bounds__(256)
k_gcw_rotate(const double* __restrict__ V, double* __restrict__ Out, int n, int64_t n9, int nb,
             int nout, const double* __restrict__ Z) {
    extern __shared__ double sz[];
    const int o = blockIdx.y;
    for (int t = threadIdx.x; t < nb * 9; t += blockDim.x) sz[t] = Z[((size_t)(t / 9) * nout + o) * 9 + (t % 9)];
    __syncthreads();
    const int node = blockIdx.x * blockDim.x + threadIdx.x;
    if (node >= n) return;
    double w[9];
#pragma unroll
    for (int q = 0; q < 9; q++) w[q] = 0.0;
    for (int b = 0; b < nb; b++) {
        const double* vb = V + (size_t)b * n9 + 9 * (int64_t)node;
        const double* cb = sz + 9 * b;
        double x[9];
#pragma unroll
        for (int q = 0; q < 9; q++) x[q] = vb[q];
#pragma unroll
        for (int col = 0; col < 3; col++)
#pragma unroll
            for (int r = 0; r < 3; r++)
                w[r + 3 * col] += x[r] * cb[3 * col] + x[r + 3] * cb[1 + 3 * col] + x[r + 6] * cb[2 + 3 * col];
    }
#pragma unroll
    for (int q = 0; q < 9; q++) Out[(size_t)o * n9 + 9 * (int64_t)node + q] = w[q];
}

__global__ void k_gcw_random(double* __restrict__ X, int64_t n9, uint64_t salt) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= n9) return;
    uint64_t z = desc_key(0x6a09e667f3bcc908ull + salt, (uint64_t)t, 7ull + salt);
    X[t] = (double)(z >> 11) * (1.0 / 9007199254740992.0) - 0.5;
}

// copy per-step results into the history the host reads at a Rayleigh-Ritz check
__global__ void k_gcw_record(const double* __restrict__ csum, int nb, const double* __restrict__ small,
                             double* __restrict__ hist_c, double* __restrict__ hist_r) {
    const int t = threadIdx.x;
    for (int x = t; x < nb * 9; x += blockDim.x) hist_c[x] = csum[x];
    if (t < 9) hist_r[t] = small[SM_R + t];
    if (t == 9) hist_r[9] = small[SM_FLAG];
}

// symmetric eigen-decomposition (cyclic Jacobi) of a small dense matrix on the host; eigenvalues
// descending.  A is d x d column-major (destroyed), Z gets the eigenvectors as columns.
static void host_jacobi_eig(std::vector<double>& A, int d, std::vector<double>& ev, std::vector<double>& Z) {
    Z.assign((size_t)d * d, 0.0);
    for (int i = 0; i < d; i++) Z[i + (size_t)i * d] = 1.0;
    for (int i = 0; i < d; i++)
        for (int j = i + 1; j < d; j++) {
            const double v = 0.5 * (A[i + (size_t)j * d] + A[j + (size_t)i * d]);
            A[i + (size_t)j * d] = A[j + (size_t)i * d] = v;
        }
    for (int sweep = 0; sweep < 100; sweep++) {
        double off = 0.0, diag = 0.0;
        for (int i = 0; i < d; i++) {
            diag += A[i + (size_t)i * d] * A[i + (size_t)i * d];
            for (int j = i + 1; j < d; j++) off += A[i + (size_t)j * d] * A[i + (size_t)j * d];
        }
        if (off <= 1e-34 * (diag + 1e-300)) break;
        for (int p = 0; p < d - 1; p++)
            for (int q = p + 1; q < d; q++) {
                const double apq = A[p + (size_t)q * d];
                if (apq == 0.0) continue;
                const double theta = (A[q + (size_t)q * d] - A[p + (size_t)p * d]) / (2.0 * apq);
                const double t = (theta >= 0.0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
                const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < d; k++) {
                    const double akp = A[k + (size_t)p * d], akq = A[k + (size_t)q * d];
                    A[k + (size_t)p * d] = c * akp - s * akq;
                    A[k + (size_t)q * d] = s * akp + c * akq;
                }
                for (int k = 0; k < d; k++) {
                    const double apk = A[p + (size_t)k * d], aqk = A[q + (size_t)k * d];
                    A[p + (size_t)k * d] = c * apk - s * aqk;
                    A[q + (size_t)k * d] = s * apk + c * aqk;
                }
                for (int k = 0; k < d; k++) {
                    const double zkp = Z[k + (size_t)p * d], zkq = Z[k + (size_t)q * d];
                    Z[k + (size_t)p * d] = c * zkp - s * zkq;
                    Z[k + (size_t)q * d] = s * zkp + c * zkq;
                }
            }
    }
    std::vector<int> ord(d);
    for (int i = 0; i < d; i++) ord[i] = i;
    std::sort(ord.begin(), ord.end(), [&](int a, int b) { return A[a + (size_t)a * d] > A[b + (size_t)b * d]; });
    ev.resize(d);
    std::vector<double> Zs((size_t)d * d);
    for (int c = 0; c < d; c++) {
        ev[c] = A[ord[c] + (size_t)ord[c] * d];
        for (int r = 0; r < d; r++) Zs[r + (size_t)c * d] = Z[r + (size_t)ord[c] * d];
    }
    Z.swap(Zs);
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

namespace {
struct Lanczos {
    desc_b200_handle* h;
    int n;
    int64_t n9;
    int n0, n1;
    std::vector<int64_t> nb9;
    double* V;       // (GCW_DMAX + 2 + GCW_KEEP + 1) blocks
    double* partial; // proj partials
    double* c;       // 9 * (GCW_DMAX+1)
    double* csum;
    double* hist_c;  // per step: 9*(GCW_DMAX+1)
    double* hist_r;  // per step: 16
    double* Zdev;    // rotation coefficients
    double* host;    // pinned staging
    int spmv_count = 0;

    double* blk(int b) const { return V + (size_t)b * n9; }

    int spmv(const double* X, double* Y) {
        const unsigned grid = (unsigned)((n1 - n0 + SPMV_WARPS - 1) / SPMV_WARPS);
        cudaEvent_t e0 = nullptr, e1 = nullptr;
        if (spmv_count < (int)h->spmv_events.size() / 2) {
            e0 = h->spmv_events[2 * spmv_count];
            e1 = h->spmv_events[2 * spmv_count + 1];
        }
        if (e0) CUDA_TRY(cudaEventRecord(e0, h->stream));
        if (grid > 0) {
            k_gcw_spmv<<<grid, SPMV_WARPS * 32, 0, h->stream>>>(h->rowstart, h->adj_nbr, h->adj_eid, h->Rij, h->gcw_coef_adj, X, Y, n0, n1);
            KERNEL_CHECK(h);
        }
        if (e1) CUDA_TRY(cudaEventRecord(e1, h->stream));
        spmv_count++;
        DESC_TRY(desc_allgather_ranges(h, Y, sizeof(double), nb9));
        return DESC_B200_OK;
    }
    // W (block index w) orthogonalised against blocks [0, nb): two passes, csum = total coefficients
    int orth_against(int nb, int w) {
        if (nb <= 0) return DESC_B200_OK;
        for (int pass = 0; pass < 2; pass++) {
            dim3 g(GCW_PROJ_BLOCKS, nb);
            k_gcw_proj<<<g, GCW_RED_TB, 0, h->stream>>>(V, blk(w), n, n9, partial);
            KERNEL_CHECK(h);
            k_gcw_proj_finish<<<(nb * 9 + 127) / 128, 128, 0, h->stream>>>(partial, nb, GCW_PROJ_BLOCKS, c, csum, pass);
            KERNEL_CHECK(h);
            k_gcw_subtract<<<(n + 255) / 256, 256, nb * 9 * sizeof(double), h->stream>>>(V, blk(w), n, n9, nb, c);
            KERNEL_CHECK(h);
        }
        return DESC_B200_OK;
    }
    // Cholesky-QR twice in place on block w; small[SM_R] = R, small[SM_FLAG] set on breakdown
    int cholqr2(int w) {
        for (int pass = 0; pass < 2; pass++) {
            k_gcw_reduce<<<GCW_RED_BLOCKS, GCW_RED_TB, 0, h->stream>>>(blk(w), blk(w), n, h->gcw_red);
            KERNEL_CHECK(h);
            k_gcw_small_orth<<<1, 32, 0, h->stream>>>(h->gcw_red, GCW_RED_BLOCKS, h->gcw_small, pass);
            KERNEL_CHECK(h);
            k_gcw_apply<<<GCW_RED_BLOCKS, GCW_RED_TB, 0, h->stream>>>(blk(w), blk(w), n, h->gcw_small, h->gcw_red);
            KERNEL_CHECK(h);
        }
        return DESC_B200_OK;
    }
    int random_block(int w, int nb, uint64_t salt) {
        k_gcw_random<<<(unsigned)((n9 + 255) / 256), 256, 0, h->stream>>>(blk(w), n9, salt);
        KERNEL_CHECK(h);
        DESC_TRY(orth_against(nb, w));
        CUDA_TRY(cudaMemsetAsync(h->gcw_small + SM_FLAG, 0, sizeof(double), h->stream));
        DESC_TRY(cholqr2(w));
        return DESC_B200_OK;
    }
};
}  // namespace

// k_gcw_apply reads X and Y separately; with X == Y (in-place Cholesky-QR) the residual partial is
// meaningless and ignored.

int desc_gcw_impl(desc_b200_handle* h, const double* d_S) {
    const int n = h->n;
    const int64_t m = h->m;
    const int64_t n9 = 9 * (int64_t)n;
    cudaStream_t st = h->stream;
    // basis capacity: at most n blocks exist in a 3n-dimensional space
    const int dmax = std::max(1, std::min(GCW_DMAX, n - 1));
    const int keep = std::max(1, std::min(GCW_KEEP, dmax - 1));
    const int total_blocks = GCW_DMAX + 2 + GCW_KEEP + 1;
    const int hist_stride = 9 * (GCW_DMAX + 2);
    if (!h->omega) {
        CUDA_TRY(cudaMalloc(&h->omega, m * sizeof(double)));
        CUDA_TRY(cudaMalloc(&h->gcw_coef, m * sizeof(double)));
        CUDA_TRY(cudaMalloc(&h->gcw_coef_adj, 2 * m * sizeof(double)));
        CUDA_TRY(cudaMalloc(&h->isd, (size_t)n * sizeof(double)));
        CUDA_TRY(cudaMalloc(&h->X[0], (size_t)total_blocks * n9 * sizeof(double)));
        CUDA_TRY(cudaMalloc(&h->X[1], n9 * sizeof(double)));
        CUDA_TRY(cudaMalloc(&h->R_est, n9 * sizeof(double)));
        const size_t red = std::max<size_t>((size_t)GCW_RED_BLOCKS * GCW_NRED, (size_t)GCW_PROJ_BLOCKS * 9 * (GCW_DMAX + 2));
        CUDA_TRY(cudaMalloc(&h->gcw_red, 2 * red * sizeof(double)));
        CUDA_TRY(cudaMalloc(&h->gcw_small, SM_SIZE * sizeof(double)));
        // c | csum | hist_c[(DMAX+2) steps] | hist_r[(DMAX+2) steps x 16] | Z
        const size_t work = (size_t)2 * hist_stride + (size_t)(GCW_DMAX + 2) * hist_stride + (size_t)(GCW_DMAX + 2) * 16 +
                            (size_t)9 * (GCW_DMAX + 2) * (GCW_KEEP + 1) + 64;
        CUDA_TRY(cudaMalloc(&h->gcw_res, work * sizeof(double)));
        CUDA_TRY(cudaMallocHost(&h->gcw_res_host, work * sizeof(double)));
    }
    std::vector<int64_t> nbd(h->world + 1, 0);
    nbd[h->world] = n;
    if (h->world > 1) {
        std::vector<int> rs(n + 1);
        CUDA_TRY(cudaMemcpyAsync(rs.data(), h->rowstart, (size_t)(n + 1) * sizeof(int), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        node_bounds(h, rs, nbd);
    }
    Lanczos L;
    L.h = h;
    L.n = n;
    L.n9 = n9;
    L.n0 = (int)nbd[h->rank];
    L.n1 = (int)nbd[h->rank + 1];
    L.nb9 = nbd;
    for (auto& v : L.nb9) v *= 9;
    L.V = h->X[0];
    const size_t red = std::max<size_t>((size_t)GCW_RED_BLOCKS * GCW_NRED, (size_t)GCW_PROJ_BLOCKS * 9 * (GCW_DMAX + 2));
    L.partial = h->gcw_red + red;
    L.c = h->gcw_res;
    L.csum = L.c + hist_stride;
    L.hist_c = L.csum + hist_stride;
    L.hist_r = L.hist_c + (size_t)(GCW_DMAX + 2) * hist_stride;
    L.Zdev = L.hist_r + (size_t)(GCW_DMAX + 2) * 16;
    L.host = h->gcw_res_host;
    double* host_hist_c = L.host + (L.hist_c - h->gcw_res);
    double* host_hist_r = L.host + (L.hist_r - h->gcw_res);
    double* host_Z = L.host + (L.Zdev - h->gcw_res);

    const unsigned gbm = (unsigned)((m + 255) / 256);
    CUDA_TRY(cudaMemsetAsync(h->gcw_small, 0, SM_SIZE * sizeof(double), st));
    CUDA_TRY(cudaMemsetAsync(h->gcw_red, 0, 2 * red * sizeof(double), st));
    k_gcw_weights<<<gbm, 256, 0, st>>>(d_S, m, h->gcw_weight_rule, 1.0 / (double)std::max(h->maxdeg, 1), h->omega);
    KERNEL_CHECK(h);
    if (h->gcw_weight_rule == 2)   // Spectral.m: no degree normalisation
        k_gcw_unit<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(h->isd, n);
    else
        k_gcw_degree<<<(unsigned)(((int64_t)n * 32 + 255) / 256), 256, 0, st>>>(h->rowstart, h->adj_eid, h->omega, n, h->isd);
    KERNEL_CHECK(h);
    k_gcw_coef<<<gbm, 256, 0, st>>>(h->ei, h->ej, h->omega, h->isd, m, h->gcw_coef);
    KERNEL_CHECK(h);
    k_gcw_coef_adj<<<(unsigned)((2 * m + 255) / 256), 256, 0, st>>>(h->adj_eid, h->gcw_coef, 2 * m, h->gcw_coef_adj);
    KERNEL_CHECK(h);
    while (h->spmv_events.size() < 2 * 32) {
        cudaEvent_t e;
        CUDA_TRY(cudaEventCreate(&e));
        h->spmv_events.push_back(e);
    }

    // projected matrix on the host: Hm is (3*cap) x (3*cap), column-major
    const int cap = GCW_DMAX + 2;
    const int ld = 3 * cap;
    std::vector<double> Hm((size_t)ld * ld, 0.0);
    auto Hset = [&](int r, int c, double v) { Hm[r + (size_t)c * ld] = v; Hm[c + (size_t)r * ld] = v; };

    uint64_t salt = 1;
    DESC_TRY(L.random_block(0, 0, salt++));
    int nproc = 0;        // processed blocks: H columns known for blocks [0, nproc)
    int nblk = 1;         // blocks in the basis (nproc processed + 1 pending, unless exhausted)
    int steps_since_check = 0, first_step_of_batch = 0, total_spmv = 0;
    bool trust_estimate = true, converged = false, exhausted = false;
    double est = INFINITY, explicit_res = INFINITY;
    std::vector<double> Rlast(9, 0.0);
    std::vector<double> ev, Z;
    const double tol = 2e-13, accept = 2e-11;
    const int max_spmv = 6000;
    int restarts = 0;

    while (!converged && total_spmv < max_spmv) {
        // ---- one Lanczos step: process the pending block (index nproc), create block nblk
        const int p = nproc;
        DESC_TRY(L.spmv(L.blk(p), L.blk(nblk)));
        total_spmv++;
        DESC_TRY(L.orth_against(nblk, nblk));
        CUDA_TRY(cudaMemsetAsync(h->gcw_small + SM_FLAG, 0, sizeof(double), st));
        DESC_TRY(L.cholqr2(nblk));
        k_gcw_record<<<1, 128, 0, st>>>(L.csum, nblk, h->gcw_small, L.hist_c + (size_t)steps_since_check * hist_stride,
                                        L.hist_r + (size_t)steps_since_check * 16);
        KERNEL_CHECK(h);
        if (steps_since_check == 0) first_step_of_batch = p;
        steps_since_check++;
        nproc++;
        nblk++;
        const bool full = nproc >= dmax;
        const bool check = full || steps_since_check >= 4 || nblk > n;
        if (!check) continue;

        // ---- pull the step results, extend H, Rayleigh-Ritz on the host
        CUDA_TRY(cudaMemcpyAsync(host_hist_c, L.hist_c, (size_t)steps_since_check * hist_stride * sizeof(double), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaMemcpyAsync(host_hist_r, L.hist_r, (size_t)steps_since_check * 16 * sizeof(double), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        bool breakdown = false;
        int good_steps = steps_since_check;
        for (int s = 0; s < steps_since_check && !breakdown; s++) {
            const int pb = first_step_of_batch + s;      // block processed in this step
            const double* cs = host_hist_c + (size_t)s * hist_stride;
            for (int b = 0; b <= pb; b++)
                for (int cc = 0; cc < 3; cc++)
                    for (int a = 0; a < 3; a++) {
                        const double v = cs[9 * b + a + 3 * cc];
                        if (!(v == v)) {
                            desc_set_error("GCW: non-finite value in the Lanczos recurrence (non-finite S_vec or RijMat?)");
                            return DESC_B200_ERR_NOCONV;
                        }
                        if (b == pb) Hset(3 * b + a, 3 * pb + cc, 0.5 * (v + cs[9 * b + cc + 3 * a]));
                        else Hset(3 * b + a, 3 * pb + cc, v);
                    }
            const double* hr = host_hist_r + (size_t)s * 16;
            for (int x = 0; x < 9; x++) Rlast[x] = hr[x];
            if (hr[9] != 0.0) {
                breakdown = true;     // the block created in this step is numerically rank deficient;
                good_steps = s + 1;   // later steps of the batch built on it and are dropped
            }
        }
        if (breakdown) {
            // keep the blocks processed up to and including the breakdown step; replace the unusable
            // new block with a fresh random one (or stop if the whole space is spanned)
            nproc = first_step_of_batch + good_steps;
            nblk = nproc;
            std::fill(Rlast.begin(), Rlast.end(), 0.0);
            trust_estimate = false;
            if (nblk >= n) exhausted = true;
        }
        steps_since_check = 0;
        const int d = 3 * nproc;
        std::vector<double> A((size_t)d * d);
        for (int cc = 0; cc < d; cc++)
            for (int r = 0; r < d; r++) A[r + (size_t)cc * d] = Hm[r + (size_t)cc * ld];
        host_jacobi_eig(A, d, ev, Z);
        est = 0.0;
        for (int i = 0; i < 3 && i < d; i++) {
            double rn = 0.0;
            for (int r = 0; r < 3; r++) {
                double v = 0.0;
                for (int k = 0; k < 3; k++) v += Rlast[r + 3 * k] * Z[(d - 3 + k) + (size_t)i * d];
                rn += v * v;
            }
            est = std::max(est, std::sqrt(rn));
        }
        const bool must_restart = full || breakdown || exhausted;
        const bool verify_now = exhausted || (trust_estimate ? est <= tol : must_restart);
        if (!verify_now && !must_restart) continue;

        // ---- rotate to Ritz blocks: top `kr` blocks into the scratch area
        const int kr = std::min(keep, nproc);
        const int scratch = GCW_DMAX + 2;   // first scratch block
        for (int b = 0; b < nproc; b++)
            for (int o = 0; o < kr; o++)
                for (int cc = 0; cc < 3; cc++)
                    for (int r = 0; r < 3; r++)
                        host_Z[((size_t)b * kr + o) * 9 + r + 3 * cc] = Z[(3 * b + r) + (size_t)(3 * o + cc) * d];
        CUDA_TRY(cudaMemcpyAsync(L.Zdev, host_Z, (size_t)nproc * kr * 9 * sizeof(double), cudaMemcpyHostToDevice, st));
        {
            dim3 g((n + 255) / 256, kr);
            k_gcw_rotate<<<g, 256, nproc * 9 * sizeof(double), st>>>(L.V, L.blk(scratch), n, n9, nproc, kr, L.Zdev);
            KERNEL_CHECK(h);
        }
        if (verify_now) {
            // explicit verification on the three leading Ritz vectors: Y = N X, H3 = X'Y, ||Y - X H3||_F
            double* X = L.blk(scratch);
            double* Y = h->X[1];
            DESC_TRY(L.spmv(X, Y));
            total_spmv++;
            k_gcw_reduce<<<GCW_RED_BLOCKS, GCW_RED_TB, 0, st>>>(X, Y, n, h->gcw_red);
            KERNEL_CHECK(h);
            k_gcw_small_orth<<<1, 32, 0, st>>>(h->gcw_red, GCW_RED_BLOCKS, h->gcw_small, 0);
            KERNEL_CHECK(h);
            k_gcw_residual<<<GCW_RED_BLOCKS, GCW_RED_TB, 0, st>>>(X, Y, n, h->gcw_small, h->gcw_red);
            KERNEL_CHECK(h);
            k_gcw_small_res<<<1, 32, 0, st>>>(h->gcw_red, GCW_RED_BLOCKS, h->gcw_small, L.c, 0);
            KERNEL_CHECK(h);
            CUDA_TRY(cudaMemcpyAsync(L.host, L.c, sizeof(double), cudaMemcpyDeviceToHost, st));
            CUDA_TRY(cudaStreamSynchronize(st));
            explicit_res = L.host[0];
            if (explicit_res <= accept) {
                converged = true;
                for (int i = 0; i < 3; i++) h->gcw_theta[i] = i < d ? ev[i] : 0.0;
                CUDA_TRY(cudaMemcpyAsync(h->X[1], X, n9 * sizeof(double), cudaMemcpyDeviceToDevice, st));
                break;
            }
            trust_estimate = false;
            if (exhausted) {
                desc_set_error("GCW: Krylov space exhausted but residual is %.3e", explicit_res);
                return DESC_B200_ERR_NOCONV;
            }
        }
        if (!must_restart) continue;
        // ---- thick restart: basis = [kr Ritz blocks, pending block]; H = diag(theta) on the kept part
        const int pending_old = nblk - 1;   // valid only when no breakdown happened
        for (int o = 0; o < kr; o++)
            CUDA_TRY(cudaMemcpyAsync(L.blk(o), L.blk(scratch + o), n9 * sizeof(double), cudaMemcpyDeviceToDevice, st));
        if (!breakdown) {
            if (pending_old != kr)
                CUDA_TRY(cudaMemcpyAsync(L.blk(kr), L.blk(pending_old), n9 * sizeof(double), cudaMemcpyDeviceToDevice, st));
        } else {
            DESC_TRY(L.random_block(kr, kr, salt++));
        }
        std::fill(Hm.begin(), Hm.end(), 0.0);
        for (int i = 0; i < 3 * kr; i++) Hm[i + (size_t)i * ld] = ev[i];
        nproc = kr;
        nblk = kr + 1;
        restarts++;
    }
    h->tm.gcw_iters = total_spmv;
    h->gcw_last_res = explicit_res;
    if (!converged) {
        desc_set_error("GCW block Lanczos did not converge: estimate %.3e, explicit residual %.3e after %d SpMVs (%d restarts)",
                       est, explicit_res, total_spmv, restarts);
        return DESC_B200_ERR_NOCONV;
    }
    // back-transform V = D^-1/2 X, normalise columns, sign rule, project every node block to SO(3)
    {
        const double ident[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
        for (int x = 0; x < 9; x++) L.host[x] = ident[x];
        CUDA_TRY(cudaMemcpyAsync(h->gcw_small + SM_Z, L.host, 9 * sizeof(double), cudaMemcpyHostToDevice, st));
    }
    double* X = h->X[1];
    double* Vout = L.blk(0);
    k_gcw_ritz_apply<<<GCW_RED_BLOCKS, GCW_RED_TB, 0, st>>>(X, h->isd, n, h->gcw_small, Vout, h->gcw_red);
    KERNEL_CHECK(h);
    k_gcw_small_final<<<1, 32, 0, st>>>(h->gcw_red, GCW_RED_BLOCKS, Vout, h->gcw_small);
    KERNEL_CHECK(h);
    k_gcw_project<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(Vout, h->gcw_small, n, h->R_est);
    KERNEL_CHECK(h);
    CUDA_TRY(cudaStreamSynchronize(st));
    {   // mean device time of one SpMV kernel (bench: roofline of the block-sparse product)
        double tot = 0.0;
        int cnt = 0;
        for (int k = 0; k < std::min(L.spmv_count, (int)h->spmv_events.size() / 2); k++) {
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, h->spmv_events[2 * k], h->spmv_events[2 * k + 1]) == cudaSuccess) {
                tot += ms;
                cnt++;
            }
        }
        h->tm.gcw_spmv_ms = cnt > 0 ? tot / cnt : 0.0;
    }
    return DESC_B200_OK;
}
