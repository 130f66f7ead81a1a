// Small SO(3) device helpers shared by cycle.cu, gcw.cu and diag.cu.
#pragma once
#include <cuda_runtime.h>

// abs(acos(x)) with MATLAB's complex branches for |x|>1 (DESC.m:147): acosh(x) for x>1,
// sqrt(pi^2+acosh(-x)^2) for x<-1.
__device__ __forceinline__ double abs_acos_dev(double x) {
    if (x > 1.0) {
        double t = x - 1.0;
        return log1p(t + sqrt(t * (t + 2.0)));
    }
    if (x < -1.0) {
        double t = -x - 1.0;
        double a = log1p(t + sqrt(t * (t + 2.0)));
        const double pi = 3.14159265358979323846;
        return sqrt(pi * pi + a * a);
    }
    return acos(x);
}


// nearest rotation in the reference's sense: [U,~,V]=svd(A); U*diag(1,1,det(U*V'))*V'
// one-sided (Hestenes) Jacobi SVD of a 3x3, singular values sorted descending like LAPACK.
__device__ inline void proj_so3_dev(const double* Ain, double* Rout) {
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

