// Microbenchmark: latency/throughput of the private-table read-modify-write chain used by the PGD
// scatter (LDS.64 random -> DADD -> STS.64 -> __syncwarp), as a function of warps per SM, and of a
// divergent 8-byte global gather from an L2-resident table.  Build: nvcc -arch=sm_100a -O3.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void k_rmw(int steps, int tstride, long long* out, double* sink) {
    extern __shared__ double T[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double* Tw = T + (size_t)warp * tstride;
    for (int i = lane; i < tstride; i += 32) Tw[i] = 0.0;
    __syncwarp();
    uint32_t x = 1234567u * (blockIdx.x * blockDim.x + threadIdx.x + 1);
    // distinct ranks per lane: lane-striped random offsets
    const int per = tstride / 32;
    long long t0 = clock64();
    double acc = 0.0;
    for (int s = 0; s < steps; s++) {
        x = x * 1664525u + 1013904223u;
        const int idx = lane * per + (int)(((x >> 16) * (uint32_t)per) >> 16);
        const double t = Tw[idx];
        Tw[idx] = t + 1.0;
        __syncwarp();
    }
    long long t1 = clock64();
    if (lane == 0) out[blockIdx.x * (blockDim.x >> 5) + warp] = t1 - t0;
    for (int i = lane; i < tstride; i += 32) acc += Tw[i];
    if (acc == -1.0) *sink = acc;
}
__global__ void k_gather(int steps, const double* __restrict__ S, int nS, long long* out, double* sink) {
    uint32_t x = 1234567u * (blockIdx.x * blockDim.x + threadIdx.x + 1);
    double acc = 0.0;
    long long t0 = clock64();
    for (int s = 0; s < steps; s += 4) {
        double v[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            x = x * 1664525u + 1013904223u;
            v[u] = S[(x >> 4) % (uint32_t)nS];
        }
        acc += v[0] + v[1] + v[2] + v[3];
    }
    long long t1 = clock64();
    if ((threadIdx.x & 31) == 0) out[(blockIdx.x * blockDim.x + threadIdx.x) >> 5] = t1 - t0;
    if (acc == -1.0) *sink = acc;
}
int main() {
    long long* d_out; double* d_sink; double* d_S;
    const int nS = 5000000;
    cudaMalloc(&d_out, 148 * 64 * sizeof(long long));
    cudaMalloc(&d_sink, 8);
    cudaMalloc(&d_S, nS * 8);
    cudaMemset(d_S, 0, nS * 8);
    const int tstride = 1152, steps = 20000;
    for (int wpb : {1, 2, 4, 8, 16, 24}) {
        size_t smem = (size_t)wpb * tstride * 8;
        cudaFuncSetAttribute(k_rmw, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        k_rmw<<<148, wpb * 32, smem>>>(steps, tstride, d_out, d_sink);
        cudaDeviceSynchronize();
        long long h[148 * 32];
        cudaMemcpy(h, d_out, 148 * wpb * 8, cudaMemcpyDeviceToHost);
        double avg = 0; for (int i = 0; i < 148 * wpb; i++) avg += h[i]; avg /= 148 * wpb;
        printf("rmw: %2d warps/SM: %.1f cycles/step/warp -> %.3f slots/cycle/SM (30 lanes/step)\n", wpb, avg / steps, wpb * 30.0 / (avg / steps));
    }
    for (int wpb : {4, 8, 16, 32, 64}) {
        int blocks = 148 * (wpb > 32 ? 2 : 1), tpb = (wpb > 32 ? wpb / 2 : wpb) * 32;
        k_gather<<<blocks, tpb>>>(2000, d_S, nS, d_out, d_sink);
        cudaDeviceSynchronize();
        long long h[148 * 64];
        cudaMemcpy(h, d_out, 148 * wpb * 8, cudaMemcpyDeviceToHost);
        double avg = 0; for (int i = 0; i < 148 * wpb; i++) avg += h[i]; avg /= 148 * wpb;
        printf("gather: %2d warps/SM: %.1f cycles/gather-instr/warp -> %.3f gathers/cycle/SM\n", wpb, avg / 2000, wpb * 32.0 / (avg / 2000));
    }
    cudaError_t e = cudaGetLastError();
    printf("status %s\n", cudaGetErrorString(e));
    return 0;
}
