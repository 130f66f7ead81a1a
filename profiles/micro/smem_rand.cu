// Microbenchmark: cost (LSU cycles per warp instruction) of random shared-memory table accesses with
// 30 active lanes: 64-bit words vs split 32-bit hi/lo arrays, load-only and read-modify-write.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(int steps, int tstride, long long* out, double* sink) {
    extern __shared__ double T[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double* Tw = T + (size_t)warp * tstride;
    uint32_t* Th = reinterpret_cast<uint32_t*>(Tw);
    uint32_t* Tl = Th + tstride;
    for (int i = lane; i < tstride; i += 32) Tw[i] = 0.0;
    __syncwarp();
    uint32_t x = 1234567u * (blockIdx.x * blockDim.x + threadIdx.x + 1);
    const int per = tstride / 32;
    double acc = 0.0;
    uint32_t acci = 0;
    long long t0 = clock64();
    for (int s = 0; s < steps; s++) {
        x = x * 1664525u + 1013904223u;
        const int idx = lane * per + (int)(((x >> 16) * (uint32_t)per) >> 16);
        if (lane < 30) {
            if (MODE == 0) acc += Tw[idx];                                   // 64-bit load
            if (MODE == 1) { acci += Th[idx]; acci += Tl[idx]; }             // 2 x 32-bit loads
            if (MODE == 2) { Tw[idx] = Tw[idx] + 1.0; }                      // 64-bit rmw
            if (MODE == 3) {                                                 // split rmw
                double v = __hiloint2double((int)Th[idx], (int)Tl[idx]) + 1.0;
                Th[idx] = (uint32_t)__double2hiint(v);
                Tl[idx] = (uint32_t)__double2loint(v);
            }
        }
        if (MODE >= 2) __syncwarp();
    }
    long long t1 = clock64();
    if (lane == 0) out[blockIdx.x * (blockDim.x >> 5) + warp] = t1 - t0;
    if (acc == -1.0 || acci == 77u) *sink = acc;
}
template <int MODE>
void run(const char* name, long long* d_out, double* d_sink) {
    const int tstride = 1152, steps = 20000, wpb = 16;
    size_t smem = (size_t)wpb * tstride * 8;
    cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k<MODE><<<148, wpb * 32, smem>>>(steps, tstride, d_out, d_sink);
    cudaDeviceSynchronize();
    long long h[148 * 32];
    cudaMemcpy(h, d_out, 148 * wpb * 8, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < 148 * wpb; i++) avg += h[i]; avg /= 148 * wpb;
    printf("%s: %.2f SM-cycles per warp-step at saturation (16 warps/SM)\n", name, avg / steps / wpb);
}
int main() {
    long long* d_out; double* d_sink;
    cudaMalloc(&d_out, 148 * 64 * sizeof(long long));
    cudaMalloc(&d_sink, 8);
    run<0>("load 64-bit      ", d_out, d_sink);
    run<1>("load 2x32-bit    ", d_out, d_sink);
    run<2>("rmw 64-bit       ", d_out, d_sink);
    run<3>("rmw split 32-bit ", d_out, d_sink);
    printf("status %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
