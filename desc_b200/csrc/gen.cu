// the reference's data generators on the device.
//
//   Models/Uniform_Topology.m:24-111     (corruption 'uniform' / self-consistent)
//   Models/Nonuniform_Topology.m:26-157  (crpt_type 'uniform' / 'self-consistent' / 'adv')
//   + the SfM-shaped ring topology BASELINE.json's configs[4] needs (not in the reference): pair {i,j} can be
//     an edge only if the circular distance of i and j is <= window.
//
// Every statement of the two files is kept; every draw from MATLAB's global stream (rand / randn / randperm),
// which cannot be reproduced, is a counter-based draw -- a pure function of (seed, stream, item, index) on the
// sampler's 64-bit mix (desc_key).  oracle/desc_models_ctr.py is the same arithmetic in numpy and the
// parity tests compare value for value: Ind and the corruption mask bit-exact, rotations to 1e-12.
// Because draws are counters, the sequential loops of Nonuniform_Topology.m:80-124 (randperm over nodes, randperm
// over each node's neighbours, later visits overwriting earlier ones) become per-edge rules:
//   node i is corrupted  <=> its key is among the floor(n p_node_crpt) smallest (position = its rank);
//   i corrupts neighbour j <=> (key(i,j), j) is among the floor(p_edge_crpt deg_i) smallest of i's neighbours
//                              (one threshold per node, like the cycle sampler's);
//   an edge picked from both ends keeps the value written by the endpoint that comes later in the node order.
//
// Output layout = what the C ABI takes: Ind 2m doubles (all i, then all j; 1-based, i<j, (i,j)-sorted),
// RijMat / Rij_orig 9m doubles (MATLAB 3x3xm), R_orig 9n, ErrVec m.  HBM-bound streaming (72-160 B per edge),
// one thread per edge; the 3x3 SVD projection is the one-sided Jacobi of so3.cuh.
#include "internal.cuh"
#include "so3.cuh"

#include <algorithm>
#include <cmath>

struct desc_b200_model {
    int device = 0;
    cudaStream_t stream = nullptr;
    int n = 0;
    int64_t m = 0;
    int* ei = nullptr;
    int* ej = nullptr;
    double* Ind = nullptr;
    double* RijMat = nullptr;
    double* Rij_orig = nullptr;
    double* R_orig = nullptr;
    double* ErrVec = nullptr;
    uint8_t* corrupted = nullptr;
    double gen_ms = 0.0;
    int launches = 0;
};

namespace {
enum { S_ADJ = 1, S_RORIG, S_MASK, S_NOISE, S_RCORR, S_CORR, S_NODEPERM, S_NBRPERM, S_R0, S_NOISE_OUT };

__device__ __forceinline__ uint64_t g_u64(uint64_t seed, int stream, uint64_t a, uint64_t b) {
    return desc_key(seed + (uint64_t)stream * 0xA0761D6478BD642Full, a, b);
}
__device__ __forceinline__ double g_uniform(uint64_t seed, int stream, uint64_t a, uint64_t b) {
    return (double)(g_u64(seed, stream, a, b) >> 11) * 0x1p-53;
}
// randn(3) of `item`: out[r+3c] = normal number r+3c (MATLAB fills column-major); Box-Muller pairs (2t, 2t+1)
__device__ __forceinline__ void g_randn9(uint64_t seed, int stream, uint64_t item, double* out) {
#pragma unroll
    for (int t = 0; t < 5; t++) {
        const uint64_t z1 = g_u64(seed, stream, item, 2 * t), z2 = g_u64(seed, stream, item, 2 * t + 1);
        const double u1 = ((double)(z1 >> 11) + 1.0) * 0x1p-53;
        const double u2 = (double)(z2 >> 11) * 0x1p-53;
        const double r = sqrt(-2.0 * log(u1));
        const double ang = 2.0 * 3.14159265358979323846 * u2;
        out[2 * t] = r * cos(ang);
        if (t < 4) out[2 * t + 1] = r * sin(ang);
    }
}

// is pair lo < hi allowed by the topology (ring: circular distance <= window) -- the draw itself is separate
__device__ __forceinline__ bool g_edge(uint64_t seed, double p, int n, int lo, int hi) {
    return g_uniform(seed, S_ADJ, (uint64_t)lo, (uint64_t)hi) < p;
}

// warp per row `lo`: count (FILL=false) or write (FILL=true) the edges (lo, hi), hi ascending
// (Uniform_Topology.m:29-34: G = tril(rand(n,n)<p,-1); [Ind_j,Ind_i] = find(G))
template <bool FILL>
__global__ void __launch_bounds__(256)
k_gen_rows(int n, double p, int window, uint64_t seed, int* __restrict__ cnt, const int64_t* __restrict__ rowoff,
           int* __restrict__ ei, int* __restrict__ ej, double* __restrict__ Ind, int64_t m) {
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int lo = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; lo < n; lo += warps) {
        // candidate intervals [a0,a1] and [b0,b1] (second one empty for Erdos-Renyi)
        int a0 = lo + 1, a1 = n - 1, b0 = 1, b1 = 0;
        if (window > 0 && 2 * (int64_t)window < (int64_t)n - 1) {
            a1 = min(lo + window, n - 1);
            b0 = max(lo + n - window, a1 + 1);
            b1 = n - 1;
        }
        int64_t out = FILL ? rowoff[lo] : 0;
        int total = 0;
        for (int part = 0; part < 2; part++) {
            const int s0 = part ? b0 : a0, s1 = part ? b1 : a1;
            for (int base = s0; base <= s1; base += 32) {
                const int hi = base + lane;
                const bool on = hi <= s1 && g_edge(seed, p, n, lo, hi);
                const unsigned bal = __ballot_sync(0xffffffffu, on);
                if (FILL && on) {
                    const int64_t e = out + __popc(bal & ((1u << lane) - 1u));
                    ei[e] = lo;
                    ej[e] = hi;
                    Ind[e] = (double)(lo + 1);
                    Ind[m + e] = (double)(hi + 1);
                }
                out += __popc(bal);
                total += __popc(bal);
            }
        }
        if (!FILL && lane == 0) cnt[lo] = total;
    }
}

// exclusive scan of n ints into n+1 int64 (one block; n is a node count)
__global__ void __launch_bounds__(1024)
k_gen_scan(const int* __restrict__ cnt, int n, int64_t* __restrict__ off) {
    __shared__ int64_t part[1024];
    const int per = (n + 1023) / 1024;
    const int b = threadIdx.x * per, e = min(n, b + per);
    int64_t s = 0;
    for (int i = b; i < e; i++) s += cnt[i];
    part[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        int64_t run = 0;
        for (int t = 0; t < 1024; t++) {
            const int64_t v = part[t];
            part[t] = run;
            run += v;
        }
        off[n] = run;
    }
    __syncthreads();
    s = part[threadIdx.x];
    for (int i = b; i < e; i++) {
        off[i] = s;
        s += cnt[i];
    }
}

// Q=randn(3); [U,~,V]=svd(Q); R=U*diag([1,1,det(U*V')])*V'     (Uniform_Topology.m:39-45)
__global__ void k_gen_rot(int n, uint64_t seed, int stream, double* __restrict__ R) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double q[9], r[9];
    g_randn9(seed, stream, (uint64_t)i, q);
    proj_so3_dev(q, r);
#pragma unroll
    for (int x = 0; x < 9; x++) R[9 * (int64_t)i + x] = r[x];
}

// C = A * B' for column-major 3x3
__device__ __forceinline__ void mul_abt(const double* A, const double* B, double* C) {
#pragma unroll
    for (int c = 0; c < 3; c++)
#pragma unroll
        for (int r = 0; r < 3; r++) C[r + 3 * c] = A[r] * B[c] + A[r + 3] * B[c + 3] + A[r + 6] * B[c + 6];
}

// trace(Rij_orig * RijMat') accumulated like Uniform_Topology.m:94-101, then abs(acos((tr-1)/2))/pi
__device__ __forceinline__ double err_of(const double* Ro, const double* Rm) {
    double acc[3];
#pragma unroll
    for (int r = 0; r < 3; r++) {
        double a = __dmul_rn(Ro[r], Rm[r]);
        a = __dadd_rn(a, __dmul_rn(Ro[r + 3], Rm[r + 3]));
        a = __dadd_rn(a, __dmul_rn(Ro[r + 6], Rm[r + 6]));
        acc[r] = a;
    }
    const double tr = __dadd_rn(__dadd_rn(acc[0], acc[1]), acc[2]);
    return __ddiv_rn(abs_acos_dev(__ddiv_rn(__dadd_rn(tr, -1.0), 2.0)), 3.14159265358979323846);
}

struct EdgeArgs {
    int64_t m;
    const int *ei, *ej;
    const double *R_orig, *R_corr;
    double* RijMat;
    double* Rij_orig;
    double* ErrVec;
    uint8_t* corrupted;
    uint64_t seed;
    double q, sigma, sigma_out;
    int kind;   // Uniform_Topology: 0 'uniform', 1 self-consistent; Nonuniform_Topology: 2 'uniform', 3 'self-consistent', 4 'adv'
    // Nonuniform_Topology
    const int* npos;          // position of the node in node_crpt, -1 = not corrupted
    const uint64_t* thr_key;  // per node: the largest (key, neighbour) pair it corrupts
    const int* thr_j;         // -1: corrupts nobody
};

__global__ void __launch_bounds__(128)
k_gen_edges(EdgeArgs a) {
    const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= a.m) return;
    const int i = a.ei[e], j = a.ej[e];
    double Ri[9], Rj[9], Ro[9], A[9], N[9], Rm[9];
#pragma unroll
    for (int x = 0; x < 9; x++) {
        Ri[x] = a.R_orig[9 * (int64_t)i + x];
        Rj[x] = a.R_orig[9 * (int64_t)j + x];
    }
    mul_abt(Ri, Rj, Ro);                                   // Rij_orig = R_i R_j'   (:47-51)
    bool corr;
    if (a.kind <= 1) {                                     // ---- Uniform_Topology.m
        corr = !(g_uniform(a.seed, S_MASK, (uint64_t)e, 0) >= a.q);     // noiseIndLog = rand(1,m) >= q (:53)
        if (!corr) {
            g_randn9(a.seed, S_NOISE, (uint64_t)e, N);
#pragma unroll
            for (int x = 0; x < 9; x++) A[x] = Ro[x] + a.sigma * N[x];  // :57-58
        } else if (a.kind == 0) {
            g_randn9(a.seed, S_CORR, (uint64_t)e, A);                   // :77-82
        } else {
            double Ci[9], Cj[9], Q[9];
#pragma unroll
            for (int x = 0; x < 9; x++) {
                Ci[x] = a.R_corr[9 * (int64_t)i + x];
                Cj[x] = a.R_corr[9 * (int64_t)j + x];
            }
            mul_abt(Ci, Cj, Q);
            g_randn9(a.seed, S_CORR, (uint64_t)e, N);
#pragma unroll
            for (int x = 0; x < 9; x++) A[x] = Q[x] + a.sigma * N[x];   // :84-90
        }
    } else {                                               // ---- Nonuniform_Topology.m:80-143
        const int pi = a.npos[i], pj = a.npos[j];
        bool si = false, sj = false;
        if (pi >= 0 && a.thr_j[i] >= 0) {
            const uint64_t k = g_u64(a.seed, S_NBRPERM, (uint64_t)i, (uint64_t)j);
            si = k < a.thr_key[i] || (k == a.thr_key[i] && j <= a.thr_j[i]);
        }
        if (pj >= 0 && a.thr_j[j] >= 0) {
            const uint64_t k = g_u64(a.seed, S_NBRPERM, (uint64_t)j, (uint64_t)i);
            sj = k < a.thr_key[j] || (k == a.thr_key[j] && i <= a.thr_j[j]);
        }
        corr = si || sj;
        if (corr) {
            const bool wi = si && (!sj || pi > pj);        // the later visit overwrites
            const int w = wi ? i : j, o = wi ? j : i;
            double M[9];
            if (a.kind == 2) {
                double q9[9];
                g_randn9(a.seed, S_R0, 2ull * (uint64_t)e + (wi ? 0ull : 1ull), q9);
                proj_so3_dev(q9, M);                       // :88-94
            } else {
                double Cw[9], Co[9];
                const double* src = a.kind == 3 ? a.R_corr : a.R_orig;   // :104-109 / :112-118
#pragma unroll
                for (int x = 0; x < 9; x++) {
                    Cw[x] = a.R_corr[9 * (int64_t)w + x];
                    Co[x] = src[9 * (int64_t)o + x];
                }
                mul_abt(Cw, Co, M);
            }
            g_randn9(a.seed, S_NOISE_OUT, (uint64_t)e, N);
#pragma unroll
            for (int c = 0; c < 3; c++)
#pragma unroll
                for (int r = 0; r < 3; r++)               // k>0 (w is the smaller endpoint): M, else M'
                    A[r + 3 * c] = (wi ? M[r + 3 * c] : M[c + 3 * r]) + a.sigma_out * N[r + 3 * c];   // :132-133
        } else {
            g_randn9(a.seed, S_NOISE, (uint64_t)e, N);
#pragma unroll
            for (int x = 0; x < 9; x++) A[x] = Ro[x] + a.sigma * N[x];  // :128-130
        }
    }
    proj_so3_dev(A, Rm);                                   // project back to SO(3)
#pragma unroll
    for (int x = 0; x < 9; x++) {
        a.RijMat[9 * e + x] = Rm[x];
        a.Rij_orig[9 * e + x] = Ro[x];
    }
    a.ErrVec[e] = err_of(Ro, Rm);
    a.corrupted[e] = corr ? 1 : 0;
}

// Nonuniform_Topology.m:62-64,81-86: block per node.  npos = rank of the node's key (randperm position) if it is
// among the n_node smallest, else -1; threshold = the floor(p_edge_crpt*deg)-th smallest (key, neighbour) pair.
__global__ void __launch_bounds__(256)
k_gen_node_sel(int n, double p, uint64_t seed, int n_node, double p_edge, int* __restrict__ npos,
               uint64_t* __restrict__ thr_key, int* __restrict__ thr_j) {
    extern __shared__ unsigned char smem_raw[];
    uint64_t* keys = reinterpret_cast<uint64_t*>(smem_raw);
    int* nbrs = reinterpret_cast<int*>(keys + n);
    __shared__ int s_cnt, s_rank;
    const int i = blockIdx.x;
    if (threadIdx.x == 0) {
        s_cnt = 0;
        s_rank = 0;
    }
    __syncthreads();
    const uint64_t mykey = g_u64(seed, S_NODEPERM, (uint64_t)i, 0);
    int smaller = 0;
    for (int v = threadIdx.x; v < n; v += blockDim.x) {
        const uint64_t k = g_u64(seed, S_NODEPERM, (uint64_t)v, 0);
        if (k < mykey || (k == mykey && v < i)) smaller++;
    }
    atomicAdd(&s_rank, smaller);
    __syncthreads();
    const int pos = s_rank;
    if (pos >= n_node) {
        if (threadIdx.x == 0) {
            npos[i] = -1;
            thr_j[i] = -1;
            thr_key[i] = 0;
        }
        return;
    }
    for (int v = threadIdx.x; v < n; v += blockDim.x) {
        if (v == i) continue;
        if (g_edge(seed, p, n, min(i, v), max(i, v))) {
            const int slot = atomicAdd(&s_cnt, 1);
            keys[slot] = g_u64(seed, S_NBRPERM, (uint64_t)i, (uint64_t)v);
            nbrs[slot] = v;
        }
    }
    __syncthreads();
    const int deg = s_cnt;
    const int nn = (int)floor(p_edge * (double)deg);
    if (threadIdx.x == 0) {
        npos[i] = pos;
        if (nn <= 0) {
            thr_j[i] = -1;
            thr_key[i] = 0;
        }
    }
    if (nn <= 0) return;
    for (int x = threadIdx.x; x < deg; x += blockDim.x) {
        const uint64_t kx = keys[x];
        const int jx = nbrs[x];
        int c = 0;
        for (int y = 0; y < deg; y++) {
            const uint64_t ky = keys[y];
            c += (ky < kx || (ky == kx && nbrs[y] < jx)) ? 1 : 0;
        }
        if (c == nn - 1) {   // exactly one element has this rank
            thr_key[i] = kx;
            thr_j[i] = jx;
        }
    }
}

#define GEN_CHECK(mo)                                                                         \
    do {                                                                                      \
        (mo)->launches++;                                                                     \
        CUDA_TRY(cudaGetLastError());                                                         \
    } while (0)

int generate(desc_b200_model* mo, const desc_b200_gen_opts& o) {
    const int n = o.n;
    cudaStream_t st = mo->stream;
    cudaEvent_t e0, e1;
    CUDA_TRY(cudaEventCreate(&e0));
    CUDA_TRY(cudaEventCreate(&e1));
    CUDA_TRY(cudaEventRecord(e0, st));
    int* cnt = nullptr;
    int64_t* off = nullptr;
    CUDA_TRY(cudaMalloc(&cnt, (size_t)n * sizeof(int)));
    CUDA_TRY(cudaMalloc(&off, (size_t)(n + 1) * sizeof(int64_t)));
    const int grid = DESC_SMS * 8;
    const int window = o.topology == 1 ? o.window : 0;
    k_gen_rows<false><<<grid, 256, 0, st>>>(n, o.p, window, o.seed, cnt, nullptr, nullptr, nullptr, nullptr, 0);
    GEN_CHECK(mo);
    k_gen_scan<<<1, 1024, 0, st>>>(cnt, n, off);
    GEN_CHECK(mo);
    int64_t m = 0;
    CUDA_TRY(cudaMemcpyAsync(&m, off + n, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    if (m <= 0 || m > DESC_MAX_EDGES) {
        cudaFree(cnt);
        cudaFree(off);
        desc_set_error("generator produced m=%lld edges (need 1..2^30-1)", (long long)m);
        return m <= 0 ? DESC_B200_ERR_ARG : DESC_B200_ERR_LIMIT;
    }
    mo->m = m;
    CUDA_TRY(cudaMalloc(&mo->ei, (size_t)m * sizeof(int)));
    CUDA_TRY(cudaMalloc(&mo->ej, (size_t)m * sizeof(int)));
    CUDA_TRY(cudaMalloc(&mo->Ind, (size_t)2 * m * sizeof(double)));
    CUDA_TRY(cudaMalloc(&mo->RijMat, (size_t)9 * m * sizeof(double)));
    CUDA_TRY(cudaMalloc(&mo->Rij_orig, (size_t)9 * m * sizeof(double)));
    CUDA_TRY(cudaMalloc(&mo->ErrVec, (size_t)m * sizeof(double)));
    CUDA_TRY(cudaMalloc(&mo->corrupted, (size_t)m));
    CUDA_TRY(cudaMalloc(&mo->R_orig, (size_t)9 * n * sizeof(double)));
    k_gen_rows<true><<<grid, 256, 0, st>>>(n, o.p, window, o.seed, nullptr, off, mo->ei, mo->ej, mo->Ind, m);
    GEN_CHECK(mo);
    double* R_corr = nullptr;
    CUDA_TRY(cudaMalloc(&R_corr, (size_t)9 * n * sizeof(double)));
    k_gen_rot<<<(n + 127) / 128, 128, 0, st>>>(n, o.seed, S_RORIG, mo->R_orig);
    GEN_CHECK(mo);
    k_gen_rot<<<(n + 127) / 128, 128, 0, st>>>(n, o.seed, S_RCORR, R_corr);
    GEN_CHECK(mo);
    EdgeArgs a = {};
    a.m = m;
    a.ei = mo->ei;
    a.ej = mo->ej;
    a.R_orig = mo->R_orig;
    a.R_corr = R_corr;
    a.RijMat = mo->RijMat;
    a.Rij_orig = mo->Rij_orig;
    a.ErrVec = mo->ErrVec;
    a.corrupted = mo->corrupted;
    a.seed = o.seed;
    a.q = o.q;
    a.sigma = o.sigma;
    a.sigma_out = o.sigma_out;
    a.kind = o.kind;
    int* npos = nullptr;
    uint64_t* thr_key = nullptr;
    int* thr_j = nullptr;
    if (o.kind >= 2) {
        CUDA_TRY(cudaMalloc(&npos, (size_t)n * sizeof(int)));
        CUDA_TRY(cudaMalloc(&thr_key, (size_t)n * sizeof(uint64_t)));
        CUDA_TRY(cudaMalloc(&thr_j, (size_t)n * sizeof(int)));
        const size_t smem = (size_t)n * 12;
        CUDA_TRY(cudaFuncSetAttribute(k_gen_node_sel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const int n_node = (int)std::floor((double)n * o.p_node_crpt);       // Nonuniform_Topology.m:63
        k_gen_node_sel<<<n, 256, smem, st>>>(n, o.p, o.seed, n_node, o.p_edge_crpt, npos, thr_key, thr_j);
        GEN_CHECK(mo);
        a.npos = npos;
        a.thr_key = thr_key;
        a.thr_j = thr_j;
    }
    k_gen_edges<<<(unsigned)((m + 127) / 128), 128, 0, st>>>(a);
    GEN_CHECK(mo);
    CUDA_TRY(cudaEventRecord(e1, st));
    CUDA_TRY(cudaEventSynchronize(e1));
    float ms = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
    mo->gen_ms = ms;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(cnt);
    cudaFree(off);
    cudaFree(R_corr);
    if (npos) cudaFree(npos);
    if (thr_key) cudaFree(thr_key);
    if (thr_j) cudaFree(thr_j);
    return DESC_B200_OK;
}
}  // namespace

extern "C" {

void desc_b200_model_destroy(desc_b200_model* mo) {
    if (!mo) return;
    cudaSetDevice(mo->device);
    if (mo->stream) cudaStreamSynchronize(mo->stream);
    void* ptrs[] = {mo->ei, mo->ej, mo->Ind, mo->RijMat, mo->Rij_orig, mo->R_orig, mo->ErrVec, mo->corrupted};
    for (void* p : ptrs)
        if (p) cudaFree(p);
    if (mo->stream) cudaStreamDestroy(mo->stream);
    delete mo;
}

int desc_b200_generate(const desc_b200_gen_opts* opts, desc_b200_model** out) {
    if (!out || !opts) {
        desc_set_error("null generator options / output");
        return DESC_B200_ERR_ARG;
    }
    *out = nullptr;
    const desc_b200_gen_opts o = *opts;
    const bool prob_ok = o.p > 0.0 && o.p <= 1.0 && o.sigma >= 0.0;
    if (o.n < 2 || !prob_ok || o.kind < 0 || o.kind > 4 || o.topology < 0 || o.topology > 1 ||
        (o.topology == 1 && (o.window < 1 || o.kind > 1)) ||
        (o.kind <= 1 && !(o.q >= 0.0 && o.q <= 1.0)) ||
        (o.kind >= 2 && !(o.p_node_crpt >= 0.0 && o.p_node_crpt <= 1.0 && o.p_edge_crpt >= 0.0 && o.p_edge_crpt <= 1.0 &&
                          o.sigma_out >= 0.0))) {
        desc_set_error("bad generator options (n=%d p=%g q=%g sigma=%g kind=%d topology=%d window=%d)", o.n, o.p, o.q,
                       o.sigma, o.kind, o.topology, o.window);
        return DESC_B200_ERR_ARG;
    }
    if (o.kind >= 2 && (size_t)o.n * 12 > 220 * 1024) {
        desc_set_error("Nonuniform_Topology on the device keeps a node's neighbour keys in shared memory: n <= %d", 220 * 1024 / 12);
        return DESC_B200_ERR_LIMIT;
    }
    const int ndev = desc_b200_device_count();
    if (ndev <= 0) {
        if (ndev == 0) desc_set_error("no CUDA device (this library has no CPU fallback)");
        return DESC_B200_ERR_CUDA;
    }
    int dev = o.device;
    if (dev < 0) CUDA_TRY(cudaGetDevice(&dev));
    if (dev >= ndev) {
        desc_set_error("device %d requested but only %d visible", dev, ndev);
        return DESC_B200_ERR_ARG;
    }
    CUDA_TRY(cudaSetDevice(dev));
    desc_b200_model* mo = new desc_b200_model();
    mo->device = dev;
    mo->n = o.n;
    if (cudaStreamCreateWithFlags(&mo->stream, cudaStreamNonBlocking) != cudaSuccess) {
        desc_set_error("cudaStreamCreate failed");
        delete mo;
        return DESC_B200_ERR_CUDA;
    }
    const int rc = generate(mo, o);
    if (rc != DESC_B200_OK) {
        desc_b200_model_destroy(mo);
        return rc;
    }
    *out = mo;
    return DESC_B200_OK;
}

int desc_b200_model_info(desc_b200_model* mo, int64_t info[4], double* gen_ms) {
    if (!mo || !info) {
        desc_set_error("null model");
        return DESC_B200_ERR_ARG;
    }
    info[0] = mo->n;
    info[1] = mo->m;
    info[2] = mo->launches;
    info[3] = mo->device;
    if (gen_ms) *gen_ms = mo->gen_ms;
    return DESC_B200_OK;
}

int desc_b200_model_fetch(desc_b200_model* mo, double* Ind, double* RijMat, double* R_orig, double* ErrVec,
                          double* Rij_orig, uint8_t* corrupted) {
    if (!mo) {
        desc_set_error("null model");
        return DESC_B200_ERR_ARG;
    }
    CUDA_TRY(cudaSetDevice(mo->device));
    cudaStream_t st = mo->stream;
    const size_t m = (size_t)mo->m, n = (size_t)mo->n;
    if (Ind) CUDA_TRY(cudaMemcpyAsync(Ind, mo->Ind, 2 * m * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (RijMat) CUDA_TRY(cudaMemcpyAsync(RijMat, mo->RijMat, 9 * m * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (R_orig) CUDA_TRY(cudaMemcpyAsync(R_orig, mo->R_orig, 9 * n * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (ErrVec) CUDA_TRY(cudaMemcpyAsync(ErrVec, mo->ErrVec, m * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (Rij_orig) CUDA_TRY(cudaMemcpyAsync(Rij_orig, mo->Rij_orig, 9 * m * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (corrupted) CUDA_TRY(cudaMemcpyAsync(corrupted, mo->corrupted, m, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return DESC_B200_OK;
}

int desc_b200_model_device(desc_b200_model* mo, const double** Ind, const double** RijMat, const double** R_orig,
                           const double** ErrVec) {
    if (!mo) {
        desc_set_error("null model");
        return DESC_B200_ERR_ARG;
    }
    if (Ind) *Ind = mo->Ind;
    if (RijMat) *RijMat = mo->RijMat;
    if (R_orig) *R_orig = mo->R_orig;
    if (ErrVec) *ErrVec = mo->ErrVec;
    return DESC_B200_OK;
}

}  // extern "C"
