// the CEMP+MST initialisation of MPLS on the device (reference: Algorithms/MPLS.m:152-195).
//
//   SMatij = sparse(Ind_j, Ind_i, SVec+1, n, n); Tree = minspantree(graph(SMatij,'lower'));        :154-160
//   R_1 = I; breadth-first from node 1: R_leaf = Rij * R_root (leaf < root) or Rij' * R_root        :163-186
//
// Minimum spanning tree: Boruvka rounds (every component picks its lightest outgoing edge with two atomicMin
// passes -- weight bits, then edge id among the lightest -- hooks onto the other component, mutual picks keep the
// smaller label as root, pointer jumping relabels the nodes).  `minspantree` does not document its tie-breaking;
// here ties are broken by the edge index, so the tree is the unique MST under the total order (weight, edge id)
// -- the same tree Kruskal builds in oracle/desc_oracle.py::mst_init.  <= log2(n)+1 rounds of O(m) work,
// 12 B/edge streamed per pass: HBM-bound, ~3 passes per round.
// Propagation: one CTA walks the n-1 tree edges level by level (a node's rotation is written once, by its tree
// parent, and read only in later rounds), so the result does not depend on thread scheduling.
#include "internal.cuh"

#include <algorithm>
#include <climits>

namespace {
constexpr int MST_TB = 256;

__global__ void k_mst_init(int n, int* __restrict__ comp) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v < n) comp[v] = v;
}
__global__ void k_mst_reset(int n, unsigned long long* __restrict__ best_w, int* __restrict__ best_e,
                            int* __restrict__ parent) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n) return;
    best_w[v] = ~0ull;
    best_e[v] = INT_MAX;
    parent[v] = v;
}
__device__ __forceinline__ unsigned long long mst_wbits(const double* __restrict__ S, int64_t e) {
    return (unsigned long long)__double_as_longlong(S[e] + 1.0);   // SVec+1 > 0: the bit pattern is monotone
}
__global__ void k_mst_minw(const int* __restrict__ ei, const int* __restrict__ ej, const double* __restrict__ S,
                           int64_t m, const int* __restrict__ comp, unsigned long long* __restrict__ best_w) {
    const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= m) return;
    const int a = comp[ei[e]], b = comp[ej[e]];
    if (a == b) return;
    const unsigned long long w = mst_wbits(S, e);
    atomicMin(&best_w[a], w);
    atomicMin(&best_w[b], w);
}
__global__ void k_mst_mine(const int* __restrict__ ei, const int* __restrict__ ej, const double* __restrict__ S,
                           int64_t m, const int* __restrict__ comp, const unsigned long long* __restrict__ best_w,
                           int* __restrict__ best_e) {
    const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= m) return;
    const int a = comp[ei[e]], b = comp[ej[e]];
    if (a == b) return;
    const unsigned long long w = mst_wbits(S, e);
    if (w == best_w[a]) atomicMin(&best_e[a], (int)e);
    if (w == best_w[b]) atomicMin(&best_e[b], (int)e);
}
// per component root: take the picked edge into the tree and hook onto the other component
__global__ void k_mst_hook(int n, const int* __restrict__ ei, const int* __restrict__ ej, const int* __restrict__ comp,
                           const int* __restrict__ best_e, int* __restrict__ parent, uint8_t* __restrict__ in_tree,
                           int* __restrict__ merged) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n || comp[c] != c) return;
    const int e = best_e[c];
    if (e == INT_MAX) return;
    const int a = comp[ei[e]], b = comp[ej[e]];
    const int other = a == c ? b : a;
    in_tree[e] = 1;
    if (best_e[other] == e && c < other) return;   // mutual pick: the smaller label stays the root
    parent[c] = other;
    atomicAdd(merged, 1);
}
__global__ void k_mst_jump(int n, const int* __restrict__ comp, const int* __restrict__ parent, int* __restrict__ comp_out) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n) return;
    int r = comp[v];
    while (parent[r] != r) r = parent[r];
    comp_out[v] = r;
}
__global__ void k_mst_compact(const uint8_t* __restrict__ in_tree, int64_t m, int* __restrict__ list, int* __restrict__ count) {
    const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e < m && in_tree[e]) list[atomicAdd(count, 1)] = (int)e;
}

// MPLS.m:163-186 -- level-synchronous walk of the tree from node 0, one CTA
__global__ void __launch_bounds__(1024)
k_mst_propagate(int n, int ntree, const int* __restrict__ list, const int* __restrict__ ei, const int* __restrict__ ej,
                const double* __restrict__ Rij, int* level, double* R, int* __restrict__ n_added) {
    for (int v = threadIdx.x; v < n; v += blockDim.x) level[v] = v == 0 ? 0 : -1;
    if (threadIdx.x < 9) R[threadIdx.x] = (threadIdx.x % 4 == 0) ? 1.0 : 0.0;   // R_est(:,:,1) = eye(3)
    __syncthreads();
    for (int round = 0; round < n; round++) {
        int mine = 0;
        for (int t = threadIdx.x; t < ntree; t += blockDim.x) {
            const int e = list[t];
            const int i = ei[e], j = ej[e];
            const int li = level[i], lj = level[j];
            int root, leaf;
            if (li >= 0 && li <= round && lj < 0) {
                root = i;
                leaf = j;
            } else if (lj >= 0 && lj <= round && li < 0) {
                root = j;
                leaf = i;
            } else {
                continue;
            }
            // IndMat(leaf, root) > 0 <=> leaf < root: R_leaf = Rij * R_root, else Rij' * R_root    (:178-182)
            const bool tr = !(leaf < root);
            const double* A = Rij + 9 * (int64_t)e;
            const double* B = R + 9 * (int64_t)root;
            double* O = R + 9 * (int64_t)leaf;
            for (int c = 0; c < 3; c++)
                for (int r = 0; r < 3; r++) {
                    double s = 0.0;
                    for (int x = 0; x < 3; x++) s += (tr ? A[x + 3 * r] : A[r + 3 * x]) * B[x + 3 * c];
                    O[r + 3 * c] = s;
                }
            level[leaf] = round + 1;
            mine++;
        }
        __threadfence_block();
        if (__syncthreads_count(mine > 0) == 0) break;
    }
    __syncthreads();
    int cnt = 0;
    for (int v = threadIdx.x; v < n; v += blockDim.x) cnt += level[v] >= 0 ? 1 : 0;
    atomicAdd(n_added, cnt);
}
}  // namespace

int desc_mst_init_impl(desc_b200_handle* h, const double* d_S, double* d_R) {
    // multi-GPU handles: replicated on every rank (O(m log n) work on the replicated edge list, 7 ms at cfg 4)
    const int n = h->n;
    const int64_t m = h->m;
    cudaStream_t st = h->stream;
    int *comp[2], *parent, *best_e, *list, *level, *ctr;
    unsigned long long* best_w;
    uint8_t* in_tree;
    CUDA_TRY(cudaMalloc(&comp[0], (size_t)n * sizeof(int)));
    CUDA_TRY(cudaMalloc(&comp[1], (size_t)n * sizeof(int)));
    CUDA_TRY(cudaMalloc(&parent, (size_t)n * sizeof(int)));
    CUDA_TRY(cudaMalloc(&best_e, (size_t)n * sizeof(int)));
    CUDA_TRY(cudaMalloc(&list, (size_t)n * sizeof(int)));
    CUDA_TRY(cudaMalloc(&level, (size_t)n * sizeof(int)));
    CUDA_TRY(cudaMalloc(&ctr, 4 * sizeof(int)));
    CUDA_TRY(cudaMalloc(&best_w, (size_t)n * sizeof(unsigned long long)));
    CUDA_TRY(cudaMalloc(&in_tree, (size_t)m));
    void* to_free[] = {comp[0], comp[1], parent, best_e, list, level, ctr, best_w, in_tree};
    auto cleanup = [&]() {
        for (void* p : to_free) cudaFree(p);
    };
    auto body = [&]() -> int {
        const unsigned gn = (unsigned)((n + MST_TB - 1) / MST_TB), gm = (unsigned)((m + MST_TB - 1) / MST_TB);
        CUDA_TRY(cudaMemsetAsync(in_tree, 0, (size_t)m, st));
        k_mst_init<<<gn, MST_TB, 0, st>>>(n, comp[0]);
        KERNEL_CHECK(h);
        int cur = 0;
        for (int round = 0; round < 64; round++) {
            CUDA_TRY(cudaMemsetAsync(ctr, 0, 4 * sizeof(int), st));
            k_mst_reset<<<gn, MST_TB, 0, st>>>(n, best_w, best_e, parent);
            KERNEL_CHECK(h);
            k_mst_minw<<<gm, MST_TB, 0, st>>>(h->ei, h->ej, d_S, m, comp[cur], best_w);
            KERNEL_CHECK(h);
            k_mst_mine<<<gm, MST_TB, 0, st>>>(h->ei, h->ej, d_S, m, comp[cur], best_w, best_e);
            KERNEL_CHECK(h);
            k_mst_hook<<<gn, MST_TB, 0, st>>>(n, h->ei, h->ej, comp[cur], best_e, parent, in_tree, ctr);
            KERNEL_CHECK(h);
            int merged = 0;
            CUDA_TRY(cudaMemcpyAsync(&merged, ctr, sizeof(int), cudaMemcpyDeviceToHost, st));
            CUDA_TRY(cudaStreamSynchronize(st));
            if (merged == 0) break;
            k_mst_jump<<<gn, MST_TB, 0, st>>>(n, comp[cur], parent, comp[cur ^ 1]);
            KERNEL_CHECK(h);
            cur ^= 1;
        }
        CUDA_TRY(cudaMemsetAsync(ctr, 0, 4 * sizeof(int), st));
        k_mst_compact<<<gm, MST_TB, 0, st>>>(in_tree, m, list, ctr + 1);
        KERNEL_CHECK(h);
        int ntree = 0;
        CUDA_TRY(cudaMemcpyAsync(&ntree, ctr + 1, sizeof(int), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        if (ntree != n - 1) {
            desc_set_error("graph is not connected: the spanning forest has %d edges for %d nodes (the reference's "
                           "loop MPLS.m:171 would not terminate)", ntree, n);
            return DESC_B200_ERR_ARG;
        }
        k_mst_propagate<<<1, 1024, 0, st>>>(n, ntree, list, h->ei, h->ej, h->Rij, level, d_R, ctr + 2);
        KERNEL_CHECK(h);
        int added = 0;
        CUDA_TRY(cudaMemcpyAsync(&added, ctr + 2, sizeof(int), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        if (added != n) {
            desc_set_error("tree propagation reached %d of %d nodes", added, n);
            return DESC_B200_ERR_STATE;
        }
        return DESC_B200_OK;
    };
    const int rc = body();
    cleanup();
    return rc;
}
