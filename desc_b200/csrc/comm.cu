// Multi-GPU plumbing: one process per GPU, NCCL over NVLink/NVSwitch.
//
// The reference has no communication layer at all (single MATLAB process); what is exchanged
// here follows from sharding DESC.m's per-edge loops over contiguous edge ranges:
//   * per PGD iteration: all-reduce (sum) of the 2m partner-sum accumulators + 2 scalars
//     (DESC.m:189-190 in scatter form, DESC.m:232-233) and all-gather of the S_vec shards
//     (DESC.m:193 gathers S at arbitrary edges);
//   * per GCW power step: all-gather of the per-node 3x3 blocks (9n doubles);
//   * once in the build: all-gather of co-degrees and apex lists.
//
// NCCL is bound at run time with dlopen, so the single-GPU library (and the MEX file that wraps
// it) has no link-time dependency on NCCL; inside a torch process the already-loaded
// torch-bundled libnccl.so.2 is reused.
#include "internal.cuh"

#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <map>
#include <mutex>

namespace {
struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*ReduceScatter)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};
NcclApi g_nccl;

int nccl_load() {
    if (g_nccl.ok) return DESC_B200_OK;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
        g_nccl.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (g_nccl.lib) break;
    }
    if (!g_nccl.lib) {
        desc_set_error("multi-GPU requested but libnccl.so.2 cannot be loaded: %s", dlerror());
        return DESC_B200_ERR_NCCL;
    }
#define LOAD(field, sym)                                                              \
    *(void**)(&g_nccl.field) = dlsym(g_nccl.lib, sym);                                \
    if (!g_nccl.field) {                                                              \
        desc_set_error("libnccl is missing symbol %s", sym);                          \
        return DESC_B200_ERR_NCCL;                                                    \
    }
    LOAD(GetUniqueId, "ncclGetUniqueId");
    LOAD(CommInitRank, "ncclCommInitRank");
    LOAD(CommDestroy, "ncclCommDestroy");
    LOAD(AllReduce, "ncclAllReduce");
    LOAD(Broadcast, "ncclBroadcast");
    LOAD(AllGather, "ncclAllGather");
    LOAD(ReduceScatter, "ncclReduceScatter");
    LOAD(Send, "ncclSend");
    LOAD(Recv, "ncclRecv");
    LOAD(GroupStart, "ncclGroupStart");
    LOAD(GroupEnd, "ncclGroupEnd");
    LOAD(GetErrorString, "ncclGetErrorString");
#undef LOAD
    g_nccl.ok = true;
    return DESC_B200_OK;
}
}  // namespace

#define NCCL_TRY(expr)                                                                        \
    do {                                                                                      \
        ncclResult_t _r = (expr);                                                             \
        if (_r != ncclSuccess) {                                                              \
            desc_set_error("NCCL error at %s:%d: %s", __FILE__, __LINE__, g_nccl.GetErrorString(_r)); \
            return DESC_B200_ERR_NCCL;                                                        \
        }                                                                                     \
    } while (0)

extern "C" int desc_b200_nccl_unique_id(void* out128) {
    if (!out128) {
        desc_set_error("null output");
        return DESC_B200_ERR_ARG;
    }
    DESC_TRY(nccl_load());
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    NCCL_TRY(g_nccl.GetUniqueId(&id));
    memcpy(out128, &id, sizeof(id));
    return DESC_B200_OK;
}

// Communicators are cached per (unique id, rank, world, device): a ncclUniqueId can initialise one
// communicator only, while a caller typically solves many graphs (many handles) with the id it
// broadcast once.  Handles borrow the cached communicator; desc_b200_comm_finalize() frees them.
namespace {
std::map<std::string, ncclComm_t> g_comms;
std::mutex g_comms_mu;
}  // namespace

int desc_comm_init(desc_b200_handle* h, const void* nccl_id) {
    if (h->world <= 1) return DESC_B200_OK;
    if (!nccl_id) {
        desc_set_error("world=%d but no nccl_id in opts", h->world);
        return DESC_B200_ERR_ARG;
    }
    DESC_TRY(nccl_load());
    std::string key((const char*)nccl_id, 128);
    key += "/" + std::to_string(h->rank) + "/" + std::to_string(h->world) + "/" + std::to_string(h->device);
    std::lock_guard<std::mutex> lock(g_comms_mu);
    auto it = g_comms.find(key);
    if (it != g_comms.end()) {
        h->comm = it->second;
        return DESC_B200_OK;
    }
    ncclUniqueId id;
    memcpy(&id, nccl_id, sizeof(id));
    ncclComm_t comm;
    NCCL_TRY(g_nccl.CommInitRank(&comm, h->world, id, h->rank));
    g_comms[key] = comm;
    h->comm = comm;
    return DESC_B200_OK;
}

void desc_comm_destroy(desc_b200_handle* h) { h->comm = nullptr; }

extern "C" int desc_b200_comm_finalize(void) {
    std::lock_guard<std::mutex> lock(g_comms_mu);
    if (g_nccl.ok)
        for (auto& kv : g_comms) g_nccl.CommDestroy(kv.second);
    g_comms.clear();
    return DESC_B200_OK;
}

int desc_allreduce_sum(desc_b200_handle* h, double* buf, int64_t count) {
    if (h->world <= 1 || count <= 0) return DESC_B200_OK;
    NCCL_TRY(g_nccl.AllReduce(buf, buf, (size_t)count, ncclFloat64, ncclSum, (ncclComm_t)h->comm, h->stream));
    h->collectives++;
    return DESC_B200_OK;
}

// ---- the two per-iteration exchanges as NCCL collectives on padded, rank-major staging -------------------------
// The shards are ragged (vertex-aligned edge ranges), NCCL's all-gather / reduce-scatter want equal counts: pack the
// ranges into slots of the largest range's size, run ONE collective (NVSwitch: ring / NVLS inside NCCL, all links
// busy), unpack.  Measured against the grouped point-to-point version below (which NCCL serves with few channels
// per peer): profiles/README.md, round 2.  DESC_B200_COMM=p2p selects the point-to-point version.
static bool use_collectives(size_t staging_bytes) {
    static int mode = -1;
    if (mode < 0) {
        const char* e = getenv("DESC_B200_COMM");
        mode = (e && strcmp(e, "p2p") == 0) ? 0 : 1;
    }
    return mode == 1 && staging_bytes <= ((size_t)1 << 30);
}
static int comm_scratch_need(desc_b200_handle* h, size_t need) {
    if (h->comm_scratch_bytes < need) {
        if (h->comm_scratch) cudaFree(h->comm_scratch);
        h->comm_scratch = nullptr;
        h->comm_scratch_bytes = 0;
        CUDA_TRY(cudaMalloc(&h->comm_scratch, need));
        h->comm_scratch_bytes = need;
    }
    return DESC_B200_OK;
}
struct RangeTab {
    int64_t b[65];
    int world;
};
// words of 4 bytes: dst slot r (stride words) <- src range r, or back (pack = 1: ranges -> slots)
__global__ void k_pack_ranges(uint32_t* __restrict__ slots, uint32_t* __restrict__ flat, RangeTab t, int64_t stride,
                              int words_per_elem, int pack, int only, int skip) {
    const int r = blockIdx.y;
    if ((only >= 0 && r != only) || r == skip) return;
    const int64_t n = (t.b[r + 1] - t.b[r]) * words_per_elem;
    uint32_t* f = flat + t.b[r] * words_per_elem;
    uint32_t* s = slots + r * stride;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        if (pack)
            s[i] = f[i];
        else
            f[i] = s[i];
    }
}

static int allgather_collective(desc_b200_handle* h, void* buf, size_t elem_bytes, const std::vector<int64_t>& bounds) {
    const int W = h->world, me = h->rank;
    int64_t maxr = 0;
    for (int r = 0; r < W; r++) maxr = std::max<int64_t>(maxr, bounds[r + 1] - bounds[r]);
    if (maxr == 0) return DESC_B200_OK;
    const int wpe = (int)(elem_bytes / 4);
    const int64_t stride = maxr * wpe;                           // words per slot
    DESC_TRY(comm_scratch_need(h, (size_t)W * stride * 4));
    uint32_t* st = (uint32_t*)h->comm_scratch;
    RangeTab t;
    t.world = W;
    for (int r = 0; r <= W; r++) t.b[r] = bounds[r];
    dim3 g(DESC_SMS * 2, W);
    k_pack_ranges<<<g, 256, 0, h->stream>>>(st, (uint32_t*)buf, t, stride, wpe, 1, me, -1);
    KERNEL_CHECK(h);
    NCCL_TRY(g_nccl.AllGather(st + (int64_t)me * stride, st, (size_t)stride, ncclUint32, (ncclComm_t)h->comm, h->stream));
    k_pack_ranges<<<g, 256, 0, h->stream>>>(st, (uint32_t*)buf, t, stride, wpe, 0, -1, me);
    KERNEL_CHECK(h);
    h->collectives++;
    return DESC_B200_OK;
}

// in-place ragged all-gather: rank r owns elements [bounds[r], bounds[r+1]) of buf.  Point-to-point
// sends/receives in one group (all pairs move concurrently over NVSwitch) instead of `world`
// broadcasts, which NCCL runs one after the other.
int desc_allgather_ranges(desc_b200_handle* h, void* buf, size_t elem_bytes,
                          const std::vector<int64_t>& bounds) {
    if (h->world <= 1) return DESC_B200_OK;
    {
        int64_t maxr = 0;
        for (int r = 0; r < h->world; r++) maxr = std::max<int64_t>(maxr, bounds[r + 1] - bounds[r]);
        if (elem_bytes % 4 == 0 && h->world <= 64 && use_collectives((size_t)h->world * maxr * elem_bytes))
            return allgather_collective(h, buf, elem_bytes, bounds);
    }
    const int me = h->rank;
    char* mine = (char*)buf + (size_t)bounds[me] * elem_bytes;
    const size_t my_bytes = (size_t)(bounds[me + 1] - bounds[me]) * elem_bytes;
    NCCL_TRY(g_nccl.GroupStart());
    for (int d = 1; d < h->world; d++) {
        const int to = (me + d) % h->world, from = (me - d + h->world) % h->world;
        const size_t from_bytes = (size_t)(bounds[from + 1] - bounds[from]) * elem_bytes;
        if (my_bytes > 0) NCCL_TRY(g_nccl.Send(mine, my_bytes, ncclInt8, to, (ncclComm_t)h->comm, h->stream));
        if (from_bytes > 0)
            NCCL_TRY(g_nccl.Recv((char*)buf + (size_t)bounds[from] * elem_bytes, from_bytes, ncclInt8, from,
                                 (ncclComm_t)h->comm, h->stream));
    }
    NCCL_TRY(g_nccl.GroupEnd());
    h->collectives++;
    return DESC_B200_OK;
}

// sum over ranks of buf[bounds[r]*width .. bounds[r+1]*width) delivered to rank r only (a ragged
// reduce-scatter), plus an all-reduce of the `tail` doubles at buf[total*width ..].  Every rank
// sends its partial of range r to rank r and adds the received partials IN RANK ORDER: half the
// bytes of an all-reduce, and a summation order that does not depend on NCCL's algorithm choice.
__global__ void k_sum_partials(double* __restrict__ own, const double* __restrict__ scratch, int64_t count,
                               int64_t stride, int world, int me) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= count) return;
    double acc = 0.0;
    int slot = 0;
    for (int r = 0; r < world; r++) {
        if (r == me)
            acc += own[i];
        else
            acc += scratch[(int64_t)(slot++) * stride + i];
    }
    own[i] = acc;
}
// the same for 64-bit fixed-point partials (pgd_ell.cuh): integer sums, any order gives the same bits
__global__ void k_sum_partials_u64(unsigned long long* __restrict__ own, const unsigned long long* __restrict__ scratch,
                                   int64_t count, int64_t stride, int world) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= count) return;
    unsigned long long acc = own[i];
    for (int r = 0; r < world - 1; r++) acc += scratch[(int64_t)r * stride + i];
    own[i] = acc;
}

// slots for the reduce-scatter: slot r = [range r of buf (width doubles per element), zero padding, the `tail` doubles]
// -- every rank puts its tail partial into EVERY slot, so each rank receives the tail's sum with its own range
__global__ void k_pack_reduce(double* __restrict__ slots, const double* __restrict__ buf, RangeTab t, int64_t stride,
                              int width, int tail) {
    const int r = blockIdx.y;
    const int64_t n = (t.b[r + 1] - t.b[r]) * width;
    const double* f = buf + t.b[r] * width;
    double* s = slots + r * stride;
    const int64_t body = stride - tail;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < stride; i += (int64_t)gridDim.x * blockDim.x)
        s[i] = i < n ? f[i] : (i >= body ? buf[t.b[t.world] * width + (i - body)] : 0.0);
}
__global__ void k_unpack_reduce(const double* __restrict__ recv, double* __restrict__ buf, int64_t begin, int64_t n,
                                int64_t body, int64_t tail_at, int tail) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n + tail; i += (int64_t)gridDim.x * blockDim.x) {
        if (i < n)
            buf[begin + i] = recv[i];
        else
            buf[tail_at + (i - n)] = recv[body + (i - n)];
    }
}

int desc_reduce_to_owners(desc_b200_handle* h, double* buf, int width, const std::vector<int64_t>& bounds,
                          int tail, bool body_u64) {
    if (h->world <= 1) return DESC_B200_OK;
    const int me = h->rank, W = h->world;
    if (!body_u64 && W <= 64) {
        int64_t mx = 0;
        for (int r = 0; r < W; r++) mx = std::max<int64_t>(mx, bounds[r + 1] - bounds[r]);
        const int64_t stride = mx * width + tail;
        if (use_collectives((size_t)(W + 1) * stride * sizeof(double))) {
            DESC_TRY(comm_scratch_need(h, (size_t)(W + 1) * stride * sizeof(double)));
            double* send = (double*)h->comm_scratch;
            double* recv = send + (int64_t)W * stride;
            RangeTab t;
            t.world = W;
            for (int r = 0; r <= W; r++) t.b[r] = bounds[r];
            dim3 g(DESC_SMS * 2, W);
            k_pack_reduce<<<g, 256, 0, h->stream>>>(send, buf, t, stride, width, tail);
            KERNEL_CHECK(h);
            NCCL_TRY(g_nccl.ReduceScatter(send, recv, (size_t)stride, ncclFloat64, ncclSum, (ncclComm_t)h->comm, h->stream));
            const int64_t n = (bounds[me + 1] - bounds[me]) * width;
            k_unpack_reduce<<<DESC_SMS * 2, 256, 0, h->stream>>>(recv, buf, bounds[me] * width, n, stride - tail,
                                                                bounds[W] * width, tail);
            KERNEL_CHECK(h);
            h->collectives++;
            return DESC_B200_OK;
        }
    }
    const int64_t total = bounds[W];
    int64_t maxr = 0;
    for (int r = 0; r < W; r++) maxr = std::max<int64_t>(maxr, bounds[r + 1] - bounds[r]);
    const int64_t stride = maxr * width + tail;            // doubles per received partial
    const size_t need = (size_t)(W - 1) * stride * sizeof(double);
    if (h->comm_scratch_bytes < need) {
        if (h->comm_scratch) cudaFree(h->comm_scratch);
        h->comm_scratch = nullptr;
        CUDA_TRY(cudaMalloc(&h->comm_scratch, need));
        h->comm_scratch_bytes = need;
    }
    double* scratch = (double*)h->comm_scratch;
    const int64_t my_cnt = (bounds[me + 1] - bounds[me]) * width;
    NCCL_TRY(g_nccl.GroupStart());
    int slot_of[64];
    {
        int sl = 0;
        for (int r = 0; r < W; r++) slot_of[r] = (r == me) ? -1 : sl++;
    }
    for (int d = 1; d < W; d++) {
        const int to = (me + d) % W, from = (me - d + W) % W;
        const int64_t to_cnt = (bounds[to + 1] - bounds[to]) * width;
        if (to_cnt > 0)
            NCCL_TRY(g_nccl.Send(buf + bounds[to] * width, (size_t)to_cnt, ncclFloat64, to, (ncclComm_t)h->comm, h->stream));
        if (my_cnt > 0)
            NCCL_TRY(g_nccl.Recv(scratch + (int64_t)slot_of[from] * stride, (size_t)my_cnt, ncclFloat64, from,
                                 (ncclComm_t)h->comm, h->stream));
        if (tail > 0) {
            NCCL_TRY(g_nccl.Send(buf + total * width, (size_t)tail, ncclFloat64, to, (ncclComm_t)h->comm, h->stream));
            NCCL_TRY(g_nccl.Recv(scratch + (int64_t)slot_of[from] * stride + maxr * width, (size_t)tail, ncclFloat64, from,
                                 (ncclComm_t)h->comm, h->stream));
        }
    }
    NCCL_TRY(g_nccl.GroupEnd());
    if (my_cnt > 0) {
        if (body_u64)
            k_sum_partials_u64<<<(unsigned)((my_cnt + 255) / 256), 256, 0, h->stream>>>(
                (unsigned long long*)(buf + bounds[me] * width), (const unsigned long long*)scratch, my_cnt, stride, W);
        else
            k_sum_partials<<<(unsigned)((my_cnt + 255) / 256), 256, 0, h->stream>>>(buf + bounds[me] * width, scratch, my_cnt,
                                                                                  stride, W, me);
        KERNEL_CHECK(h);
    }
    if (tail > 0) {
        k_sum_partials<<<1, 32, 0, h->stream>>>(buf + total * width, scratch + maxr * width, tail, stride, W, me);
        KERNEL_CHECK(h);
    }
    h->collectives++;
    return DESC_B200_OK;
}
