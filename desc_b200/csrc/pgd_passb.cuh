// Pass over LARGER endpoints (included by pgd.cu after BlkArgs): the second half of a PGD iteration.
//
// One CTA per vertex v walks the slot lists of the edges (u,v), u < v, whose smaller endpoint lies in
// the local vertex range.  For a slot (uv;k) the partner edge {v,k} is incident to v, so with the
// rank of k in v's adjacency row (rk_j, 14 bits + JKI_appears) both
//   * S_t[{v,k}]          -> sjk[slot]      (read by the streamed kernel of iteration t+1 instead of
//                                            a random gather: DESC.m:193, first term) and
//   * w_t[slot] (if JKI_appears) -> partner sum "via v" of edge {v,k}      (DESC.m:185-191)
// are shared-memory table operations.  Every warp owns a private partner-sum table (the ranks inside
// one edge are distinct, so one edge per instruction is a conflict-free read-modify-write) and keeps
// the (w, rank) loads of the next two batches of edges and the headers of the third in flight in
// registers while it works on the current batch.
// No atomics: every partner-sum entry (edge, side v) is owned by vertex v's CTA.
#ifndef PB_WARPS
#define PB_WARPS 8
#endif
#define PB_TB (PB_WARPS * 32)
#ifndef PB_U
#define PB_U 4
#endif
#ifndef PB_DEPTH
#define PB_DEPTH 2                      // batches of (w, rank) loads in flight ahead of the current one
#endif

template <bool WRITE_SJK>
__global__ void __launch_bounds__(PB_TB)
k_pgd_passb(BlkArgs a, const double* __restrict__ w, const int2* __restrict__ jhdr, double* __restrict__ sjk) {
    if (a.p.ctrl[0]) return;
    extern __shared__ double sh[];
    double* T_S = sh;
    double* T_acc = sh + a.tstride;
    const int v = (int)gridDim.x - 1 - (int)blockIdx.x;   // long in-edge lists first
    const int rs = a.rowstart[v];
    const int deg = a.rowstart[v + 1] - rs;
    const int ub = min(a.v1, v);                           // neighbours u with v0 <= u < ub
    if (ub <= a.v0) return;
    const int lo = rs + (a.v0 > 0 ? desc_rank(a.bm, a.bmprefix, a.nwords, v, a.v0) : 0);
    const int hi = rs + desc_rank(a.bm, a.bmprefix, a.nwords, v, ub);
    if (hi <= lo) return;
    for (int r = threadIdx.x; r < deg; r += PB_TB) {
        if (WRITE_SJK) T_S[r] = a.p.S_next[a.adj_eid[rs + r]];
#pragma unroll
        for (int q = 0; q < PB_WARPS; q++) T_acc[q * a.tstride + r] = 0.0;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* TA = T_acc + (size_t)warp * a.tstride;
    const uint16_t* __restrict__ RK = a.rk_j;
    const uint32_t dummy = (uint32_t)a.tstride - 1u;
    constexpr int U = PB_U;
    // Register pipeline, four batches deep.  A header is loaded by lanes < U and broadcast with
    // shuffles only one iteration LATER (a shuffle right after the load would stall on it):
    //   iteration b: issue header load of batch b+3 | broadcast header b+2, issue its (w, rank) loads |
    //                (w, rank) of b+1 in flight | work on batch b
    int2 h0[U], h1[U], h2[U], h3[U];
    double w0[U], w1[U], w2[U], w3[U];
    uint32_t r0[U], r1[U], r2[U], r3[U];
    auto load_raw = [&](int base) {
        int2 mine = make_int2(0, 0);
        if (lane < U && base + lane < hi) mine = __ldg(jhdr + base + lane);
        return mine;
    };
    auto bcast = [&](const int2 mine, int2 (&h_)[U]) {
#pragma unroll
        for (int u = 0; u < U; u++) {
            h_[u].x = __shfl_sync(0xffffffffu, mine.x, u);
            h_[u].y = __shfl_sync(0xffffffffu, mine.y, u);
        }
    };
    auto load_data = [&](const int2 (&h_)[U], double (&w_)[U], uint32_t (&r_)[U]) {
#pragma unroll
        for (int u = 0; u < U; u++) {
            w_[u] = 0.0;
            r_[u] = 0u;
            if (lane < h_[u].y) {
                w_[u] = __ldcs(w + h_[u].x + lane);
                r_[u] = __ldcs(RK + h_[u].x + lane);
            }
        }
    };
    const int stride = PB_WARPS * U;
    int base = lo + warp * U;
    {
        const int2 a0 = load_raw(base), a1 = load_raw(base + stride), a2 = load_raw(base + 2 * stride);
        bcast(a0, h0);
        bcast(a1, h1);
        if (PB_DEPTH == 3) bcast(a2, h2);
    }
    int2 raw = load_raw(base + PB_DEPTH * stride);
    load_data(h0, w0, r0);
    load_data(h1, w1, r1);
    if (PB_DEPTH == 3) load_data(h2, w2, r2);
    for (; base < hi; base += stride) {
        const int2 raw_next = load_raw(base + (PB_DEPTH + 1) * stride);
        if (PB_DEPTH == 3) {
            bcast(raw, h3);
            load_data(h3, w3, r3);
        } else {
            bcast(raw, h2);
            load_data(h2, w2, r2);
        }
        raw = raw_next;
        // branch-free: inactive lanes (past the list end / flag off) use the dummy table entry
        double ts[U];
#pragma unroll
        for (int u = 0; u < U; u++) ts[u] = T_S[lane < h0[u].y ? (r0[u] & RK_MASK) : dummy];
#pragma unroll
        for (int u = 0; u < U; u++) {
            if (WRITE_SJK) {
                if (lane < h0[u].y) __stcs(sjk + h0[u].x + lane, ts[u]);
            }
            const bool f = (r0[u] & RK_APP) != 0u;
            const uint32_t ia = f ? (r0[u] & RK_MASK) : dummy;
            const double t = TA[ia];
            TA[ia] = t + (f ? w0[u] : 0.0);
            __syncwarp();
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            if (h0[u].y > 32) {   // slot lists longer than a warp (ranks of one edge are distinct)
                for (int i2 = lane + 32; i2 < h0[u].y; i2 += 32) {
                    const uint32_t rr = RK[h0[u].x + i2];
                    if (WRITE_SJK) sjk[h0[u].x + i2] = T_S[rr & RK_MASK];
                    if (rr & RK_APP) TA[rr & RK_MASK] += w[h0[u].x + i2];
                }
                __syncwarp();
            }
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            h0[u] = h1[u];
            h1[u] = h2[u];
            w0[u] = w1[u];
            w1[u] = w2[u];
            r0[u] = r1[u];
            r1[u] = r2[u];
            if (PB_DEPTH == 3) {
                h2[u] = h3[u];
                w2[u] = w3[u];
                r2[u] = r3[u];
            }
        }
    }
    __syncthreads();
    for (int r = threadIdx.x; r < deg; r += PB_TB) {
        double x = 0.0;
#pragma unroll
        for (int q = 0; q < PB_WARPS; q++) x += T_acc[q * a.tstride + r];
        const int e2 = a.adj_eid[rs + r];
        const int k = a.adj_nbr[rs + r];
        a.p.acc_next[2 * (int64_t)e2 + (v < k ? 0 : 1)] += x;   // owned by this CTA within this kernel
    }
}


// ------------------------------------------------------------------------------------------
// TMA-fed variant (default): the (w, rank) chunks of the in-edges are scattered in memory (one
// ~240-byte piece per edge), so the direct-load version above is bound by memory latency times its
// register prefetch depth.  Here a producer warp issues two small cp.async.bulk copies per in-edge
// (lane = edge, 32 edges per tile) into a ring of shared-memory tiles; mbarriers hand tiles to the
// scatter warps (one private partner-sum table each) and back.  Bulk copies need 16-byte alignment:
// each edge's pieces are widened to 16-byte boundaries and land in fixed-size per-edge regions.
#define PT_NSW 4                        // scatter warps
#define PT_TB ((PT_NSW + 1) * 32)
#define PT_TE 32                        // in-edges per tile (one per producer lane)
#define PT_MAXSTAGES 6

struct PassbArgs {
    BlkArgs b;
    const double* w;
    const int2* jhdr;
    double* sjk;
    int wmax;      // doubles per edge region  (multiple of 2)
    int rmax;      // ranks per edge region    (multiple of 8)
    int nstages;
};
__host__ __device__ __forceinline__ size_t pt_stage_bytes(int wmax, int rmax) {
    return (size_t)PT_TE * ((size_t)wmax * 8 + (size_t)rmax * 2 + sizeof(int4));
}
__host__ __device__ __forceinline__ size_t pt_fixed_bytes(int tstride) {
    return 128 + (size_t)(1 + PT_NSW) * tstride * sizeof(double);
}

__global__ void __launch_bounds__(PT_TB)
k_pgd_passb_tma(PassbArgs pa) {
    const BlkArgs& a = pa.b;
    if (a.p.ctrl[0]) return;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw);
    uint64_t* empty = full + PT_MAXSTAGES;
    double* T_S = reinterpret_cast<double*>(smem_raw + 128);
    double* T_acc = T_S + a.tstride;
    unsigned char* stage0 = reinterpret_cast<unsigned char*>(T_acc + (size_t)PT_NSW * a.tstride);
    const int wmax = pa.wmax, rmax = pa.rmax, NST = pa.nstages;
    const size_t off_rk = (size_t)PT_TE * wmax * 8, off_hdr = off_rk + (size_t)PT_TE * rmax * 2;
    const size_t stage_bytes = pt_stage_bytes(wmax, rmax);

    const int v = (int)gridDim.x - 1 - (int)blockIdx.x;   // long in-edge lists first
    const int rs = a.rowstart[v];
    const int deg = a.rowstart[v + 1] - rs;
    const int ub = min(a.v1, v);                           // neighbours u with v0 <= u < ub
    if (ub <= a.v0) return;
    const int lo = rs + (a.v0 > 0 ? desc_rank(a.bm, a.bmprefix, a.nwords, v, a.v0) : 0);
    const int hi = rs + desc_rank(a.bm, a.bmprefix, a.nwords, v, ub);
    if (hi <= lo) return;
    const int ntiles = (hi - lo + PT_TE - 1) / PT_TE;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        for (int s = 0; s < NST; s++) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], PT_NSW);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (warp == PT_NSW) {
        // ------------------------------------------------------------------ producer (lane = in-edge)
        const uint64_t pol = l2_evict_first_policy();
        int s = 0;
        uint32_t ph = 0;
        int2 hnext = make_int2(0, 0);
        if (lo + lane < hi) hnext = __ldg(pa.jhdr + lo + lane);
        for (int t = 0; t < ntiles; t++) {
            const int2 h = hnext;
            const int pn = lo + (t + 1) * PT_TE + lane;
            hnext = make_int2(0, 0);
            if (pn < hi) hnext = __ldg(pa.jhdr + pn);      // next tile's headers fly during this tile
            unsigned char* st = stage0 + (size_t)s * stage_bytes;
            const int s0 = h.x, ns = h.y;
            const int wa = s0 & ~1, wcnt = ns > 0 ? ((s0 + ns + 1) & ~1) - wa : 0;
            const int ra = s0 & ~7, rcnt = ns > 0 ? ((s0 + ns + 7) & ~7) - ra : 0;
            uint32_t bytes = (uint32_t)(wcnt * 8 + rcnt * 2);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) bytes += __shfl_xor_sync(0xffffffffu, bytes, o);
            if (lane == 0) mbar_wait(&empty[s], ph ^ 1u);
            __syncwarp();
            reinterpret_cast<int4*>(st + off_hdr)[lane] = make_int4(s0, ns, s0 - wa, s0 - ra);
            __syncwarp();
            if (lane == 0) mbar_expect_tx(&full[s], bytes);
            __syncwarp();
            if (ns > 0) {
                bulk_g2s(st + (size_t)lane * wmax * 8, pa.w + wa, (uint32_t)(wcnt * 8), &full[s], pol);
                bulk_g2s(st + off_rk + (size_t)lane * rmax * 2, a.rk_j + ra, (uint32_t)(rcnt * 2), &full[s], pol);
            }
            if (++s == NST) {
                s = 0;
                ph ^= 1u;
            }
        }
    } else {
        // ------------------------------------------------------------------ scatter warps
        constexpr int NT = PT_NSW * 32;
        for (int r = threadIdx.x; r < deg; r += NT) {
            T_S[r] = a.p.S_next[a.adj_eid[rs + r]];
#pragma unroll
            for (int q = 0; q < PT_NSW; q++) T_acc[q * a.tstride + r] = 0.0;
        }
        asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory");
        double* TA = T_acc + (size_t)warp * a.tstride;
        const uint32_t dummy = (uint32_t)a.tstride - 1u;
        int s = 0;
        uint32_t ph = 0;
        for (int t = 0; t < ntiles; t++) {
            const unsigned char* st = stage0 + (size_t)s * stage_bytes;
            const int4* hdr = reinterpret_cast<const int4*>(st + off_hdr);
            mbar_wait(&full[s], ph);
            constexpr int U = 4;                                   // edges per batch: loads of a batch issue together
            for (int qb = warp * U; qb < PT_TE; qb += PT_NSW * U) {
                int4 h4[U];
                double wv[U], ts[U];
                uint32_t rk[U];
#pragma unroll
                for (int u = 0; u < U; u++) {
                    h4[u] = hdr[qb + u];
                    const double* sw = reinterpret_cast<const double*>(st) + (size_t)(qb + u) * wmax + h4[u].z;
                    const uint16_t* sr = reinterpret_cast<const uint16_t*>(st + off_rk) + (size_t)(qb + u) * rmax + h4[u].w;
                    const bool ok = lane < h4[u].y;
                    wv[u] = ok ? sw[lane] : 0.0;
                    rk[u] = ok ? (uint32_t)sr[lane] : 0u;
                }
#pragma unroll
                for (int u = 0; u < U; u++) ts[u] = T_S[lane < h4[u].y ? (rk[u] & RK_MASK) : dummy];
#pragma unroll
                for (int u = 0; u < U; u++) {
                    if (lane < h4[u].y) __stcs(pa.sjk + h4[u].x + lane, ts[u]);
                    const bool f = (rk[u] & RK_APP) != 0u;
                    const uint32_t ia = f ? (rk[u] & RK_MASK) : dummy;
                    const double tv = TA[ia];
                    TA[ia] = tv + (f ? wv[u] : 0.0);
                    __syncwarp();
                }
#pragma unroll
                for (int u = 0; u < U; u++) {
                    if (h4[u].y > 32) {   // slot lists longer than a warp (ranks of one edge are distinct)
                        const double* sw = reinterpret_cast<const double*>(st) + (size_t)(qb + u) * wmax + h4[u].z;
                        const uint16_t* sr = reinterpret_cast<const uint16_t*>(st + off_rk) + (size_t)(qb + u) * rmax + h4[u].w;
                        for (int i2 = lane + 32; i2 < h4[u].y; i2 += 32) {
                            const uint32_t rr = sr[i2];
                            pa.sjk[h4[u].x + i2] = T_S[rr & RK_MASK];
                            if (rr & RK_APP) TA[rr & RK_MASK] += sw[i2];
                        }
                        __syncwarp();
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);
            if (++s == NST) {
                s = 0;
                ph ^= 1u;
            }
        }
    }
    __syncthreads();
    for (int r = threadIdx.x; r < deg; r += PT_TB) {
        double x = 0.0;
#pragma unroll
        for (int q = 0; q < PT_NSW; q++) x += T_acc[q * a.tstride + r];
        const int e2 = a.adj_eid[rs + r];
        const int k = a.adj_nbr[rs + r];
        a.p.acc_next[2 * (int64_t)e2 + (v < k ? 0 : 1)] += x;   // owned by this CTA within this kernel
    }
}
