// Streamed vertex-blocked PGD iteration (included by pgd.cu after PgdArgs / BlkArgs).
//
// Same arithmetic and the same ownership rules as k_pgd_block (one CTA per vertex block, S and
// partner-sum tables in shared memory, no atomics), re-organised around the Blackwell async-copy
// machinery so that the HBM stream never waits for the arithmetic:
//
//   * one PRODUCER warp: one lane issues cp.async.bulk (TMA, 1-D) copies of the next tile's slot
//     arrays (w, S0, pk_jk, rk_i: contiguous because the edges of a vertex block and their slot
//     lists are contiguous) into a ring of shared-memory stages; completion is signalled on an
//     mbarrier (complete_tx::bytes).  L2 policy evict_first: the stream must not evict S.
//   * COMPUTE warps (G lanes per edge, as before) read the stage, gather S[e_jk] (the one
//     remaining L2 gather per slot), compute gradient / step / projection, write w_t to global
//     and, in place, into the stage, publish (slot offset, count) of their edges, and arrive on
//     the stage's `done` mbarrier.
//   * SCATTER warps own the private partner-sum tables: they re-read (w_t, rank) of a finished
//     stage with one lane per slot (distinct ranks within an edge => conflict-free read-modify-
//     write), so the serialised table update is off the compute warps' critical path and only
//     NSW tables are needed instead of one per warp (more CTAs per SM).  They release the stage to
//     the producer through the `empty` mbarrier.
//
// Bulk copies need 16-byte aligned addresses and sizes: a tile's slot range [sA, sB) is widened to
// [sA & ~7, (sB + 7) & ~7) (8 slots = 16 B of the 2-byte rank array); the slot arrays are allocated
// with 16 spare elements (build.cu) so the widened read stays inside the allocation.

#define ST_NCW 8                       // compute warps
#define ST_NSW 2                       // scatter warps
#define ST_THREADS ((ST_NCW + ST_NSW + 1) * 32)
#define ST_MAXSTAGES 4

__device__ __forceinline__ uint32_t st_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(st_smem(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(st_smem(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(st_smem(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, %2;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(st_smem(bar)),
        "r"(parity), "r"(20000u)   // suspend-time hint (ns): sleep in hardware instead of spinning
        : "memory");
}
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
        "selp.u32 %0, 1, 0, P1;\n"
        "}"
        : "=r"(ok)
        : "r"(st_smem(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void bulk_g2s_plain(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     st_smem(dst)),
                 "l"(src), "r"(bytes), "r"(st_smem(bar))
                 : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t pol) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
            st_smem(dst)),
        "l"(src), "r"(bytes), "r"(st_smem(bar)), "l"(pol)
        : "memory");
}

struct StreamArgs {
    BlkArgs b;
    int nstages;
    int tsc;          // slot capacity of one stage (multiple of 8)
    int max_ns;       // longest slot list (size of the reciprocal table)
};

// shared-memory carve-up: barriers | per-stage edge headers | 1/c table | T_S | T_acc[NSW] | stages.
// One stage: w[tsc] f64 | S0[tsc] f64 | pk_jk[tsc] u32 | rk_i[tsc] u16 | rowptr[TE+2] i64 |
//            acc_cur[2*TE] f64 | S_cur[TE+2] f64          (every offset a multiple of 16 bytes)
__host__ __device__ __forceinline__ size_t st_stage_bytes(int tsc, int te) {
    return (size_t)tsc * (8 + 8 + 4 + 2) + (size_t)(te + 2) * 8 + (size_t)te * 16 + (size_t)(te + 2) * 8;
}
__host__ __device__ __forceinline__ size_t st_rcp_bytes(int max_ns) { return ((size_t)(max_ns + 1) * 8 + 15) & ~(size_t)15; }
__host__ __device__ __forceinline__ size_t st_fixed_bytes(int te, int tstride, int max_ns) {
    return 128 + (size_t)ST_MAXSTAGES * te * sizeof(int2) + st_rcp_bytes(max_ns) +
           (size_t)(1 + ST_NSW) * tstride * sizeof(double);
}

template <int G, int EPL, int RULE>
__global__ void __launch_bounds__(ST_THREADS)
k_pgd_stream(StreamArgs sa) {
    const BlkArgs& a = sa.b;
    if (a.p.ctrl[0]) return;
    constexpr int TE = ST_NCW * 32 / G;   // edges per tile
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);
    uint64_t* full = bars;
    uint64_t* done = bars + ST_MAXSTAGES;
    uint64_t* empty = bars + 2 * ST_MAXSTAGES;
    int2* hdr = reinterpret_cast<int2*>(smem_raw + 128);
    double* rcp = reinterpret_cast<double*>(smem_raw + 128 + (size_t)ST_MAXSTAGES * TE * sizeof(int2));
    double* T_S = reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(rcp) + st_rcp_bytes(sa.max_ns));
    double* T_acc = T_S + a.tstride;
    unsigned char* stage0 = reinterpret_cast<unsigned char*>(T_acc + (size_t)ST_NSW * a.tstride);
    const int tsc = sa.tsc;
    const size_t stage_bytes = st_stage_bytes(tsc, TE);
    const size_t off_d = (size_t)tsc * 8, off_pk = (size_t)tsc * 16, off_rk = (size_t)tsc * 20;
    const size_t off_rp = (size_t)tsc * 22, off_acc = off_rp + (size_t)(TE + 2) * 8, off_so = off_acc + (size_t)TE * 16;
    const int NST = sa.nstages;

    const int v = a.v0 + blockIdx.x;
    const int rs = a.rowstart[v];
    const int deg = a.rowstart[v + 1] - rs;
    const int e_lo = a.estart[v], e_hi = a.estart[v + 1];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (e_hi <= e_lo) {   // no own edges: nothing to update, nothing to scatter
        if (threadIdx.x == 0) {
            a.partial[2 * blockIdx.x] = 0.0;
            a.partial[2 * blockIdx.x + 1] = 0.0;
        }
        return;
    }
    const int ntiles = (e_hi - e_lo + TE - 1) / TE;
    const int eoff = e_lo & 1;   // the per-edge arrays are copied from an even edge index (16-byte alignment)
    if (threadIdx.x == 0) {
        for (int s = 0; s < NST; s++) {
            mbar_init(&full[s], 1);
            mbar_init(&done[s], ST_NCW);
            mbar_init(&empty[s], ST_NSW);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    double objp = 0.0, chgp = 0.0;
    if (warp == ST_NCW + ST_NSW) {
        // ------------------------------------------------------------------ producer
        // lanes prefetch the slot boundaries of 32 tiles at a time; lane 0 issues the bulk copies
        const uint64_t pol = l2_evict_first_policy();
        int s = 0;
        uint32_t ph = 0;
        for (int tb = 0; tb < ntiles; tb += 32) {
            const int tl = min(tb + lane, ntiles - 1);
            const int le0 = e_lo + tl * TE, le1 = min(le0 + TE, e_hi);
            const int64_t lA = a.p.rowptr[le0] - a.p.slot_base;
            const int64_t lB = a.p.rowptr[le1] - a.p.slot_base;
            const int nu = min(32, ntiles - tb);
            for (int u = 0; u < nu; u++) {
                const int64_t sA = __shfl_sync(0xffffffffu, lA, u);
                const int64_t sB = __shfl_sync(0xffffffffu, lB, u);
                if (lane == 0) {
                    const int t = tb + u;
                    const int e0 = e_lo + t * TE, e1 = min(e0 + TE, e_hi);
                    const int e0a = e0 - eoff;
                    const int rp_cnt = (e1 + 1 - e0a + 1) & ~1;
                    const int so_cnt = (e1 - e0a + 1) & ~1;
                    const int64_t base = sA & ~(int64_t)7;
                    const int64_t cnt = ((sB + 7) & ~(int64_t)7) - base;
                    unsigned char* st = stage0 + (size_t)s * stage_bytes;
                    mbar_wait(&empty[s], ph ^ 1u);
                    mbar_expect_tx(&full[s], (uint32_t)(cnt * 22 + rp_cnt * 8 + (e1 - e0) * 16 + so_cnt * 8));
                    bulk_g2s_plain(st + off_rp, a.p.rowptr + e0a, (uint32_t)(rp_cnt * 8), &full[s]);
                    bulk_g2s_plain(st + off_acc, a.p.acc_cur + 2 * (int64_t)e0, (uint32_t)((e1 - e0) * 16), &full[s]);
                    bulk_g2s_plain(st + off_so, a.p.S_cur + e0a, (uint32_t)(so_cnt * 8), &full[s]);
                    if (cnt > 0) {
                        bulk_g2s(st, a.p.w_cur + base, (uint32_t)(cnt * 8), &full[s], pol);
                        bulk_g2s(st + off_d, a.p.S0 + base, (uint32_t)(cnt * 8), &full[s], pol);
                        bulk_g2s(st + off_pk, a.p.pk_jk + base, (uint32_t)(cnt * 4), &full[s], pol);
                        bulk_g2s(st + off_rk, a.rk_i + base, (uint32_t)(cnt * 2), &full[s], pol);
                    }
                }
                if (++s == NST) {
                    s = 0;
                    ph ^= 1u;
                }
                __syncwarp();
            }
        }
    } else {
        // ------------------------------------------------------------------ tables (compute + scatter warps)
        constexpr int NT = (ST_NCW + ST_NSW) * 32;
        for (int c = threadIdx.x; c <= sa.max_ns; c += NT) rcp[c] = c > 0 ? 1.0 / (double)c : 0.0;
        for (int r = threadIdx.x; r < deg; r += NT) {
            T_S[r] = a.p.S_cur[a.adj_eid[rs + r]];
#pragma unroll
            for (int q = 0; q < ST_NSW; q++) T_acc[q * a.tstride + r] = 0.0;
        }
        asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory");
        if (warp < ST_NCW) {
            // -------------------------------------------------------------- compute
            // Branch-free inner loop: a lane's EPL slots are idx = r + x*G; lanes past the end of the
            // slot list (ok[x] false) read harmless stage bytes and are neutralised with selects.
            const int r = threadIdx.x & (G - 1);
            const int q = threadIdx.x / G;           // edge within the tile
            const double nlr = -a.p.lr;
            const double* __restrict__ Sg = a.p.S_cur;
            int s = 0;
            uint32_t ph = 0;
            unsigned char* st = stage0;
            int e = e_lo + q;
            for (int t = 0; t < ntiles; t++, e += TE) {
                const int64_t* rp = reinterpret_cast<const int64_t*>(st + off_rp) + eoff;
                mbar_wait(&full[s], ph);
                int sl = 0, ns = 0;
                int64_t s0 = 0;
                double A = 0.0, B = 0.0, Sold = 0.0;
                if (e < e_hi) {
                    const int64_t rq = rp[q];
                    s0 = rq - a.p.slot_base;
                    ns = (int)(rp[q + 1] - rq);
                    sl = (int)(s0 - ((rp[0] - a.p.slot_base) & ~(int64_t)7));
                    const double2 ab = reinterpret_cast<const double2*>(st + off_acc)[q];
                    A = ab.x;
                    B = ab.y;
                    Sold = (reinterpret_cast<const double*>(st + off_so) + eoff)[q];
                }
                double* pw = reinterpret_cast<double*>(st) + sl + r;
                const double* pd = reinterpret_cast<const double*>(st + off_d) + sl + r;
                const uint32_t* ppk = reinterpret_cast<const uint32_t*>(st + off_pk) + sl + r;
                const uint16_t* prk = reinterpret_cast<const uint16_t*>(st + off_rk) + sl + r;
                double w[EPL], d[EPL], sj[EPL], ts[EPL];
                bool ok[EPL], fa[EPL], fb[EPL];
#pragma unroll
                for (int x = 0; x < EPL; x++) {
                    ok[x] = r + x * G < ns;
                    const uint32_t pj = ppk[x * G];
                    const uint32_t rk = prk[x * G];
                    fa[x] = (rk & RK_APP) != 0u;
                    fb[x] = (pj & PK_APP) != 0u;
                    sj[x] = Sg[ok[x] ? (pj & PK_MASK) : 0u];
                    ts[x] = T_S[ok[x] ? (rk & RK_MASK) : 0u];
                    w[x] = ok[x] ? pw[x * G] : 0.0;
                    d[x] = ok[x] ? pd[x * G] : 0.0;
                }
                double g[EPL];
                double gsum = 0.0, wsum = 0.0;
#pragma unroll
                for (int x = 0; x < EPL; x++) {
                    const double sg = sj[x] + ts[x];
                    objp = fma(w[x], sg, objp);
                    const double part = (fa[x] ? A : 0.0) + (fb[x] ? B : 0.0);
                    g[x] = ok[x] ? fma(part, d[x], sg) : 0.0;
                    gsum += g[x];
                    wsum += w[x];
                }
                if (RULE == 0) {
                    // constant / piecewise step: sum(w + step) = sum(w) - lr (sum(g) - ns mean(g)); one
                    // reduction round trip serves both sums
#pragma unroll
                    for (int o = G / 2; o > 0; o >>= 1) {
                        gsum += __shfl_xor_sync(0xffffffffu, gsum, o);
                        wsum += __shfl_xor_sync(0xffffffffu, wsum, o);
                    }
                } else {
                    gsum = group_sum<G>(gsum);
                }
                const double fns = (double)ns;
                const double rns = rcp[ns];          // 1/ns (0 when the edge has no slots)
                const double gmean = gsum * rns;
                if (RULE == 0) {
                    wsum = fma(nlr, gsum - fns * gmean, wsum);
#pragma unroll
                    for (int x = 0; x < EPL; x++) w[x] = ok[x] ? fma(nlr, g[x] - gmean, w[x]) : -1e300;
                } else {
                    wsum = 0.0;
#pragma unroll
                    for (int x = 0; x < EPL; x++) {
                        if (ok[x]) {
                            const double gr = g[x] - gmean;
                            const int64_t sg = s0 + r + x * G;
                            const double mt = a.p.beta1 * a.p.adam_m[sg] + (1.0 - a.p.beta1) * gr;
                            const double vt = a.p.beta2 * a.p.adam_v[sg] + (1.0 - a.p.beta2) * (gr * gr);
                            a.p.adam_m[sg] = mt;
                            a.p.adam_v[sg] = vt;
                            w[x] = w[x] + -a.p.lr * (mt / a.p.corr1) / (sqrt(vt / a.p.corr2) + 1e-8);
                            wsum += w[x];
                        } else {
                            w[x] = -1e300;
                        }
                    }
                    wsum = group_sum<G>(wsum);
                }
                // Michelot: T <- (sum_{w>T} w - 1)/#{w>T} until the active set stops shrinking
                int cnt = ns;
                double T = (wsum - 1.0) * rns;
                for (int mit = 0; mit < G * EPL + 2; mit++) {
                    double s2 = 0.0;
                    int c2 = 0;
#pragma unroll
                    for (int x = 0; x < EPL; x++) {
                        const bool in = w[x] > T;
                        s2 += in ? w[x] : 0.0;
                        c2 += in ? 1 : 0;
                    }
#pragma unroll
                    for (int o = G / 2; o > 0; o >>= 1) {
                        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
                        c2 += __shfl_xor_sync(0xffffffffu, c2, o);
                    }
                    const bool changed = (c2 != cnt) && (c2 > 0);
                    if (changed) {
                        T = (s2 - 1.0) * rcp[c2];
                        cnt = c2;
                    }
                    if (!__any_sync(0xffffffffu, changed)) break;
                }
                double snew = 0.0;
                double* __restrict__ pout = a.p.w_next + s0 + r;
#pragma unroll
                for (int x = 0; x < EPL; x++) {
                    const double wo = fmax(w[x] - T, 0.0);
                    snew = fma(wo, d[x], snew);
                    if (ok[x]) {
                        __stcs(pout + x * G, wo);
                        pw[x * G] = wo;
                    }
                }
                snew = group_sum<G>(snew);
                if (r == 0) {
                    hdr[s * TE + q] = make_int2(sl, ns);
                    if (ns > 0) {
                        a.p.S_next[e] = snew;
                        chgp += fabs(snew - Sold);
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&done[s]);
                st += stage_bytes;
                if (++s == NST) {
                    s = 0;
                    ph ^= 1u;
                    st = stage0;
                }
            }
        } else {
            // -------------------------------------------------------------- scatter
            const int swi = warp - ST_NCW;
            double* Tw = T_acc + (size_t)swi * a.tstride;
            int s = 0;
            uint32_t ph = 0;
            unsigned char* st = stage0;
            for (int t = 0; t < ntiles; t++) {
                const double* sw = reinterpret_cast<const double*>(st);
                const uint16_t* srk = reinterpret_cast<const uint16_t*>(st + off_rk);
                mbar_wait(&done[s], ph);
                constexpr int U = 4;
                for (int qb = swi * U; qb < TE; qb += ST_NSW * U) {
                    int2 h2[U];
                    double wv[U];
                    uint32_t rk[U];
#pragma unroll
                    for (int u = 0; u < U; u++) {
                        h2[u] = (qb + u < TE) ? hdr[s * TE + qb + u] : make_int2(0, 0);
                        wv[u] = 0.0;
                        rk[u] = 0u;
                        if (lane < h2[u].y) {
                            wv[u] = sw[h2[u].x + lane];
                            rk[u] = srk[h2[u].x + lane];
                        }
                    }
#pragma unroll
                    for (int u = 0; u < U; u++) {
                        if (rk[u] & RK_APP) Tw[rk[u] & RK_MASK] += wv[u];
                        __syncwarp();
                        for (int i2 = lane + 32; i2 < h2[u].y; i2 += 32) {   // slot lists longer than a warp
                            const uint32_t rr = srk[h2[u].x + i2];           // (ranks of one edge are distinct)
                            if (rr & RK_APP) Tw[rr & RK_MASK] += sw[h2[u].x + i2];
                        }
                        __syncwarp();
                    }
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[s]);
                st += stage_bytes;
                if (++s == NST) {
                    s = 0;
                    ph ^= 1u;
                    st = stage0;
                }
            }
        }
    }
    __syncthreads();
    // flush the private tables: every (edge, side) entry is owned by this CTA within this kernel
    for (int r = threadIdx.x; r < deg; r += ST_THREADS) {
        double x = 0.0;
#pragma unroll
        for (int q = 0; q < ST_NSW; q++) x += T_acc[q * a.tstride + r];
        const int e2 = a.adj_eid[rs + r];
        const int k = a.adj_nbr[rs + r];
        a.p.acc_next[2 * (int64_t)e2 + (v < k ? 0 : 1)] += x;
    }
    objp = group_sum<32>(objp);
    chgp = group_sum<32>(chgp);
    __shared__ double red[2 * ST_NCW];
    if (lane == 0 && warp < ST_NCW) {
        red[2 * warp] = objp;
        red[2 * warp + 1] = chgp;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double o = 0.0, c = 0.0;
#pragma unroll
        for (int q = 0; q < ST_NCW; q++) {
            o += red[2 * q];
            c += red[2 * q + 1];
        }
        a.partial[2 * blockIdx.x] = o;
        a.partial[2 * blockIdx.x + 1] = c;
    }
}
