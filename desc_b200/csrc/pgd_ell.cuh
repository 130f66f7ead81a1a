// Lane-per-edge PGD iteration on a tile-interleaved ("ELL") copy of the slot arrays (included by pgd.cu after
// PgdArgs).  One kernel per iteration.  EXPERIMENTAL path (DESC_B200_PGD_PATH=ell), parity-tested but NOT the
// default: measured on B200 at cfg 4 it takes 4.8-5.0 ms per iteration against 2.2 ms of the two-pass streamed
// path (profiles/README.md, round 2): the 120 MB random working set (S + the 2m fixed-point accumulators) does not
// stay in L2 while 4.5 GB of slot data stream through it (L2 hit rate 38 %, DRAM traffic 9.8 GB), and both the
// gathers (1.4 ms) and the 9e7 REDG.64 per iteration (2.0 ms) are latency-bound at 16 warps/SM.
//
// Why another layout.  The CSR slot arrays keep an edge's slots contiguous, so a kernel that gives an edge to a
// group of lanes has to shuffle every per-edge quantity (gradient mean, projection threshold, new S) across the
// group and can keep only a handful of edges per warp in flight: the streamed two-pass kernels (pgd_stream.cuh,
// pgd_passb.cuh) are bound by that per-edge dependent chain at ~18 warps/SM, not by bytes (profiles/README.md).
// Here the slots of 32/G consecutive edges of one vertex block are interleaved: slot s of the tile's edge q sits
// at  tile_base + (s / G) * 32 + q * G + (s % G),  so lane (q*G + s%G) of a warp walks "its" edge with stride 32
// and every warp-level load / store is one fully coalesced 256-byte (f64), 128-byte (u32) or 64-byte (u16)
// access.  For the common slot budgets (n_sample <= 32) G = 1: a lane owns a whole edge, the tangent projection,
// the step and Michelot's simplex projection (DESC.m:195-224) are straight-line per-lane code without a single
// shuffle, and a warp has 32 edges x 30 independent gathers in flight.
//
// Per slot (ij;k) of vertex block i:
//   S[e_ik]  -> shared-memory table of the block (rank of k in i's adjacency row, streamed as 16 bits)
//   S[e_jk]  -> 8-byte gather from the L2-resident S vector (40 MB at cfg 4)
//   partner sums (DESC.m:185-191, scatter form): w_t is added to the accumulators of {i,k} via i and {j,k} via j
//            with 64-bit FIXED-POINT reductions (red.global.add.u64, scale 2^50): integer addition is
//            associative, so the sums are bit-reproducible for ANY execution order and for any number of GPUs,
//            without private tables, flush passes or a second kernel.  w in [0,1], at most 2^13 addends per
//            accumulator: 63 bits suffice; the quantisation (2^-51 per addend) is below the rounding noise of
//            the FP64 sums it replaces and far inside the 1e-10 parity bar.
// Streamed bytes per slot: w in 8 + w out 8 + S0 8 (+8 when it is re-read from L2) + rank 2 + e_jk 4 = 30 B,
// against the 52 B of the two-pass path and the 40 B of SURVEY 8d.
//
// MODE 0: one iteration (state t-1 -> t).  MODE 1: state 0 (DESC.m:148-157).  MODE 2: objective of a state.

#ifndef ELL_TB
#define ELL_TB 128
#endif
#define ELL_WARPS (ELL_TB / 32)
#define ELL_FIX 1125899906842624.0             // 2^50
#define ELL_UNFIX (1.0 / 1125899906842624.0)
// timing-experiment switches (profiles/build_variant.sh); any of them breaks the results
#ifndef ELL_ABL_NORED
#define ELL_ABL_NORED 0
#endif
#ifndef ELL_ABL_NOGATHER
#define ELL_ABL_NOGATHER 0
#endif
#ifndef ELL_ABL_NOPROJ
#define ELL_ABL_NOPROJ 0
#endif
#ifndef ELL_KEEPD
#define ELL_KEEPD 0                            // 1: keep S0 of the lane's slots in registers instead of re-reading it
#endif

struct EllArgs {
    PgdArgs p;                  // w_cur / w_next / adam_* are in ELL order here; acc_* hold u64 fixed-point sums
    const int4* tiles;          // {base lo, base hi, first edge, steps}
    const int* vtile;           // first tile of every local vertex block
    const double* d;            // S0, ELL order
    const uint16_t* rk;         // rk_i words, ELL order
    const uint32_t* pj;         // pk_jk words, ELL order
    const int *rowstart, *adj_nbr, *adj_eid, *estart;
    int v0;
    int tstride;                // padded max degree
    int max_ns;
    double* partial;            // 2 per CTA: objective / change partials
};

__host__ __device__ __forceinline__ size_t ell_smem_bytes(int tstride, int max_ns) {
    return (size_t)tstride * 8 + (size_t)((max_ns + 2 + 1) & ~1) * 8 + (size_t)tstride * 4;
}

__device__ __forceinline__ unsigned long long ell_fix(double w) {
    return (unsigned long long)__double2ll_rn(w * ELL_FIX);
}
__device__ __forceinline__ double ell_unfix(unsigned long long q) { return __ll2double_rn((long long)q) * ELL_UNFIX; }

template <int G, int EPL, int RULE, int MODE>
__global__ void __launch_bounds__(ELL_TB)
k_pgd_ell(EllArgs a) {
    if (a.p.ctrl[0]) return;
    extern __shared__ __align__(16) unsigned char ell_smem[];
    double* T_S = reinterpret_cast<double*>(ell_smem);
    double* rcp = T_S + a.tstride;
    int* T_E = reinterpret_cast<int*>(rcp + ((a.max_ns + 2 + 1) & ~1));
    __shared__ double red[2 * ELL_WARPS];
    const int v = a.v0 + blockIdx.x;
    const int t0 = a.vtile[blockIdx.x], t1 = a.vtile[blockIdx.x + 1];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double objp = 0.0, chgp = 0.0;
    if (t1 > t0) {
        const int rs = a.rowstart[v];
        const int deg = a.rowstart[v + 1] - rs;
        const int e_hi = a.estart[v + 1];
        // tables of the vertex block: S of every incident edge and the accumulator index of (edge, via v)
        for (int r = threadIdx.x; r < deg; r += ELL_TB) {
            const int e2 = a.adj_eid[rs + r];
            if (MODE != 1) T_S[r] = a.p.S_cur[e2];
            if (MODE != 2) T_E[r] = 2 * e2 + (v < a.adj_nbr[rs + r] ? 0 : 1);
        }
        if (MODE == 0)
            for (int c = threadIdx.x; c <= a.max_ns; c += ELL_TB) rcp[c] = c > 0 ? 1.0 / (double)c : 0.0;
        __syncthreads();
        const int q = lane / G, r = lane % G;
        const unsigned long long* accq = reinterpret_cast<const unsigned long long*>(a.p.acc_cur);
        unsigned long long* accn = reinterpret_cast<unsigned long long*>(a.p.acc_next);
        const double nlr = -a.p.lr;
        for (int t = t0 + warp; t < t1; t += ELL_WARPS) {
            const int4 tl = a.tiles[t];
            const int64_t base = (int64_t)(((uint64_t)(uint32_t)tl.y << 32) | (uint64_t)(uint32_t)tl.x) + lane;
            const int steps = tl.w;
            const int e = tl.z + q;
            int ns = 0;
            if (e < e_hi) ns = (int)(a.p.rowptr[e + 1] - a.p.rowptr[e]);
            const int nx = (ns - r + G - 1) / G;       // this lane owns steps x < nx of its edge
            if (MODE == 1) {
                // ---- state 0: uniform weights, S = mean of the edge's S0, partner sums of state 0
                const double w0 = ns > 0 ? 1.0 / (double)ns : 0.0;
                const unsigned long long f0 = ell_fix(w0);
                double sn = 0.0;
#pragma unroll
                for (int x = 0; x < EPL; x++) {
                    if (x < steps && x < nx) {
                        const int64_t pos = base + x * 32;
                        sn = fma(w0, __ldg(a.d + pos), sn);
                        a.p.w_next[pos] = w0;
                        const uint32_t rkv = __ldg(a.rk + pos), pjv = __ldg(a.pj + pos);
                        if (rkv & RK_APP) atomicAdd(accn + T_E[rkv & RK_MASK], f0);
                        if (pjv & PK_APP) atomicAdd(accn + 2 * (int64_t)(pjv & PK_MASK) + ((pjv & PK_SEL) ? 0 : 1), f0);
                    }
                }
                sn = group_sum<G>(sn);
                if (r == 0 && ns > 0) a.p.S_next[e] = sn;
                continue;
            }
            if (MODE == 2) {
#pragma unroll
                for (int x = 0; x < EPL; x++) {
                    if (x < steps && x < nx) {
                        const int64_t pos = base + x * 32;
                        const uint32_t rkv = __ldg(a.rk + pos), pjv = __ldg(a.pj + pos);
                        objp = fma(__ldcs(a.p.w_cur + pos), a.p.S_cur[pjv & PK_MASK] + T_S[rkv & RK_MASK], objp);
                    }
                }
                continue;
            }
            // ---- one iteration
            double A = 0.0, B = 0.0, Sold = 0.0;
            if (ns > 0) {
                const ulonglong2 ab = *reinterpret_cast<const ulonglong2*>(accq + 2 * (int64_t)e);
                A = ell_unfix(ab.x);
                B = ell_unfix(ab.y);
                Sold = a.p.S_cur[e];
            }
            double u[EPL];
#if ELL_KEEPD
            double dk[EPL];
#endif
            double gsum = 0.0, wsum = 0.0;
#pragma unroll
            for (int x = 0; x < EPL; x++) {
                u[x] = -1e300;
#if ELL_KEEPD
                dk[x] = 0.0;
#endif
                if (x < steps) {
                    const bool ok = x < nx;
                    const int64_t pos = base + x * 32;
                    uint32_t rkv = 0u, pjv = 0u;
                    double w = 0.0, dd = 0.0, sj = 0.0;
                    if (ok) {
                        rkv = __ldg(a.rk + pos);
                        pjv = __ldg(a.pj + pos);
                        w = __ldcs(a.p.w_cur + pos);
                        dd = __ldg(a.d + pos);
                        sj = ELL_ABL_NOGATHER ? 0.25 : a.p.S_cur[pjv & PK_MASK];
                    }
                    const double sg = sj + T_S[rkv & RK_MASK];
                    objp = fma(w, sg, objp);
                    const double part = ((rkv & RK_APP) ? A : 0.0) + ((rkv & RK_APP2) ? B : 0.0);
                    const double g = ok ? fma(part, dd, sg) : 0.0;
                    gsum += g;
                    wsum += w;
#if ELL_KEEPD
                    dk[x] = dd;
#endif
                    if (ok) u[x] = RULE == 0 ? fma(nlr, g, w) : g;   // w - lr g (the mean is added back below)
                }
            }
            if (G > 1) {
                gsum = group_sum<G>(gsum);
                wsum = group_sum<G>(wsum);
            }
            const double fns = (double)ns;
            const double rns = rcp[ns];                  // 1/ns (0 when the edge has no slots)
            const double gmean = gsum * rns;             // tangent projection = mean removal (DESC.m:195-204)
            if (RULE == 0) {
                wsum = fma(nlr, gsum - fns * gmean, wsum);
                const double c = a.p.lr * gmean;
#pragma unroll
                for (int x = 0; x < EPL; x++)
                    if (x < nx) u[x] += c;
            } else {
                wsum = 0.0;
#pragma unroll
                for (int x = 0; x < EPL; x++) {
                    if (x < nx) {
                        const int64_t pos = base + x * 32;
                        const double gr = u[x] - gmean;
                        const double mt = a.p.beta1 * a.p.adam_m[pos] + (1.0 - a.p.beta1) * gr;
                        const double vt = a.p.beta2 * a.p.adam_v[pos] + (1.0 - a.p.beta2) * (gr * gr);
                        a.p.adam_m[pos] = mt;
                        a.p.adam_v[pos] = vt;
                        u[x] = a.p.w_cur[pos] + -a.p.lr * (mt / a.p.corr1) / (sqrt(vt / a.p.corr2) + 1e-8);
                        wsum += u[x];
                    }
                }
                if (G > 1) wsum = group_sum<G>(wsum);
            }
            // Michelot's active-set iteration: T <- (sum_{u>T} u - 1)/#{u>T} until the set stops shrinking; the
            // fixed point is the threshold of the reference's sort-and-scan (DESC.m:215-223)
            int cnt = ns;
            double T = (wsum - 1.0) * rns;
            for (int mit = 0; mit < (ELL_ABL_NOPROJ ? 0 : G * EPL + 2); mit++) {
                double s2 = 0.0;
                int c2 = 0;
#pragma unroll
                for (int x = 0; x < EPL; x++) {
                    const bool in = u[x] > T;
                    s2 += in ? u[x] : 0.0;
                    c2 += in ? 1 : 0;
                }
                if (G > 1) {
                    s2 = group_sum<G>(s2);
                    c2 = group_sum_int<G>(c2);
                }
                const bool changed = (c2 != cnt) && (c2 > 0);
                if (changed) {
                    T = (s2 - 1.0) * rcp[c2];
                    cnt = c2;
                }
                if (G == 1) {
                    if (!changed) break;                 // a lane owns its edge: no warp-wide agreement needed
                } else if (!__any_sync(0xffffffffu, changed)) {
                    break;
                }
            }
            double snew = 0.0;
#pragma unroll
            for (int x = 0; x < EPL; x++) {
                if (x < steps && x < nx) {
                    const int64_t pos = base + x * 32;
                    const double wo = fmax(u[x] - T, 0.0);
#if ELL_KEEPD
                    snew = fma(wo, dk[x], snew);
#else
                    snew = fma(wo, __ldg(a.d + pos), snew);
#endif
                    __stcs(a.p.w_next + pos, wo);
                    if (wo > 0.0 && !ELL_ABL_NORED) {
                        const uint32_t rkv = __ldg(a.rk + pos), pjv = __ldg(a.pj + pos);
                        const unsigned long long f = ell_fix(wo);
                        if (rkv & RK_APP) atomicAdd(accn + T_E[rkv & RK_MASK], f);
                        if (pjv & PK_APP) atomicAdd(accn + 2 * (int64_t)(pjv & PK_MASK) + ((pjv & PK_SEL) ? 0 : 1), f);
                    }
                }
            }
            if (G > 1) snew = group_sum<G>(snew);
            if (r == 0 && ns > 0) {
                a.p.S_next[e] = snew;
                chgp += fabs(snew - Sold);
            }
        }
    }
    if (MODE == 1) return;
    objp = group_sum<32>(objp);
    chgp = group_sum<32>(chgp);
    if (lane == 0) {
        red[2 * warp] = objp;
        red[2 * warp + 1] = chgp;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double o = 0.0, c = 0.0;
#pragma unroll
        for (int k = 0; k < ELL_WARPS; k++) {
            o += red[2 * k];
            c += red[2 * k + 1];
        }
        a.partial[2 * blockIdx.x] = o;
        a.partial[2 * blockIdx.x + 1] = c;
    }
}

// ---- layout construction -----------------------------------------------------------------------------------
// steps of every tile = ceil(longest slot list of its edges / G); sizes[t] = steps * 32
__global__ void k_ell_steps(const int* __restrict__ te0, const int* __restrict__ tcnt, int ntiles,
                            const int64_t* __restrict__ rowptr, int G, int* __restrict__ sizes) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ntiles) return;
    int mx = 0;
    const int e0 = te0[t];
    for (int k = 0; k < tcnt[t]; k++) mx = max(mx, (int)(rowptr[e0 + k + 1] - rowptr[e0 + k]));
    sizes[t] = ((mx + G - 1) / G) * 32;
}
__global__ void k_ell_tiles(const int* __restrict__ te0, const int* __restrict__ sizes, const int64_t* __restrict__ tbase,
                            int ntiles, int4* __restrict__ tiles) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ntiles) return;
    const uint64_t b = (uint64_t)tbase[t];
    tiles[t] = make_int4((int)(uint32_t)(b & 0xffffffffu), (int)(uint32_t)(b >> 32), te0[t], sizes[t] / 32);
}
// one warp per tile: CSR slot arrays -> ELL order (WHAT 1: rank / partner words, 2: S0, 4: ELL weights back to CSR)
template <int WHAT>
__global__ void __launch_bounds__(256)
k_ell_permute(const int4* __restrict__ tiles, const int* __restrict__ tcnt, int ntiles, int G,
              const int64_t* __restrict__ rowptr, int64_t slot_base, const uint16_t* __restrict__ rk_i,
              const uint32_t* __restrict__ pk_jk, const double* __restrict__ S0, uint16_t* __restrict__ ell_rk,
              uint32_t* __restrict__ ell_pj, double* __restrict__ ell_d, const double* __restrict__ ell_w,
              double* __restrict__ w_csr) {
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    const int q = lane / G, r = lane % G;
    for (int t = warp; t < ntiles; t += nwarps) {
        const int4 tl = tiles[t];
        const int64_t base = (int64_t)(((uint64_t)(uint32_t)tl.y << 32) | (uint64_t)(uint32_t)tl.x) + lane;
        int ns = 0;
        int64_t s0 = 0;
        if (q < tcnt[t]) {
            s0 = rowptr[tl.z + q];
            ns = (int)(rowptr[tl.z + q + 1] - s0);
            s0 -= slot_base;
        }
        for (int x = 0; x < tl.w; x++) {
            const int s = x * G + r;
            const bool ok = s < ns;
            const int64_t pos = base + (int64_t)x * 32;
            if (WHAT & 1) {
                ell_rk[pos] = ok ? rk_i[s0 + s] : (uint16_t)0;
                ell_pj[pos] = ok ? pk_jk[s0 + s] : 0u;
            }
            if (WHAT & 2) ell_d[pos] = ok ? S0[s0 + s] : 0.0;
            if (WHAT & 4) {
                if (ok) w_csr[s0 + s] = ell_w[pos];
            }
        }
    }
}
