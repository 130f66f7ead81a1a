/* oracle/desc_full.c -- TEST INFRASTRUCTURE, not the product.
 *
 * Plain-C / OpenMP restatement of every stage of the reference's solver path EXCEPT the projected-gradient loop
 * (that one is oracle/desc_pgd.c):
 *
 *   desc_c_graph / desc_c_codeg      Algorithms/DESC.m:19-54   n, adjacency, co-degree (A*A).*A, n_sample rule
 *   desc_c_fill / desc_c_recip       Algorithms/DESC.m:56-127  slot lists (find + sampler), Ind_jk / Ind_ki, IKJ / JKI
 *   desc_c_cycle                     Algorithms/DESC.m:129-147 d_ijk in the reference's unfused operation order
 *   desc_c_gcw_weights / _matvec     Utils/GCW.m:13-27         the normalised block operator the eigen-solver applies
 *
 * It exists for two reasons: (a) the numpy oracle (oracle/desc_oracle.py) cannot hold the full-size configurations
 * (1.5e8 slots at cfg 4), so element-wise parity tests of the CUDA path at the headline sizes compare against this
 * port, which is pinned against the numpy oracle on the small fixtures (tests/test_oracle_c.py); (b) it is the
 * multi-core CPU baseline of bench.py (`cpu_baseline`, `--impl reference`): all stages threaded with OpenMP, so the
 * CPU arm is not dominated by single-threaded numpy.
 *
 * Semantics kept from the reference (see the numpy oracle for the long form): 0-based indices here; candidates of
 * an edge in ascending apex order (find, DESC.m:82); `len >= n_sample` samples (DESC.m:83) -- with the shared
 * counter-based key (the stand-in for datasample, identical to csrc/internal.cuh desc_key and
 * desc_oracle.sampler_keys) keeping the n_sample smallest (key, apex) pairs, stored ascending; MATLAB median
 * (mean of the middle two); RijMat4d(:,:,a,b) = stored matrix if a<b else its transpose (DESC.m:63-66); products
 * accumulated column by column from zero with separate multiply / add roundings (DESC.m:137-146; build with
 * -ffp-contract=off); abs(acos(x)) with MATLAB's complex branches for |x| > 1.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may load this.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static void set_threads(int threads) {
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#else
    (void)threads;
#endif
}

/* ---- sampler key (csrc/internal.cuh desc_key, desc_oracle.sampler_keys) ------------------------------------ */
static inline uint64_t desc_key(uint64_t seed, uint64_t edge, uint64_t k) {
    uint64_t z = seed + 0x9E3779B97F4A7C15ull * (edge + 1ull);
    z ^= 0xD1B54A32D192ED03ull * (k + 1ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

/* ---- A1: symmetric CSR adjacency with edge ids (IndMat, DESC.m:67-68) + bitmap (AdjMat, DESC.m:23-24) ------ */
/* rowstart: n+1, nbr/eid: 2m (neighbours ascending within a row), bm: n * nw64 words */
int desc_c_graph(int64_t n, int64_t m, const int32_t* ei, const int32_t* ej, int64_t* rowstart, int32_t* nbr,
                 int32_t* eid, uint64_t* bm, int64_t nw64) {
    memset(rowstart, 0, sizeof(int64_t) * (size_t)(n + 1));
    for (int64_t e = 0; e < m; e++) {
        if (ei[e] < 0 || ej[e] <= ei[e] || ej[e] >= n) return -1;
        rowstart[ei[e] + 1]++;
        rowstart[ej[e] + 1]++;
    }
    for (int64_t v = 0; v < n; v++) rowstart[v + 1] += rowstart[v];
    int64_t* fill = (int64_t*)malloc(sizeof(int64_t) * (size_t)n);
    memcpy(fill, rowstart, sizeof(int64_t) * (size_t)n);
    /* edges are (i,j)-sorted: for row v, the neighbours u<v arrive in increasing u (edges (u,v) by u), then the
       neighbours j>v in increasing j -- but the two groups interleave in time, so fill lower and upper parts apart */
    int64_t* nlow = (int64_t*)calloc((size_t)n, sizeof(int64_t));
    for (int64_t e = 0; e < m; e++) nlow[ej[e]]++;
    int64_t* fill_hi = (int64_t*)malloc(sizeof(int64_t) * (size_t)n);
    for (int64_t v = 0; v < n; v++) fill_hi[v] = rowstart[v] + nlow[v];
    for (int64_t e = 0; e < m; e++) {
        const int32_t i = ei[e], j = ej[e];
        nbr[fill_hi[i]] = j;   /* upper part of row i */
        eid[fill_hi[i]++] = (int32_t)e;
        nbr[fill[j]] = i;      /* lower part of row j */
        eid[fill[j]++] = (int32_t)e;
    }
    free(fill);
    free(fill_hi);
    free(nlow);
    memset(bm, 0, sizeof(uint64_t) * (size_t)(n * nw64));
    for (int64_t e = 0; e < m; e++) {
        bm[(int64_t)ei[e] * nw64 + (ej[e] >> 6)] |= 1ull << (ej[e] & 63);
        bm[(int64_t)ej[e] * nw64 + (ei[e] >> 6)] |= 1ull << (ei[e] & 63);
    }
    return 0;
}

/* ---- A2: co-degree of every edge = |N(i) & N(j)|  ((A*A).*A, DESC.m:29) ---------------------------------- */
void desc_c_codeg(int64_t m, const int32_t* ei, const int32_t* ej, const uint64_t* bm, int64_t nw64, int32_t* codeg,
                  int threads) {
    set_threads(threads);
#pragma omp parallel for schedule(static)
    for (int64_t e = 0; e < m; e++) {
        const uint64_t* a = bm + (int64_t)ei[e] * nw64;
        const uint64_t* b = bm + (int64_t)ej[e] * nw64;
        int c = 0;
        for (int64_t w = 0; w < nw64; w++) c += __builtin_popcountll(a[w] & b[w]);
        codeg[e] = c;
    }
}

/* edge id of {a,b} (IndMat lookup): binary search in row a */
static inline int32_t edge_of(const int64_t* rowstart, const int32_t* nbr, const int32_t* eid, int32_t a, int32_t b) {
    int64_t lo = rowstart[a], hi = rowstart[a + 1];
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (nbr[mid] < b)
            lo = mid + 1;
        else
            hi = mid;
    }
    return eid[lo];
}

/* k-th smallest (0-based) of a[0..n) by Hoare quickselect; permutes a */
static uint64_t kth_smallest(uint64_t* a, int n, int k) {
    int lo = 0, hi = n - 1;
    while (lo < hi) {
        const uint64_t piv = a[lo + ((hi - lo) >> 1)];
        int i = lo, j = hi;
        while (i <= j) {
            while (a[i] < piv) i++;
            while (a[j] > piv) j--;
            if (i <= j) {
                const uint64_t t = a[i];
                a[i] = a[j];
                a[j] = t;
                i++;
                j--;
            }
        }
        if (k <= j)
            hi = j;
        else if (k >= i)
            lo = i;
        else
            break;
    }
    return a[k];
}

/* ---- A3: slot lists (DESC.m:56-96).  rowptr_all (m+1) = exclusive scan of min(codeg, n_sample) over ALL edges.
   Outputs per slot: apex k, Ind_jk, Ind_ki (edge ids). -------------------------------------------------------- */
void desc_c_fill(int64_t n, int64_t m, const int32_t* ei, const int32_t* ej, const uint64_t* bm, int64_t nw64,
                 const int64_t* rowstart, const int32_t* nbr, const int32_t* eid, const int32_t* codeg,
                 const int64_t* rowptr_all, int n_sample, uint64_t seed, int32_t* apex, int32_t* e_jk, int32_t* e_ki,
                 int threads) {
    set_threads(threads);
    int maxc = 1;
    for (int64_t e = 0; e < m; e++)
        if (codeg[e] > maxc) maxc = codeg[e];
#pragma omp parallel
    {
        int32_t* cand = (int32_t*)malloc(sizeof(int32_t) * (size_t)maxc);
        uint64_t* key = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)maxc);
        uint64_t* tmp = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)maxc);
        int32_t* keep = (int32_t*)malloc(sizeof(int32_t) * (size_t)maxc);
#pragma omp for schedule(dynamic, 256)
        for (int64_t e = 0; e < m; e++) {
            const int c = codeg[e];
            if (c == 0) continue;
            const int32_t i = ei[e], j = ej[e];
            const uint64_t* a = bm + (int64_t)i * nw64;
            const uint64_t* b = bm + (int64_t)j * nw64;
            int cnt = 0;                                       /* find(AdjMat(:,i).*AdjMat(:,j)): ascending k */
            for (int64_t w = 0; w < nw64; w++) {
                uint64_t x = a[w] & b[w];
                while (x) {
                    const int bit = __builtin_ctzll(x);
                    x &= x - 1;
                    cand[cnt++] = (int32_t)(w * 64 + bit);
                }
            }
            int ns = c;
            if (c > n_sample) {
                /* DESC.m:83-85 with the shared key sampler: keep the n_sample smallest (key, apex) pairs, in
                   ascending apex order.  Threshold = n_sample-th smallest key; ties (never seen with 64-bit keys,
                   handled anyway) go to the smallest apices, which is the (key, apex) order. */
                for (int q = 0; q < c; q++) tmp[q] = key[q] = desc_key(seed, (uint64_t)e, (uint64_t)cand[q]);
                const uint64_t thr = kth_smallest(tmp, c, n_sample - 1);
                int below = 0;
                for (int q = 0; q < c; q++) below += key[q] < thr;
                int ties = n_sample - below;
                ns = 0;
                for (int q = 0; q < c; q++) {
                    if (key[q] < thr)
                        keep[ns++] = cand[q];
                    else if (key[q] == thr && ties > 0) {
                        keep[ns++] = cand[q];
                        ties--;
                    }
                }
            } else {
                for (int q = 0; q < ns; q++) keep[q] = cand[q];
            }
            const int64_t r0 = rowptr_all[e];
            for (int q = 0; q < ns; q++) {
                const int32_t k = keep[q];
                apex[r0 + q] = k;
                e_jk[r0 + q] = edge_of(rowstart, nbr, eid, j, k);   /* Ind_jk = IndMat(j,k), DESC.m:87 */
                e_ki[r0 + q] = edge_of(rowstart, nbr, eid, i, k);   /* Ind_ki = IndMat(k,i), DESC.m:88 */
            }
        }
        free(cand);
        free(key);
        free(tmp);
        free(keep);
    }
}

/* ---- A4: reciprocal slots (DESC.m:98-127): IKJ(c) = slot of (ik; j) or -1, JKI(c) = slot of (jk; i) or -1.
   Apex lists are ascending (desc_c_fill) -> binary search; `sorted` = 0 searches linearly (explicit lists). ----- */
static inline int64_t slot_of(const int64_t* rowptr_all, const int32_t* apex, int64_t edge, int32_t v, int sorted) {
    int64_t lo = rowptr_all[edge], hi = rowptr_all[edge + 1];
    if (!sorted) {
        for (int64_t s = lo; s < hi; s++)
            if (apex[s] == v) return s;
        return -1;
    }
    const int64_t end = hi;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (apex[mid] < v)
            lo = mid + 1;
        else
            hi = mid;
    }
    return (lo < end && apex[lo] == v) ? lo : -1;
}
void desc_c_recip(int64_t m, const int32_t* ei, const int32_t* ej, const int64_t* rowptr_all, const int32_t* apex,
                  const int32_t* e_jk, const int32_t* e_ki, int sorted, int64_t* IKJ, int64_t* JKI, int threads) {
    set_threads(threads);
#pragma omp parallel for schedule(dynamic, 1024)
    for (int64_t e = 0; e < m; e++) {
        for (int64_t s = rowptr_all[e]; s < rowptr_all[e + 1]; s++) {
            IKJ[s] = slot_of(rowptr_all, apex, e_ki[s], ej[e], sorted);
            JKI[s] = slot_of(rowptr_all, apex, e_jk[s], ei[e], sorted);
        }
    }
}

/* ---- A5: cycle inconsistency (DESC.m:129-147) ------------------------------------------------------------- */
static double abs_acos(double x) {   /* abs(acos(x)) with MATLAB's complex branches */
    if (x >= -1.0 && x <= 1.0) return acos(x);
    if (x > 1.0) {
        const double t = x - 1.0;
        return log1p(t + sqrt(t * (t + 2.0)));
    }
    if (x < -1.0) {
        const double t = -x - 1.0;
        const double a = log1p(t + sqrt(t * (t + 2.0)));
        return sqrt(M_PI * M_PI + a * a);
    }
    return x;   /* NaN */
}
#define MAT(p, r, c, tr) ((tr) ? (p)[(c) + 3 * (r)] : (p)[(r) + 3 * (c)])
/* Rij: MATLAB 3x3xm (element (r,c) of edge e at 9e + r + 3c) */
void desc_c_cycle(int64_t m, const int32_t* ei, const int32_t* ej, const int64_t* rowptr_all, const int32_t* apex,
                  const int32_t* e_jk, const int32_t* e_ki, const double* Rij, double* S0, int threads) {
    set_threads(threads);
#pragma omp parallel for schedule(dynamic, 1024)
    for (int64_t e = 0; e < m; e++) {
        const double* A = Rij + 9 * e;
        const int32_t i = ei[e], j = ej[e];
        for (int64_t s = rowptr_all[e]; s < rowptr_all[e + 1]; s++) {
            const int32_t k = apex[s];
            const double* B = Rij + 9 * (int64_t)e_jk[s];
            const double* C = Rij + 9 * (int64_t)e_ki[s];
            const int tb = !(j < k);   /* RijMat4d(:,:,j,k): stored (j,k) if j<k else transpose (DESC.m:65-66) */
            const int tc = !(k < i);   /* RijMat4d(:,:,k,i) */
            double d[3];
            for (int q = 0; q < 3; q++) {
                double c0[3];
                for (int t = 0; t < 3; t++) {   /* R_cycle0(q,t) = ((0 + A(q,1)B(1,t)) + A(q,2)B(2,t)) + A(q,3)B(3,t) */
                    double acc = A[q + 0] * MAT(B, 0, t, tb);
                    acc = acc + A[q + 3] * MAT(B, 1, t, tb);
                    acc = acc + A[q + 6] * MAT(B, 2, t, tb);
                    c0[t] = acc;
                }
                double acc = c0[0] * MAT(C, 0, q, tc);   /* diagonal of R_cycle (DESC.m:141-146) */
                acc = acc + c0[1] * MAT(C, 1, q, tc);
                acc = acc + c0[2] * MAT(C, 2, q, tc);
                d[q] = acc;
            }
            const double tr = (d[0] + d[1]) + d[2];
            S0[s] = abs_acos((tr - 1.0) / 2.0) / M_PI;
        }
    }
}

/* ---- A13: GCW (Utils/GCW.m:13-27).  coef_e = w_e / sqrt(d_i d_j), w_e = 1/(s^power + 1e-8), d = row sums ---- */
void desc_c_gcw_weights(int64_t n, int64_t m, const int32_t* ei, const int32_t* ej, const double* S_vec, int rule,
                        double* coef, double* isd) {
    /* rule 0: GCW.m:20 (s^1.5); 1: CEMP_GCW.m:141 (s); 2: Spectral.m (unweighted, un-normalised) */
    double* d = (double*)calloc((size_t)n, sizeof(double));
    for (int64_t e = 0; e < m; e++) {
        double w = 1.0;
        if (rule == 0) w = 1.0 / (S_vec[e] * sqrt(S_vec[e]) + 1e-8);
        if (rule == 1) w = 1.0 / (S_vec[e] + 1e-8);
        coef[e] = w;
        d[ei[e]] += w;
        d[ej[e]] += w;
    }
    for (int64_t v = 0; v < n; v++) isd[v] = rule == 2 ? 1.0 : 1.0 / sqrt(d[v]);
    for (int64_t e = 0; e < m; e++) coef[e] = coef[e] * isd[ei[e]] * isd[ej[e]];
    free(d);
}
/* y = N x for ncol right-hand sides, N = D^-1/2 (W o R) D^-1/2 (symmetric, similar to GCW.m:25's D^-1 (W o R));
   x, y: 3n x ncol column-major.  Row-parallel over the symmetric adjacency: deterministic, no atomics. */
void desc_c_gcw_matvec(int64_t n, const int64_t* rowstart, const int32_t* nbr, const int32_t* eid, const double* Rij,
                       const double* coef, const double* x, double* y, int ncol, int threads) {
    set_threads(threads);
#pragma omp parallel for schedule(static)
    for (int64_t v = 0; v < n; v++) {
        for (int c = 0; c < ncol; c++) {
            const double* xc = x + (int64_t)c * 3 * n;
            double a0 = 0.0, a1 = 0.0, a2 = 0.0;
            for (int64_t p = rowstart[v]; p < rowstart[v + 1]; p++) {
                const int32_t u = nbr[p];
                const int64_t e = eid[p];
                const double* R = Rij + 9 * e;
                const double w = coef[e];
                const double x0 = xc[3 * u], x1 = xc[3 * u + 1], x2 = xc[3 * u + 2];
                if (v < u) {   /* block (v,u) = w R */
                    a0 += w * (R[0] * x0 + R[3] * x1 + R[6] * x2);
                    a1 += w * (R[1] * x0 + R[4] * x1 + R[7] * x2);
                    a2 += w * (R[2] * x0 + R[5] * x1 + R[8] * x2);
                } else {       /* block (v,u) = w R' */
                    a0 += w * (R[0] * x0 + R[1] * x1 + R[2] * x2);
                    a1 += w * (R[3] * x0 + R[4] * x1 + R[5] * x2);
                    a2 += w * (R[6] * x0 + R[7] * x1 + R[8] * x2);
                }
            }
            double* yc = y + (int64_t)c * 3 * n;
            yc[3 * v] = a0;
            yc[3 * v + 1] = a1;
            yc[3 * v + 2] = a2;
        }
    }
}
