/* oracle/desc_pgd.c -- TEST INFRASTRUCTURE, not the product.
 *
 * Plain-C restatement of the reference's projected-gradient loop, Algorithms/DESC.m:148-261, on the CSR
 * incidence of oracle/desc_oracle.py (same arrays, 0-based).  It follows the reference statement by
 * statement -- gather form of the partner sums through IKJ/JKI (DESC.m:185-191), gradient (:193),
 * tangent projection (:195-204), constant / piecewise step (:207, Utils/ConstantStepSize.m,
 * PiecewiseStepSize.m), ascending sort + threshold search for the simplex projection (:215-224),
 * S update (:229), bookkeeping and the patience-30 early stop (:232-256) -- with the per-edge loops
 * run by OpenMP threads.  Used (a) to cross-check the numpy oracle (tests/test_oracle.py) and (b) as the
 * multi-core CPU baseline of bench.py.  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs
 * may load it.
 *
 * Build: make -C oracle   (gcc -O2 -fopenmp -shared -fPIC)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static int cmp_double(const void* a, const void* b) {
    const double x = *(const double*)a, y = *(const double*)b;
    return (x > y) - (x < y);
}

/* ascending sort of the edge's candidate weights (DESC.m:215): insertion sort for the short lists the sampler
   produces, qsort beyond */
static void sort_asc(double* a, int n) {
    if (n > 64) {
        qsort(a, (size_t)n, sizeof(double), cmp_double);
        return;
    }
    for (int i = 1; i < n; i++) {
        const double v = a[i];
        int j = i - 1;
        while (j >= 0 && a[j] > v) {
            a[j + 1] = a[j];
            j--;
        }
        a[j + 1] = v;
    }
}

/* returns iterations run.  hist: 2*iters doubles [average_change, obj] per iteration.
 * rule_kind 0: step = -lr*grad ; 1: t++, step = -lr/(fix(t/decay)+1)*grad (t_io is read and advanced) */
int desc_c_pgd(int64_t m, int64_t m_pos, const int64_t* pos_edges, const int64_t* rowptr,
               const int32_t* e_jk, const int32_t* e_ki, const int64_t* IKJ, const int64_t* JKI,
               const double* S0, int iters, int rule_kind, double lr, double decay_interval, int64_t* t_io,
               int patience, double tol, int threads, double* S_vec, double* wijk, double* hist) {
    const int64_t m_cycle = rowptr[m_pos];
    int max_ns = 0;
    for (int64_t l = 0; l < m_pos; l++) {
        const int ns = (int)(rowptr[l + 1] - rowptr[l]);
        if (ns > max_ns) max_ns = ns;
    }
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#endif
    double* grad = (double*)malloc(sizeof(double) * (size_t)(m_cycle > 0 ? m_cycle : 1));
    double* wnew = (double*)malloc(sizeof(double) * (size_t)(m_cycle > 0 ? m_cycle : 1));
    double* S_last = (double*)malloc(sizeof(double) * (size_t)(m > 0 ? m : 1));
    double* S_next = (double*)malloc(sizeof(double) * (size_t)(m > 0 ? m : 1));
    for (int64_t e = 0; e < m; e++) S_vec[e] = 1.0;                       /* DESC.m:148 */
#pragma omp parallel for schedule(static)
    for (int64_t l = 0; l < m_pos; l++) {                                 /* DESC.m:151-157 */
        const int64_t a = rowptr[l], b = rowptr[l + 1];
        const double w0 = 1.0 / (double)(b - a);
        double s = 0.0;
        for (int64_t c = a; c < b; c++) {
            wijk[c] = w0;
            s += w0 * S0[c];
        }
        S_vec[pos_edges[l]] = s;
    }
    memcpy(S_last, S_vec, sizeof(double) * (size_t)m);
    int misses = 0, iters_run = 0;
    double obj_prev = 0.0;
    int64_t t = t_io ? *t_io : 0;
    for (int it = 1; it <= iters; it++) {
        double lr_eff = lr;
        if (rule_kind == 1) {                                             /* PiecewiseStepSize.m:13-18 */
            t += 1;
            lr_eff = lr / (trunc((double)t / decay_interval) + 1.0);
        }
#pragma omp parallel
        {
            double* srt = (double*)malloc(sizeof(double) * (size_t)(max_ns > 0 ? max_ns : 1));
#pragma omp for schedule(static)
            for (int64_t l = 0; l < m_pos; l++) {
                const int64_t a = rowptr[l], b = rowptr[l + 1];
                const int ns = (int)(b - a);
                double A = 0.0, B = 0.0;                                  /* DESC.m:189-190 */
                for (int64_t c = a; c < b; c++) {
                    if (IKJ[c] >= 0) A += wijk[IKJ[c]];
                    if (JKI[c] >= 0) B += wijk[JKI[c]];
                }
                const double nv = 1.0 / sqrt((double)ns);
                double dot = 0.0;
                for (int64_t c = a; c < b; c++) {                         /* DESC.m:193 */
                    const double g = S_vec[e_jk[c]] + S_vec[e_ki[c]] +
                                     ((IKJ[c] >= 0 ? A : 0.0) + (JKI[c] >= 0 ? B : 0.0)) * S0[c];
                    grad[c] = g;
                    dot += g * nv;                                        /* DESC.m:201 */
                }
                for (int64_t c = a; c < b; c++) {
                    const double g = grad[c] - dot * nv;                  /* DESC.m:202 */
                    wnew[c] = wijk[c] + (-lr_eff * g);                    /* DESC.m:207 */
                    srt[c - a] = wnew[c];
                }
                sort_asc(srt, ns);                                        /* DESC.m:215 */
                double T = 0.0;
                for (int Ti = 0; Ti < ns; Ti++) {                         /* DESC.m:216-223 */
                    double acc = 0.0;
                    for (int q = Ti; q < ns; q++) acc += srt[q] - srt[Ti];
                    if (acc < 1.0) {
                        T = srt[Ti] - (1.0 - acc) / (double)(ns - Ti);
                        break;
                    }
                }
                double s = 0.0;
                for (int64_t c = a; c < b; c++) {                         /* DESC.m:224, 229 */
                    const double w = wnew[c] - T;
                    wnew[c] = w > 0.0 ? w : 0.0;
                    s += wnew[c] * S0[c];
                }
                S_next[pos_edges[l]] = s;
            }
            free(srt);
        }
        /* the reference updates wijk / S_vec only after all edges were processed with the OLD values */
        memcpy(wijk, wnew, sizeof(double) * (size_t)m_cycle);
        double change = 0.0, obj = 0.0;
#pragma omp parallel for schedule(static)
        for (int64_t l = 0; l < m_pos; l++) S_vec[pos_edges[l]] = S_next[pos_edges[l]];
#pragma omp parallel for schedule(static) reduction(+ : change)
        for (int64_t e = 0; e < m; e++) change += fabs(S_vec[e] - S_last[e]);       /* DESC.m:232 */
#pragma omp parallel for schedule(static) reduction(+ : obj)
        for (int64_t c = 0; c < m_cycle; c++) obj += wijk[c] * (S_vec[e_jk[c]] + S_vec[e_ki[c]]);   /* :233 */
        hist[2 * (it - 1)] = change / (double)m;
        hist[2 * (it - 1) + 1] = obj;
        iters_run = it;
        if (it > 1 && obj_prev - obj < tol) {                             /* DESC.m:243-256 */
            misses += 1;
            if (misses >= patience) break;
        } else {
            misses = 0;
        }
        obj_prev = obj;
        memcpy(S_last, S_vec, sizeof(double) * (size_t)m);
    }
    if (t_io) *t_io = t;
    free(grad);
    free(wnew);
    free(S_last);
    free(S_next);
    return iters_run;
}

int desc_c_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
