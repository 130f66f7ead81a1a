// A5: cycle inconsistency d_ijk (reference: DESC.m:129-147).
//
//   S0_long(c) = abs(acos((trace(R_ij * R_jk * R_ki) - 1) / 2)) / pi
//
// R_jk = RijMat4d(:,:,j,k) is the stored matrix of edge {j,k} if j<k, else its transpose
// (DESC.m:65-66); likewise R_ki = RijMat4d(:,:,k,i).
//
// Bit-level parity: the reference accumulates the two 3x3 products with separate
// multiply and add roundings, from zero, index j = 1,2,3 (bsxfun loops DESC.m:137-143), and the
// trace as (C11+C22)+C33 (DESC.m:146).  The kernel does exactly that with __dmul_rn/__dadd_rn so
// no FMA contraction can change (trace-1)/2 -- which matters because acos is infinitely
// ill-conditioned at 1 and exactly-consistent cycles land there.  For arguments a few ulp outside
// [-1,1] MATLAB's acos goes complex and abs() takes the modulus: acosh(x) for x>1,
// sqrt(pi^2+acosh(-x)^2) for x<-1.
//
// Layout: one group of G lanes per edge (G chosen from the slot budget), lanes stride over the
// edge's slots; R_ij is loaded once per lane (broadcast within the group), R_jk / R_ki are the
// two gathers.  HBM-bound; tensor cores deliberately unused (3x3 gather arithmetic).
#include "internal.cuh"
#include "so3.cuh"

// M(r,c) of a stored column-major 3x3, optionally transposed
#define MAT(p, r, c, tr) ((tr) ? (p)[(c) + 3 * (r)] : (p)[(r) + 3 * (c)])

template <int G>
__global__ void __launch_bounds__(256)
k_cycle(const double* __restrict__ Rij, const int64_t* __restrict__ rowptr,
        const uint32_t* __restrict__ pk_jk, const uint32_t* __restrict__ pk_ki,
        double* __restrict__ S0, int64_t e0, int64_t e1, int64_t slot_base) {
    const int r = threadIdx.x & (G - 1);
    const int64_t grp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) / G;
    const int64_t ngrp = ((int64_t)gridDim.x * blockDim.x) / G;
    for (int64_t e = e0 + grp; e < e1; e += ngrp) {
        const int64_t s0 = rowptr[e], s1 = rowptr[e + 1];
        if (s0 == s1) continue;
        double a[9];
        const double* pa = Rij + 9 * e;
#pragma unroll
        for (int x = 0; x < 9; x++) a[x] = __ldg(pa + x);  // a[r+3c] = R_ij(r,c)
        for (int64_t s = s0 + r; s < s1; s += G) {
            const uint32_t pj = pk_jk[s - slot_base], pi = pk_ki[s - slot_base];
            const double* pb = Rij + 9 * (int64_t)(pj & PK_MASK);
            const double* pc = Rij + 9 * (int64_t)(pi & PK_MASK);
            const bool tb = !(pj & PK_SEL);  // j>k: stored (k,j) -> transpose
            const bool tc = (pi & PK_SEL);   // i<k: stored (i,k), need (k,i) -> transpose
            double b[9], c[9];
#pragma unroll
            for (int x = 0; x < 9; x++) {
                b[x] = __ldg(pb + x);
                c[x] = __ldg(pc + x);
            }
            double d[3];
#pragma unroll
            for (int q = 0; q < 3; q++) {
                // C0(q,t) = ((0 + A(q,0)B(0,t)) + A(q,1)B(1,t)) + A(q,2)B(2,t)
                double c0[3];
#pragma unroll
                for (int t = 0; t < 3; t++) {
                    double acc = __dmul_rn(a[q + 0], MAT(b, 0, t, tb));
                    acc = __dadd_rn(acc, __dmul_rn(a[q + 3], MAT(b, 1, t, tb)));
                    acc = __dadd_rn(acc, __dmul_rn(a[q + 6], MAT(b, 2, t, tb)));
                    c0[t] = acc;
                }
                // C(q,q) = ((0 + C0(q,0)Rki(0,q)) + C0(q,1)Rki(1,q)) + C0(q,2)Rki(2,q)
                double acc = __dmul_rn(c0[0], MAT(c, 0, q, tc));
                acc = __dadd_rn(acc, __dmul_rn(c0[1], MAT(c, 1, q, tc)));
                acc = __dadd_rn(acc, __dmul_rn(c0[2], MAT(c, 2, q, tc)));
                d[q] = acc;
            }
            const double tr = __dadd_rn(__dadd_rn(d[0], d[1]), d[2]);
            const double x = __ddiv_rn(__dadd_rn(tr, -1.0), 2.0);
            S0[s - slot_base] = __ddiv_rn(abs_acos_dev(x), 3.14159265358979323846);
        }
    }
}

int desc_cycle_impl(desc_b200_handle* h) {
    if (!h->built) {
        desc_set_error("cycle_inconsistency before build_incidence");
        return DESC_B200_ERR_STATE;
    }
    if (h->n_slots > 0) {
        const int grid = DESC_SMS * 8;
        if (h->max_ns <= 8)
            k_cycle<8><<<grid, 256, 0, h->stream>>>(h->Rij, h->rowptr, h->pk_jk, h->pk_ki, h->S0, h->e_begin, h->e_end, h->slot_base);
        else if (h->max_ns <= 16)
            k_cycle<16><<<grid, 256, 0, h->stream>>>(h->Rij, h->rowptr, h->pk_jk, h->pk_ki, h->S0, h->e_begin, h->e_end, h->slot_base);
        else
            k_cycle<32><<<grid, 256, 0, h->stream>>>(h->Rij, h->rowptr, h->pk_jk, h->pk_ki, h->S0, h->e_begin, h->e_end, h->slot_base);
        KERNEL_CHECK(h);
    }
    h->have_s0 = true;
    h->ell_have_d = false;   // the lane-per-edge PGD layout re-reads S0 at its next call
    return DESC_B200_OK;
}
