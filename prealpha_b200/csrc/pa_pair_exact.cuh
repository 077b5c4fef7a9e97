// pa_pair_exact.cuh -- literal per-pair evaluation shared by the general, the exception and the
// cell-sorted general kernels (device code, included by pa_kernels.cu and pa_cells.cu).
#ifndef PA_PAIR_EXACT_CUH
#define PA_PAIR_EXACT_CUH
#include <cuda_runtime.h>
#include "pa_internal.h"

static __device__ __forceinline__ unsigned long long pa_dbits(double x) { return (unsigned long long)__double_as_longlong(x); }

// Fortran INT() of a REAL(4) as the reference's x86-64 build executes it (cvttss2si):
// NaN / out of range -> INT_MIN; then the +1 wraps like the ADD instruction.
static __device__ __forceinline__ int pa_int_of_float_x86(float x)
{
    if (!(x > -2147483904.0f && x < 2147483648.0f)) return (int)0x80000000;
    return __float2int_rz(x);
}
static __device__ __forceinline__ int pa_plus1_wrap(int v) { return (int)((unsigned)v + 1u); }

// ---------------------------------------------------------------------------
// Bin of a pair inside the cutoff from its separable minimum m (fp64) and its link vector (row F,
// Distribution:796-827).  Shared by the literal evaluation below and by the in-register path of the
// cell-sorted general kernel, so both produce the same bits by construction.
// returns 1 = binned (index in *bin), 2 = out of bounds.
// thr_a / thr_b: the job's class tables (global memory) or a copy of them in shared memory.
static __device__ __forceinline__ int pa_bin_from_link(const PaPass &p, double m, const double (&link)[3], int *bin,
                                                       const unsigned long long *thr_a, const double *thr_b)
{
    const int na = p.job.bin_a, nb = p.job.bin_b;
    int ia, ib;
    if (p.job.mode == PA_MODE_CDF) {                         // Distribution:796-809, Angles:104-115
        const double *R = p.job.rot;
        const double rx = __dadd_rn(__dadd_rn(__dmul_rn(link[0], R[0]), __dmul_rn(link[1], R[1])), __dmul_rn(link[2], R[2]));
        const double ry = __dadd_rn(__dadd_rn(__dmul_rn(link[0], R[3]), __dmul_rn(link[1], R[4])), __dmul_rn(link[2], R[5]));
        const double rz = __dadd_rn(__dadd_rn(__dmul_rn(link[0], R[6]), __dmul_rn(link[1], R[7])), __dmul_rn(link[2], R[8]));
        const float a = __double2float_rn(__dsqrt_rn(__dadd_rn(__dmul_rn(rx, rx), __dmul_rn(ry, ry))));
        const float b = __double2float_rn(rz);
        ia = pa_plus1_wrap(pa_int_of_float_x86(__fdiv_rn(a, p.job.step_a)));
        const float bs = __fadd_rn(b, p.job.shift);
        ib = pa_plus1_wrap(pa_int_of_float_x86(__fdiv_rn(bs, p.job.step_b)));
        if (ib == 1 && bs < 0.0f) ib = 0;
    } else {                                                 // Distribution:812-827
        // a-class = largest k with thr_a[k] <= m (as bit patterns): a float32 estimate of SQRT(m)/step_a seeds the
        // search, the monotone table decides (the result does not depend on the seed)
        const unsigned long long db = pa_dbits(m);
        int lo = min(max((int)(__fsqrt_rn(__double2float_rn(m)) * p.job.guess_scale), 0), na + 1);
        while (lo < na + 1 && thr_a[lo + 1] <= db) ++lo;
        while (lo > 0 && thr_a[lo] > db) --lo;
        ia = lo + 1;                                         // class na -> binpos_a = na+1 (out of bounds)
        const double len = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(link[0], link[0]), __dmul_rn(link[1], link[1])), __dmul_rn(link[2], link[2])));
        const double lx = __ddiv_rn(link[0], len), ly = __ddiv_rn(link[1], len), lz = __ddiv_rn(link[2], len);
        const double dot = __dadd_rn(__dadd_rn(__dadd_rn(0.0, __dmul_rn(lx, p.job.ref_vec[0])), __dmul_rn(ly, p.job.ref_vec[1])),
                                     __dmul_rn(lz, p.job.ref_vec[2]));
        if (!(dot >= -1.0 && dot <= 1.0)) ib = -1;           // ACOS -> NaN -> INT() = INT_MIN -> binpos_b < 1
        else {
            // thr_b[k] (k=1..nb) = largest dot with INT(REAL4(ACOS(dot))/step_b) >= k, decreasing in k
            // largest k with dot <= thr_b[k], seeded by a float32 acos (again: the table decides)
            int l2 = min(max((int)(acosf(__double2float_rn(dot)) / p.job.step_b), 0), nb);
            while (l2 < nb && dot <= thr_b[l2 + 1]) ++l2;
            while (l2 > 0 && !(dot <= thr_b[l2])) --l2;
            ib = l2 + 1;
        }
    }
    if (ia > 0 && ia <= na && ib > 0 && ib <= nb) { *bin = (ib - 1) * na + (ia - 1); return 1; }
    return 2;
}

// ---------------------------------------------------------------------------
// literal per-pair evaluation (rows D + F), used by the general and the exception kernel.
// returns 0 = beyond cutoff (not counted), 1 = binned (index in *bin), 2 = out of bounds.
static __device__ __noinline__ int pa_eval_pair_exact(const PaPass &p, long long i_abs, long long j_abs, int *bin)
{
    const PaPoints &P = p.pts;
    const double p1[3] = { P.px[i_abs], P.py[i_abs], P.pz[i_abs] };
    const double p2[3] = { P.px[j_abs], P.py[j_abs], P.pz[j_abs] };
    double v[3][3], mv[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const double L = p.box.L[c];
        const double dm = __dsub_rn(__dadd_rn(p2[c], -L), p1[c]);      // FLOAT(-1)*L = -L
        const double d0 = __dsub_rn(__dadd_rn(p2[c], 0.0), p1[c]);
        const double dp = __dsub_rn(__dadd_rn(p2[c], L), p1[c]);
        v[c][0] = __dmul_rn(dm, dm); v[c][1] = __dmul_rn(d0, d0); v[c][2] = __dmul_rn(dp, dp);
        double m = v[c][0];
        if (v[c][1] < m) m = v[c][1];
        if (v[c][2] < m) m = v[c][2];
        mv[c] = m;
    }
    const double m = __dadd_rn(__dadd_rn(mv[0], mv[1]), mv[2]);
    // best starts at maximum_distance_squared, strict '<' (Molecular:1438,1460)
    const bool beaten = (m < p.box.max_dist_sq);
    // REAL4(min(m, maximum_distance_squared)) < maxdist**2, as a threshold on the fp64 m
    // (the host-built tables already contain the cap, see pa::a_class_of)
    if (!(pa_dbits(m) < p.job.cut_bits)) return 0;
    // first strict minimiser in loop order a(x) outer, c(z) inner (Molecular:1451-1467)
    double s[3] = { 0.0, 0.0, 0.0 };
    if (beaten) {
        int ka = 1, kb = 1, kc = 1;
        for (int k = 0; k < 3; ++k) if (__dadd_rn(__dadd_rn(v[0][k], mv[1]), mv[2]) == m) { ka = k; break; }
        for (int k = 0; k < 3; ++k) if (__dadd_rn(__dadd_rn(v[0][ka], v[1][k]), mv[2]) == m) { kb = k; break; }
        for (int k = 0; k < 3; ++k) if (__dadd_rn(__dadd_rn(v[0][ka], v[1][kb]), v[2][k]) == m) { kc = k; break; }
        s[0] = (double)(ka - 1) * p.box.L[0]; s[1] = (double)(kb - 1) * p.box.L[1]; s[2] = (double)(kc - 1) * p.box.L[2];
    }
    double link[3];
    {
        const double w1[3] = { P.wx[i_abs], P.wy[i_abs], P.wz[i_abs] };
        const double w2[3] = { P.wx[j_abs], P.wy[j_abs], P.wz[j_abs] };
        double c1[3], c2[3];
        if (p.job.use_coc) {
            c1[0] = P.qx[i_abs]; c1[1] = P.qy[i_abs]; c1[2] = P.qz[i_abs];
            c2[0] = P.qx[j_abs]; c2[1] = P.qy[j_abs]; c2[2] = P.qz[j_abs];
        } else {
            c1[0] = P.cx[i_abs]; c1[1] = P.cy[i_abs]; c1[2] = P.cz[i_abs];
            c2[0] = P.cx[j_abs]; c2[1] = P.cy[j_abs]; c2[2] = P.cz[j_abs];
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            double tr = s[c];
            if (!p.box.wrap_trajectory) tr = __dadd_rn(tr, __dsub_rn(w2[c], w1[c]));   // Molecular:1449,1468
            link[c] = __dsub_rn(__dadd_rn(tr, c2[c]), c1[c]);                             // Distribution:786-792
        }
    }
    return pa_bin_from_link(p, m, link, bin, p.job.thr_a, p.job.thr_b);
}


#endif
