// pa_kernels.cu -- hand-written sm_100a kernels of the pair-distribution hot path.
//
//   pa_prepare_kernel      rows C, C', Wv of SURVEY 8a: centre of mass / charge in the
//                          reference's sequential fp64 order (Molecular:2804-2834, 2869-2890,
//                          2953-2963) and wrap_vector (Molecular:1560-1586), once per molecule
//                          per frame instead of once per pair.
//   pa_rdf_fast_kernel     rows D + F for the RDF case (pdf, bin_count_b = 1, reference vector
//                          along +-z): fp64 minimum image in the reference's operation order
//                          (no FMA contraction), exact float32 binning through a threshold
//                          table on the fp64 squared distance, per-CTA shared-memory histogram.
//   pa_general_kernel      rows D + F for everything else (2-D pdf, cdf, centre of charge,
//                          arbitrary reference vectors): same pair loop, literal per-pair
//                          evaluation for pairs inside the cutoff.
//   pa_exception_kernel    literal evaluation of the few pairs the fast kernel cannot decide
//                          (link vector (anti)parallel to the reference axis, coincident atoms).
//
// Exactness notes (see DESIGN.md "Arithmetic contract"):
//   * every fp64 operation that the reference performs is issued as an explicit
//     __dadd_rn/__dsub_rn/__dmul_rn/__ddiv_rn/__dsqrt_rn (never contracted to FMA);
//   * min over the 27 images of fl(fl(dx^2+dy^2)+dz^2) equals fl(fl(min dx^2+min dy^2)+min dz^2)
//     because rounding is monotone (Molecular:1451-1467 restated separably), and
//     min_k fl(d_k^2) = fl((min_k |d_k|)^2) for the same reason;
//   * INT(SQRT(REAL4(d2))/step_a) is a monotone step function of the fp64 d2, so the bin is
//     found by comparing the bit pattern of d2 against host-built thresholds; the device never
//     evaluates sqrtf/acos for a bin decision (MUFU.SQRT only seeds the table lookup).
#include <cuda_runtime.h>
#include <cstdint>
#include "pa_internal.h"
#include "pa_pair_exact.cuh"

#define PA_THREADS 128
#define PA_JC 128            // observed molecules staged per shared-memory chunk
#define PA_JREC 10           // doubles per staged record: (x-L,x,x+L, y-L,y,y+L, z-L,z,z+L, pad)

static __device__ __forceinline__ unsigned long long dbits(double x) { return (unsigned long long)__double_as_longlong(x); }


// ---------------------------------------------------------------------------
// rows C, C', Wv
__global__ void __launch_bounds__(256) pa_prepare_kernel(PaPrepareArgs a)
{
    const long long total = (long long)a.nframes * a.pts.M;
    for (long long gid = blockIdx.x * (long long)blockDim.x + threadIdx.x; gid < total;
         gid += (long long)gridDim.x * blockDim.x) {
        const int frame = a.frame0 + (int)(gid / a.pts.M);
        const int slot = (int)(gid % a.pts.M);
        int t = 0;
        while (t + 1 < a.ntypes && slot >= a.types[t + 1].offset) ++t;
        const PaDevType ty = a.types[t];
        if (ty.xyz == nullptr) continue;                     // type not referenced by any job
        const int mol = slot - ty.offset;
        const float *r = ty.xyz + ((size_t)frame * ty.nmol + mol) * (size_t)ty.natoms * 3;
        double c[3], q[3] = { 0.0, 0.0, 0.0 };
        if (a.box.com_boost) {                               // Molecular:2810-2817
            c[0] = (double)r[0]; c[1] = (double)r[1]; c[2] = (double)r[2];
        } else {                                             // Molecular:2819-2833
            c[0] = c[1] = c[2] = 0.0;
            for (int at = 0; at < ty.natoms; ++at) {
                const double m = ty.atom_mass[at];
#pragma unroll
                for (int k = 0; k < 3; ++k) c[k] = __dadd_rn(c[k], __dmul_rn((double)r[at * 3 + k], m));
            }
#pragma unroll
            for (int k = 0; k < 3; ++k) c[k] = __ddiv_rn(c[k], ty.mass);
        }
        if (a.need_coc) {                                    // Molecular:2869-2890, 2953-2963
            for (int at = 0; at < ty.natoms; ++at) {
                const double qa = ty.atom_charge[at];
#pragma unroll
                for (int k = 0; k < 3; ++k) q[k] = __dadd_rn(q[k], __dmul_rn((double)r[at * 3 + k], qa));
            }
            if (ty.charge != 0) {
#pragma unroll
                for (int k = 0; k < 3; ++k) q[k] = __ddiv_rn(q[k], ty.realcharge);
            }
        }
        double p[3] = { c[0], c[1], c[2] }, w[3] = { 0.0, 0.0, 0.0 };
        if (!a.box.wrap_trajectory) {                        // Molecular:1442-1449 + wrap_vector
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                for (int it = 0; it < 1000; ++it) {
                    if (p[k] < a.box.lo[k]) { p[k] = __dadd_rn(p[k], a.box.L[k]); w[k] = __dadd_rn(w[k], a.box.L[k]); }
                    else if (p[k] > a.box.hi[k]) { p[k] = __dsub_rn(p[k], a.box.L[k]); w[k] = __dsub_rn(w[k], a.box.L[k]); }
                    else break;
                }
            }
        }
        const long long o = (long long)frame * a.pts.M + slot;
        a.pts.px[o] = p[0]; a.pts.py[o] = p[1]; a.pts.pz[o] = p[2];
        a.pts.wx[o] = w[0]; a.pts.wy[o] = w[1]; a.pts.wz[o] = w[2];
        a.pts.cx[o] = c[0]; a.pts.cy[o] = c[1]; a.pts.cz[o] = c[2];
        if (a.need_coc && a.pts.qx) { a.pts.qx[o] = q[0]; a.pts.qy[o] = q[1]; a.pts.qz[o] = q[2]; }
    }
}

// ---------------------------------------------------------------------------
// shared helpers of the two pair kernels

struct PaItem { int frame, i0, j0, j1; };

static __device__ __forceinline__ PaItem pa_decode_item(const PaPass &p, long long item)
{
    PaItem it;
    const int js = (int)(item % p.n_jsplits);
    const long long r = item / p.n_jsplits;
    const int ti = (int)(r % p.n_itiles);
    it.frame = p.frame0 + (int)(r / p.n_itiles);
    it.j0 = js * p.jsplit_len;
    it.j1 = min(p.nobs, it.j0 + p.jsplit_len);
    it.i0 = ti;
    return it;
}

// stage observed molecules [jbeg, jbeg+PA_JC) of one frame as (p-L, p, p+L) triples.
static __device__ __forceinline__ void pa_stage_j(const PaPass &p, long long fbase, int jbeg, int jend, double *s_j)
{
    for (int t = threadIdx.x; t < PA_JC; t += blockDim.x) {
        const int j = jbeg + t;
        double x, y, z;
        if (j < jend) {
            const long long ja = fbase + p.obs_off + j;
            x = p.pts.px[ja]; y = p.pts.py[ja]; z = p.pts.pz[ja];
        } else {
            x = y = z = __longlong_as_double(0x7ff8000000000000LL);   // NaN: never inside the cutoff
        }
        double *rec = s_j + t * PA_JREC;
        rec[0] = __dadd_rn(x, -p.box.L[0]); rec[1] = x; rec[2] = __dadd_rn(x, p.box.L[0]);
        rec[3] = __dadd_rn(y, -p.box.L[1]); rec[4] = y; rec[5] = __dadd_rn(y, p.box.L[1]);
        rec[6] = __dadd_rn(z, -p.box.L[2]); rec[7] = z; rec[8] = __dadd_rn(z, p.box.L[2]);
        rec[9] = 0.0;
    }
}

// min over the three images of one component, squared: fl((min_k |(p2+s_k)-p1|)^2)
static __device__ __forceinline__ double pa_min_image_sq(double jm, double j0, double jp, double p1)
{
    // selects work on the signed values so that |.| stays an operand modifier of DSETP
    const double dm = __dsub_rn(jm, p1);
    const double d0 = __dsub_rn(j0, p1);
    const double dp = __dsub_rn(jp, p1);
    double m = fabs(d0) < fabs(dm) ? d0 : dm;
    m = fabs(dp) < fabs(m) ? dp : m;
    return __dmul_rn(m, m);
}

// ---------------------------------------------------------------------------
// RDF fast path.  RI reference molecules per thread in registers, observed molecules broadcast
// from shared memory.  Shared histogram: one 32-bit counter per class at an 8-byte stride
// (classes 0..na-1 = bins, na = out of bounds in a, na+1 = beyond cutoff, never touched).
template <int RI, bool VERIFY>
__global__ void __launch_bounds__(PA_THREADS) pa_rdf_fast_kernel(PaPass p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int na = p.job.bin_a;
    double *s_j = reinterpret_cast<double *>(smem_raw);                                   // PA_JC*PA_JREC doubles
    unsigned long long *s_thr = reinterpret_cast<unsigned long long *>(s_j + PA_JC * PA_JREC);   // na+3
    unsigned int *s_hist = reinterpret_cast<unsigned int *>(s_thr + (na + 3));           // 2*(na+2) words

    for (int k = threadIdx.x; k < na + 3; k += blockDim.x) s_thr[k] = p.job.thr_a[k];
    for (int k = threadIdx.x; k < 2 * (na + 2); k += blockDim.x) s_hist[k] = 0u;
    __syncthreads();

    const float scale = p.job.guess_scale;
    const unsigned long long cut = p.job.cut_bits;
    const float cls_max = (float)(na + 1);
    const int hi_flag = 0x3E100000;                          // hi word of 2^-30: |link_xy|^2 below this -> literal path
    const long long n_items = (long long)p.nframes * p.n_itiles * p.n_jsplits;
    unsigned long long since_flush = 0;
    constexpr int TI = PA_THREADS * RI;

    // tile shard r of n takes the items w with w % n == r (pa_exec.tile_shard_*)
    for (long long item = (long long)blockIdx.x * p.tshard_count + p.tshard_rank; item < n_items; item += (long long)gridDim.x * p.tshard_count) {
        const PaItem it = pa_decode_item(p, item);
        const long long fbase = (long long)it.frame * p.pts.M;
        const int i0 = it.i0 * TI;
        double p1x[RI], p1y[RI], p1z[RI];
#pragma unroll
        for (int r = 0; r < RI; ++r) {
            const int i = i0 + r * PA_THREADS + threadIdx.x;
            if (i < p.nref) {
                const long long ia = fbase + p.ref_off + i;
                p1x[r] = p.pts.px[ia]; p1y[r] = p.pts.py[ia]; p1z[r] = p.pts.pz[ia];
            } else {
                // padding lanes: far away but finite in float range, so the class seed stays valid
                p1x[r] = p1y[r] = p1z[r] = 1.0e15;
            }
        }
        const unsigned long long item_pairs = (unsigned long long)TI * (unsigned long long)(it.j1 - it.j0);
        if (since_flush + item_pairs > 0x7fffffffull) {      // 32-bit shared counters: flush before they could wrap
            __syncthreads();
            unsigned long long within = 0;
            for (int k = threadIdx.x; k < na + 1; k += blockDim.x) {
                const unsigned int c = s_hist[2 * k];
                if (c) {
                    if (k < na) { atomicAdd(&p.job.acc[k], (unsigned long long)((long long)c * p.weight)); within += c; }
                    else atomicAdd(&p.job.acc[na + 1], (unsigned long long)c);
                    s_hist[2 * k] = 0u;
                }
            }
            if (within) atomicAdd(&p.job.acc[na], within);
            since_flush = 0;
        }
        since_flush += item_pairs;

        for (int jc = it.j0; jc < it.j1; jc += PA_JC) {
            __syncthreads();
            pa_stage_j(p, fbase, jc, it.j1, s_j);
            __syncthreads();
            const int cnt = min(PA_JC, it.j1 - jc);
#pragma unroll 2
            for (int jj = 0; jj < cnt; ++jj) {
                const double2 *rec = reinterpret_cast<const double2 *>(s_j + jj * PA_JREC);
                const double2 r0 = rec[0], r1 = rec[1], r2 = rec[2], r3 = rec[3], r4 = rec[4];
                // r0=(xm,x) r1=(xp,ym) r2=(y,yp) r3=(zm,z) r4=(zp,pad)
#pragma unroll
                for (int r = 0; r < RI; ++r) {
                    const double sx = pa_min_image_sq(r0.x, r0.y, r1.x, p1x[r]);
                    const double sy = pa_min_image_sq(r1.y, r2.x, r2.y, p1y[r]);
                    const double sz = pa_min_image_sq(r3.x, r3.y, r4.x, p1z[r]);
                    const double s1 = __dadd_rn(sx, sy);
                    const double d2 = __dadd_rn(s1, sz);
                    // seed for the class: x ~ sqrt(d2)/step_a from the top 32 bits of d2
                    const int hi = __double2hiint(d2);
                    const float f = __uint_as_float((unsigned)hi * 8u + 0x40000000u);   // (hi - 0x38000000) << 3
                    float rt;
                    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(rt) : "f"(f));
                    float y = fmaf(rt, scale, -0.75f);
                    y = fminf(fmaxf(y, 0.0f), cls_max);
                    const int n = __float_as_int(y + 8388608.0f) - 0x4B000000;   // rint(y) in [0, na+1]
                    const unsigned long long db = dbits(d2);
                    // inside the cutoff the class is min(floor(x), na), so it is n or n+1; beyond the
                    // cutoff the class jumps to na+1 and the seed says nothing: test the cutoff itself.
                    const bool inside = db < cut;
                    const int cls = n + (db >= s_thr[n + 1] ? 1 : 0);
                    if (VERIFY) {
                        // invariant of the seed: true class is n or n+1
                        const bool ok = (s_thr[n] <= db) && (n + 2 > na + 2 || db < s_thr[n + 2]);
                        if (inside && !ok && __double2hiint(s1) >= hi_flag) atomicAdd(p.verify_fail, 1u);
                    }
                    if (__double2hiint(s1) < hi_flag) {
                        // |link_xy| < 3e-5 A: (anti)parallel to the reference axis, coincident or the
                        // molecule itself -> decided by the literal evaluation, not here.
                        const int i = i0 + r * PA_THREADS + threadIdx.x;
                        const int j = jc + jj;
                        if (!(p.same_type && i == j)) {
                            const unsigned int slot = atomicAdd(p.exc_count, 1u);
                            if (slot < p.exc_cap)
                                p.exc_list[slot] = make_uint2((unsigned int)(fbase + p.ref_off + i), (unsigned int)(fbase + p.obs_off + j));
                        }
                    } else if (inside) {
                        atomicAdd(&s_hist[2 * cls], 1u);
                    }
                }
            }
        }
    }
    __syncthreads();
    unsigned long long within = 0;
    for (int k = threadIdx.x; k < na + 1; k += blockDim.x) {
        const unsigned int c = s_hist[2 * k];
        if (c) {
            if (k < na) { atomicAdd(&p.job.acc[k], (unsigned long long)((long long)c * p.weight)); within += c; }
            else atomicAdd(&p.job.acc[na + 1], (unsigned long long)c);
        }
    }
    if (within) atomicAdd(&p.job.acc[na], within);
}

// ---------------------------------------------------------------------------
// General path: any mode / reference vector / centre of charge.  Pairs inside the cutoff get
// the literal evaluation.  SMEM_HIST: 32-bit counters in shared memory (nbins+1 words, last =
// out of bounds), otherwise 64-bit atomics straight into the accumulators in L2.
template <int RI, bool SMEM_HIST>
__global__ void __launch_bounds__(PA_THREADS) pa_general_kernel(const __grid_constant__ PaPass p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int nbins = p.job.bin_a * p.job.bin_b;
    double *s_j = reinterpret_cast<double *>(smem_raw);
    unsigned int *s_hist = reinterpret_cast<unsigned int *>(s_j + PA_JC * PA_JREC);      // nbins+1 (SMEM_HIST)
    if (SMEM_HIST) {
        for (int k = threadIdx.x; k < nbins + 1; k += blockDim.x) s_hist[k] = 0u;
    }
    __syncthreads();
    const unsigned long long cut = p.job.cut_bits;
    const long long n_items = (long long)p.nframes * p.n_itiles * p.n_jsplits;
    unsigned long long since_flush = 0;
    constexpr int TI = PA_THREADS * RI;
    unsigned long long g_within = 0, g_oob = 0;               // !SMEM_HIST: per-thread counters

    // tile shard r of n takes the items w with w % n == r (pa_exec.tile_shard_*)
    for (long long item = (long long)blockIdx.x * p.tshard_count + p.tshard_rank; item < n_items; item += (long long)gridDim.x * p.tshard_count) {
        const PaItem it = pa_decode_item(p, item);
        const long long fbase = (long long)it.frame * p.pts.M;
        const int i0 = it.i0 * TI;
        double p1x[RI], p1y[RI], p1z[RI];
#pragma unroll
        for (int r = 0; r < RI; ++r) {
            const int i = i0 + r * PA_THREADS + threadIdx.x;
            if (i < p.nref) {
                const long long ia = fbase + p.ref_off + i;
                p1x[r] = p.pts.px[ia]; p1y[r] = p.pts.py[ia]; p1z[r] = p.pts.pz[ia];
            } else {
                p1x[r] = p1y[r] = p1z[r] = __longlong_as_double(0x7ff8000000000000LL);
            }
        }
        if (SMEM_HIST) {
            const unsigned long long item_pairs = (unsigned long long)TI * (unsigned long long)(it.j1 - it.j0);
            if (since_flush + item_pairs > 0x7fffffffull) {
                __syncthreads();
                unsigned long long within = 0;
                for (int k = threadIdx.x; k < nbins + 1; k += blockDim.x) {
                    const unsigned int c = s_hist[k];
                    if (c) {
                        if (k < nbins) { atomicAdd(&p.job.acc[k], (unsigned long long)((long long)c * p.weight)); within += c; }
                        else atomicAdd(&p.job.acc[nbins + 1], (unsigned long long)c);
                        s_hist[k] = 0u;
                    }
                }
                if (within) atomicAdd(&p.job.acc[nbins], within);
                since_flush = 0;
            }
            since_flush += item_pairs;
        }
        for (int jc = it.j0; jc < it.j1; jc += PA_JC) {
            __syncthreads();
            pa_stage_j(p, fbase, jc, it.j1, s_j);
            __syncthreads();
            const int cnt = min(PA_JC, it.j1 - jc);
            for (int jj = 0; jj < cnt; ++jj) {
                const double2 *rec = reinterpret_cast<const double2 *>(s_j + jj * PA_JREC);
                const double2 r0 = rec[0], r1 = rec[1], r2 = rec[2], r3 = rec[3], r4 = rec[4];
#pragma unroll
                for (int r = 0; r < RI; ++r) {
                    const double sx = pa_min_image_sq(r0.x, r0.y, r1.x, p1x[r]);
                    const double sy = pa_min_image_sq(r1.y, r2.x, r2.y, p1y[r]);
                    const double sz = pa_min_image_sq(r3.x, r3.y, r4.x, p1z[r]);
                    const double d2 = __dadd_rn(__dadd_rn(sx, sy), sz);
                    if (dbits(d2) < cut) {
                        const int i = i0 + r * PA_THREADS + threadIdx.x;
                        const int j = jc + jj;
                        if (p.same_type && i == j) continue;                     // Distribution:778
                        int bin = 0;
                        const int st = pa_eval_pair_exact(p, fbase + p.ref_off + i, fbase + p.obs_off + j, &bin);
                        if (SMEM_HIST) {
                            if (st == 1) atomicAdd(&s_hist[bin], 1u);
                            else if (st == 2) atomicAdd(&s_hist[nbins], 1u);
                        } else {
                            if (st == 1) { atomicAdd(&p.job.acc[bin], (unsigned long long)p.weight); ++g_within; }
                            else if (st == 2) ++g_oob;
                        }
                    }
                }
            }
        }
    }
    __syncthreads();
    if (SMEM_HIST) {
        unsigned long long within = 0;
        for (int k = threadIdx.x; k < nbins + 1; k += blockDim.x) {
            const unsigned int c = s_hist[k];
            if (c) {
                if (k < nbins) { atomicAdd(&p.job.acc[k], (unsigned long long)((long long)c * p.weight)); within += c; }
                else atomicAdd(&p.job.acc[nbins + 1], (unsigned long long)c);
            }
        }
        if (within) atomicAdd(&p.job.acc[nbins], within);
    } else {
        if (g_within) atomicAdd(&p.job.acc[nbins], g_within);
        if (g_oob) atomicAdd(&p.job.acc[nbins + 1], g_oob);
    }
}

// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(128) pa_exception_kernel(const __grid_constant__ PaPass p)
{
    const unsigned int n = min(*p.exc_count, p.exc_cap);
    const int nbins = p.job.bin_a * p.job.bin_b;
    for (unsigned int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
        const uint2 ij = p.exc_list[e];
        int bin = 0;
        const int st = pa_eval_pair_exact(p, (long long)ij.x, (long long)ij.y, &bin);
        if (st == 1) {
            atomicAdd(&p.job.acc[bin], (unsigned long long)p.weight);
            atomicAdd(&p.job.acc[nbins], 1ull);
        } else if (st == 2) {
            atomicAdd(&p.job.acc[nbins + 1], 1ull);
        }
    }
}

// ---------------------------------------------------------------------------
// charge_arm mode: fill_histogram_charge_arm (Distribution:847-895) with charge_arm()
// (Molecular:2926-2937) and compute_squared_radius_of_gyration (Molecular:945-972).
__global__ void __launch_bounds__(128) pa_charge_arm_kernel(PaChargeArmArgs a)
{
    const PaDevType ty = a.types[a.type];
    const long long total = (long long)a.nframes * ty.nmol;
    const int na = a.job.bin_a, nb = a.job.bin_b, nbins = na * nb;
    for (long long id = blockIdx.x * (long long)blockDim.x + threadIdx.x; id < total; id += (long long)gridDim.x * blockDim.x) {
        const int frame = a.frame0 + (int)(id / ty.nmol), mol = (int)(id % ty.nmol);
        const float *r = ty.xyz + ((size_t)frame * ty.nmol + mol) * (size_t)ty.natoms * 3;
        double qd[3] = { 0.0, 0.0, 0.0 }, c[3] = { 0.0, 0.0, 0.0 }, link[3];
        for (int at = 0; at < ty.natoms; ++at) {
            const double q = ty.atom_charge[at], m = ty.atom_mass[at];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const double x = (double)r[at * 3 + k];
                qd[k] = __dadd_rn(qd[k], __dmul_rn(x, q));
                c[k] = __dadd_rn(c[k], __dmul_rn(x, m));
            }
        }
        if (a.com_boost) { c[0] = (double)r[0]; c[1] = (double)r[1]; c[2] = (double)r[2]; }   // Molecular:2810-2817
        else {
#pragma unroll
            for (int k = 0; k < 3; ++k) c[k] = __ddiv_rn(c[k], ty.mass);
        }
        if (ty.charge == 0) { link[0] = qd[0]; link[1] = qd[1]; link[2] = qd[2]; }
        else {
#pragma unroll
            for (int k = 0; k < 3; ++k) link[k] = __dmul_rn(ty.realcharge, __dsub_rn(__ddiv_rn(qd[k], ty.realcharge), c[k]));
        }
        if (a.normalise_clm) {
            double rgy = 0.0;
            for (int at = 0; at < ty.natoms; ++at) {
                const double dx = __dsub_rn((double)r[at * 3 + 0], c[0]), dy = __dsub_rn((double)r[at * 3 + 1], c[1]), dz = __dsub_rn((double)r[at * 3 + 2], c[2]);
                const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
                rgy = __dadd_rn(rgy, __dmul_rn(ty.atom_mass[at], d2));
            }
            rgy = __ddiv_rn(rgy, ty.mass);
            const double den = __dmul_rn(ty.mass, rgy);
#pragma unroll
            for (int k = 0; k < 3; ++k) link[k] = __ddiv_rn(link[k], den);
        }
        const double S = __dadd_rn(__dadd_rn(__dmul_rn(link[0], link[0]), __dmul_rn(link[1], link[1])), __dmul_rn(link[2], link[2]));
        const unsigned long long sb = dbits(S);
        bool oob = !(sb >= a.tiny_bits) || !(S == S);              // a < 1E-8 (or NaN)
        int ia = 0, ib = 0;
        if (!oob) {
            int lo = 0, hi = na;                                   // class in [0, na]; na = binpos_a out of range
            while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (a.job.thr_a[mid] <= sb) lo = mid; else hi = mid - 1; }
            ia = lo;
            const double af = (double)__double2float_rn(__dsqrt_rn(S));
            const double lx = __ddiv_rn(link[0], af), ly = __ddiv_rn(link[1], af), lz = __ddiv_rn(link[2], af);
            const double dot = __dadd_rn(__dadd_rn(__dadd_rn(0.0, __dmul_rn(lx, a.job.ref_vec[0])), __dmul_rn(ly, a.job.ref_vec[1])), __dmul_rn(lz, a.job.ref_vec[2]));
            if (!(dot >= -1.0 && dot <= 1.0)) oob = true;
            else {
                int l2 = 0, h2 = nb;
                while (l2 < h2) { const int mid = (l2 + h2 + 1) >> 1; if (dot <= a.job.thr_b[mid]) l2 = mid; else h2 = mid - 1; }
                ib = l2;
            }
            if (ia >= na || ib >= nb) oob = true;
        }
        if (oob) atomicAdd(&a.job.acc[nbins + 1], 1ull);
        else { atomicAdd(&a.job.acc[ib * na + ia], 1ull); atomicAdd(&a.job.acc[nbins], 1ull); }
    }
}

int pa_launch_charge_arm(const PaChargeArmArgs &a, void *stream)
{
    pa_charge_arm_kernel<<<148 * 4, 128, 0, (cudaStream_t)stream>>>(a);
    return (int)cudaGetLastError();
}

// ---------------------------------------------------------------------------
// launch wrappers

static int pa_grid_for(long long n_items, int sm_count, int ctas_per_sm)
{
    long long g = (long long)sm_count * ctas_per_sm;
    if (n_items < g) g = n_items;                            // (a tile shard owns about n_items / tshard_count of them: a few idle CTAs at most)
    if (g < 1) g = 1;
    return (int)g;
}

int pa_launch_prepare(const PaPrepareArgs &a, void *stream)
{
    const long long total = (long long)a.nframes * a.pts.M;
    if (total <= 0) return 0;
    long long blocks = (total + 255) / 256;
    if (blocks > 148 * 32) blocks = 148 * 32;
    pa_prepare_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(a);
    return (int)cudaGetLastError();
}

template <typename K>
static int pa_occupancy(K kernel, size_t smem)
{
    int n = 0;
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, PA_THREADS, smem) != cudaSuccess || n < 1) n = 1;
    return n;
}

int pa_rdf_ri(int nref) { return nref > 2 * PA_THREADS ? 4 : (nref > PA_THREADS ? 2 : 1); }

int pa_launch_rdf_fast(const PaPass &p, int sm_count, void *stream)
{
    const int na = p.job.bin_a;
    const size_t smem = (size_t)PA_JC * PA_JREC * 8 + (size_t)(na + 3) * 8 + (size_t)(na + 2) * 8;
    const long long n_items = (long long)p.nframes * p.n_itiles * p.n_jsplits;
    const bool verify = p.verify_fail != nullptr;
    const int ri = pa_rdf_ri(p.nref);
#define PA_LAUNCH_FAST(RI_, V_)                                                                     \
    do {                                                                                            \
        const int occ = pa_occupancy(pa_rdf_fast_kernel<RI_, V_>, smem);                            \
        pa_rdf_fast_kernel<RI_, V_><<<pa_grid_for(n_items, sm_count, occ), PA_THREADS, smem,        \
                                      (cudaStream_t)stream>>>(p);                                   \
    } while (0)
    if (verify) { if (ri == 4) PA_LAUNCH_FAST(4, true); else if (ri == 2) PA_LAUNCH_FAST(2, true); else PA_LAUNCH_FAST(1, true); }
    else { if (ri == 4) PA_LAUNCH_FAST(4, false); else if (ri == 2) PA_LAUNCH_FAST(2, false); else PA_LAUNCH_FAST(1, false); }
#undef PA_LAUNCH_FAST
    return (int)cudaGetLastError();
}

int pa_general_ri(int nref) { return nref > PA_THREADS ? 2 : 1; }
#define PA_SMEM_HIST_MAX_BINS 40000

int pa_launch_general(const PaPass &p, int sm_count, void *stream)
{
    const int nbins = p.job.bin_a * p.job.bin_b;
    const bool smem_hist = nbins <= PA_SMEM_HIST_MAX_BINS;
    const size_t smem = (size_t)PA_JC * PA_JREC * 8 + (smem_hist ? (size_t)(nbins + 1) * 4 : 0);
    const long long n_items = (long long)p.nframes * p.n_itiles * p.n_jsplits;
    const int ri = pa_general_ri(p.nref);
#define PA_LAUNCH_GEN(RI_, S_)                                                                      \
    do {                                                                                            \
        const int occ = pa_occupancy(pa_general_kernel<RI_, S_>, smem);                             \
        pa_general_kernel<RI_, S_><<<pa_grid_for(n_items, sm_count, occ), PA_THREADS, smem,         \
                                     (cudaStream_t)stream>>>(p);                                    \
    } while (0)
    if (smem_hist) { if (ri == 2) PA_LAUNCH_GEN(2, true); else PA_LAUNCH_GEN(1, true); }
    else { if (ri == 2) PA_LAUNCH_GEN(2, false); else PA_LAUNCH_GEN(1, false); }
#undef PA_LAUNCH_GEN
    return (int)cudaGetLastError();
}

// bookkeeping after the exception kernel of a pass: sticky overflow flag, running total, reset
__global__ void pa_exception_finish_kernel(unsigned int *count, unsigned int *flags, unsigned int cap)
{
    const unsigned int n = *count;
    if (n > cap) flags[1] = 1u;
    flags[3] += min(n, cap);
    *count = 0u;
}

int pa_launch_exceptions(const PaPass &p, unsigned int /*count_hint*/, void *stream)
{
    pa_exception_kernel<<<148, 128, 0, (cudaStream_t)stream>>>(p);
    pa_exception_finish_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(p.exc_count, p.flags, p.exc_cap);
    return (int)cudaGetLastError();
}

// FP64-pipe issue rate without FMA credit (independent DSUB/DMUL chains, the pair kernel's mix):
// the denominator of this path's roofline, measured on the device it runs on.
__global__ void __launch_bounds__(256) pa_fp64_rate_kernel(double *out, double a, double b, int iters)
{
    double x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = a + threadIdx.x * 1e-9 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) { x[i] = __dsub_rn(x[i], b); x[i] = __dmul_rn(x[i], a); }
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

extern "C" int pa_debug_fp64_issue_rate(int device, double *inst_per_s)
{
    if (cudaSetDevice(device) != cudaSuccess) return PA_ERR_NO_DEVICE;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return PA_ERR_CUDA;
    const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 4096;
    double *d = nullptr;
    if (cudaMalloc(&d, sizeof(double) * blocks * threads) != cudaSuccess) return PA_ERR_CUDA;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 6; ++r) {
        cudaEventRecord(e0);
        pa_fp64_rate_kernel<<<blocks, threads>>>(d, 1.0000001, 1e-7, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (r >= 2 && ms < best) best = ms;
    }
    const cudaError_t err = cudaGetLastError();
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d);
    if (err != cudaSuccess) return PA_ERR_CUDA;
    *inst_per_s = (double)blocks * threads * iters * 16.0 / (best * 1e-3);
    return 0;
}

// Shared-memory atomic rate with the access pattern of the cell-sorted RDF kernels: every lane of a warp increments a
// different 4-byte counter inside a 128-class window of a 1000-class histogram (one observed molecule against 32 reference
// molecules), 128-thread CTAs at the pair kernel's occupancy.  The second denominator of this path's roofline, measured
// on the device it runs on (same kernel as prealpha_b200/csrc/pa_microbench2.cu, layout 0).
__global__ void __launch_bounds__(128) pa_smem_atomic_rate_kernel(unsigned int *out, int nbins, int window, int iters, unsigned int seed)
{
    extern __shared__ unsigned int h[];
    for (int k = threadIdx.x; k < nbins; k += blockDim.x) h[k] = 0;
    __syncthreads();
    unsigned int s = seed ^ (blockIdx.x * 9781u + threadIdx.x * 6271u + 1u);
    unsigned int sw = seed ^ (blockIdx.x * 7919u + (threadIdx.x >> 5) * 104729u + 7u);      // warp-uniform stream: window position
    unsigned int acc = 0;
    for (int it = 0; it < iters; ++it) {
        sw = sw * 1664525u + 1013904223u;
        const int base = (int)(((unsigned long long)(sw >> 8) * (unsigned)(nbins - window)) >> 24);
        s = s * 1664525u + 1013904223u;
        const int bin = base + (int)((s >> 8) & (unsigned)(window - 1));
        asm volatile("red.shared.add.u32 [%0], 1;" ::"r"((unsigned)__cvta_generic_to_shared(h + bin)) : "memory");
        acc += bin;
    }
    __syncthreads();
    unsigned int t = acc;
    for (int k = threadIdx.x; k < nbins; k += blockDim.x) t += h[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = t;
}

extern "C" int pa_debug_smem_atomic_rate(int device, double *atomics_per_s)
{
    if (cudaSetDevice(device) != cudaSuccess) return PA_ERR_NO_DEVICE;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return PA_ERR_CUDA;
    const int blocks = prop.multiProcessorCount * 6, threads = 128, iters = 8192, nbins = 1000;
    unsigned int *d = nullptr;
    if (cudaMalloc(&d, sizeof(unsigned int) * blocks * threads) != cudaSuccess) return PA_ERR_CUDA;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 6; ++r) {
        cudaEventRecord(e0);
        pa_smem_atomic_rate_kernel<<<blocks, threads, nbins * 4>>>(d, nbins, 128, iters, 12345u);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (r >= 2 && ms < best) best = ms;
    }
    const cudaError_t err = cudaGetLastError();
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d);
    if (err != cudaSuccess) return PA_ERR_CUDA;
    *atomics_per_s = (double)blocks * threads * iters / (best * 1e-3);
    return 0;
}

int pa_kernels_selfcheck()
{
    cudaFuncAttributes attr;
    return (int)cudaFuncGetAttributes(&attr, pa_prepare_kernel);
}
