/*
 * pa_oracle.c -- TEST INFRASTRUCTURE ONLY.  CPU restatement of prealpha's
 * pair-distribution path, written to be checked against, never shipped.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this.  The product (prealpha_b200/) never does.
 *
 * Parity status: PINNED.  Every routine here is validated bit-for-bit against
 * the reference's own shipped binary (oracle/_ref/prealpha, run through
 * oracle/build_ref.sh) on the known-answer cases of SURVEY.md 8c; the outputs
 * of those runs are committed under tests/golden/ (see tests/golden/make_golden.py).
 *
 * Each function cites the reference lines it follows.  Shorthand:
 *   Distribution:n = Development/Module_Distribution.f90:n   (pre-alpha.f03 +14884)
 *   Molecular:n    = Development/Module_Molecular.f90:n      (pre-alpha.f03 +1644)
 *   Angles:n       = Development/Module_Angles.f90:n         (pre-alpha.f03 +1461)
 *   SETTINGS:n     = Development/Module_SETTINGS.f90:n       (pre-alpha.f03 +1)
 *
 * Arithmetic contract: coordinates float32, COM and
 * minimum image float64 without FMA, squared distance rounded to float32,
 * everything after that float32.  Build with -ffp-contract=off, no fast-math.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <limits.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "../include/prealpha_b200.h"

/* SETTINGS:46 -- `pi` is REAL(8) but the literal has no d0 suffix, so it is the
 * float32 value of pi widened to double; twopi = 2.0*pi (SETTINGS:47). */
static const double ORC_PI = (double)3.14159274101257324f;
#define ORC_TWOPI (2.0 * ORC_PI)
static const double ORC_DEGREES = 57.295779513082320876798154814105; /* SETTINGS:41 */
static const double ORC_AVOGADRO = 6.02214076e23;                   /* SETTINGS:42 (true double) */
static const double ORC_ECHARGE = (double)1.602176634E-19f;         /* SETTINGS:43 (REAL4 literal) */
static const double ORC_VACPERM = (double)8.8541878128E-12f;        /* SETTINGS:44 (REAL4 literal) */

/* Fortran INT() of a REAL(4) as the x86-64 -O0 build does it (cvttss2si):
 * NaN and out-of-range give INT_MIN ("integer indefinite"). */
static int orc_int_of_float(float x)
{
    if (!(x > -2147483904.0f && x < 2147483648.0f)) return INT_MIN;
    return (int)x;
}
/* wrapping +1 like the two's-complement ADD the binary executes */
static int orc_plus1(int v) { return (int)((unsigned)v + 1u); }

/* ------------------------------------------------------------------------ */
/* give_center_of_mass, Molecular:2804-2834 */
static void orc_center_of_mass(const pa_traj *t, int step, int type, int mol, double c[3])
{
    const pa_type *ty = &t->types[type];
    const float *r = ty->xyz + ((size_t)step * ty->nmol + mol) * (size_t)ty->natoms * 3;
    if (t->com_boost) {               /* use_firstatom_as_com, Molecular:2810-2817 */
        c[0] = (double)r[0]; c[1] = (double)r[1]; c[2] = (double)r[2];
        return;
    }
    c[0] = c[1] = c[2] = 0.0;
    for (int a = 0; a < ty->natoms; ++a) {
        double m = ty->atom_mass[a];
        for (int k = 0; k < 3; ++k) {
            double w = (double)r[a * 3 + k];
            w = w * m;                /* weighted_pos = weighted_pos*mass     */
            c[k] = c[k] + w;          /* com = com + weighted_pos             */
        }
    }
    for (int k = 0; k < 3; ++k) c[k] = c[k] / ty->mass;
}

/* give_qd_vector + give_center_of_charge, Molecular:2869-2890, 2953-2963 */
static void orc_center_of_charge(const pa_traj *t, int step, int type, int mol, double c[3])
{
    const pa_type *ty = &t->types[type];
    const float *r = ty->xyz + ((size_t)step * ty->nmol + mol) * (size_t)ty->natoms * 3;
    c[0] = c[1] = c[2] = 0.0;
    for (int a = 0; a < ty->natoms; ++a) {
        double q = ty->atom_charge[a];
        for (int k = 0; k < 3; ++k) {
            double w = (double)r[a * 3 + k];
            w = w * q;
            c[k] = c[k] + w;
        }
    }
    if (ty->charge != 0)
        for (int k = 0; k < 3; ++k) c[k] = c[k] / (double)ty->realcharge;
}

/* wrap_vector, Molecular:1560-1586 (strict compares, at most 1000 moves) */
static void orc_wrap_vector(const pa_traj *t, double p[3], double shift[3])
{
    for (int k = 0; k < 3; ++k) {
        double L = t->box_hi[k] - t->box_lo[k];      /* box_size, Molecular:937 */
        shift[k] = 0.0;
        for (int it = 0; it < 1000; ++it) {
            if (p[k] < t->box_lo[k]) { p[k] = p[k] + L; shift[k] = shift[k] + L; }
            else if (p[k] > t->box_hi[k]) { p[k] = p[k] - L; shift[k] = shift[k] - L; }
            else break;
        }
    }
}

/* give_smallest_distance_squared, Molecular:1430-1469: literal 27-image loop,
 * a outermost, strict `<` keeps the first minimum and its shift. */
static double orc_smallest_distance_squared(const pa_traj *t, const double c1[3], const double c2[3],
                                            double translation[3])
{
    double p1[3] = { c1[0], c1[1], c1[2] }, p2[3] = { c2[0], c2[1], c2[2] };
    double sh1[3], wrapshift[3] = { 0, 0, 0 };
    double best = t->max_dist_sq;
    /* `translation` is unassigned in the reference when no image beats the
     * initial value; zero is our (documented) stand-in. */
    translation[0] = translation[1] = translation[2] = 0.0;
    if (!t->wrap_trajectory) {
        orc_wrap_vector(t, p1, sh1);
        orc_wrap_vector(t, p2, wrapshift);
        for (int k = 0; k < 3; ++k) wrapshift[k] = wrapshift[k] - sh1[k];
    }
    double L[3];
    for (int k = 0; k < 3; ++k) L[k] = t->box_hi[k] - t->box_lo[k];
    for (int a = -1; a <= 1; ++a)
        for (int b = -1; b <= 1; ++b)
            for (int c = -1; c <= 1; ++c) {
                double s[3] = { (double)(float)a * L[0], (double)(float)b * L[1], (double)(float)c * L[2] };
                double d0 = (p2[0] + s[0]) - p1[0];
                double d1 = (p2[1] + s[1]) - p1[1];
                double d2 = (p2[2] + s[2]) - p1[2];
                double clip = (d0 * d0 + d1 * d1) + d2 * d2;   /* SUM(x**2), sequential */
                if (clip < best) {
                    best = clip;
                    translation[0] = s[0]; translation[1] = s[1]; translation[2] = s[2];
                }
            }
    if (!t->wrap_trajectory)
        for (int k = 0; k < 3; ++k) translation[k] = translation[k] + wrapshift[k];
    return best;
}

/* normalize3D, SETTINGS:1052-1059 (length 0 gives NaN components, error 1 is only printed) */
static void orc_normalize3d(double v[3])
{
    double len = sqrt((v[0] * v[0] + v[1] * v[1]) + v[2] * v[2]);
    v[0] = v[0] / len; v[1] = v[1] / len; v[2] = v[2] / len;
}

/* ------------------------------------------------------------------------ */
/* prepare_rotation + make_rotation_matrix, Angles:51-115, target = z axis. */
static void orc_prepare_rotation(const double start[3], double rot[9], int *aligned)
{
    const double tolerance = 0.0001;                   /* Angles:8 */
    double a[3] = { start[0], start[1], start[2] }, b[3] = { 0.0, 0.0, 1.0 };
    double u[3];
    orc_normalize3d(a);
    orc_normalize3d(b);
    double cosa = ((0.0 + a[0] * b[0]) + a[1] * b[1]) + a[2] * b[2];
    double sina = sqrt(1 - cosa * cosa);
    double angle = acos(cosa) * ORC_DEGREES;
    *aligned = 0;
    if (angle <= tolerance) {
        *aligned = 1;
        for (int i = 0; i < 9; ++i) rot[i] = 0.0;
        rot[0] = rot[4] = rot[8] = 1.0;
        return;
    }
    if ((180.0 - angle) <= tolerance) {                /* REAL4 180.0 widened */
        double dummy[3] = { 0, 0, 0 };
        if (a[0] <= (1.0 - tolerance) && a[0] >= (-1.0 + tolerance)) dummy[0] = 1.0; else dummy[2] = 1.0;
        u[0] = a[1] * dummy[2] - a[2] * dummy[1];      /* crossproduct, SETTINGS:1045-1050 */
        u[1] = a[2] * dummy[0] - a[0] * dummy[2];
        u[2] = a[0] * dummy[1] - a[1] * dummy[0];
    } else {
        u[0] = a[1] * b[2] - a[2] * b[1];
        u[1] = a[2] * b[0] - a[0] * b[2];
        u[2] = a[0] * b[1] - a[1] * b[0];
    }
    orc_normalize3d(u);
    double omc = 1.0 - cosa;
    /* row-major R(k,:) ; Angles:53-61 */
    rot[0] = cosa + (u[0] * u[0]) * omc;
    rot[1] = u[0] * u[1] * omc - u[2] * sina;
    rot[2] = u[0] * u[2] * omc + u[1] * sina;
    rot[3] = u[0] * u[1] * omc + u[2] * sina;
    rot[4] = cosa + (u[1] * u[1]) * omc;
    rot[5] = u[1] * u[2] * omc - u[0] * sina;
    rot[6] = u[0] * u[2] * omc - u[1] * sina;
    rot[7] = u[2] * u[1] * omc + u[0] * sina;
    rot[8] = cosa + (u[2] * u[2]) * omc;
}

/* Row S: Distribution:353-377 (step sizes), :526-527 (maxdist_optimize),
 * :656 (shift), :665-667 (reference vector, rotation). */
float orc_charge_arm_maxdist(const pa_traj *t, int ref_type, int normalise_clm);
int orc_job_setup(const pa_traj *t, const pa_job_request *rq, pa_job *job)
{
    memset(job, 0, sizeof *job);
    if (rq->ref_type < 1 || rq->ref_type > t->ntypes) return PA_E_TYPE_INVALID;
    if (rq->obs_type > t->ntypes) return PA_E_TYPE_INVALID;
    job->mode = rq->mode;
    job->ref_type = rq->ref_type; job->obs_type = rq->obs_type;
    job->bin_a = rq->bin_a < 1 ? 1 : (rq->bin_a > 1000 ? 1000 : rq->bin_a);   /* :601-614 */
    job->bin_b = rq->bin_b < 1 ? 1 : (rq->bin_b > 1000 ? 1000 : rq->bin_b);
    job->sampling_interval = rq->sampling_interval;
    job->weigh_charge = rq->weigh_charge; job->use_coc = rq->use_coc;
    job->subtract_uniform = rq->subtract_uniform;
    float maxdist = rq->maxdist;
    job->normalise_clm = rq->normalise_clm;
    if (rq->mode == PA_MODE_CHARGE_ARM) {
        if (rq->maxdist_optimize) maxdist = orc_charge_arm_maxdist(t, rq->ref_type, rq->normalise_clm);
    } else if (rq->maxdist_optimize) {
        double lim = t->box_hi[0] - t->box_lo[0];                 /* give_box_limit = MINVAL(box_size) */
        for (int k = 1; k < 3; ++k) { double l = t->box_hi[k] - t->box_lo[k]; if (l < lim) lim = l; }
        maxdist = (float)(lim / 2.0);                             /* REAL8/REAL4 -> REAL4 */
        if (maxdist < 0.0f) maxdist = 10.0f;
    }
    job->maxdist = maxdist;
    const float ratio = 0.70710678118654752440f;                  /* 2.0**(-0.5) folded in REAL4 */
    if (rq->mode == PA_MODE_CDF) {
        job->step_a = (maxdist * ratio) / (float)job->bin_a;
        job->step_b = ((2.0f * maxdist) * ratio) / (float)job->bin_b;
    } else {
        job->step_a = maxdist / (float)job->bin_a;
        job->step_b = (float)(ORC_PI / (double)(float)job->bin_b);
    }
    job->shift = maxdist * ratio;                                 /* :656 (used by cdf only) */
    for (int k = 0; k < 3; ++k) { job->ref_xyz[k] = rq->ref_xyz[k]; job->ref_vec[k] = (double)(float)rq->ref_xyz[k]; }
    if (rq->ref_xyz[0] == 0 && rq->ref_xyz[1] == 0 && rq->ref_xyz[2] == 0) return PA_ERR_UNSUPPORTED;
    orc_normalize3d(job->ref_vec);
    orc_prepare_rotation(job->ref_vec, job->rot, &job->aligned);
    return 0;
}

/* ------------------------------------------------------------------------ */
/* fill_histogram, Distribution:762-845, for one frame and one observed type. */
static void orc_fill_histogram_charge_arm(const pa_traj *t, const pa_job *job, int step, int64_t *hist, int64_t *within, int64_t *oob);
static void orc_fill_histogram_range(const pa_traj *t, const pa_job *job, int step, int obs_type0, int i_begin, int i_end,
                                     int64_t *hist, int64_t *within, int64_t *oob);
static void orc_fill_histogram(const pa_traj *t, const pa_job *job, int step, int obs_type0,
                               int64_t *hist, int64_t *within, int64_t *oob)
{
    orc_fill_histogram_range(t, job, step, obs_type0, 0, t->types[job->ref_type - 1].nmol, hist, within, oob);
}
/* the same loop nest for the reference molecules i_begin <= i < i_end (0-based): the iterations of the outer loop
 * (Distribution:773) are independent and the results are integer sums, so any partition of it adds up to the whole */
static void orc_fill_histogram_range(const pa_traj *t, const pa_job *job, int step, int obs_type0, int i_begin, int i_end,
                                     int64_t *hist, int64_t *within, int64_t *oob)
{
    const int ref0 = job->ref_type - 1;
    const float maxdist_squared = job->maxdist * job->maxdist;    /* maxdist**2, :659 */
    const int nobs = t->types[obs_type0].nmol;
    const int w = job->weigh_charge ? t->types[obs_type0].charge : 1;
    for (int i = i_begin; i < i_end; ++i) {
        double cref[3];
        orc_center_of_mass(t, step, ref0, i, cref);
        for (int j = 0; j < nobs; ++j) {
            if (ref0 == obs_type0 && i == j) continue;
            double cobs[3], link[3];
            orc_center_of_mass(t, step, obs_type0, j, cobs);
            float d2 = (float)orc_smallest_distance_squared(t, cref, cobs, link);
            if (!(d2 < maxdist_squared)) continue;
            if (job->use_coc) {
                double qo[3], qr[3];
                orc_center_of_charge(t, step, obs_type0, j, qo);
                orc_center_of_charge(t, step, ref0, i, qr);
                for (int k = 0; k < 3; ++k) link[k] = (link[k] + qo[k]) - qr[k];
            } else {
                for (int k = 0; k < 3; ++k) link[k] = (link[k] + cobs[k]) - cref[k];
            }
            int ia, ib;
            if (job->mode == PA_MODE_CDF) {
                const double *R = job->rot;                        /* rotate, Angles:104-115 */
                double x = link[0], y = link[1], z = link[2];
                double rx = (x * R[0] + y * R[1]) + z * R[2];
                double ry = (x * R[3] + y * R[4]) + z * R[5];
                double rz = (x * R[6] + y * R[7]) + z * R[8];
                float a = (float)sqrt(rx * rx + ry * ry);
                float b = (float)rz;
                ia = orc_plus1(orc_int_of_float(a / job->step_a));
                float bs = b + job->shift;
                ib = orc_plus1(orc_int_of_float(bs / job->step_b));
                if (ib == 1 && bs < 0) ib = 0;
            } else {
                float a = sqrtf(d2);
                orc_normalize3d(link);
                double dot = ((0.0 + link[0] * job->ref_vec[0]) + link[1] * job->ref_vec[1]) + link[2] * job->ref_vec[2];
                float b = (float)acos(dot);
                ia = orc_plus1(orc_int_of_float(a / job->step_a));
                ib = orc_plus1(orc_int_of_float(b / job->step_b));
            }
            if (ia > 0 && ia <= job->bin_a && ib > 0 && ib <= job->bin_b) {
                *within += 1;
                hist[(size_t)(ib - 1) * job->bin_a + (ia - 1)] += w;
            } else {
                *oob += 1;
            }
        }
    }
}

/* iterate_timesteps_parallelised, Distribution:696-760.  Frames are independent;
 * OpenMP over frames with private histograms like the reference.  first/stride
 * select a shard of the sampled frames (stride 1, first 0 = everything). */
int orc_distribution_histogram_shard(const pa_traj *t, const pa_job *job, int shard_rank, int shard_count,
                                     int nthreads, int64_t *hist, int64_t *within_bin, int64_t *out_of_bounds)
{
    const size_t nb = (size_t)job->bin_a * job->bin_b;
    memset(hist, 0, nb * sizeof(int64_t));
    int64_t within = 0, oob = 0;
    const int dt = job->sampling_interval < 1 ? 1 : job->sampling_interval;
    const int nsamp = (t->nsteps + dt - 1) / dt;
    if (shard_count < 1) shard_count = 1;
#ifdef _OPENMP
    if (nthreads < 1) nthreads = omp_get_max_threads();
#pragma omp parallel num_threads(nthreads)
#endif
    {
        int64_t *hl = (int64_t *)calloc(nb, sizeof(int64_t));
        int64_t wl = 0, ol = 0;
#ifdef _OPENMP
#pragma omp for schedule(static, 1)
#endif
        for (int k = 0; k < nsamp; ++k) {
            if (k % shard_count != shard_rank) continue;
            int step = k * dt;
            if (job->mode == PA_MODE_CHARGE_ARM) {
                orc_fill_histogram_charge_arm(t, job, step, hl, &wl, &ol);
            } else if (job->obs_type < 1) {
                for (int ot = 0; ot < t->ntypes; ++ot) orc_fill_histogram(t, job, step, ot, hl, &wl, &ol);
            } else {
                orc_fill_histogram(t, job, step, job->obs_type - 1, hl, &wl, &ol);
            }
        }
#ifdef _OPENMP
#pragma omp critical
#endif
        {
            for (size_t i = 0; i < nb; ++i) hist[i] += hl[i];
            within += wl; oob += ol;
        }
        free(hl);
    }
    *within_bin = within; *out_of_bounds = oob;
    return 0;
}

/* ONE frame (0-based step), OpenMP over blocks of reference molecules instead of over frames: same loop body, same
 * integers, but a single huge frame (the 1M-atom benchmark frame) uses every host core.  The reference itself has no
 * such mode (its parallel loop is over time steps, Distribution:717); this exists so that the tests can pin the GPU
 * path on frames of the benchmark's size. */
int orc_distribution_histogram_frame(const pa_traj *t, const pa_job *job, int step, int nthreads,
                                     int64_t *hist, int64_t *within_bin, int64_t *out_of_bounds)
{
    const size_t nb = (size_t)job->bin_a * job->bin_b;
    memset(hist, 0, nb * sizeof(int64_t));
    int64_t within = 0, oob = 0;
    if (step < 0 || step >= t->nsteps || job->mode == PA_MODE_CHARGE_ARM) return -3;
    const int nref = t->types[job->ref_type - 1].nmol;
    const int block = 16, nblocks = (nref + block - 1) / block;
#ifdef _OPENMP
    if (nthreads < 1) nthreads = omp_get_max_threads();
#pragma omp parallel num_threads(nthreads)
#endif
    {
        int64_t *hl = (int64_t *)calloc(nb, sizeof(int64_t));
        int64_t wl = 0, ol = 0;
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 1)
#endif
        for (int b = 0; b < nblocks; ++b) {
            const int i0 = b * block, i1 = (i0 + block < nref) ? i0 + block : nref;
            if (job->obs_type < 1) {
                for (int ot = 0; ot < t->ntypes; ++ot) orc_fill_histogram_range(t, job, step, ot, i0, i1, hl, &wl, &ol);
            } else {
                orc_fill_histogram_range(t, job, step, job->obs_type - 1, i0, i1, hl, &wl, &ol);
            }
        }
#ifdef _OPENMP
#pragma omp critical
#endif
        {
            for (size_t i = 0; i < nb; ++i) hist[i] += hl[i];
            within += wl; oob += ol;
        }
        free(hl);
    }
    *within_bin = within; *out_of_bounds = oob;
    return 0;
}

int orc_distribution_histogram(const pa_traj *t, const pa_job *job, int nthreads,
                               int64_t *hist, int64_t *within_bin, int64_t *out_of_bounds)
{
    return orc_distribution_histogram_shard(t, job, 0, 1, nthreads, hist, within_bin, out_of_bounds);
}

/* ------------------------------------------------------------------------ */
/* transfer_histogram, Distribution:897-977.  g is (bin_a,bin_b) with a fastest. */
int orc_transfer_histogram(const pa_traj *t, const pa_job *job, const int64_t *hist, int64_t within_bin,
                           int64_t out_of_bounds, int verbose, float *g, float *ideal_density_out)
{
    (void)out_of_bounds;
    const int na = job->bin_a, nb = job->bin_b;
    float within_real;
    int rc = 0;
    if (within_bin < 0) {                        /* :904-906, error 101 */
        rc = PA_E_BIN_OVERFLOW;
        within_real = 0.0f;
        for (int b = 0; b < nb; ++b) for (int a = 0; a < na; ++a) within_real = within_real + (float)hist[(size_t)b * na + a];
    } else within_real = (float)within_bin;
    float total_volume;
    if (job->mode == PA_MODE_CDF) {
        float s3 = (job->shift * job->shift) * job->shift;
        total_volume = (float)(ORC_TWOPI * (double)s3);                        /* :918 */
    } else {
        float m3 = (job->maxdist * job->maxdist) * job->maxdist;
        total_volume = (float)(((double)(4.0f / 3.0f) * ORC_PI) * (double)m3); /* :920 */
    }
    float ideal_density = within_real / total_volume;                          /* :924 */
    for (size_t i = 0; i < (size_t)na * nb; ++i) g[i] = (float)hist[i] / ideal_density;   /* :926 */
    if (verbose) {                                                             /* :927-931 */
        const int dt = job->sampling_interval;
        int denom = (t->nsteps / dt) * t->types[job->ref_type - 1].nmol;
        ideal_density = ideal_density / (float)denom;
    }
    for (int ia = 1; ia <= na; ++ia) {
        float a = (float)ia * job->step_a, chunk_area;
        if (job->mode == PA_MODE_CDF) {
            float am = a - job->step_a;
            chunk_area = (float)(ORC_PI * (double)((a * a) - (am * am)));      /* :942 */
        } else {
            float am = a - job->step_a;
            chunk_area = ((a * a) * a) - ((am * am) * am);                     /* :946 */
            chunk_area = (float)((ORC_TWOPI / (double)3.0f) * (double)chunk_area);   /* :947 */
        }
        for (int ib = 1; ib <= nb; ++ib) {
            float chunk_volume;
            if (job->mode == PA_MODE_CDF) chunk_volume = chunk_area * job->step_b;   /* :958 */
            else {
                float b = (float)ib * job->step_b;
                if (job->mode == PA_MODE_CHARGE_ARM) chunk_volume = cosf(b - job->step_b) - cosf(b);    /* :967 */
                else chunk_volume = chunk_area * (cosf(b - job->step_b) - cosf(b));  /* :962 */
            }
            size_t idx = (size_t)(ib - 1) * na + (ia - 1);
            g[idx] = g[idx] / chunk_volume;                                          /* :972 */
        }
    }
    if (job->mode == PA_MODE_CHARGE_ARM) {                                           /* :976, SUM in array element order */
        float sum = 0.0f;
        for (size_t i = 0; i < (size_t)na * nb; ++i) sum = sum + g[i];
        for (size_t i = 0; i < (size_t)na * nb; ++i) g[i] = g[i] / sum;
    }
    *ideal_density_out = ideal_density;
    return rc;
}

/* ------------------------------------------------------------------------ */
/* gfortran list-directed output of one REAL(4) item: 1 blank + G16.9E2 with
 * scale factor 1 in the E branch (libgfortran write_real / FMT_G rules);
 * observed with `cat -A` on the reference binary's output. */
int orc_format_real4(float value, char *buf)
{
    double m = fabs((double)value);
    char tmp[64];
    if (isnan(value)) { strcpy(buf, "              NaN"); return 17; }
    if (isinf(value)) { strcpy(buf, value > 0 ? "         Infinity" : "        -Infinity"); return 17; }
    int use_f = 0, e10 = 0;
    if (m == 0.0) { use_f = 1; e10 = 0; }
    else {
        snprintf(tmp, sizeof tmp, "%.8e", m);      /* 9 significant digits, correctly rounded */
        e10 = atoi(strchr(tmp, 'e') + 1);
        use_f = (e10 >= -1 && e10 <= 8);
    }
    char body[64];
    if (use_f) {
        int decimals = (m == 0.0) ? 8 : 8 - e10;   /* F12.(9-k), k digits before the point */
        snprintf(body, sizeof body, "%.*f", decimals, (double)value);
        /* field: w-4 = 12 right-justified, then 4 blanks */
        snprintf(tmp, sizeof tmp, "%12s    ", body);
    } else {
        snprintf(body, sizeof body, "%.8E", (double)value);      /* d.ddddddddE+dd */
        snprintf(tmp, sizeof tmp, "%16s", body);
    }
    buf[0] = ' ';
    strcpy(buf + 1, tmp);
    return (int)strlen(buf);
}

static void orc_write_reals(FILE *f, const float *v, int n)
{
    char buf[64];
    for (int i = 0; i < n; ++i) { orc_format_real4(v[i], buf); fputs(buf, f); }
    fputc('\n', f);
}

/* number_integral_trapezoid, Distribution:1159-1169 */
static float orc_number_trapezoid(float g_a, float g_b, float r_a, float r_b)
{
    float n = (r_b * g_a - r_a * g_b) / (r_b - r_a);
    float m = (g_b - n) / r_b;
    float rb2 = r_b * r_b, ra2 = r_a * r_a;
    float upper = (m / 4.0f) * (rb2 * rb2) + (n / 3.0f) * (rb2 * r_b);
    float lower = (m / 4.0f) * (ra2 * ra2) + (n / 3.0f) * (ra2 * r_a);
    return (float)(((double)4.0f * ORC_PI) * (double)(upper - lower));
}
/* energy_integral_trapezoid, Distribution:1171-1183 */
static float orc_energy_trapezoid(int reference_charge, float g_a, float g_b, float r_a, float r_b)
{
    float n = (r_b * g_a - r_a * g_b) / (r_b - r_a);
    float m = (g_b - n) / r_b;
    float rb2 = r_b * r_b, ra2 = r_a * r_a;
    float upper = (m / 3.0f) * (rb2 * r_b) + (n / 2.0f) * rb2;
    float lower = (m / 3.0f) * (ra2 * r_a) + (n / 2.0f) * ra2;
    double v = ((double)(float)reference_charge * (ORC_ECHARGE * ORC_ECHARGE)) / ((double)2.0f * ORC_VACPERM);
    v = v * (double)(upper - lower);
    v = v * ORC_AVOGADRO;
    v = v * (double)1.0e10f;
    return (float)v;
}

/* write_distribution_functions, Distribution:982-1185 (cdf and pdf). */
int orc_write_distribution(const pa_traj *t, const pa_job *job, int reference_number, const char *path_prefix,
                           const char *input_file_name, float *g, float ideal_density)
{
    const int na = job->bin_a, nb = job->bin_b;
    char base[2048], header[1024], fname[2200];
    const int reference_charge = t->types[job->ref_type - 1].charge;
    const int write_energy = (reference_charge != 0) && job->weigh_charge;
    snprintf(base, sizeof base, "%sreference_%d_", path_prefix, reference_number);
    snprintf(header, sizeof header, "reference vector %+d %+d %+d, ", job->ref_xyz[0], job->ref_xyz[1], job->ref_xyz[2]);
    /* TRIM(headerline) drops the trailing blank before the next piece is appended (:1029,:1034) */
    size_t hl = strlen(header); while (hl && header[hl - 1] == ' ') header[--hl] = 0;
    if (job->mode == PA_MODE_CHARGE_ARM) {                                           /* :1008-1024 */
        const char *what = reference_charge != 0 ? "charge_arm" : "dipole_moment";
        snprintf(base + strlen(base), sizeof base - strlen(base), "%s%s_type_%d", job->normalise_clm ? "CLM-normalised_" : "", what, job->ref_type);
        snprintf(header + hl, sizeof header - hl, " %s for molecule type %d,", reference_charge != 0 ? "charge arm" : "dipole moment", job->ref_type);
    } else if (job->obs_type < 1) {
        snprintf(base + strlen(base), sizeof base - strlen(base), "all_around_%d", job->ref_type);
        snprintf(header + hl, sizeof header - hl, " all molecules around type %d,", job->ref_type);
    } else {
        snprintf(base + strlen(base), sizeof base - strlen(base), "type_%d_around_%d", job->obs_type, job->ref_type);
        snprintf(header + hl, sizeof header - hl, " all molecules of type %d around type %d,", job->obs_type, job->ref_type);
    }
    float shift_b;
    const char *unitline;
    FILE *f3;
    char mainname[2200];
    if (job->mode == PA_MODE_CDF) {
        shift_b = job->maxdist * 0.70710678118654752440f;                       /* :1040 */
        snprintf(mainname, sizeof mainname, "%s%s_cdf", base, job->weigh_charge ? "_xcharge" : "");
        strcat(header, " cylindrical distribution function");
        unitline = "a(perpendicular_distance) b(collinear_distance) g(a,b)";
    } else if (job->mode == PA_MODE_CHARGE_ARM) {                                    /* :1112-1139 */
        shift_b = 0.0f;
        char trace_name[2200];
        snprintf(trace_name, sizeof trace_name, "%s_rdf", base);
        snprintf(mainname, sizeof mainname, "%s%s_pdf", base, job->subtract_uniform ? "_-trace" : "");
        FILE *ft = fopen(trace_name, "w");
        if (!ft) return 26;
        fprintf(ft, "%s radial distribution function (trace of pdf)\n", header);
        fprintf(ft, "distance g(r)\n");
        for (int ia = 1; ia <= na; ++ia) {
            float trace = 0.0f;
            for (int ib = 0; ib < nb; ++ib) trace = trace + g[(size_t)ib * na + (ia - 1)];
            trace = trace / (float)nb;
            float a = (float)ia * job->step_a;
            a = a - 0.5f * job->step_a;
            float rowt[2] = { a, trace };
            orc_write_reals(ft, rowt, 2);
            if (job->subtract_uniform)
                for (int ib = 0; ib < nb; ++ib) g[(size_t)ib * na + (ia - 1)] = g[(size_t)ib * na + (ia - 1)] - trace;
        }
        fclose(ft);
        unitline = "a(distance) b(polar_angle) g(a,b)";
        strcat(header, " polar distribution function");
    } else {
        shift_b = 0.0f;
        char trace_name[2200];
        snprintf(trace_name, sizeof trace_name, "%s_rdf%s", base, job->weigh_charge ? "_xcharge" : "");
        snprintf(mainname, sizeof mainname, "%s%s%s_pdf", base, job->weigh_charge ? "_xcharge" : "",
                 job->subtract_uniform ? "_-trace" : "");
        FILE *ft = fopen(trace_name, "w");
        snprintf(fname, sizeof fname, "%s_numberintegral", trace_name);
        FILE *fn = fopen(fname, "w");
        FILE *fe = NULL;
        if (!ft || !fn) { if (ft) fclose(ft); if (fn) fclose(fn); return 26; }
        fprintf(ft, "%s radial distribution function (trace of pdf)\n", header);
        fprintf(ft, "distance g(r)\n");
        fprintf(fn, "%s number integral INT(4*pi*r^2*rho*g(r))dR\n", header);
        fprintf(fn, "distance n(r)\n");
        float number_integral = 0.0f, energy_integral = 0.0f, zeros[2] = { 0.0f, 0.0f };
        if (write_energy) {
            snprintf(fname, sizeof fname, "%s_energyintegral", trace_name);
            fe = fopen(fname, "w");
            if (!fe) { fclose(ft); fclose(fn); return 26; }
            fprintf(fe, "%s (Coulomb) energy integral INT(z*e^2*r*rho*g(r)/(2*epsilon0))dR\n", header);
            fprintf(fe, "distance E(r)[J/mol]\n");
            orc_write_reals(fe, zeros, 2);
        }
        orc_write_reals(fn, zeros, 2);
        float previous = 0.0f;
        for (int ia = 1; ia <= na; ++ia) {
            float trace = 0.0f;
            for (int ib = 0; ib < nb; ++ib) trace = trace + g[(size_t)ib * na + (ia - 1)];
            trace = trace / (float)nb;                                           /* :1088 */
            float a = (float)ia * job->step_a;
            number_integral = number_integral + orc_number_trapezoid(previous, trace, a - job->step_a, a) * ideal_density;
            float row[2] = { a, number_integral };
            orc_write_reals(fn, row, 2);
            if (write_energy) {
                energy_integral = energy_integral +
                    orc_energy_trapezoid(reference_charge, previous, trace, a - job->step_a, a) * ideal_density;
                float rowe[2] = { a, energy_integral };
                orc_write_reals(fe, rowe, 2);
            }
            a = a - 0.5f * job->step_a;
            float rowt[2] = { a, trace };
            orc_write_reals(ft, rowt, 2);
            previous = trace;
            if (job->subtract_uniform)
                for (int ib = 0; ib < nb; ++ib) g[(size_t)ib * na + (ia - 1)] = g[(size_t)ib * na + (ia - 1)] - trace;
        }
        fclose(ft); fclose(fn); if (fe) fclose(fe);
        unitline = "a(distance) b(polar_angle) g(a,b)";
        strcat(header, " polar distribution function");
    }
    if (job->weigh_charge && job->mode != PA_MODE_CHARGE_ARM) strcat(header, ", weighed by charge");
    f3 = fopen(mainname, "w");
    if (!f3) return 26;
    fprintf(f3, "Output in this file is based on the input file '%s':\n", input_file_name);
    fprintf(f3, "%s\n%s\n", header, unitline);
    for (int ia = 1; ia <= na; ++ia)
        for (int ib = 1; ib <= nb; ++ib) {
            float a = ((float)ia * job->step_a) - 0.5f * job->step_a;
            float b = (((float)ib * job->step_b) - 0.5f * job->step_b) - shift_b;
            if (job->mode != PA_MODE_CDF) b = (float)((double)b * ORC_DEGREES);
            float row[3] = { a, b, g[(size_t)(ib - 1) * na + (ia - 1)] };
            orc_write_reals(f3, row, 3);
        }
    fclose(f3);
    return 0;
}

/* ------------------------------------------------------------------------ */
/* give_smallest_atom_distance_squared, Molecular:1380-1427: like the COM
 * version but on two atoms and ALWAYS wrapping both. */
double orc_smallest_atom_distance_squared(const pa_traj *t, const float r1[3], const float r2[3])
{
    pa_traj tt = *t; tt.wrap_trajectory = 0;
    double c1[3] = { (double)r1[0], (double)r1[1], (double)r1[2] };
    double c2[3] = { (double)r2[0], (double)r2[1], (double)r2[2] };
    double tr[3];
    return orc_smallest_distance_squared(&tt, c1, c2, tr);
}

int orc_abi_version(void) { return PA_ABI_VERSION; }

/* ------------------------------------------------------------------------ */
/* Minimal xyz body reader, Molecular:4872-4881 (2 header lines per frame) and
 * :5032-5046 (`READ(3,*) element_name, coordinates`): element token, then three
 * REAL(4) items converted decimal->float32 directly (strtof), rest of line ignored.
 * out: [nsteps][natoms_total][3]; elements: [natoms_total][4] from frame 1. */
int orc_read_traj(const char *path, int natoms_total, int nsteps, int header_lines, float *out, char *elements);
int orc_read_xyz(const char *path, int natoms_total, int nsteps, float *out, char *elements)
{
    return orc_read_traj(path, natoms_total, nsteps, 2, out, elements);
}
/* header_lines = 2 (xyz, Molecular:4873) or 9 (lmp, Molecular:4835) */
int orc_read_traj(const char *path, int natoms_total, int nsteps, int header_lines, float *out, char *elements)
{
    FILE *f = fopen(path, "r");
    if (!f) return PA_ERR_IO;
    char *line = NULL; size_t cap = 0;
    for (int s = 0; s < nsteps; ++s) {
        for (int h = 0; h < header_lines; ++h)
            if (getline(&line, &cap, f) < 0) { fclose(f); free(line); return 84; }
        for (int a = 0; a < natoms_total; ++a) {
            if (getline(&line, &cap, f) < 0) { fclose(f); free(line); return 84; }
            char *p = line;
            while (*p == ' ' || *p == '\t') ++p;
            char *e = p;
            while (*e && *e != ' ' && *e != '\t' && *e != ',' && *e != '\n') ++e;
            if (s == 0 && elements) {
                int n = (int)(e - p); if (n > 3) n = 3;
                memset(elements + a * 4, 0, 4); memcpy(elements + a * 4, p, n);
            }
            p = e;
            for (int k = 0; k < 3; ++k) {
                while (*p == ' ' || *p == '\t' || *p == ',') ++p;
                char *q; out[((size_t)s * natoms_total + a) * 3 + k] = strtof(p, &q); p = q;
            }
        }
    }
    free(line); fclose(f);
    return 0;
}

/* ------------------------------------------------------------------------ */
/* Row X: compute_intermolecular_distances / compute_intramolecular_distances
 * (Distances:976-1091 / 856-974), first pass and stdev pass, then the post-processing of
 * perform_distance_analysis (Distances:1161-1243).  Literal loop order. */
static float orc_atom_d2(const pa_traj *t, int step, int ty1, int mol1, int at1, int ty2, int mol2, int at2)
{
    const pa_type *a = &t->types[ty1], *b = &t->types[ty2];
    const float *r1 = a->xyz + (((size_t)step * a->nmol + mol1) * a->natoms + at1) * 3;
    const float *r2 = b->xyz + (((size_t)step * b->nmol + mol2) * b->natoms + at2) * 3;
    return (float)orc_smallest_atom_distance_squared(t, r1, r2);      /* REAL(4) distance_squared */
}

/* contact_distance -> print_distances, Debug:586-623: give_intramolecular_distances (Molecular:1271-1292)
 * and give_intermolecular_contact_distance (Molecular:1295-1367; note the EXIT at :1355 -- within the type
 * under study only molecules with a lower index than the current one are visited).  OpenMP over the outer
 * molecule loop here as in the reference (min is order independent).
 * out = { largest intramolecular, smallest intramolecular, smallest intermolecular }. */
int orc_contact_distance(const pa_traj *t, int timestep, int type, int nthreads, float out[3])
{
    if (timestep < 1 || timestep > t->nsteps || type < 1 || type > t->ntypes) return PA_ERR_BAD_ARG;
    const int step = timestep - 1, ty1 = type - 1;
    const pa_type *a = &t->types[ty1];
    float smallest = (float)t->max_dist_sq, largest = 0.0f;
    for (int m = 0; m < a->nmol; ++m)
        for (int a1 = 0; a1 + 1 < a->natoms; ++a1)
            for (int a2 = a1 + 1; a2 < a->natoms; ++a2) {
                const float cur = orc_atom_d2(t, step, ty1, m, a1, ty1, m, a2);
                if (cur > largest) largest = cur;
                if (cur < smallest) smallest = cur;
            }
    out[0] = sqrtf(largest); out[1] = sqrtf(smallest);
    float inter = (float)t->max_dist_sq;
    if (nthreads < 1) nthreads = 1;
#pragma omp parallel num_threads(nthreads)
    {
        float local = (float)t->max_dist_sq;
#pragma omp for schedule(dynamic, 1)
        for (int m1 = 0; m1 < a->nmol; ++m1)
            for (int a1 = 0; a1 < a->natoms; ++a1)
                for (int ty2 = 0; ty2 < t->ntypes; ++ty2)
                    for (int m2 = 0; m2 < t->types[ty2].nmol; ++m2) {
                        if (ty2 == ty1 && m2 == m1) break;
                        for (int a2 = 0; a2 < t->types[ty2].natoms; ++a2) {
                            const float cur = orc_atom_d2(t, step, ty1, m1, a1, ty2, m2, a2);
                            if (cur < local) local = cur;
                        }
                    }
#pragma omp critical
        if (local < inter) inter = local;
    }
    out[2] = sqrtf(inter);
    return 0;
}

/* Neighbour test of the cluster / speciation analyses (Cluster:1092-1100, Speciation:1407-1431): literal double
 * loop over two lists of atom instances, REAL(4) give_smallest_atom_distance_squared < REAL(4) squared cutoff.
 * Pairs come out in loop order (ia outer, ib inner). */
int orc_neighbour_pairs(const pa_traj *t, const pa_neigh_request *rq, pa_neigh_pair *out, int64_t capacity, int64_t *count)
{
    if (rq->timestep < 1 || rq->timestep > t->nsteps) return PA_ERR_BAD_ARG;
    const int step = rq->timestep - 1;
    int64_t n = 0;
    for (int ia = 0; ia < rq->n_a; ++ia) {
        const int ty1 = rq->a[3 * ia] - 1, m1 = rq->a[3 * ia + 1] - 1, a1 = rq->a[3 * ia + 2] - 1;
        const int row = (rq->a_class ? rq->a_class[ia] : 0) * rq->n_class_b;
        for (int ib = rq->same_list ? ia + 1 : 0; ib < rq->n_b; ++ib) {
            const int ty2 = rq->b[3 * ib] - 1, m2 = rq->b[3 * ib + 1] - 1, a2 = rq->b[3 * ib + 2] - 1;
            if (rq->skip_same_molecule && ty1 == ty2 && m1 == m2) continue;
            const float cut = rq->cutoff_sq[row + (rq->b_class ? rq->b_class[ib] : 0)];
            if (orc_atom_d2(t, step, ty1, m1, a1, ty2, m2, a2) < cut) {
                if (n < capacity) { out[n].ia = ia; out[n].ib = ib; }
                ++n;
            }
        }
    }
    *count = n;
    return 0;
}

static void orc_distance_pass(const pa_traj *t, const pa_dist_job *job, int stdev_run, pa_dist_result *res)
{
    const float maxdist_squared = job->maxdist * job->maxdist;
    for (int step = 0; step < job->nsteps; step += job->sampling_interval) {
        for (int k = 0; k < job->nsubjobs; ++k) {
            const pa_dist_subjob *sj = &job->subjobs[k];
            int64_t localweight = 0, localweight_ffc = 0;
            double local_closest = 0, local_exponential = 0, weight_exponential = 0, local_exponential_stdev = 0, local_ffc = 0, local_ffcexp = 0;
            for (int e1 = 0; e1 < sj->n_ref; ++e1) {
                const int ty1 = sj->ref[2 * e1] - 1, at1 = sj->ref[2 * e1 + 1] - 1;
                for (int m1 = 0; m1 < t->types[ty1].nmol; ++m1) {
                    float closest = maxdist_squared;
                    int valid = 0;
                    if (stdev_run && job->calc_exp) { local_exponential = 0; weight_exponential = 0; }
                    for (int e2 = 0; e2 < sj->n_obs; ++e2) {
                        const int ty2 = sj->obs[2 * e2] - 1, at2 = sj->obs[2 * e2 + 1] - 1;
                        int m2lo, m2hi;
                        if (job->intermolecular) { m2lo = 0; m2hi = t->types[ty2].nmol; }
                        else { if (ty1 != ty2 || at1 == at2) continue; m2lo = m1; m2hi = m1 + 1; }
                        for (int m2 = m2lo; m2 < m2hi; ++m2) {
                            if (job->intermolecular && m1 == m2 && ty1 == ty2) continue;
                            const float d2 = orc_atom_d2(t, step, ty1, m1, at1, ty2, m2, at2);
                            if (!(d2 < maxdist_squared)) continue;
                            valid = 1;
                            if (d2 < closest) closest = d2;
                            double ef = 0.0;
                            if (job->calc_exp) {
                                const float distance = sqrtf(d2);
                                ef = (double)expf(-job->exponent * distance);
                                local_exponential = local_exponential + ef * (double)distance;
                                weight_exponential = weight_exponential + ef;
                            }
                            if (job->calc_ffc) {
                                float ffc;
                                if (job->intermolecular) { const float distance = sqrtf(d2); ffc = 1.0f / ((distance * distance) * distance); }
                                else ffc = 1.0f / ((d2 * d2) * d2);
                                local_ffc = local_ffc + (double)ffc;
                                localweight_ffc += 1;
                                if (job->calc_exp) local_ffcexp = local_ffcexp + ef * (double)ffc;
                            }
                        }
                    }
                    if (valid) {
                        localweight += 1;
                        if (stdev_run) {
                            const double dv = (double)sqrtf(closest) - res[k].closest_average;
                            local_closest = local_closest + dv * dv;
                            if (job->calc_exp) {
                                const double de = local_exponential / weight_exponential - res[k].exponential_average;
                                local_exponential_stdev = local_exponential_stdev + de * de;
                            }
                        } else local_closest = local_closest + (double)sqrtf(closest);
                    } else res[k].missed_neighbours += 1;     /* also during the stdev pass (Distances:948,1067) */
                }
            }
            const double lw = (double)(float)localweight;
            if (stdev_run) {
                res[k].closest_stdev += local_closest / lw;
                if (job->calc_exp) res[k].exponential_stdev += local_exponential_stdev / lw;
            } else {
                res[k].closest_average += local_closest / lw;
                if (job->calc_exp) { res[k].exponential_average += local_exponential; res[k].exponential_weight += weight_exponential; }
                if (job->calc_ffc) {
                    res[k].ffc_average += local_ffc;
                    if (job->calc_exp) res[k].ffcexp_average += local_ffcexp;
                    res[k].ffc_weight += (double)(float)localweight_ffc;
                }
            }
        }
    }
}

int orc_distance_subjobs(const pa_traj *t, const pa_dist_job *job, pa_dist_result *res)
{
    memset(res, 0, sizeof(pa_dist_result) * job->nsubjobs);
    orc_distance_pass(t, job, 0, res);
    int nav = (job->nsteps - 1 + job->sampling_interval) / job->sampling_interval;
    if (nav < 0) nav = 0;
    const double navf = (double)(float)nav;
    for (int k = 0; k < job->nsubjobs; ++k) {
        res[k].closest_average = res[k].closest_average / navf;
        if (job->calc_exp) res[k].exponential_average = res[k].exponential_average / res[k].exponential_weight;
        if (job->calc_ffc) {
            res[k].ffc_average = res[k].ffc_average / res[k].ffc_weight;
            if (job->calc_exp) res[k].ffcexp_average = res[k].ffcexp_average / res[k].exponential_weight;
            /* REAL8 ** REAL4 constant: (-1.0/6.0) and (-1.0/3.0) are REAL(4) quotients */
            const double ex = job->intermolecular ? (double)(-1.0f / 3.0f) : (double)(-1.0f / 6.0f);
            res[k].ffc_average = pow(res[k].ffc_average, ex);
            if (job->calc_exp) res[k].ffcexp_average = pow(res[k].ffcexp_average, ex);
        }
    }
    if (job->calc_stdev) {
        /* the stdev pass of the reference adds to missed_neighbours a second time (Distances:948,1067):
         * reproduce by letting the pass count again, as the binary's report shows the doubled number */
        orc_distance_pass(t, job, 1, res);
        for (int k = 0; k < job->nsubjobs; ++k) {
            res[k].closest_stdev = sqrt(res[k].closest_stdev / navf);
            if (job->calc_exp) res[k].exponential_stdev = sqrt(res[k].exponential_stdev / navf);
        }
    }
    return 0;
}

/* ------------------------------------------------------------------------ */
/* wrap_full, Molecular:1703-1765: REAL(4) centre of mass; whole molecules are moved by
 * +-box_size until the centre of mass is inside; coordinates are rewritten in REAL(4). */
int orc_wrap_full(pa_traj *t)
{
    for (int step = 0; step < t->nsteps; ++step)
        for (int ty = 0; ty < t->ntypes; ++ty) {
            const pa_type *T = &t->types[ty];
            for (int m = 0; m < T->nmol; ++m) {
                double c[3];
                orc_center_of_mass(t, step, ty, m, c);
                float *r = (float *)T->xyz + ((size_t)step * T->nmol + m) * (size_t)T->natoms * 3;
                for (int k = 0; k < 3; ++k) {
                    float com = (float)c[k];
                    const double L = t->box_hi[k] - t->box_lo[k];
                    for (int it = 0; it < 1000; ++it) {
                        if ((double)com < t->box_lo[k]) {
                            com = (float)((double)com + L);
                            for (int a = 0; a < T->natoms; ++a) r[a * 3 + k] = (float)((double)r[a * 3 + k] + L);
                        } else if ((double)com > t->box_hi[k]) {
                            com = (float)((double)com - L);
                            for (int a = 0; a < T->natoms; ++a) r[a * 3 + k] = (float)((double)r[a * 3 + k] - L);
                        } else break;
                    }
                }
            }
        }
    return 0;
}

/* unwrap_full_SEQUENTIAL, Molecular:1590-1699 (called with WRAP_TRAJECTORY = F). */
int orc_unwrap_full(pa_traj *t)
{
    int nmol_total = 0;
    for (int ty = 0; ty < t->ntypes; ++ty) nmol_total += t->types[ty].nmol;
    double *acc = (double *)calloc((size_t)nmol_total * 3, sizeof(double));
    char *jumped = (char *)calloc((size_t)nmol_total, 1);
    double thr[3];
    for (int k = 0; k < 3; ++k) thr[k] = (t->box_hi[k] - t->box_lo[k]) / 2.0;
    for (int step = 0; step + 1 < t->nsteps; ++step) {
        int idx = 0;
        for (int ty = 0; ty < t->ntypes; ++ty) {
            const pa_type *T = &t->types[ty];
            for (int m = 0; m < T->nmol; ++m, ++idx) {
                float *rn = (float *)T->xyz + ((size_t)(step + 1) * T->nmol + m) * (size_t)T->natoms * 3;
                double *as = acc + (size_t)idx * 3;
                if (jumped[idx]) {
                    if ((as[0] * as[0] + as[1] * as[1]) + as[2] * as[2] < 0.001) jumped[idx] = 0;
                    else for (int a = 0; a < T->natoms; ++a) for (int k = 0; k < 3; ++k) rn[a * 3 + k] = (float)((double)rn[a * 3 + k] + as[k]);
                }
                double c1[3], c2[3], link[3];
                orc_center_of_mass(t, step, ty, m, c1);
                orc_center_of_mass(t, step + 1, ty, m, c2);
                (void)orc_smallest_distance_squared(t, c1, c2, link);
                if (fabs(link[0]) > thr[0] || fabs(link[1]) > thr[1] || fabs(link[2]) > thr[2]) {
                    for (int a = 0; a < T->natoms; ++a) for (int k = 0; k < 3; ++k) rn[a * 3 + k] = (float)((double)rn[a * 3 + k] + link[k]);
                    jumped[idx] = 1;
                    for (int k = 0; k < 3; ++k) as[k] = as[k] + link[k];
                }
            }
        }
    }
    free(acc); free(jumped);
    return 0;
}

/* ------------------------------------------------------------------------ */
/* charge_arm mode.  charge_arm(): Molecular:2926-2937; compute_squared_radius_of_gyration:
 * Molecular:945-972; optimise_charge_arm: Distribution:1232-1269; fill_histogram_charge_arm:
 * Distribution:847-895. */
static void orc_charge_arm_vec(const pa_traj *t, int step, int type, int mol, double out[3])
{
    const pa_type *ty = &t->types[type];
    const float *r = ty->xyz + ((size_t)step * ty->nmol + mol) * (size_t)ty->natoms * 3;
    double qd[3] = { 0, 0, 0 };
    for (int a = 0; a < ty->natoms; ++a)
        for (int k = 0; k < 3; ++k) { double w = (double)r[a * 3 + k]; w = w * ty->atom_charge[a]; qd[k] = qd[k] + w; }
    if (ty->charge == 0) { out[0] = qd[0]; out[1] = qd[1]; out[2] = qd[2]; return; }
    double c[3];
    orc_center_of_mass(t, step, type, mol, c);
    for (int k = 0; k < 3; ++k) {
        const double ca = qd[k] / (double)ty->realcharge;
        out[k] = (double)ty->realcharge * (ca - c[k]);
    }
}

static double orc_rgy_sq(const pa_traj *t, int step, int type, int mol)
{
    const pa_type *ty = &t->types[type];
    const float *r = ty->xyz + ((size_t)step * ty->nmol + mol) * (size_t)ty->natoms * 3;
    double c[3], rgy = 0.0;
    orc_center_of_mass(t, step, type, mol, c);
    for (int a = 0; a < ty->natoms; ++a) {
        double d[3];
        for (int k = 0; k < 3; ++k) d[k] = (double)r[a * 3 + k] - c[k];
        const double d2 = (d[0] * d[0] + d[1] * d[1]) + d[2] * d[2];
        rgy = rgy + ty->atom_mass[a] * d2;
    }
    return rgy / ty->mass;
}

/* optimise_charge_arm (first time step only): returns the REAL(4) maxdist */
float orc_charge_arm_maxdist(const pa_traj *t, int ref_type, int normalise_clm)
{
    const int ty = ref_type - 1;
    float maxdist = 0.0f;
    for (int m = 0; m < t->types[ty].nmol; ++m) {
        double v[3];
        orc_charge_arm_vec(t, 0, ty, m, v);
        float chargearm = (float)sqrt((v[0] * v[0] + v[1] * v[1]) + v[2] * v[2]);
        if (normalise_clm) chargearm = (float)((double)chargearm / (t->types[ty].mass * orc_rgy_sq(t, 0, ty, m)));
        if (chargearm > maxdist) maxdist = chargearm;
    }
    if (normalise_clm) { maxdist = maxdist * 100000.0f; maxdist = (float)(int)ceilf(maxdist); maxdist = maxdist / 100000.0f; }
    else { maxdist = maxdist * 10.0f; maxdist = (float)(int)ceilf(maxdist); maxdist = maxdist / 10.0f; }
    if (maxdist < 0.0f) maxdist = 10.0f;
    return maxdist;
}

static void orc_fill_histogram_charge_arm(const pa_traj *t, const pa_job *job, int step, int64_t *hist, int64_t *within, int64_t *oob)
{
    const int ty = job->ref_type - 1;
    for (int m = 0; m < t->types[ty].nmol; ++m) {
        double link[3];
        orc_charge_arm_vec(t, step, ty, m, link);
        if (job->normalise_clm) {
            const double den = t->types[ty].mass * orc_rgy_sq(t, step, ty, m);
            for (int k = 0; k < 3; ++k) link[k] = link[k] / den;
        }
        const float a = (float)sqrt((link[0] * link[0] + link[1] * link[1]) + link[2] * link[2]);
        if (a < 1E-8f) { *oob += 1; continue; }
        for (int k = 0; k < 3; ++k) link[k] = link[k] / (double)a;
        const double dot = ((0.0 + link[0] * job->ref_vec[0]) + link[1] * job->ref_vec[1]) + link[2] * job->ref_vec[2];
        const float b = (float)acos(dot);
        const int ia = orc_plus1(orc_int_of_float(a / job->step_a)), ib = orc_plus1(orc_int_of_float(b / job->step_b));
        if (ia > 0 && ia <= job->bin_a && ib > 0 && ib <= job->bin_b) { *within += 1; hist[(size_t)(ib - 1) * job->bin_a + (ia - 1)] += 1; }
        else *oob += 1;
    }
}
