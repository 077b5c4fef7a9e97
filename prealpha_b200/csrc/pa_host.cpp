// pa_host.cpp -- host side of the drop-in boundary: everything around the hot loop that the
// reference does in REAL(4)/REAL(8) scalar Fortran and that must be reproduced bit for bit:
//   row S  job set-up (step sizes, reference vector, rotation matrix)    Distribution:353-377,656-675
//   class tables for exact binning on the device
//   row N  transfer_histogram                                             Distribution:897-977
//   row W  write_distribution_functions + list-directed REAL(4) output   Distribution:982-1185
// Compiled with -ffp-contract=off; no fast-math.  No compute fallback lives here: the O(N^2)
// pair loop exists only as CUDA kernels.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <climits>
#include <string>
#include <vector>
#include <algorithm>
#include "pa_internal.h"

namespace {

// SETTINGS:46: REAL(8) pi initialised from a default-kind literal = float32 pi widened.
const double kPi = (double)3.14159274101257324f;
const double kTwoPi = 2.0 * kPi;                       // SETTINGS:47
const double kDegrees = 57.295779513082320876798154814105;   // SETTINGS:41
const double kAvogadro = 6.02214076e23;                // SETTINGS:42
const double kECharge = (double)1.602176634E-19f;      // SETTINGS:43 (REAL4 literal)
const double kVacPerm = (double)8.8541878128E-12f;     // SETTINGS:44 (REAL4 literal)
const float kFoolsproof = 0.70710678118654752440f;     // 2.0**(-0.5), Distribution:353

inline int int_of_float_x86(float x)                   // cvttss2si
{
    if (!(x > -2147483904.0f && x < 2147483648.0f)) return INT_MIN;
    return (int)x;
}

void normalize3(double v[3])                            // SETTINGS:1052-1059
{
    const double len = std::sqrt((v[0] * v[0] + v[1] * v[1]) + v[2] * v[2]);
    v[0] = v[0] / len; v[1] = v[1] / len; v[2] = v[2] / len;
}

void cross3(const double a[3], const double b[3], double out[3])   // SETTINGS:1042-1050
{
    out[0] = a[1] * b[2] - a[2] * b[1];
    out[1] = a[2] * b[0] - a[0] * b[2];
    out[2] = a[0] * b[1] - a[1] * b[0];
}

// prepare_rotation(start, z) + make_rotation_matrix, Angles:51-115
void prepare_rotation_to_z(const double start[3], double R[9], int *aligned)
{
    const double tol = 0.0001;
    double a[3] = { start[0], start[1], start[2] }, b[3] = { 0.0, 0.0, 1.0 }, u[3];
    normalize3(a);
    normalize3(b);
    double cosa = 0.0;
    for (int k = 0; k < 3; ++k) cosa = cosa + a[k] * b[k];
    const double sina = std::sqrt(1 - cosa * cosa);
    const double angle = std::acos(cosa) * kDegrees;
    *aligned = 0;
    if (angle <= tol) {
        *aligned = 1;
        for (int k = 0; k < 9; ++k) R[k] = 0.0;
        R[0] = R[4] = R[8] = 1.0;
        return;
    }
    if ((180.0 - angle) <= tol) {
        double dummy[3] = { 0.0, 0.0, 0.0 };
        if (a[0] <= (1.0 - tol) && a[0] >= (-1.0 + tol)) dummy[0] = 1.0; else dummy[2] = 1.0;
        cross3(a, dummy, u);
    } else {
        cross3(a, b, u);
    }
    normalize3(u);
    const double omc = 1.0 - cosa;
    R[0] = cosa + (u[0] * u[0]) * omc;
    R[1] = u[0] * u[1] * omc - u[2] * sina;
    R[2] = u[0] * u[2] * omc + u[1] * sina;
    R[3] = u[0] * u[1] * omc + u[2] * sina;
    R[4] = cosa + (u[1] * u[1]) * omc;
    R[5] = u[1] * u[2] * omc - u[0] * sina;
    R[6] = u[0] * u[2] * omc - u[1] * sina;
    R[7] = u[2] * u[1] * omc + u[0] * sina;
    R[8] = cosa + (u[2] * u[2]) * omc;
}

inline unsigned long long bits_of(double x) { unsigned long long b; memcpy(&b, &x, 8); return b; }
inline double double_of(unsigned long long b) { double x; memcpy(&x, &b, 8); return x; }
// order-preserving map of all doubles onto uint64
inline unsigned long long key_of(double x)
{
    unsigned long long b = bits_of(x);
    return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}
inline double double_of_key(unsigned long long k)
{
    return double_of((k & 0x8000000000000000ull) ? (k & 0x7fffffffffffffffull) : ~k);
}

}  // namespace

// ---------------------------------------------------------------------------
// classes

// a-class of a pair whose separable minimum over the 27 images is m (fp64):
//   0..bin_a-1  bin index INT(SQRT(d2f)/step_a)      (Distribution:823,826)
//   bin_a       inside the cutoff but binpos_a out of range (counted out_of_bounds)
//   bin_a+1     not inside the cutoff (Distribution:783)
// with d2f = REAL4(min(m, maximum_distance_squared)) (Molecular:1438,1460; Distribution:780).
// For cdf jobs only the cutoff matters: class 0 = inside, 1 = beyond.
int pa::a_class_of(const pa_job &job, double max_dist_sq, double m)
{
    const double d2 = (m < max_dist_sq) ? m : max_dist_sq;
    const float d2f = (float)d2;
    const float m2 = job.maxdist * job.maxdist;
    const int na = job.mode == PA_MODE_PDF ? job.bin_a : 0;
    if (!(d2f < m2)) return na + 1;
    if (job.mode != PA_MODE_PDF) return 0;
    const float a = sqrtf(d2f);
    const int ia = int_of_float_x86(a / job.step_a);
    if (ia < 0 || ia >= na) return na;
    return ia;
}

// thr[k] = bit pattern of the smallest non-negative double m with a_class_of(m) >= k
// (k = 0..na+2; thr[0] = 0, thr[na+2] = never).  The class is monotone in m because min,
// the fp64->fp32 rounding, sqrtf, the division by step_a > 0 and INT() all are.
void pa::build_thr_a(const pa_job &job, double max_dist_sq, std::vector<unsigned long long> &thr)
{
    const int na = job.mode == PA_MODE_PDF ? job.bin_a : 0;
    const unsigned long long inf_bits = 0x7ff0000000000000ull, never = ~0ull;
    thr.assign((size_t)na + 3, never);
    thr[0] = 0;
    for (int k = 1; k <= na + 1; ++k) {
        if (pa::a_class_of(job, max_dist_sq, double_of(inf_bits)) < k) { thr[k] = never; continue; }
        unsigned long long lo = thr[k - 1] == never ? inf_bits : thr[k - 1], hi = inf_bits;   // answer in [lo, hi]
        if (pa::a_class_of(job, max_dist_sq, double_of(lo)) >= k) { thr[k] = lo; continue; }
        while (hi - lo > 1) {                     // invariant: class(lo) < k <= class(hi)
            const unsigned long long mid = lo + (hi - lo) / 2;
            if (pa::a_class_of(job, max_dist_sq, double_of(mid)) >= k) hi = mid; else lo = mid;
        }
        thr[k] = hi;
    }
}

// b-class INT(REAL4(ACOS(dot))/step_b) for dot in [-1,1] (Distribution:825,827); the caller
// treats |dot| > 1 / NaN as out of bounds (ACOS -> NaN -> INT() = INT_MIN).
int pa::b_class_of(const pa_job &job, double dot)
{
    const float b = (float)std::acos(dot);
    const int ib = int_of_float_x86(b / job.step_b);
    return ib < 0 ? INT_MAX : ib;
}

// thr[k], k = 1..bin_b: largest dot in [-1,1] whose b-class is >= k (class decreases with dot);
// -2.0 when no dot reaches class k.  thr[0] unused (= +2).
void pa::build_thr_b(const pa_job &job, std::vector<double> &thr)
{
    const int nb = job.bin_b;
    thr.assign((size_t)nb + 1, -2.0);
    thr[0] = 2.0;
    const unsigned long long kmin = key_of(-1.0), kmax = key_of(1.0);
    for (int k = 1; k <= nb; ++k) {
        if (pa::b_class_of(job, -1.0) < k) { thr[k] = -2.0; continue; }
        if (pa::b_class_of(job, 1.0) >= k) { thr[k] = 1.0; continue; }
        unsigned long long lo = kmin, hi = kmax;      // class(lo) >= k > class(hi)
        while (hi - lo > 1) {
            const unsigned long long mid = lo + (hi - lo) / 2;
            if (pa::b_class_of(job, double_of_key(mid)) >= k) lo = mid; else hi = mid;
        }
        thr[k] = double_of_key(lo);
    }
}

// charge_arm mode: a-class of the squared arm length S: INT(REAL4(SQRT(S))/step_a), clamped to
// [0, bin_a] (bin_a = out of range); lengths below 1E-8 are out of bounds before binning
// (Distribution:861-866), reported through tiny_bits = smallest S with REAL4(SQRT(S)) >= 1E-8.
static int charge_arm_class_of(const pa_job &job, double S)
{
    const float a = (float)std::sqrt(S);
    const int ia = int_of_float_x86(a / job.step_a);
    return (ia < 0 || ia >= job.bin_a) ? job.bin_a : ia;
}

void pa::build_thr_charge_arm(const pa_job &job, std::vector<unsigned long long> &thr, unsigned long long &tiny_bits)
{
    const int na = job.bin_a;
    const unsigned long long inf_bits = 0x7ff0000000000000ull, never = ~0ull;
    thr.assign((size_t)na + 3, never);
    thr[0] = 0;
    {   // smallest S with !(a < 1E-8)
        unsigned long long lo = 0, hi = inf_bits;      // a(lo) < 1E-8 <= a(hi)
        while (hi - lo > 1) {
            const unsigned long long mid = lo + (hi - lo) / 2;
            if ((float)std::sqrt(double_of(mid)) < 1E-8f) lo = mid; else hi = mid;
        }
        tiny_bits = hi;
    }
    // INT(a/step_a) is monotone in S for step_a > 0; for step_a <= 0 or NaN every length is out of range.
    if (!(job.step_a > 0.0f)) { for (int k = 1; k <= na; ++k) thr[k] = 0; return; }
    for (int k = 1; k <= na; ++k) {
        if (charge_arm_class_of(job, double_of(inf_bits)) < k) { thr[k] = never; continue; }
        unsigned long long lo = thr[k - 1] == never ? inf_bits : thr[k - 1], hi = inf_bits;
        if (charge_arm_class_of(job, double_of(lo)) >= k) { thr[k] = lo; continue; }
        while (hi - lo > 1) {
            const unsigned long long mid = lo + (hi - lo) / 2;
            if (charge_arm_class_of(job, double_of(mid)) >= k) hi = mid; else lo = mid;
        }
        thr[k] = hi;
    }
}

// The RDF fast kernel decides the a-bin only and sends pairs whose link vector is within
// |link_xy|^2 < 2^-30 A^2 of the z axis to the literal evaluation.  That is sufficient when the
// b-dimension has a single bin, the reference vector is +-z, the link vector is the
// centre-of-mass one, and 1e-15*maxdist^2 stays far below 2^-30.
bool pa::job_is_rdf_fast(const pa_job &job, const pa_traj &traj)
{
    (void)traj;
    if (job.mode != PA_MODE_PDF || job.bin_b != 1 || job.use_coc) return false;
    if (!(job.ref_vec[0] == 0.0 && job.ref_vec[1] == 0.0 && std::fabs(job.ref_vec[2]) == 1.0)) return false;
    if (!(job.maxdist > 0.0f && job.maxdist <= 400.0f)) return false;
    if (!(job.step_a >= 1.0e-6f) || !std::isfinite(job.step_a)) return false;
    return true;
}

// charge_arm helpers on the host (first time step only, for optimise_charge_arm, Distribution:1232-1269)
namespace {

void host_com(const pa_traj &t, const pa_type &ty, const float *r, double c[3])       // Molecular:2804-2833
{
    if (t.com_boost) { c[0] = (double)r[0]; c[1] = (double)r[1]; c[2] = (double)r[2]; return; }
    c[0] = c[1] = c[2] = 0.0;
    for (int a = 0; a < ty.natoms; ++a)
        for (int k = 0; k < 3; ++k) { double w = (double)r[a * 3 + k]; w = w * ty.atom_mass[a]; c[k] = c[k] + w; }
    for (int k = 0; k < 3; ++k) c[k] = c[k] / ty.mass;
}

void host_charge_arm(const pa_traj &t, const pa_type &ty, const float *r, double out[3])   // Molecular:2926-2937
{
    double qd[3] = { 0.0, 0.0, 0.0 };
    for (int a = 0; a < ty.natoms; ++a)
        for (int k = 0; k < 3; ++k) { double w = (double)r[a * 3 + k]; w = w * ty.atom_charge[a]; qd[k] = qd[k] + w; }
    if (ty.charge == 0) { out[0] = qd[0]; out[1] = qd[1]; out[2] = qd[2]; return; }
    double c[3];
    host_com(t, ty, r, c);
    for (int k = 0; k < 3; ++k) { const double ca = qd[k] / (double)ty.realcharge; out[k] = (double)ty.realcharge * (ca - c[k]); }
}

double host_rgy_sq(const pa_traj &t, const pa_type &ty, const float *r)                   // Molecular:945-972
{
    double c[3], rgy = 0.0;
    host_com(t, ty, r, c);
    for (int a = 0; a < ty.natoms; ++a) {
        const double dx = (double)r[a * 3] - c[0], dy = (double)r[a * 3 + 1] - c[1], dz = (double)r[a * 3 + 2] - c[2];
        const double d2 = (dx * dx + dy * dy) + dz * dz;
        rgy = rgy + ty.atom_mass[a] * d2;
    }
    return rgy / ty.mass;
}

float charge_arm_maxdist(const pa_traj &t, int type, bool clm)
{
    const pa_type &ty = t.types[type];
    float maxdist = 0.0f;
    for (int m = 0; m < ty.nmol; ++m) {
        const float *r = ty.xyz + (size_t)m * ty.natoms * 3;
        double v[3];
        host_charge_arm(t, ty, r, v);
        float arm = (float)std::sqrt((v[0] * v[0] + v[1] * v[1]) + v[2] * v[2]);
        if (clm) arm = (float)((double)arm / (ty.mass * host_rgy_sq(t, ty, r)));
        if (arm > maxdist) maxdist = arm;
    }
    const float scale = clm ? 100000.0f : 10.0f;                                          // :1256-1264
    maxdist = maxdist * scale;
    maxdist = (float)int_of_float_x86(std::ceil(maxdist));
    maxdist = maxdist / scale;
    if (maxdist < 0.0f) maxdist = 10.0f;
    return maxdist;
}

}  // namespace

// ---------------------------------------------------------------------------
extern "C" {

int pa_job_setup(const pa_traj *traj, const pa_job_request *rq, pa_job *job)
{
    if (!traj || !rq || !job) return PA_ERR_BAD_ARG;
    memset(job, 0, sizeof *job);
    if (rq->mode != PA_MODE_PDF && rq->mode != PA_MODE_CDF && rq->mode != PA_MODE_CHARGE_ARM) { pa::set_error("unknown operation mode (error 99)"); return PA_E_DIST_INPUT_FORMAT; }
    if (rq->ref_type < 1 || rq->ref_type > traj->ntypes || rq->obs_type > traj->ntypes) {
        pa::set_error("molecule type index out of range (error 33)");           // Distribution:415-427
        return PA_E_TYPE_INVALID;
    }
    job->mode = rq->mode;
    job->ref_type = rq->ref_type; job->obs_type = rq->obs_type;
    job->bin_a = std::min(std::max(rq->bin_a, 1), 1000);                          // Distribution:601-614
    job->bin_b = std::min(std::max(rq->bin_b, 1), 1000);
    job->sampling_interval = rq->sampling_interval;
    job->weigh_charge = rq->weigh_charge ? 1 : 0;
    job->use_coc = rq->use_coc ? 1 : 0;
    job->subtract_uniform = (rq->subtract_uniform && rq->mode != PA_MODE_CDF) ? 1 : 0;
    job->normalise_clm = (rq->normalise_clm && rq->mode == PA_MODE_CHARGE_ARM) ? 1 : 0;
    float maxdist = rq->maxdist;
    if (rq->mode == PA_MODE_CHARGE_ARM) {
        if (rq->maxdist_optimize) {
            const pa_type &ty = traj->types[rq->ref_type - 1];
            if (!ty.xyz || traj->nsteps < 1 || !ty.atom_mass || !ty.atom_charge) {
                pa::set_error("charge_arm maxdist_optimize needs the first time step, masses and charges"); return PA_ERR_BAD_ARG;
            }
            maxdist = charge_arm_maxdist(*traj, rq->ref_type - 1, job->normalise_clm != 0);
        }
    } else if (rq->maxdist_optimize) {                                                   // Distribution:526-534
        double lim = traj->box_hi[0] - traj->box_lo[0];
        for (int k = 1; k < 3; ++k) lim = std::min(lim, traj->box_hi[k] - traj->box_lo[k]);
        maxdist = (float)(lim / 2.0);
        if (maxdist < 0.0f) maxdist = 10.0f;
    }
    job->maxdist = maxdist;
    if (rq->mode == PA_MODE_CDF) {                                                // Distribution:362-365
        job->step_a = (maxdist * kFoolsproof) / (float)job->bin_a;
        job->step_b = ((2.0f * maxdist) * kFoolsproof) / (float)job->bin_b;
    } else {                                                                      // Distribution:375-377
        job->step_a = maxdist / (float)job->bin_a;
        job->step_b = (float)(kPi / (double)(float)job->bin_b);
    }
    job->shift = maxdist * kFoolsproof;                                           // Distribution:656
    for (int k = 0; k < 3; ++k) { job->ref_xyz[k] = rq->ref_xyz[k]; job->ref_vec[k] = (double)(float)rq->ref_xyz[k]; }
    if (rq->ref_xyz[0] == 0 && rq->ref_xyz[1] == 0 && rq->ref_xyz[2] == 0) {
        // Distribution:660-663: per-pair RAND() reference inside an OpenMP loop, not reproducible
        // in the reference itself.
        pa::set_error("randomised reference vector (0 0 0) is not supported");
        return PA_ERR_UNSUPPORTED;
    }
    if (job->sampling_interval < 1) { pa::set_error("sampling_interval < 1"); return PA_ERR_BAD_ARG; }
    if (rq->mode == PA_MODE_CHARGE_ARM && maxdist < 0.0f) { pa::set_error("charge_arm: negative maxdist"); return PA_ERR_UNSUPPORTED; }
    normalize3(job->ref_vec);                                                     // Distribution:665-667
    prepare_rotation_to_z(job->ref_vec, job->rot, &job->aligned);
    return 0;
}

int pa_transfer_histogram(const pa_traj *traj, const pa_job *job, const int64_t *hist, int64_t within_bin,
                          int64_t out_of_bounds, int32_t verbose, float *g, float *ideal_density_out)
{
    (void)out_of_bounds;
    if (!traj || !job || !hist || !g || !ideal_density_out) return PA_ERR_BAD_ARG;
    const int na = job->bin_a, nb = job->bin_b;
    const size_t n = (size_t)na * nb;
    int rc = 0;
    float within_real;
    if (within_bin < 0) {                                   // Distribution:904-906
        rc = PA_E_BIN_OVERFLOW;
        within_real = 0.0f;
        for (size_t i = 0; i < n; ++i) within_real = within_real + (float)hist[i];
    } else {
        within_real = (float)within_bin;
    }
    float total_volume;
    if (job->mode == PA_MODE_CDF) {
        const float s3 = (job->shift * job->shift) * job->shift;
        total_volume = (float)(kTwoPi * (double)s3);                              // :918
    } else {
        const float m3 = (job->maxdist * job->maxdist) * job->maxdist;
        total_volume = (float)(((double)(4.0f / 3.0f) * kPi) * (double)m3);       // :920
    }
    float rho = within_real / total_volume;                                       // :924
    for (size_t i = 0; i < n; ++i) g[i] = (float)hist[i] / rho;                   // :926
    if (verbose) {                                                                // :927-931
        const int denom = (traj->nsteps / job->sampling_interval) * traj->types[job->ref_type - 1].nmol;
        rho = rho / (float)denom;
    }
    for (int ia = 1; ia <= na; ++ia) {
        const float a = (float)ia * job->step_a, am = a - job->step_a;
        float chunk_area;
        if (job->mode == PA_MODE_CDF) {
            chunk_area = (float)(kPi * (double)((a * a) - (am * am)));            // :942
        } else {
            chunk_area = ((a * a) * a) - ((am * am) * am);                        // :946
            chunk_area = (float)((kTwoPi / (double)3.0f) * (double)chunk_area);   // :947
        }
        for (int ib = 1; ib <= nb; ++ib) {
            float chunk_volume;
            if (job->mode == PA_MODE_CDF) {
                chunk_volume = chunk_area * job->step_b;                          // :958
            } else {
                const float b = (float)ib * job->step_b;
                if (job->mode == PA_MODE_CHARGE_ARM) chunk_volume = cosf(b - job->step_b) - cosf(b);   // :967
                else chunk_volume = chunk_area * (cosf(b - job->step_b) - cosf(b));                   // :962
            }
            float &v = g[(size_t)(ib - 1) * na + (ia - 1)];
            v = v / chunk_volume;                                                 // :972
        }
    }
    if (job->mode == PA_MODE_CHARGE_ARM) {                                        // :976 (SUM in array element order)
        float sum = 0.0f;
        for (size_t i = 0; i < n; ++i) sum = sum + g[i];
        for (size_t i = 0; i < n; ++i) g[i] = g[i] / sum;
    }
    *ideal_density_out = rho;
    return rc;
}

// One list-directed REAL(4) item as libgfortran prints it: blank + G16.9E2, i.e. F format with
// 9 significant digits and 4 trailing blanks for 0.1 <= |x| < 1e9 (and for 0), otherwise
// d.ddddddddE+dd.  Verified against the reference binary's files (cat -A).
int pa_format_real4(float value, char *buf)
{
    if (std::isnan(value)) return snprintf(buf, 32, "%17s", "NaN");
    if (std::isinf(value)) return snprintf(buf, 32, "%17s", value > 0 ? "Infinity" : "-Infinity");
    const double m = std::fabs((double)value);
    char digits[48], body[48];
    int e10 = 0;
    bool use_f = (m == 0.0);
    if (!use_f) {
        snprintf(digits, sizeof digits, "%.8e", m);
        e10 = atoi(strchr(digits, 'e') + 1);
        use_f = (e10 >= -1 && e10 <= 8);
    }
    if (use_f) {
        snprintf(body, sizeof body, "%.*f", (m == 0.0) ? 8 : 8 - e10, (double)value);
        return snprintf(buf, 32, " %12s    ", body);
    }
    snprintf(body, sizeof body, "%.8E", (double)value);
    return snprintf(buf, 32, " %16s", body);
}

}  // extern "C"

namespace {

void write_reals(FILE *f, std::initializer_list<float> vals)
{
    char buf[32];
    for (float v : vals) { pa_format_real4(v, buf); fputs(buf, f); }
    fputc('\n', f);
}

float number_trapezoid(float g_a, float g_b, float r_a, float r_b)                // Distribution:1159-1169
{
    const float n = (r_b * g_a - r_a * g_b) / (r_b - r_a);
    const float m = (g_b - n) / r_b;
    const float rb2 = r_b * r_b, ra2 = r_a * r_a;
    const float upper = (m / 4.0f) * (rb2 * rb2) + (n / 3.0f) * (rb2 * r_b);
    const float lower = (m / 4.0f) * (ra2 * ra2) + (n / 3.0f) * (ra2 * r_a);
    return (float)(((double)4.0f * kPi) * (double)(upper - lower));
}

float energy_trapezoid(int ref_charge, float g_a, float g_b, float r_a, float r_b)   // Distribution:1171-1183
{
    const float n = (r_b * g_a - r_a * g_b) / (r_b - r_a);
    const float m = (g_b - n) / r_b;
    const float rb2 = r_b * r_b, ra2 = r_a * r_a;
    const float upper = (m / 3.0f) * (rb2 * r_b) + (n / 2.0f) * rb2;
    const float lower = (m / 3.0f) * (ra2 * r_a) + (n / 2.0f) * ra2;
    double v = ((double)(float)ref_charge * (kECharge * kECharge)) / ((double)2.0f * kVacPerm);
    v = v * (double)(upper - lower);
    v = v * kAvogadro;
    v = v * (double)1.0e10f;
    return (float)v;
}

std::string rtrim(std::string s) { while (!s.empty() && s.back() == ' ') s.pop_back(); return s; }

}  // namespace

extern "C" int pa_write_distribution(const pa_traj *traj, const pa_job *job, int32_t reference_number,
                                     const char *path_prefix, const char *input_file_name, float *g, float rho)
{
    if (!traj || !job || !path_prefix || !input_file_name || !g) return PA_ERR_BAD_ARG;
    const int na = job->bin_a, nb = job->bin_b;
    const int ref_charge = traj->types[job->ref_type - 1].charge;
    const bool energy = (ref_charge != 0) && job->weigh_charge;                   // :992
    char tmp[256];
    std::string base = std::string(path_prefix) + "reference_" + std::to_string(reference_number) + "_";   // :1004
    snprintf(tmp, sizeof tmp, "reference vector %+d %+d %+d, ", job->ref_xyz[0], job->ref_xyz[1], job->ref_xyz[2]);
    std::string header = tmp;
    const bool arm = job->mode == PA_MODE_CHARGE_ARM;
    if (arm) {                                                                    // :1008-1024
        base += std::string(job->normalise_clm ? "CLM-normalised_" : "") + (ref_charge != 0 ? "charge_arm" : "dipole_moment") +
                "_type_" + std::to_string(job->ref_type);
        header = rtrim(header) + (ref_charge != 0 ? " charge arm" : " dipole moment") + " for molecule type " +
                 std::to_string(job->ref_type) + ",";
    } else if (job->obs_type < 1) {                                               // :1025-1036
        base += "all_around_" + std::to_string(job->ref_type);
        header = rtrim(header) + " all molecules around type " + std::to_string(job->ref_type) + ", ";
    } else {
        base += "type_" + std::to_string(job->obs_type) + "_around_" + std::to_string(job->ref_type);
        header = rtrim(header) + " all molecules of type " + std::to_string(job->obs_type) + " around type " +
                 std::to_string(job->ref_type) + ", ";
    }
    const std::string xc = job->weigh_charge ? "_xcharge" : "";
    std::string main_name, unitline;
    float shift_b;
    if (job->mode == PA_MODE_CDF) {                                               // :1039-1047
        shift_b = job->maxdist * kFoolsproof;
        main_name = base + xc + "_cdf";
        header = rtrim(header) + " cylindrical distribution function";
        unitline = "a(perpendicular_distance) b(collinear_distance) g(a,b)";
    } else if (arm) {                                                             // :1112-1139
        shift_b = 0.0f;
        const std::string trace_name = base + "_rdf";
        main_name = base + (job->subtract_uniform ? "_-trace" : "") + "_pdf";
        FILE *ft = fopen(trace_name.c_str(), "w");
        if (!ft) { pa::set_error("cannot open output file " + trace_name + " (error 26)"); return 26; }
        fprintf(ft, "%s radial distribution function (trace of pdf)\ndistance g(r)\n", rtrim(header).c_str());
        for (int ia = 1; ia <= na; ++ia) {
            float trace = 0.0f;
            for (int ib = 0; ib < nb; ++ib) trace = trace + g[(size_t)ib * na + (ia - 1)];
            trace = trace / (float)nb;
            float a = (float)ia * job->step_a;
            a = a - 0.5f * job->step_a;
            write_reals(ft, { a, trace });
            if (job->subtract_uniform)
                for (int ib = 0; ib < nb; ++ib) g[(size_t)ib * na + (ia - 1)] = g[(size_t)ib * na + (ia - 1)] - trace;
        }
        fclose(ft);
        unitline = "a(distance) b(polar_angle) g(a,b)";
        header = rtrim(header) + " polar distribution function";
    } else {                                                                      // :1048-1111
        shift_b = 0.0f;
        const std::string trace_name = base + "_rdf" + xc;
        main_name = base + xc + (job->subtract_uniform ? "_-trace" : "") + "_pdf";
        FILE *ft = fopen(trace_name.c_str(), "w");
        FILE *fn = fopen((trace_name + "_numberintegral").c_str(), "w");
        FILE *fe = energy ? fopen((trace_name + "_energyintegral").c_str(), "w") : nullptr;
        if (!ft || !fn || (energy && !fe)) {
            if (ft) fclose(ft); if (fn) fclose(fn); if (fe) fclose(fe);
            pa::set_error("cannot open output file " + trace_name + " (error 26)");
            return 26;
        }
        const std::string h = rtrim(header);
        fprintf(ft, "%s radial distribution function (trace of pdf)\ndistance g(r)\n", h.c_str());
        fprintf(fn, "%s number integral INT(4*pi*r^2*rho*g(r))dR\ndistance n(r)\n", h.c_str());
        float number_integral = 0.0f, energy_integral = 0.0f, previous = 0.0f;
        if (fe) {
            fprintf(fe, "%s (Coulomb) energy integral INT(z*e^2*r*rho*g(r)/(2*epsilon0))dR\ndistance E(r)[J/mol]\n", h.c_str());
            write_reals(fe, { 0.0f, 0.0f });
        }
        write_reals(fn, { 0.0f, 0.0f });
        for (int ia = 1; ia <= na; ++ia) {
            float trace = 0.0f;
            for (int ib = 0; ib < nb; ++ib) trace = trace + g[(size_t)ib * na + (ia - 1)];
            trace = trace / (float)nb;                                            // :1088
            float a = (float)ia * job->step_a;
            number_integral = number_integral + number_trapezoid(previous, trace, a - job->step_a, a) * rho;
            write_reals(fn, { a, number_integral });
            if (fe) {
                energy_integral = energy_integral + energy_trapezoid(ref_charge, previous, trace, a - job->step_a, a) * rho;
                write_reals(fe, { a, energy_integral });
            }
            a = a - 0.5f * job->step_a;
            write_reals(ft, { a, trace });
            previous = trace;
            if (job->subtract_uniform)
                for (int ib = 0; ib < nb; ++ib) g[(size_t)ib * na + (ia - 1)] = g[(size_t)ib * na + (ia - 1)] - trace;
        }
        fclose(ft); fclose(fn); if (fe) fclose(fe);
        unitline = "a(distance) b(polar_angle) g(a,b)";
        header = rtrim(header) + " polar distribution function";
    }
    if (job->weigh_charge && !arm) header = rtrim(header) + ", weighed by charge";   // :1143
    FILE *f = fopen(main_name.c_str(), "w");
    if (!f) { pa::set_error("cannot open output file " + main_name + " (error 26)"); return 26; }
    fprintf(f, "Output in this file is based on the input file '%s':\n%s\n%s\n", input_file_name, rtrim(header).c_str(), unitline.c_str());
    for (int ia = 1; ia <= na; ++ia)
        for (int ib = 1; ib <= nb; ++ib) {                                        // :1147-1154
            const float a = ((float)ia * job->step_a) - 0.5f * job->step_a;
            float b = (((float)ib * job->step_b) - 0.5f * job->step_b) - shift_b;
            if (job->mode != PA_MODE_CDF) b = (float)((double)b * kDegrees);
            write_reals(f, { a, b, g[(size_t)(ib - 1) * na + (ia - 1)] });
        }
    fclose(f);
    return 0;
}

// ---- introspection for the tests: the class tables the kernels bin with (host-built) ----
extern "C" int pa_debug_thr_a(const pa_job *job, double max_dist_sq, unsigned long long *out, int32_t n)
{
    std::vector<unsigned long long> t;
    pa::build_thr_a(*job, max_dist_sq, t);
    if (!out) return (int)t.size();
    for (int i = 0; i < n && i < (int)t.size(); ++i) out[i] = t[i];
    return (int)t.size();
}
extern "C" int pa_debug_thr_b(const pa_job *job, double *out, int32_t n)
{
    std::vector<double> t;
    pa::build_thr_b(*job, t);
    if (!out) return (int)t.size();
    for (int i = 0; i < n && i < (int)t.size(); ++i) out[i] = t[i];
    return (int)t.size();
}
extern "C" int pa_debug_a_class(const pa_job *job, double max_dist_sq, double m) { return pa::a_class_of(*job, max_dist_sq, m); }
extern "C" int pa_debug_b_class(const pa_job *job, double dot) { return pa::b_class_of(*job, dot); }
extern "C" int pa_debug_is_rdf_fast(const pa_job *job) { pa_traj t{}; return pa::job_is_rdf_fast(*job, t) ? 1 : 0; }
