// pa_distance.cu -- row X: the `distance` analysis of Module_Distances on the GPU.
//
// Reference: compute_intermolecular_distances / compute_intramolecular_distances
// (Distances:976-1091 / 856-974) are serial five-deep loops
//   step -> subjob -> reference atom index -> reference molecule -> observed atom index -> observed molecule
// around give_smallest_atom_distance_squared (Molecular:1380-1427: both atoms are always wrapped,
// then the 27-image loop).  Here one warp owns one reference atom instance (atom index x molecule)
// and its lanes stride over the observed atom instances; per instance the closest squared distance
// (float32, exact) and the fp64 partial sums are written out, and the host finishes in the
// reference's loop order (sum over reference atoms, averages, stdev pass from the same per-atom
// values, FFC powers -- Distances:1188-1235).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
#include <cmath>
#include <vector>
#include <string>
#include "pa_internal.h"

namespace {

#define PD_CK(call)                                                                              \
    do {                                                                                         \
        cudaError_t e__ = (call);                                                                \
        if (e__ != cudaSuccess) {                                                                \
            pa::set_error(std::string(#call) + ": " + cudaGetErrorString(e__));                  \
            rc = (e__ == cudaErrorNoDevice || e__ == cudaErrorInsufficientDriver ||              \
                  e__ == cudaErrorNoKernelImageForDevice) ? PA_ERR_NO_DEVICE : PA_ERR_CUDA;      \
            goto done;                                                                           \
        }                                                                                        \
    } while (0)

struct PdType { const float *xyz; int natoms, nmol; };          // one frame: [mol][atom][3]
struct PdEntry { int type, atom; };                               // 0-based
struct PdOut {                                                    // per reference atom instance
    float closest_d2; int valid;
    double sum_exp_d, sum_exp, sum_ffc, sum_ffcexp;
    long long nffc;
};
struct PdArgs {
    const PdType *types;
    const PdEntry *ref, *obs;
    const int *ref_first;       // [n_ref+1] prefix of reference instances (entry e owns nmol(type) instances)
    int n_ref, n_obs, n_inst;
    int intermolecular, calc_exp, calc_ffc;
    float maxdist_sq, exponent;
    double lo[3], hi[3], L[3], mds;
    PdOut *out;
};

__device__ __forceinline__ void pd_load_wrapped(const PdArgs &a, int type, int mol, int atom, double p[3])
{
    const PdType ty = a.types[type];
    const float *r = ty.xyz + ((size_t)mol * ty.natoms + atom) * 3;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        double v = (double)r[k];
        for (int it = 0; it < 1000; ++it) {                       // wrap_vector, Molecular:1560-1586
            if (v < a.lo[k]) v = __dadd_rn(v, a.L[k]);
            else if (v > a.hi[k]) v = __dsub_rn(v, a.L[k]);
            else break;
        }
        p[k] = v;
    }
}

// REAL(4) distance_squared of Distances:1015: min over the 27 images, separable (see pa_kernels.cu)
__device__ __forceinline__ float pd_distance_squared(const PdArgs &a, const double p1[3], const double p2[3])
{
    double s[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const double dm = __dsub_rn(__dadd_rn(p2[k], -a.L[k]), p1[k]);
        const double d0 = __dsub_rn(p2[k], p1[k]);
        const double dp = __dsub_rn(__dadd_rn(p2[k], a.L[k]), p1[k]);
        double m = fabs(d0) < fabs(dm) ? d0 : dm;
        m = fabs(dp) < fabs(m) ? dp : m;
        s[k] = __dmul_rn(m, m);
    }
    const double m = __dadd_rn(__dadd_rn(s[0], s[1]), s[2]);
    return __double2float_rn(m < a.mds ? m : a.mds);
}

struct PdAcc {
    float closest; int valid; double se_d, se, sf, sfe; long long nf;
};

__device__ __forceinline__ void pd_pair(const PdArgs &a, float d2, PdAcc &c)
{
    if (!(d2 < a.maxdist_sq)) return;
    c.valid = 1;
    if (d2 < c.closest) c.closest = d2;
    double ef = 0.0;
    float distance = 0.f;
    if (a.calc_exp || (a.calc_ffc && a.intermolecular)) distance = __fsqrt_rn(d2);
    if (a.calc_exp) {
        // exp of a REAL(4): the reference calls expf; correctly rounded here via fp64 exp
        ef = (double)__double2float_rn(exp((double)(-__fmul_rn(a.exponent, distance))));
        c.se_d = __dadd_rn(c.se_d, __dmul_rn(ef, (double)distance));
        c.se = __dadd_rn(c.se, ef);
    }
    if (a.calc_ffc) {
        float ffc;
        if (a.intermolecular) ffc = __fdiv_rn(1.0f, __fmul_rn(__fmul_rn(distance, distance), distance));   // 1/r^3
        else ffc = __fdiv_rn(1.0f, __fmul_rn(__fmul_rn(d2, d2), d2));                                      // 1/(r^2)^3
        c.sf = __dadd_rn(c.sf, (double)ffc);
        c.nf += 1;
        if (a.calc_exp) c.sfe = __dadd_rn(c.sfe, __dmul_rn(ef, (double)ffc));
    }
}

__global__ void __launch_bounds__(128) pa_distance_kernel(PdArgs a)
{
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= a.n_inst) return;
    // reference instance -> (entry, molecule)
    int e1 = 0;
    while (e1 + 1 < a.n_ref && warp >= a.ref_first[e1 + 1]) ++e1;
    const int m1 = warp - a.ref_first[e1];
    const PdEntry r = a.ref[e1];
    double p1[3];
    pd_load_wrapped(a, r.type, m1, r.atom, p1);
    PdAcc c; c.closest = a.maxdist_sq; c.valid = 0; c.se_d = c.se = c.sf = c.sfe = 0.0; c.nf = 0;
    if (a.intermolecular) {
        for (int e2 = 0; e2 < a.n_obs; ++e2) {
            const PdEntry o = a.obs[e2];
            const int nmol = a.types[o.type].nmol;
            for (int m2 = lane; m2 < nmol; m2 += 32) {
                if (m2 == m1 && o.type == r.type) continue;                  // Distances:1014
                double p2[3];
                pd_load_wrapped(a, o.type, m2, o.atom, p2);
                pd_pair(a, pd_distance_squared(a, p1, p2), c);
            }
        }
    } else {
        for (int e2 = lane; e2 < a.n_obs; e2 += 32) {
            const PdEntry o = a.obs[e2];
            if (o.type != r.type || o.atom == r.atom) continue;              // Distances:891-893
            double p2[3];
            pd_load_wrapped(a, o.type, m1, o.atom, p2);
            pd_pair(a, pd_distance_squared(a, p1, p2), c);
        }
    }
    // fixed-order warp reduction (deterministic)
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const float oc = __shfl_down_sync(0xffffffffu, c.closest, off);
        const int ov = __shfl_down_sync(0xffffffffu, c.valid, off);
        c.closest = fminf(c.closest, oc); c.valid |= ov;
        c.se_d = __dadd_rn(c.se_d, __shfl_down_sync(0xffffffffu, c.se_d, off));
        c.se = __dadd_rn(c.se, __shfl_down_sync(0xffffffffu, c.se, off));
        c.sf = __dadd_rn(c.sf, __shfl_down_sync(0xffffffffu, c.sf, off));
        c.sfe = __dadd_rn(c.sfe, __shfl_down_sync(0xffffffffu, c.sfe, off));
        c.nf += __shfl_down_sync(0xffffffffu, c.nf, off);
    }
    if (lane == 0) {
        PdOut o; o.closest_d2 = c.closest; o.valid = c.valid; o.sum_exp_d = c.se_d; o.sum_exp = c.se; o.sum_ffc = c.sf; o.sum_ffcexp = c.sfe; o.nffc = c.nf;
        a.out[warp] = o;
    }
}

}  // namespace

extern "C" int pa_distance_subjobs(const pa_traj *traj, const pa_dist_job *job, const pa_exec *exec, pa_dist_result *results)
{
    if (!traj || !job || !results || job->nsubjobs < 1 || !job->subjobs || job->sampling_interval < 1) { pa::set_error("bad argument"); return PA_ERR_BAD_ARG; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) { cudaGetLastError(); pa::set_error("no CUDA device (this path has no CPU fallback)"); return PA_ERR_NO_DEVICE; }
    const int device = exec ? exec->device : 0;
    if (cudaSetDevice(device) != cudaSuccess || pa_kernels_selfcheck() != 0) { cudaGetLastError(); pa::set_error("no sm_100a kernel image for this device"); return PA_ERR_NO_DEVICE; }
    const int nsteps = std::min(job->nsteps, traj->nsteps);
    for (int k = 0; k < job->nsubjobs; ++k) {
        const pa_dist_subjob &sj = job->subjobs[k];
        if (sj.n_ref < 1 || sj.n_obs < 1 || !sj.ref || !sj.obs) { pa::set_error("empty subjob"); return PA_ERR_BAD_ARG; }
        for (int e = 0; e < sj.n_ref + sj.n_obs; ++e) {
            const int32_t *p = e < sj.n_ref ? sj.ref + 2 * e : sj.obs + 2 * (e - sj.n_ref);
            if (p[0] < 1 || p[0] > traj->ntypes) { pa::set_error("invalid molecule type in distance subjob (error 110)"); return 110; }
            if (p[1] < 1 || p[1] > traj->types[p[0] - 1].natoms) { pa::set_error("invalid atom index in distance subjob (error 111)"); return 111; }
        }
    }
    int rc = 0;
    const int nt = traj->ntypes;
    std::vector<float *> d_xyz(nt, nullptr);
    PdType *d_types = nullptr;
    std::vector<PdEntry *> d_ref(job->nsubjobs, nullptr), d_obs(job->nsubjobs, nullptr);
    std::vector<int *> d_first(job->nsubjobs, nullptr);
    std::vector<std::vector<int>> first(job->nsubjobs);
    std::vector<PdOut *> d_out(job->nsubjobs, nullptr);
    std::vector<int> ninst(job->nsubjobs, 0);
    // per (step, subjob) per-instance records, kept for the stdev pass
    std::vector<std::vector<std::vector<PdOut>>> recs;
    std::vector<PdType> h_types(nt);
    {
        for (int t = 0; t < nt; ++t) {
            const size_t n = (size_t)traj->types[t].nmol * traj->types[t].natoms * 3;
            PD_CK(cudaMalloc(&d_xyz[t], sizeof(float) * n));
            h_types[t].xyz = d_xyz[t]; h_types[t].natoms = traj->types[t].natoms; h_types[t].nmol = traj->types[t].nmol;
        }
        PD_CK(cudaMalloc(&d_types, sizeof(PdType) * nt));
        PD_CK(cudaMemcpy(d_types, h_types.data(), sizeof(PdType) * nt, cudaMemcpyHostToDevice));
        for (int k = 0; k < job->nsubjobs; ++k) {
            const pa_dist_subjob &sj = job->subjobs[k];
            std::vector<PdEntry> hr(sj.n_ref), ho(sj.n_obs);
            first[k].assign(sj.n_ref + 1, 0);
            for (int e = 0; e < sj.n_ref; ++e) {
                hr[e].type = sj.ref[2 * e] - 1; hr[e].atom = sj.ref[2 * e + 1] - 1;
                first[k][e + 1] = first[k][e] + traj->types[hr[e].type].nmol;
            }
            for (int e = 0; e < sj.n_obs; ++e) { ho[e].type = sj.obs[2 * e] - 1; ho[e].atom = sj.obs[2 * e + 1] - 1; }
            ninst[k] = first[k][sj.n_ref];
            PD_CK(cudaMalloc(&d_ref[k], sizeof(PdEntry) * sj.n_ref));
            PD_CK(cudaMalloc(&d_obs[k], sizeof(PdEntry) * sj.n_obs));
            PD_CK(cudaMalloc(&d_first[k], sizeof(int) * (sj.n_ref + 1)));
            PD_CK(cudaMalloc(&d_out[k], sizeof(PdOut) * ninst[k]));
            PD_CK(cudaMemcpy(d_ref[k], hr.data(), sizeof(PdEntry) * sj.n_ref, cudaMemcpyHostToDevice));
            PD_CK(cudaMemcpy(d_obs[k], ho.data(), sizeof(PdEntry) * sj.n_obs, cudaMemcpyHostToDevice));
            PD_CK(cudaMemcpy(d_first[k], first[k].data(), sizeof(int) * (sj.n_ref + 1), cudaMemcpyHostToDevice));
        }
        for (int step = 0; step < nsteps; step += job->sampling_interval) {
            for (int t = 0; t < nt; ++t) {
                const size_t n = (size_t)traj->types[t].nmol * traj->types[t].natoms * 3;
                PD_CK(cudaMemcpy(d_xyz[t], traj->types[t].xyz + (size_t)step * n, sizeof(float) * n, cudaMemcpyHostToDevice));
            }
            recs.emplace_back(job->nsubjobs);
            for (int k = 0; k < job->nsubjobs; ++k) {
                PdArgs a{};
                a.types = d_types; a.ref = d_ref[k]; a.obs = d_obs[k]; a.ref_first = d_first[k];
                a.n_ref = job->subjobs[k].n_ref; a.n_obs = job->subjobs[k].n_obs; a.n_inst = ninst[k];
                a.intermolecular = job->intermolecular; a.calc_exp = job->calc_exp; a.calc_ffc = job->calc_ffc;
                a.maxdist_sq = job->maxdist * job->maxdist; a.exponent = job->exponent;
                for (int c = 0; c < 3; ++c) { a.lo[c] = traj->box_lo[c]; a.hi[c] = traj->box_hi[c]; a.L[c] = traj->box_hi[c] - traj->box_lo[c]; }
                a.mds = traj->max_dist_sq;
                a.out = d_out[k];
                const int blocks = (ninst[k] + 3) / 4;
                pa_distance_kernel<<<blocks, 128>>>(a);
                PD_CK(cudaGetLastError());
                recs.back()[k].resize(ninst[k]);
                PD_CK(cudaMemcpy(recs.back()[k].data(), d_out[k], sizeof(PdOut) * ninst[k], cudaMemcpyDeviceToHost));
            }
        }
    }
    // host: the reference's accumulation over reference atoms, steps, and its post-processing
    {
        int nav = (job->nsteps - 1 + job->sampling_interval) / job->sampling_interval;
        if (nav < 0) nav = 0;
        const double navf = (double)(float)nav;
        memset(results, 0, sizeof(pa_dist_result) * job->nsubjobs);
        for (int pass = 0; pass < (job->calc_stdev ? 2 : 1); ++pass) {
            for (size_t s = 0; s < recs.size(); ++s)
                for (int k = 0; k < job->nsubjobs; ++k) {
                    pa_dist_result &r = results[k];
                    int localweight = 0; long long lw_ffc = 0;
                    double local_closest = 0, local_exp = 0, w_exp = 0, local_exp_stdev = 0, local_ffc = 0, local_ffcexp = 0;
                    for (const PdOut &o : recs[s][k]) {
                        if (o.valid) {
                            ++localweight;
                            const double cd = (double)sqrtf(o.closest_d2);
                            if (pass == 1) {
                                local_closest = local_closest + (cd - r.closest_average) * (cd - r.closest_average);
                                if (job->calc_exp) { const double de = o.sum_exp_d / o.sum_exp - r.exponential_average; local_exp_stdev = local_exp_stdev + de * de; }
                            } else local_closest = local_closest + cd;
                        } else r.missed_neighbours += 1;
                        local_exp = local_exp + o.sum_exp_d; w_exp = w_exp + o.sum_exp;
                        local_ffc = local_ffc + o.sum_ffc; local_ffcexp = local_ffcexp + o.sum_ffcexp; lw_ffc += o.nffc;
                    }
                    const double lw = (double)(float)localweight;
                    if (pass == 1) {
                        r.closest_stdev += local_closest / lw;
                        if (job->calc_exp) r.exponential_stdev += local_exp_stdev / lw;
                    } else {
                        r.closest_average += local_closest / lw;
                        if (job->calc_exp) { r.exponential_average += local_exp; r.exponential_weight += w_exp; }
                        if (job->calc_ffc) { r.ffc_average += local_ffc; if (job->calc_exp) r.ffcexp_average += local_ffcexp; r.ffc_weight += (double)(float)lw_ffc; }
                    }
                }
            for (int k = 0; k < job->nsubjobs; ++k) {
                pa_dist_result &r = results[k];
                if (pass == 0) {
                    r.closest_average = r.closest_average / navf;
                    if (job->calc_exp) r.exponential_average = r.exponential_average / r.exponential_weight;
                    if (job->calc_ffc) {
                        r.ffc_average = r.ffc_average / r.ffc_weight;
                        if (job->calc_exp) r.ffcexp_average = r.ffcexp_average / r.exponential_weight;
                        const double ex = job->intermolecular ? (double)(-1.0f / 3.0f) : (double)(-1.0f / 6.0f);
                        r.ffc_average = std::pow(r.ffc_average, ex);
                        if (job->calc_exp) r.ffcexp_average = std::pow(r.ffcexp_average, ex);
                    }
                } else {
                    r.closest_stdev = std::sqrt(r.closest_stdev / navf);
                    if (job->calc_exp) r.exponential_stdev = std::sqrt(r.exponential_stdev / navf);
                }
            }
        }
    }
done:
    for (float *p : d_xyz) cudaFree(p);
    cudaFree(d_types);
    for (int k = 0; k < job->nsubjobs; ++k) { cudaFree(d_ref[k]); cudaFree(d_obs[k]); cudaFree(d_first[k]); cudaFree(d_out[k]); }
    return rc;
}
