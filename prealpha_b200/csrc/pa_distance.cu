// pa_distance.cu -- row X: the `distance` analysis of Module_Distances on the GPU.
//
// Reference: compute_intermolecular_distances / compute_intramolecular_distances
// (Distances:976-1091 / 856-974) are serial five-deep loops
//   step -> subjob -> reference atom index -> reference molecule -> observed atom index -> observed molecule
// around give_smallest_atom_distance_squared (Molecular:1380-1427: both atoms are always wrapped,
// then the 27-image loop).  Here one warp owns one reference atom instance (atom index x molecule)
// and its lanes take 32 consecutive observed atom instances of the reference's loop order at a time;
// the squared distances below the cutoff come back in loop order and the host replays the reference's
// sequential accumulation on them (sums over pairs, over reference atoms, averages, stdev pass, FFC
// powers -- Distances:1188-1235), which makes every REAL(8) result bit-identical to the reference's.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include <string>
#include <algorithm>
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
struct PdOut;                                                     // (unused: the sums are formed on the host, in the reference's order)
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

// One warp per reference atom instance; its lanes take 32 consecutive observed instances of the reference's loop order
// (observed atom index, then observed molecule) at a time.  The squared distances below the cutoff are the only thing the
// sums of the reference depend on, so the kernel hands them back IN LOOP ORDER (ballot + popc compaction keeps the order
// inside a group of 32) and the host replays the reference's sequential fp64 accumulation on them -- bit for bit, because
// it is the same chain of additions with the same libm expf.  Pass 1 (WRITE = false) counts, pass 2 writes at the offsets
// the host derived from the counts.
template <bool WRITE>
__global__ void __launch_bounds__(128) pa_distance_kernel(PdArgs a, int inst0, int ninst, int *count, const long long *offset, float *list)
{
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= ninst) return;
    const int inst = inst0 + w;
    const unsigned int lane_lt = (1u << lane) - 1u;
    // reference instance -> (entry, molecule)
    int e1 = 0;
    while (e1 + 1 < a.n_ref && inst >= a.ref_first[e1 + 1]) ++e1;
    const int m1 = inst - a.ref_first[e1];
    const PdEntry r = a.ref[e1];
    double p1[3];
    pd_load_wrapped(a, r.type, m1, r.atom, p1);
    int cnt = 0;
    float *out = WRITE ? list + offset[w] : nullptr;
    if (a.intermolecular) {
        for (int e2 = 0; e2 < a.n_obs; ++e2) {
            const PdEntry o = a.obs[e2];
            const int nmol = a.types[o.type].nmol;
            for (int base = 0; base < nmol; base += 32) {
                const int m2 = base + lane;
                bool valid = false;
                float d2 = 0.f;
                if (m2 < nmol && !(m2 == m1 && o.type == r.type)) {          // Distances:1014
                    double p2[3];
                    pd_load_wrapped(a, o.type, m2, o.atom, p2);
                    d2 = pd_distance_squared(a, p1, p2);
                    valid = d2 < a.maxdist_sq;
                }
                const unsigned int m = __ballot_sync(0xffffffffu, valid);
                if (WRITE && valid) out[cnt + __popc(m & lane_lt)] = d2;
                cnt += __popc(m);
            }
        }
    } else {
        for (int base = 0; base < a.n_obs; base += 32) {
            const int e2 = base + lane;
            bool valid = false;
            float d2 = 0.f;
            if (e2 < a.n_obs) {
                const PdEntry o = a.obs[e2];
                if (o.type == r.type && o.atom != r.atom) {                   // Distances:891-893
                    double p2[3];
                    pd_load_wrapped(a, o.type, m1, o.atom, p2);
                    d2 = pd_distance_squared(a, p1, p2);
                    valid = d2 < a.maxdist_sq;
                }
            }
            const unsigned int m = __ballot_sync(0xffffffffu, valid);
            if (WRITE && valid) out[cnt + __popc(m & lane_lt)] = d2;
            cnt += __popc(m);
        }
    }
    if (!WRITE && lane == 0) count[w] = cnt;
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
    std::vector<int> ninst(job->nsubjobs, 0);
    int *d_count = nullptr;
    long long *d_offset = nullptr;
    float *d_list = nullptr;
    size_t list_cap = 0;
    int max_inst = 0;
    // per (step, subjob): for every reference atom instance, in loop order, the squared distances below the cutoff in
    // loop order (lists[...] concatenated, counts[...] per instance)
    struct StepRec { std::vector<int> counts; std::vector<float> d2; };
    std::vector<std::vector<StepRec>> recs;
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
            max_inst = std::max(max_inst, ninst[k]);
            PD_CK(cudaMalloc(&d_ref[k], sizeof(PdEntry) * sj.n_ref));
            PD_CK(cudaMalloc(&d_obs[k], sizeof(PdEntry) * sj.n_obs));
            PD_CK(cudaMalloc(&d_first[k], sizeof(int) * (sj.n_ref + 1)));
            PD_CK(cudaMemcpy(d_ref[k], hr.data(), sizeof(PdEntry) * sj.n_ref, cudaMemcpyHostToDevice));
            PD_CK(cudaMemcpy(d_obs[k], ho.data(), sizeof(PdEntry) * sj.n_obs, cudaMemcpyHostToDevice));
            PD_CK(cudaMemcpy(d_first[k], first[k].data(), sizeof(int) * (sj.n_ref + 1), cudaMemcpyHostToDevice));
        }
        PD_CK(cudaMalloc(&d_count, sizeof(int) * std::max(max_inst, 1)));
        PD_CK(cudaMalloc(&d_offset, sizeof(long long) * std::max(max_inst, 1)));
        std::vector<long long> h_off(std::max(max_inst, 1));
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
                a.out = nullptr;
                StepRec &rec = recs.back()[k];
                rec.counts.resize(ninst[k]);
                const int blocks = (ninst[k] + 3) / 4;
                pa_distance_kernel<false><<<blocks, 128>>>(a, 0, ninst[k], d_count, nullptr, nullptr);
                PD_CK(cudaGetLastError());
                PD_CK(cudaMemcpy(rec.counts.data(), d_count, sizeof(int) * ninst[k], cudaMemcpyDeviceToHost));
                long long total = 0;
                for (int i = 0; i < ninst[k]; ++i) total += rec.counts[i];
                rec.d2.resize((size_t)total);
                // pass 2 in pieces of at most 64 Mi list entries (256 MB)
                const long long piece = (long long)64 << 20;
                int i0 = 0;
                long long done = 0;
                while (i0 < ninst[k]) {
                    int i1 = i0;
                    long long n = 0;
                    while (i1 < ninst[k] && (i1 == i0 || n + rec.counts[i1] <= piece)) { h_off[i1 - i0] = n; n += rec.counts[i1]; ++i1; }
                    if ((size_t)n > list_cap) {
                        if (d_list) cudaFree(d_list);
                        d_list = nullptr;
                        list_cap = (size_t)std::max<long long>(n, 1 << 20);
                        PD_CK(cudaMalloc(&d_list, sizeof(float) * list_cap));
                    }
                    if (n > 0) {
                        PD_CK(cudaMemcpy(d_offset, h_off.data(), sizeof(long long) * (i1 - i0), cudaMemcpyHostToDevice));
                        pa_distance_kernel<true><<<(i1 - i0 + 3) / 4, 128>>>(a, i0, i1 - i0, nullptr, d_offset, d_list);
                        PD_CK(cudaGetLastError());
                        PD_CK(cudaMemcpy(rec.d2.data() + done, d_list, sizeof(float) * (size_t)n, cudaMemcpyDeviceToHost));
                    }
                    done += n;
                    i0 = i1;
                }
            }
        }
    }
    // host: the reference's loop nest (Distances:1000-1075 / 880-960) over the squared distances the kernel kept, in the
    // same order, with the same REAL(4) / REAL(8) expressions -- every sum is the same chain of additions
    {
        int nav = (job->nsteps - 1 + job->sampling_interval) / job->sampling_interval;
        if (nav < 0) nav = 0;
        const double navf = (double)(float)nav;
        const float maxdist_squared = job->maxdist * job->maxdist;
        memset(results, 0, sizeof(pa_dist_result) * job->nsubjobs);
        for (int pass = 0; pass < (job->calc_stdev ? 2 : 1); ++pass) {
            const bool stdev_run = pass == 1;
            for (size_t s = 0; s < recs.size(); ++s)
                for (int k = 0; k < job->nsubjobs; ++k) {
                    pa_dist_result &r = results[k];
                    const StepRec &rec = recs[s][k];
                    int localweight = 0; long long localweight_ffc = 0;
                    double local_closest = 0, local_exponential = 0, weight_exponential = 0, local_exponential_stdev = 0, local_ffc = 0, local_ffcexp = 0;
                    size_t pos = 0;
                    for (int i = 0; i < ninst[k]; ++i) {
                        float closest = maxdist_squared;
                        const int n = rec.counts[i];
                        if (stdev_run && job->calc_exp) { local_exponential = 0; weight_exponential = 0; }
                        for (int q = 0; q < n; ++q) {
                            const float d2 = rec.d2[pos + q];
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
                        pos += (size_t)n;
                        if (n > 0) {
                            localweight += 1;
                            if (stdev_run) {
                                const double dv = (double)sqrtf(closest) - r.closest_average;
                                local_closest = local_closest + dv * dv;
                                if (job->calc_exp) {
                                    const double de = local_exponential / weight_exponential - r.exponential_average;
                                    local_exponential_stdev = local_exponential_stdev + de * de;
                                }
                            } else local_closest = local_closest + (double)sqrtf(closest);
                        } else r.missed_neighbours += 1;                   // also during the stdev pass (Distances:948,1067)
                    }
                    const double lw = (double)(float)localweight;
                    if (stdev_run) {
                        r.closest_stdev += local_closest / lw;
                        if (job->calc_exp) r.exponential_stdev += local_exponential_stdev / lw;
                    } else {
                        r.closest_average += local_closest / lw;
                        if (job->calc_exp) { r.exponential_average += local_exponential; r.exponential_weight += weight_exponential; }
                        if (job->calc_ffc) {
                            r.ffc_average += local_ffc;
                            if (job->calc_exp) r.ffcexp_average += local_ffcexp;
                            r.ffc_weight += (double)(float)localweight_ffc;
                        }
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
    cudaFree(d_types); cudaFree(d_count); cudaFree(d_offset); cudaFree(d_list);
    for (int k = 0; k < job->nsubjobs; ++k) { cudaFree(d_ref[k]); cudaFree(d_obs[k]); cudaFree(d_first[k]); }
    return rc;
}

// ---------------------------------------------------------------------------
// contact distances: contact_distance / print_distances, Debug:586-623.
//
// One frame.  All atoms are wrapped once (wrap_vector, as give_smallest_atom_distance_squared does
// for both atoms of every call) into one SoA table in type / molecule / atom order.  The
// intermolecular search keeps the reference's evaluation order -- the atom of the molecule type
// under study is always pos_1, and within the same type only molecules with a LOWER index are
// visited (the EXIT at Molecular:1355 leaves the molecule loop) -- because (p2+s)-p1 and (p1+s)-p2
// can round differently.  Minima / maxima are taken on the fp64 squared distance; the REAL(4)
// rounding and the clip to maximum_distance_squared are monotone and applied once at the end.
namespace {

#define PC_THREADS 128
#define PC_CHUNK 128

struct PcArgs {
    const double *x, *y, *z;            // wrapped positions of all atoms of the frame
    int n_all;
    int t_begin, t_end;                 // atom index range of the molecule type under study
    int natoms;                         // atoms per molecule of that type
    int jsplit;                         // observed atoms per blockIdx.y
    double L[3];
    unsigned long long *out;            // [0] min inter, [1] min intra, [2] max intra (bit patterns of doubles >= 0)
};

__global__ void pa_contact_wrap_kernel(const float *xyz, int n, int first, double3 lo, double3 hi, double3 L,
                                       double *x, double *y, double *z)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double l[3] = { lo.x, lo.y, lo.z }, h[3] = { hi.x, hi.y, hi.z }, len[3] = { L.x, L.y, L.z };
    double p[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        double v = (double)xyz[(size_t)i * 3 + k];
        for (int it = 0; it < 1000; ++it) {                       // wrap_vector, Molecular:1560-1586
            if (v < l[k]) v = __dadd_rn(v, len[k]);
            else if (v > h[k]) v = __dsub_rn(v, len[k]);
            else break;
        }
        p[k] = v;
    }
    x[first + i] = p[0]; y[first + i] = p[1]; z[first + i] = p[2];
}

__device__ __forceinline__ double pc_component(double xm, double x0, double xp, double xi)
{
    const double dm = __dsub_rn(xm, xi), d0 = __dsub_rn(x0, xi), dp = __dsub_rn(xp, xi);
    double m = fabs(d0) < fabs(dm) ? d0 : dm;
    m = fabs(dp) < fabs(m) ? dp : m;
    return __dmul_rn(m, m);
}

__global__ void __launch_bounds__(PC_THREADS) pa_contact_inter_kernel(PcArgs a)
{
    __shared__ double sj[9][PC_CHUNK];
    __shared__ unsigned long long smin;
    const int i = a.t_begin + blockIdx.x * PC_THREADS + threadIdx.x;
    const bool live = i < a.t_end;
    const double xi = live ? a.x[i] : 0.0, yi = live ? a.y[i] : 0.0, zi = live ? a.z[i] : 0.0;
    // observed atoms allowed for this reference atom: everything outside [excl_lo, t_end)
    const int excl_lo = live ? a.t_begin + ((i - a.t_begin) / a.natoms) * a.natoms : 0;
    const int j_begin = blockIdx.y * a.jsplit, j_end = min(a.n_all, j_begin + a.jsplit);
    double best = __longlong_as_double(0x7ff0000000000000ll);
    if (threadIdx.x == 0) smin = 0x7ff0000000000000ull;
    __syncthreads();
    for (int j0 = j_begin; j0 < j_end; j0 += PC_CHUNK) {
        __syncthreads();
        {
            const int j = j0 + threadIdx.x;
            double px = 1e300, py = 1e300, pz = 1e300;                 // padding: never the minimum
            if (j < j_end) { px = a.x[j]; py = a.y[j]; pz = a.z[j]; }
            sj[0][threadIdx.x] = __dadd_rn(px, -a.L[0]); sj[1][threadIdx.x] = px; sj[2][threadIdx.x] = __dadd_rn(px, a.L[0]);
            sj[3][threadIdx.x] = __dadd_rn(py, -a.L[1]); sj[4][threadIdx.x] = py; sj[5][threadIdx.x] = __dadd_rn(py, a.L[1]);
            sj[6][threadIdx.x] = __dadd_rn(pz, -a.L[2]); sj[7][threadIdx.x] = pz; sj[8][threadIdx.x] = __dadd_rn(pz, a.L[2]);
        }
        __syncthreads();
        const int nj = min(PC_CHUNK, j_end - j0);
        // chunk entirely allowed for every thread of this CTA?  (block-uniform test on the CTA's exclusion hull)
        const int cta_lo = a.t_begin + ((blockIdx.x * PC_THREADS) / a.natoms) * a.natoms;
        const bool clean = (j0 + nj <= cta_lo) || (j0 >= a.t_end);
        if (clean) {
#pragma unroll 4
            for (int jj = 0; jj < nj; ++jj) {
                const double s = __dadd_rn(__dadd_rn(pc_component(sj[0][jj], sj[1][jj], sj[2][jj], xi),
                                                     pc_component(sj[3][jj], sj[4][jj], sj[5][jj], yi)),
                                           pc_component(sj[6][jj], sj[7][jj], sj[8][jj], zi));
                best = s < best ? s : best;
            }
        } else {
            for (int jj = 0; jj < nj; ++jj) {
                const int j = j0 + jj;
                if (j >= excl_lo && j < a.t_end) continue;              // same molecule, or same type with a higher index
                const double s = __dadd_rn(__dadd_rn(pc_component(sj[0][jj], sj[1][jj], sj[2][jj], xi),
                                                     pc_component(sj[3][jj], sj[4][jj], sj[5][jj], yi)),
                                           pc_component(sj[6][jj], sj[7][jj], sj[8][jj], zi));
                best = s < best ? s : best;
            }
        }
    }
    if (live && best == best) atomicMin(&smin, (unsigned long long)__double_as_longlong(best));
    __syncthreads();
    if (threadIdx.x == 0) atomicMin(&a.out[0], smin);
}

__global__ void pa_contact_intra_kernel(PcArgs a)
{
    const int mol = blockIdx.x * blockDim.x + threadIdx.x;
    const int nmol = (a.t_end - a.t_begin) / a.natoms;
    if (mol >= nmol) return;
    const int base = a.t_begin + mol * a.natoms;
    double lo = __longlong_as_double(0x7ff0000000000000ll), hi = 0.0;
    bool any = false;
    for (int a1 = 0; a1 + 1 < a.natoms; ++a1) {                         // Molecular:1281-1289
        const double xi = a.x[base + a1], yi = a.y[base + a1], zi = a.z[base + a1];
        for (int a2 = a1 + 1; a2 < a.natoms; ++a2) {
            const double xj = a.x[base + a2], yj = a.y[base + a2], zj = a.z[base + a2];
            const double s = __dadd_rn(__dadd_rn(pc_component(__dadd_rn(xj, -a.L[0]), xj, __dadd_rn(xj, a.L[0]), xi),
                                                 pc_component(__dadd_rn(yj, -a.L[1]), yj, __dadd_rn(yj, a.L[1]), yi)),
                                       pc_component(__dadd_rn(zj, -a.L[2]), zj, __dadd_rn(zj, a.L[2]), zi));
            lo = s < lo ? s : lo; hi = s > hi ? s : hi; any = true;
        }
    }
    if (any) {
        atomicMin(&a.out[1], (unsigned long long)__double_as_longlong(lo));
        atomicMax(&a.out[2], (unsigned long long)__double_as_longlong(hi));
    }
}

}  // namespace

extern "C" int pa_contact_distance(const pa_traj *traj, int32_t timestep, int32_t type, const pa_exec *exec, pa_contact_result *result)
{
    if (!traj || !result || timestep < 1 || timestep > traj->nsteps) { pa::set_error("bad argument (time step out of range: error 57)"); return PA_ERR_BAD_ARG; }
    if (type < 1 || type > traj->ntypes) { pa::set_error("molecule type index out of range (error 33)"); return PA_E_TYPE_INVALID; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) { cudaGetLastError(); pa::set_error("no CUDA device (this path has no CPU fallback)"); return PA_ERR_NO_DEVICE; }
    const int device = exec ? exec->device : 0;
    if (cudaSetDevice(device) != cudaSuccess || pa_kernels_selfcheck() != 0) { cudaGetLastError(); pa::set_error("no sm_100a kernel image for this device"); return PA_ERR_NO_DEVICE; }
    int rc = 0;
    const int nt = traj->ntypes;
    std::vector<int> first(nt + 1, 0);
    for (int t = 0; t < nt; ++t) {
        const long long n = (long long)traj->types[t].nmol * traj->types[t].natoms;
        if (first[t] + n > 0x7fffffffll) { pa::set_error("too many atoms"); return PA_ERR_BAD_ARG; }
        first[t + 1] = first[t] + (int)n;
    }
    const int n_all = first[nt];
    double *d_pos = nullptr;
    float *d_raw = nullptr;
    unsigned long long *d_out = nullptr;
    unsigned long long h_out[3] = { 0x7ff0000000000000ull, 0x7ff0000000000000ull, 0ull };
    memset(result, 0, sizeof *result);
    {
        size_t max_raw = 1;
        for (int t = 0; t < nt; ++t) max_raw = std::max(max_raw, (size_t)(first[t + 1] - first[t]) * 3);
        PD_CK(cudaMalloc(&d_pos, sizeof(double) * 3 * (size_t)std::max(n_all, 1)));
        PD_CK(cudaMalloc(&d_raw, sizeof(float) * max_raw));
        PD_CK(cudaMalloc(&d_out, sizeof h_out));
        PD_CK(cudaMemcpy(d_out, h_out, sizeof h_out, cudaMemcpyHostToDevice));
        double *x = d_pos, *y = d_pos + n_all, *z = d_pos + 2 * (size_t)n_all;
        const double3 lo = make_double3(traj->box_lo[0], traj->box_lo[1], traj->box_lo[2]);
        const double3 hi = make_double3(traj->box_hi[0], traj->box_hi[1], traj->box_hi[2]);
        const double3 L = make_double3(traj->box_hi[0] - traj->box_lo[0], traj->box_hi[1] - traj->box_lo[1], traj->box_hi[2] - traj->box_lo[2]);
        for (int t = 0; t < nt; ++t) {
            const int n = first[t + 1] - first[t];
            if (n < 1) continue;
            if (!traj->types[t].xyz) { pa::set_error("coordinates missing"); rc = PA_ERR_BAD_ARG; goto done; }
            PD_CK(cudaMemcpy(d_raw, traj->types[t].xyz + (size_t)(timestep - 1) * n * 3, sizeof(float) * (size_t)n * 3, cudaMemcpyHostToDevice));
            pa_contact_wrap_kernel<<<(n + 255) / 256, 256>>>(d_raw, n, first[t], lo, hi, L, x, y, z);
            PD_CK(cudaGetLastError());
        }
        PcArgs a{};
        a.x = x; a.y = y; a.z = z; a.n_all = n_all;
        a.t_begin = first[type - 1]; a.t_end = first[type]; a.natoms = traj->types[type - 1].natoms;
        a.L[0] = L.x; a.L[1] = L.y; a.L[2] = L.z; a.out = d_out;
        const int n_ref = a.t_end - a.t_begin;
        if (n_ref > 0) {
            const int bx = (n_ref + PC_THREADS - 1) / PC_THREADS;
            int splits = std::max(1, (148 * 8 + bx - 1) / bx);
            const int max_splits = std::max(1, (n_all + PC_CHUNK - 1) / PC_CHUNK);
            splits = std::min(std::min(splits, max_splits), 65535);
            a.jsplit = (((n_all + splits - 1) / splits + PC_CHUNK - 1) / PC_CHUNK) * PC_CHUNK;
            splits = (n_all + a.jsplit - 1) / a.jsplit;
            pa_contact_inter_kernel<<<dim3(bx, splits), PC_THREADS>>>(a);
            PD_CK(cudaGetLastError());
            const int nmol = traj->types[type - 1].nmol;
            pa_contact_intra_kernel<<<(nmol + 127) / 128, 128>>>(a);
            PD_CK(cudaGetLastError());
        }
        PD_CK(cudaMemcpy(h_out, d_out, sizeof h_out, cudaMemcpyDeviceToHost));
        auto finish = [&](unsigned long long bits, bool is_min) {
            double m; memcpy(&m, &bits, 8);
            float v = is_min ? (float)traj->max_dist_sq : 0.0f;            // smallest=maximum_distance_squared / largest=0.0
            const float c = (float)(m < traj->max_dist_sq ? m : traj->max_dist_sq);   // REAL(4) function result, Molecular:1400,1418
            if (is_min ? (c < v) : (c > v)) v = c;
            return sqrtf(v);
        };
        result->smallest_intermolecular = finish(h_out[0], true);
        result->smallest_intramolecular = finish(h_out[1], true);
        result->largest_intramolecular = (h_out[1] == 0x7ff0000000000000ull && h_out[2] == 0) ? 0.0f : finish(h_out[2], false);
        const long long nm = traj->types[type - 1].nmol, na = a.natoms;
        result->pairs_evaluated = (long long)n_ref * (n_all - n_ref) + na * na * (nm * (nm - 1) / 2);
    }
done:
    cudaFree(d_pos); cudaFree(d_raw); cudaFree(d_out);
    return rc;
}

// ---------------------------------------------------------------------------
// neighbour pairs: the REAL(4) test give_smallest_atom_distance_squared(...) < cutoff**2 of the
// cluster / speciation analyses (Cluster:1094-1100, Speciation:1407-1431) for every pair of two atom lists.
// Fast path: a cell list (further down).  Fallback for short lists and boxes of fewer than 3 x 3 x 3 cells: brute force
// like the reference (every a-instance against every b-instance, observed atoms staged as (x-L, x, x+L) triples), hits
// appended with one atomic per warp and sorted on the host into the order of the reference's nested loops.
namespace {

struct PnArgs {
    const double *x, *y, *z;            // wrapped positions of all atoms of the frame
    const int *ia_atom, *ib_atom;       // global atom index of every instance
    const int *ia_mol, *ib_mol;         // global molecule id of every instance (skip_same_molecule)
    const int *a_class, *b_class;       // may be null
    const float *cutoff_sq;
    int n_a, n_b, n_class_b;
    int same_list, skip_same_molecule;
    int jsplit;
    double L[3], mds;
    int2 *out;
    unsigned long long capacity;
    unsigned long long *count;
};

__global__ void __launch_bounds__(PC_THREADS) pa_neighbour_kernel(PnArgs a)
{
    __shared__ double sj[9][PC_CHUNK];
    __shared__ int s_mol[PC_CHUNK], s_cls[PC_CHUNK];
    const int ia = blockIdx.x * PC_THREADS + threadIdx.x;
    const bool live = ia < a.n_a;
    double xi = 0.0, yi = 0.0, zi = 0.0;
    int moli = -1, rowi = 0;
    if (live) {
        const int at = a.ia_atom[ia];
        xi = a.x[at]; yi = a.y[at]; zi = a.z[at];
        moli = a.ia_mol[ia];
        rowi = (a.a_class ? a.a_class[ia] : 0) * a.n_class_b;
    }
    const int lane = threadIdx.x & 31;
    const int j_begin = blockIdx.y * a.jsplit, j_end = min(a.n_b, j_begin + a.jsplit);
    for (int j0 = j_begin; j0 < j_end; j0 += PC_CHUNK) {
        __syncthreads();
        {
            const int j = j0 + threadIdx.x;
            double px = 1e300, py = 1e300, pz = 1e300;
            int mol = -2, cls = 0;
            if (j < j_end) {
                const int at = a.ib_atom[j];
                px = a.x[at]; py = a.y[at]; pz = a.z[at];
                mol = a.ib_mol[j]; cls = a.b_class ? a.b_class[j] : 0;
            }
            sj[0][threadIdx.x] = __dadd_rn(px, -a.L[0]); sj[1][threadIdx.x] = px; sj[2][threadIdx.x] = __dadd_rn(px, a.L[0]);
            sj[3][threadIdx.x] = __dadd_rn(py, -a.L[1]); sj[4][threadIdx.x] = py; sj[5][threadIdx.x] = __dadd_rn(py, a.L[1]);
            sj[6][threadIdx.x] = __dadd_rn(pz, -a.L[2]); sj[7][threadIdx.x] = pz; sj[8][threadIdx.x] = __dadd_rn(pz, a.L[2]);
            s_mol[threadIdx.x] = mol; s_cls[threadIdx.x] = cls;
        }
        __syncthreads();
        const int nj = min(PC_CHUNK, j_end - j0);
        for (int jj = 0; jj < nj; ++jj) {
            const int j = j0 + jj;
            bool hit = false;
            if (live && !(a.same_list && j <= ia) && !(a.skip_same_molecule && s_mol[jj] == moli)) {
                const double s = __dadd_rn(__dadd_rn(pc_component(sj[0][jj], sj[1][jj], sj[2][jj], xi),
                                                     pc_component(sj[3][jj], sj[4][jj], sj[5][jj], yi)),
                                           pc_component(sj[6][jj], sj[7][jj], sj[8][jj], zi));
                // REAL(4) function result (Molecular:1400,1418) against the REAL(4) squared cutoff
                hit = __double2float_rn(s < a.mds ? s : a.mds) < a.cutoff_sq[rowi + s_cls[jj]];
            }
            const unsigned int m = __ballot_sync(0xffffffffu, hit);
            if (m == 0u) continue;
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(a.count, (unsigned long long)__popc(m));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (hit) {
                const unsigned long long slot = base + __popc(m & ((1u << lane) - 1u));
                if (slot < a.capacity) a.out[slot] = make_int2(ia, j);
            }
        }
    }
}

// ---- cell list (the fast path): the b-instances are binned into a grid whose cells are at least as wide as the largest
// cutoff, and every a-instance only meets the b-instances of the 27 cells around its own (periodic).  The test on a
// candidate pair is the same literal REAL(4) predicate as in the brute-force kernel, so the set of pairs is identical --
// a pair closer than the cutoff differs by less than one cell width per component in its minimum image, hence sits in
// adjacent cells.  Needs >= 3 cells per component and every wrapped position inside the box (else: brute force).
struct PnGrid { int g[3]; double lo[3], inv_w[3]; int ncell; };

__global__ void pn_cell_count_kernel(PnArgs a, PnGrid G, int *cell_of_b, int *cell_count, unsigned int *err)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= a.n_b) return;
    const int at = a.ib_atom[j];
    const double p[3] = { a.x[at], a.y[at], a.z[at] };
    int c[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const double u = (p[k] - G.lo[k]) * G.inv_w[k];
        if (!(u >= 0.0 && u <= (double)G.g[k])) { atomicOr(err, 1u); c[k] = 0; continue; }     // outside the box / NaN
        c[k] = min(G.g[k] - 1, (int)u);
    }
    const int cell = (c[2] * G.g[1] + c[1]) * G.g[0] + c[0];
    cell_of_b[j] = cell;
    atomicAdd(&cell_count[cell], 1);
}

// exclusive scan of cell_count[0..n) into cell_start[0..n] (one block)
__global__ void __launch_bounds__(1024) pn_scan_kernel(const int *cell_count, int *cell_start, int n)
{
    __shared__ int s_part[1024];
    const int per = (n + 1023) / 1024;
    const int b = min(n, (int)threadIdx.x * per), e = min(n, b + per);
    int sum = 0;
    for (int i = b; i < e; ++i) sum += cell_count[i];
    s_part[threadIdx.x] = sum;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {
        int v = threadIdx.x >= off ? s_part[threadIdx.x - off] : 0;
        __syncthreads();
        s_part[threadIdx.x] += v;
        __syncthreads();
    }
    int run = threadIdx.x ? s_part[threadIdx.x - 1] : 0;
    for (int i = b; i < e; ++i) { cell_start[i] = run; run += cell_count[i]; }
    if (threadIdx.x == 1023) cell_start[n] = s_part[1023];
}

__global__ void pn_scatter_kernel(PnArgs a, const int *cell_of_b, const int *cell_start, int *cell_fill, int *sorted_b)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= a.n_b) return;
    const int cell = cell_of_b[j];
    sorted_b[cell_start[cell] + atomicAdd(&cell_fill[cell], 1)] = j;
}

// Two passes over the same candidates, so that the pair list leaves the device already in the order of the reference's
// nested loops (a-instance, then b-instance) and the host has nothing to sort: FILL = false counts the hits of every
// a-instance, an exclusive scan turns the counts into segment starts, FILL = true writes each a-instance's hits into its
// segment and orders the (few) entries of the segment by b-instance.
template <bool FILL>
__global__ void __launch_bounds__(128) pn_cells_kernel(PnArgs a, PnGrid G, const int *cell_start, const int *sorted_b, unsigned int *err,
                                                       unsigned long long *tests, int *hits, const int *seg_start)
{
    const int ia = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long ntest = 0;
    if (ia < a.n_a) {
        const int at = a.ia_atom[ia];
        const double xi = a.x[at], yi = a.y[at], zi = a.z[at];
        const int moli = a.ia_mol[ia];
        const int rowi = (a.a_class ? a.a_class[ia] : 0) * a.n_class_b;
        int c[3];
        const double p[3] = { xi, yi, zi };
        bool ok = true;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const double u = (p[k] - G.lo[k]) * G.inv_w[k];
            if (!(u >= 0.0 && u <= (double)G.g[k])) { ok = false; c[k] = 0; continue; }
            c[k] = min(G.g[k] - 1, (int)u);
        }
        int nhit = 0;
        const unsigned long long seg = FILL ? (unsigned long long)seg_start[ia] : 0ull;
        if (!ok) atomicOr(err, 1u);
        else {
            for (int dz = -1; dz <= 1; ++dz) {
                const int cz = (c[2] + dz + G.g[2]) % G.g[2];
                for (int dy = -1; dy <= 1; ++dy) {
                    const int cy = (c[1] + dy + G.g[1]) % G.g[1];
                    for (int dx = -1; dx <= 1; ++dx) {
                        const int cx = (c[0] + dx + G.g[0]) % G.g[0];
                        const int cell = (cz * G.g[1] + cy) * G.g[0] + cx;
                        for (int q = cell_start[cell]; q < cell_start[cell + 1]; ++q) {
                            const int j = sorted_b[q];
                            if ((a.same_list && j <= ia) || (a.skip_same_molecule && a.ib_mol[j] == moli)) continue;
                            const int bt = a.ib_atom[j];
                            const double px = a.x[bt], py = a.y[bt], pz = a.z[bt];
                            const double s = __dadd_rn(__dadd_rn(pc_component(__dadd_rn(px, -a.L[0]), px, __dadd_rn(px, a.L[0]), xi),
                                                                 pc_component(__dadd_rn(py, -a.L[1]), py, __dadd_rn(py, a.L[1]), yi)),
                                                       pc_component(__dadd_rn(pz, -a.L[2]), pz, __dadd_rn(pz, a.L[2]), zi));
                            ++ntest;
                            // REAL(4) function result (Molecular:1400,1418) against the REAL(4) squared cutoff
                            if (__double2float_rn(s < a.mds ? s : a.mds) < a.cutoff_sq[rowi + (a.b_class ? a.b_class[j] : 0)]) {
                                if (FILL) {
                                    // insertion into the ordered segment (a handful of entries)
                                    int k = nhit;
                                    for (; k > 0; --k) {
                                        const int jp = a.out[seg + (unsigned)(k - 1)].y;
                                        if (jp < j) break;
                                        a.out[seg + (unsigned)k] = make_int2(ia, jp);
                                    }
                                    a.out[seg + (unsigned)k] = make_int2(ia, j);
                                }
                                ++nhit;
                            }
                        }
                    }
                }
            }
        }
        if (!FILL) hits[ia] = nhit;
        if (!FILL) ntest |= (unsigned long long)nhit << 40;                  // (tests of one thread < 2^40) hits ride along
    }
    if (!FILL) {
        // distance tests carried out (statistics) and the 64-bit number of pairs (the scan of `hits` is 32-bit)
        unsigned long long nh = ntest >> 40;
        ntest &= (1ull << 40) - 1ull;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { ntest += __shfl_xor_sync(0xffffffffu, ntest, o); nh += __shfl_xor_sync(0xffffffffu, nh, o); }
        if ((threadIdx.x & 31) == 0 && ntest) atomicAdd(tests, ntest);
        if ((threadIdx.x & 31) == 0 && nh) atomicAdd(a.count, nh);
    }
}

}  // namespace

// Work buffers of pa_neighbour_pairs: stream-ordered allocations from the device's default memory pool (kept by the driver
// between calls -- the release threshold is raised once per device), so that a cluster analysis calling this once per
// time step does not pay eight cudaMalloc / cudaFree round trips every time.  Everything runs on the default stream.
template <typename T> static cudaError_t pn_alloc(T **ptr, size_t bytes)
{
    static bool pool_set[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !pool_set[dev]) {
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
            unsigned long long keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        pool_set[dev] = true;
    }
    return cudaMallocAsync(reinterpret_cast<void **>(ptr), std::max<size_t>(bytes, 16), 0);
}

extern "C" int pa_neighbour_pairs(const pa_traj *traj, const pa_neigh_request *rq, const pa_exec *exec,
                                  pa_neigh_pair *out, int64_t capacity, int64_t *count, int64_t *tests)
{
    if (!traj || !rq || !count || capacity < 0 || (capacity > 0 && !out) || rq->n_a < 0 || rq->n_b < 0 || !rq->cutoff_sq ||
        rq->n_class_a < 1 || rq->n_class_b < 1 || (rq->n_a > 0 && !rq->a) || (rq->n_b > 0 && !rq->b)) {
        pa::set_error("bad argument"); return PA_ERR_BAD_ARG;
    }
    if (rq->timestep < 1 || rq->timestep > traj->nsteps) { pa::set_error("time step out of range (error 57)"); return PA_ERR_BAD_ARG; }
    const int nt = traj->ntypes;
    std::vector<int> first(nt + 1, 0), first_mol(nt + 1, 0);
    for (int t = 0; t < nt; ++t) {
        const long long n = (long long)traj->types[t].nmol * traj->types[t].natoms;
        if (first[t] + n > 0x7fffffffll) { pa::set_error("too many atoms"); return PA_ERR_BAD_ARG; }
        first[t + 1] = first[t] + (int)n;
        first_mol[t + 1] = first_mol[t] + traj->types[t].nmol;
    }
    // instance lists -> global atom / molecule indices (host; validates the triples: errors 110 / 111 / 33)
    std::vector<int> h_at[2], h_mol[2];
    for (int side = 0; side < 2; ++side) {
        const int n = side ? rq->n_b : rq->n_a;
        const int32_t *lst = side ? rq->b : rq->a;
        const int32_t *cls = side ? rq->b_class : rq->a_class;
        h_at[side].resize(n); h_mol[side].resize(n);
        for (int i = 0; i < n; ++i) {
            const int ty = lst[3 * i], mol = lst[3 * i + 1], at = lst[3 * i + 2];
            if (ty < 1 || ty > nt) { pa::set_error("invalid molecule type in atom list (error 33)"); return PA_E_TYPE_INVALID; }
            if (mol < 1 || mol > traj->types[ty - 1].nmol) { pa::set_error("invalid molecule index in atom list (error 110)"); return 110; }
            if (at < 1 || at > traj->types[ty - 1].natoms) { pa::set_error("invalid atom index in atom list (error 111)"); return 111; }
            if (cls && (cls[i] < 0 || cls[i] >= (side ? rq->n_class_b : rq->n_class_a))) { pa::set_error("cutoff class out of range"); return PA_ERR_BAD_ARG; }
            h_at[side][i] = first[ty - 1] + (mol - 1) * traj->types[ty - 1].natoms + (at - 1);
            h_mol[side][i] = first_mol[ty - 1] + (mol - 1);
        }
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) { cudaGetLastError(); pa::set_error("no CUDA device (this path has no CPU fallback)"); return PA_ERR_NO_DEVICE; }
    const int device = exec ? exec->device : 0;
    if (cudaSetDevice(device) != cudaSuccess || pa_kernels_selfcheck() != 0) { cudaGetLastError(); pa::set_error("no sm_100a kernel image for this device"); return PA_ERR_NO_DEVICE; }
    *count = 0;
    if (tests) *tests = 0;
    if (rq->n_a == 0 || rq->n_b == 0) return 0;
    int rc = 0;
    const int n_all = first[nt];
    double *d_pos = nullptr;
    float *d_raw = nullptr, *d_cut = nullptr;
    int *d_idx = nullptr;
    int2 *d_out = nullptr;
    unsigned long long *d_count = nullptr, h_count = 0;
    int *d_cellbuf = nullptr;
    unsigned long long *d_aux = nullptr;
    {
        size_t max_raw = 1;
        for (int t = 0; t < nt; ++t) max_raw = std::max(max_raw, (size_t)(first[t + 1] - first[t]) * 3);
        PD_CK(pn_alloc(&d_pos, sizeof(double) * 3 * (size_t)std::max(n_all, 1)));
        PD_CK(pn_alloc(&d_raw, sizeof(float) * max_raw));
        double *x = d_pos, *y = d_pos + n_all, *z = d_pos + 2 * (size_t)n_all;
        const double3 lo = make_double3(traj->box_lo[0], traj->box_lo[1], traj->box_lo[2]);
        const double3 hi = make_double3(traj->box_hi[0], traj->box_hi[1], traj->box_hi[2]);
        const double3 L = make_double3(traj->box_hi[0] - traj->box_lo[0], traj->box_hi[1] - traj->box_lo[1], traj->box_hi[2] - traj->box_lo[2]);
        for (int t = 0; t < nt; ++t) {
            const int n = first[t + 1] - first[t];
            if (n < 1) continue;
            if (!traj->types[t].xyz) { pa::set_error("coordinates missing"); rc = PA_ERR_BAD_ARG; goto done; }
            PD_CK(cudaMemcpy(d_raw, traj->types[t].xyz + (size_t)(rq->timestep - 1) * n * 3, sizeof(float) * (size_t)n * 3, cudaMemcpyHostToDevice));
            pa_contact_wrap_kernel<<<(n + 255) / 256, 256>>>(d_raw, n, first[t], lo, hi, L, x, y, z);
            PD_CK(cudaGetLastError());
        }
        // index block: [a atoms][a mols][a classes][b atoms][b mols][b classes]
        const size_t na = (size_t)rq->n_a, nb = (size_t)rq->n_b;
        PD_CK(pn_alloc(&d_idx, sizeof(int) * 3 * (na + nb)));
        PD_CK(cudaMemcpy(d_idx, h_at[0].data(), sizeof(int) * na, cudaMemcpyHostToDevice));
        PD_CK(cudaMemcpy(d_idx + na, h_mol[0].data(), sizeof(int) * na, cudaMemcpyHostToDevice));
        if (rq->a_class) PD_CK(cudaMemcpy(d_idx + 2 * na, rq->a_class, sizeof(int) * na, cudaMemcpyHostToDevice));
        PD_CK(cudaMemcpy(d_idx + 3 * na, h_at[1].data(), sizeof(int) * nb, cudaMemcpyHostToDevice));
        PD_CK(cudaMemcpy(d_idx + 3 * na + nb, h_mol[1].data(), sizeof(int) * nb, cudaMemcpyHostToDevice));
        if (rq->b_class) PD_CK(cudaMemcpy(d_idx + 3 * na + 2 * nb, rq->b_class, sizeof(int) * nb, cudaMemcpyHostToDevice));
        const size_t ncut = (size_t)rq->n_class_a * rq->n_class_b;
        PD_CK(pn_alloc(&d_cut, sizeof(float) * ncut));
        PD_CK(cudaMemcpy(d_cut, rq->cutoff_sq, sizeof(float) * ncut, cudaMemcpyHostToDevice));
        PD_CK(pn_alloc(&d_count, sizeof(unsigned long long)));
        PD_CK(cudaMemset(d_count, 0, sizeof(unsigned long long)));
        PnArgs a{};
        a.x = x; a.y = y; a.z = z;
        a.ia_atom = d_idx; a.ia_mol = d_idx + na; a.a_class = rq->a_class ? d_idx + 2 * na : nullptr;
        a.ib_atom = d_idx + 3 * na; a.ib_mol = d_idx + 3 * na + nb; a.b_class = rq->b_class ? d_idx + 3 * na + 2 * nb : nullptr;
        a.cutoff_sq = d_cut; a.n_a = rq->n_a; a.n_b = rq->n_b; a.n_class_b = rq->n_class_b;
        a.same_list = rq->same_list ? 1 : 0; a.skip_same_molecule = rq->skip_same_molecule ? 1 : 0;
        a.L[0] = L.x; a.L[1] = L.y; a.L[2] = L.z; a.mds = traj->max_dist_sq;
        a.out = nullptr; a.capacity = (unsigned long long)capacity; a.count = d_count;
        // cell list when the box holds at least 3 x 3 x 3 cells of the largest cutoff (flag 0x800 of pa_exec: brute force)
        bool used_cells = false;
        unsigned long long h_tests = 0;
        {
            float cmax2 = 0.f;
            for (size_t i = 0; i < ncut; ++i) cmax2 = std::max(cmax2, rq->cutoff_sq[i]);
            const double cmax = std::sqrt((double)cmax2) * (1.0 + 1e-9) + 1e-9;     // margin >> rounding of the cell index
            PnGrid G{};
            const double Lk[3] = { L.x, L.y, L.z };
            int min_b = 1024;                                  // below that the brute-force kernel is done before the grid is built
            if (const char *e = getenv("PA_NEIGH_MIN_CELLS")) min_b = atoi(e);       // test hook
            bool grid_ok = cmax2 > 0.f && std::isfinite(cmax) && rq->n_b >= min_b && !(exec && (exec->flags & 0x800));
            long long ncell = 1;
            for (int k = 0; k < 3; ++k) {
                G.g[k] = (int)std::min(128.0, std::floor(Lk[k] / cmax));
                if (!(G.g[k] >= 3)) grid_ok = false;
                G.lo[k] = traj->box_lo[k]; G.inv_w[k] = (double)G.g[k] / Lk[k];
                ncell *= std::max(G.g[k], 1);
            }
            if (grid_ok) {
                G.ncell = (int)ncell;
                PD_CK(pn_alloc(&d_cellbuf, sizeof(int) * ((size_t)2 * rq->n_b + 3 * (size_t)ncell + 2 + 2 * (size_t)rq->n_a + 1)));
                int *cell_of_b = d_cellbuf, *sorted_b = d_cellbuf + rq->n_b, *cell_count = sorted_b + rq->n_b, *cell_start = cell_count + ncell,
                    *cell_fill = cell_start + ncell + 1, *hits = cell_fill + ncell + 1, *seg_start = hits + rq->n_a;
                PD_CK(cudaMemset(cell_count, 0, sizeof(int) * (size_t)ncell));
                PD_CK(cudaMemset(cell_fill, 0, sizeof(int) * (size_t)ncell));
                PD_CK(pn_alloc(&d_aux, sizeof(unsigned long long) * 2));
                PD_CK(cudaMemset(d_aux, 0, sizeof(unsigned long long) * 2));
                unsigned int *d_err = reinterpret_cast<unsigned int *>(d_aux);
                unsigned long long *d_tests = d_aux + 1;
                pn_cell_count_kernel<<<(rq->n_b + 255) / 256, 256>>>(a, G, cell_of_b, cell_count, d_err);
                pn_scan_kernel<<<1, 1024>>>(cell_count, cell_start, (int)ncell);
                pn_scatter_kernel<<<(rq->n_b + 255) / 256, 256>>>(a, cell_of_b, cell_start, cell_fill, sorted_b);
                pn_cells_kernel<false><<<(rq->n_a + 127) / 128, 128>>>(a, G, cell_start, sorted_b, d_err, d_tests, hits, nullptr);
                pn_scan_kernel<<<1, 1024>>>(hits, seg_start, rq->n_a);
                PD_CK(cudaGetLastError());
                unsigned long long aux[2], total = 0;
                PD_CK(cudaMemcpy(aux, d_aux, sizeof aux, cudaMemcpyDeviceToHost));
                PD_CK(cudaMemcpy(&total, d_count, sizeof total, cudaMemcpyDeviceToHost));
                if ((unsigned int)(aux[0] & 0xffffffffu) != 0u) {
                    PD_CK(cudaMemset(d_count, 0, sizeof(unsigned long long)));      // a position outside the box: start over, brute force
                } else {
                    if (total > 0x7fffffffull) { pa::set_error("more than 2^31 neighbour pairs"); rc = PA_ERR_UNSUPPORTED; goto done; }
                    used_cells = true; h_tests = aux[1];
                    h_count = total;
                    if (total > 0 && total <= (unsigned long long)capacity) {
                        PD_CK(pn_alloc(&d_out, sizeof(int2) * (size_t)total));
                        a.out = d_out;
                        pn_cells_kernel<true><<<(rq->n_a + 127) / 128, 128>>>(a, G, cell_start, sorted_b, d_err, d_tests, nullptr, seg_start);
                        PD_CK(cudaGetLastError());
                        static_assert(sizeof(pa_neigh_pair) == sizeof(int2), "layout");
                        PD_CK(cudaMemcpy(out, d_out, sizeof(int2) * (size_t)total, cudaMemcpyDeviceToHost));     // already in loop order
                    }
                }
            }
        }
        if (!used_cells) {
            // brute force (few instances, or a box of fewer than 3 cells per component): pairs are appended as they are found
            // and ordered on the host
            PD_CK(pn_alloc(&d_out, sizeof(int2) * (size_t)std::max<int64_t>(capacity, 1)));
            a.out = d_out;
            const int bx = (rq->n_a + PC_THREADS - 1) / PC_THREADS;
            int splits = std::max(1, (148 * 8 + bx - 1) / bx);
            splits = std::min(std::min(splits, std::max(1, (rq->n_b + PC_CHUNK - 1) / PC_CHUNK)), 65535);
            a.jsplit = (((rq->n_b + splits - 1) / splits + PC_CHUNK - 1) / PC_CHUNK) * PC_CHUNK;
            splits = (rq->n_b + a.jsplit - 1) / a.jsplit;
            pa_neighbour_kernel<<<dim3(bx, splits), PC_THREADS>>>(a);
            PD_CK(cudaGetLastError());
            PD_CK(cudaMemcpy(&h_count, d_count, sizeof h_count, cudaMemcpyDeviceToHost));
            const size_t ncopy = (size_t)std::min<unsigned long long>(h_count, (unsigned long long)capacity);
            if (ncopy > 0) {
                static_assert(sizeof(pa_neigh_pair) == sizeof(int2), "layout");
                PD_CK(cudaMemcpy(out, d_out, sizeof(int2) * ncopy, cudaMemcpyDeviceToHost));
                if (h_count <= (unsigned long long)capacity)       // complete list: the order of the reference's nested loops
                    std::sort(out, out + ncopy, [](const pa_neigh_pair &p, const pa_neigh_pair &q) { return p.ia != q.ia ? p.ia < q.ia : p.ib < q.ib; });
            }
        }
        *count = (int64_t)h_count;
        if (tests) {
            long long t = (long long)rq->n_a * rq->n_b;
            if (rq->same_list) t = (long long)rq->n_a * (rq->n_a - 1) / 2;
            *tests = used_cells ? (long long)h_tests : t;       // brute force: upper bound when skip_same_molecule removes pairs
        }
    }
done:
    for (void *ptr : { (void *)d_pos, (void *)d_raw, (void *)d_cut, (void *)d_idx, (void *)d_out, (void *)d_count, (void *)d_cellbuf, (void *)d_aux })
        if (ptr) cudaFreeAsync(ptr, 0);
    cudaStreamSynchronize(0);
    return rc;
}
