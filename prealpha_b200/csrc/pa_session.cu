// pa_session.cu -- device sessions and the compute entry points of the C ABI.
//
// Replaces iterate_timesteps_parallelised (Distribution:696-760): where the reference hands
// frames to OpenMP threads with private histograms and merges them in a CRITICAL section,
// a session streams batches of frames through pinned staging buffers (cudaMemcpyAsync on a
// copy stream, double buffered), runs the prepare + pair kernels on a compute stream and keeps
// one int64 accumulator block per job in HBM.  Frames are independent, so a shard of the
// sampled frames can be given to each GPU; integer accumulation makes the sum order-free.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <chrono>
#include <mutex>
#include <thread>
#include <dlfcn.h>
#include <algorithm>
#include <string>
#include <vector>
#include "pa_internal.h"

namespace {

thread_local std::string g_last_error;

#define PA_CK(call)                                                                              \
    do {                                                                                         \
        cudaError_t e__ = (call);                                                                \
        if (e__ != cudaSuccess) {                                                                \
            pa::set_error(std::string(#call) + ": " + cudaGetErrorString(e__));                  \
            return e__ == cudaErrorNoDevice || e__ == cudaErrorInsufficientDriver ||             \
                           e__ == cudaErrorNoKernelImageForDevice ? PA_ERR_NO_DEVICE : PA_ERR_CUDA; \
        }                                                                                        \
    } while (0)

struct TypeMeta {
    int32_t charge, natoms, nmol, offset;
    double mass, realcharge;
    bool needed;
    double *d_mass = nullptr, *d_charge = nullptr;
    size_t frame_floats() const { return (size_t)nmol * natoms * 3; }
};

struct TypeGrid {
    bool enabled = false;
    PaCellGrid g{};
    // two deterministic tile lists (pa_cells.cu): cell-aligned CTA tiles of the general kernel and 64-molecule warp tiles
    // of the RDF kernel
    PaTile *tiles = nullptr, *wtiles = nullptr;
    unsigned int *counters = nullptr;   // [0] ntiles, [1] work counter, [2] number of warp tiles
    unsigned int tile_cap = 0, wtile_cap = 0;
    bool tiles_ready = false, wtiles_ready = false;
    int *row_off = nullptr;             // scratch of the tile builder: one int per (frame, row of cells)
};

struct JobState {
    pa_job job;
    PaDevJob dj;
    bool fast;
    double delta_top = 0.0;   // relative distance of the cutoff / the start of class na from the boundary of bin na (variant 0)
    int f32_cnt = 0;     // > 0: float32 segments of the cell kernel enabled (counter classes needed), see pa_cells.cu
    int cells_variant;   // 0: approximate class + rare exact fix-up, 1: exact table for every pair (see pa_cells.cu)
    size_t acc_off;      // offset (words) inside the accumulator block
    size_t nbins;
    unsigned long long *d_thr_a = nullptr;
    double *d_thr_b = nullptr;
    unsigned long long tiny_bits = 0;   // charge_arm only
};

constexpr unsigned int kExcCap = 4u << 20;
constexpr int kCellsMinMolecules = 2000;
constexpr double kCellsMinPairs = 2.0e7;

}  // namespace

void pa::set_error(const std::string &msg) { g_last_error = msg; }

// Page-locking host memory costs tens to hundreds of milliseconds per call; the staging buffers of the last
// sessions are kept (at most 4 buffers, 1 GiB in total) and handed to the next session that asks for the same size.
namespace {
struct PinnedEntry { void *ptr; size_t bytes; };
std::mutex g_pool_mutex;
std::vector<PinnedEntry> g_pinned_pool;

void *pinned_pool_take(size_t bytes)
{
    std::lock_guard<std::mutex> lock(g_pool_mutex);
    for (size_t i = 0; i < g_pinned_pool.size(); ++i)
        if (g_pinned_pool[i].bytes == bytes) { void *p = g_pinned_pool[i].ptr; g_pinned_pool.erase(g_pinned_pool.begin() + i); return p; }
    return nullptr;
}

void pinned_pool_give(void *ptr, size_t bytes)
{
    std::lock_guard<std::mutex> lock(g_pool_mutex);
    size_t total = bytes;
    for (const PinnedEntry &e : g_pinned_pool) total += e.bytes;
    if (g_pinned_pool.size() >= 4 || total > (size_t(1) << 30)) { cudaFreeHost(ptr); return; }
    g_pinned_pool.push_back({ ptr, bytes });
}
}  // namespace

// Device memory of finished sessions is kept as well (exact-size blocks, at most 4 GiB per process): a one-shot call
// allocates ~40 buffers and frees them again, and cudaFree synchronises the device -- tens to hundreds of milliseconds
// per call that repeated analyses of same-shaped trajectories need not pay.  pa_trim() returns everything.
namespace {
struct DevBlock { void *ptr; size_t bytes; int dev; };
std::mutex g_dev_mutex;
std::vector<DevBlock> g_dev_pool;
std::vector<DevBlock> g_dev_live;
size_t g_dev_pool_bytes = 0;

cudaError_t pool_malloc_raw(void **out, size_t bytes)
{
    int dev = 0;
    cudaGetDevice(&dev);
    if (bytes == 0) bytes = 1;
    {
        std::lock_guard<std::mutex> lock(g_dev_mutex);
        for (size_t i = 0; i < g_dev_pool.size(); ++i)
            if (g_dev_pool[i].dev == dev && g_dev_pool[i].bytes == bytes) {
                *out = g_dev_pool[i].ptr;
                g_dev_live.push_back(g_dev_pool[i]);
                g_dev_pool_bytes -= bytes;
                g_dev_pool.erase(g_dev_pool.begin() + i);
                return cudaSuccess;
            }
    }
    cudaError_t e = cudaMalloc(out, bytes);
    if (e != cudaSuccess) {
        // out of memory with blocks parked in the pool: give them back and try once more
        cudaGetLastError();
        std::vector<DevBlock> drop;
        { std::lock_guard<std::mutex> lock(g_dev_mutex); drop.swap(g_dev_pool); g_dev_pool_bytes = 0; }
        for (const DevBlock &b : drop) { cudaSetDevice(b.dev); cudaFree(b.ptr); }
        cudaSetDevice(dev);
        e = cudaMalloc(out, bytes);
        if (e != cudaSuccess) return e;
    }
    std::lock_guard<std::mutex> lock(g_dev_mutex);
    g_dev_live.push_back({ *out, bytes, dev });
    return cudaSuccess;
}
template <typename T> cudaError_t pool_malloc(T **out, size_t bytes) { return pool_malloc_raw(reinterpret_cast<void **>(out), bytes); }

void pool_free(void *ptr)
{
    if (!ptr) return;
    DevBlock b{ ptr, 0, 0 };
    bool keep = false;
    {
        std::lock_guard<std::mutex> lock(g_dev_mutex);
        for (size_t i = 0; i < g_dev_live.size(); ++i)
            if (g_dev_live[i].ptr == ptr) { b = g_dev_live[i]; g_dev_live.erase(g_dev_live.begin() + i); break; }
        if (b.bytes && g_dev_pool.size() < 512 && g_dev_pool_bytes + b.bytes <= (size_t(4) << 30)) {
            g_dev_pool.push_back(b); g_dev_pool_bytes += b.bytes; keep = true;
        }
    }
    if (!keep) cudaFree(ptr);
}

void pool_trim()
{
    std::vector<DevBlock> drop;
    { std::lock_guard<std::mutex> lock(g_dev_mutex); drop.swap(g_dev_pool); g_dev_pool_bytes = 0; }
    int dev = 0;
    cudaGetDevice(&dev);
    for (const DevBlock &b : drop) { cudaSetDevice(b.dev); cudaFree(b.ptr); }
    cudaSetDevice(dev);
}
}  // namespace

struct pa_session {
    int device = 0, sm_count = 148;
    pa_exec exec{};
    std::vector<TypeMeta> types;
    PaBox box{};
    int64_t M = 0;
    std::vector<JobState> jobs;
    bool need_coc = false, force_general = false, verify = false;
    int max_frames = 0, nframes = 0;       // capacity / resident frames
    int cur = 0;                            // raw buffer slot holding the resident frames
    cudaStream_t s_compute = nullptr, s_copy = nullptr;
    cudaEvent_t ev_h2d[2] = { nullptr, nullptr }, ev_done[2] = { nullptr, nullptr };
    bool slot_used[2] = { false, false };
    std::vector<float *> d_raw[2];
    float *h_pinned[2] = { nullptr, nullptr };
    size_t h_pinned_bytes[2] = { 0, 0 };
    size_t pinned_floats = 0;
    PaDevType *d_types[2] = { nullptr, nullptr };
    double *d_pts = nullptr;
    PaPoints pts{};
    unsigned long long *d_acc = nullptr;
    size_t acc_words = 0;
    uint2 *d_exc = nullptr;
    unsigned int *d_exc_count = nullptr;    // [0] count, [1] sticky error flag, [2] verify failures, [3] total, [4] count of list 2
    unsigned long long *d_kstats = nullptr; // counters of the cell kernels, see PaPass::kstats
    pa_stats stats{};
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> kernel_events;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> f32_events;     // around every float32 launch of the cell-sorted RDF kernel
    std::vector<TypeGrid> grids;            // cell-sorted copies of the types the RDF fast path pairs
    bool use_cells = false;
    unsigned int exc_cap = kExcCap;         // entries of the exception list (both halves together)
    bool exact_ok = false;                  // resident batch: every coordinate difference is exact in fp64 (see upload)
};

namespace {

int upload_types_table(pa_session *s, int slot)
{
    std::vector<PaDevType> h(s->types.size());
    for (size_t t = 0; t < s->types.size(); ++t) {
        const TypeMeta &m = s->types[t];
        h[t].xyz = m.needed ? s->d_raw[slot][t] : nullptr;
        h[t].atom_mass = m.d_mass; h[t].atom_charge = m.d_charge;
        h[t].mass = m.mass; h[t].realcharge = m.realcharge;
        h[t].natoms = m.natoms; h[t].nmol = m.nmol; h[t].charge = m.charge; h[t].offset = m.offset;
    }
    PA_CK(cudaMemcpy(s->d_types[slot], h.data(), h.size() * sizeof(PaDevType), cudaMemcpyHostToDevice));
    return 0;
}

}  // namespace

int pa_multi_device_histogram(const pa_traj *traj, const pa_job *jobs, int32_t njobs, pa_exec ex,
                              int64_t *const *hists, int64_t *within_bin, int64_t *out_of_bounds, pa_stats *stats);
void pa_multi_device_trim();
int pa_sessions_reduce_to_first(pa_session *const *sess, int n, const std::vector<void *> *comms_in);
void pa_sessions_merge_stats(pa_session *const *sess, int n, bool by_frames, pa_stats *out);

extern "C" {

int pa_abi_version(void) { return PA_ABI_VERSION; }

void pa_trim(void)
{
    {
        std::lock_guard<std::mutex> lock(g_pool_mutex);
        for (const PinnedEntry &e : g_pinned_pool) cudaFreeHost(e.ptr);
        g_pinned_pool.clear();
    }
    pool_trim();
    pa_multi_device_trim();
    // the stream-ordered work buffers of pa_neighbour_pairs (default memory pool of every device this process has used)
    int n = 0, cur = 0;
    if (cudaGetDeviceCount(&n) == cudaSuccess && cudaGetDevice(&cur) == cudaSuccess) {
        for (int d = 0; d < n; ++d) {
            cudaMemPool_t pool;
            if (cudaDeviceGetDefaultMemPool(&pool, d) == cudaSuccess) cudaMemPoolTrimTo(pool, 0);
        }
        cudaSetDevice(cur);
    }
    cudaGetLastError();
}
const char *pa_last_error(void) { return g_last_error.c_str(); }

int pa_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    int ok = 0;
    for (int d = 0; d < n; ++d) {
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, d) == cudaSuccess && prop.major == 10) ++ok;
    }
    return ok;
}

// Every failure path of the set-up below returns through pa_session_create, which destroys the partially built
// session (pa_session_destroy tolerates null members), so no device buffer, pinned buffer, stream or event leaks.
static int session_create_impl(pa_session *s, const pa_traj *traj, const pa_job *jobs, int32_t njobs, const pa_exec &ex, int32_t max_frames);

int pa_session_create(const pa_traj *traj, const pa_job *jobs, int32_t njobs, const pa_exec *exec,
                      int32_t max_frames, pa_session **out)
{
    if (!traj || !jobs || njobs < 1 || !out || max_frames < 1 || traj->ntypes < 1) { pa::set_error("bad argument"); return PA_ERR_BAD_ARG; }
    pa_exec ex{};
    if (exec) ex = *exec;
    if (ex.shard_count < 1) ex.shard_count = 1;
    if (ex.tile_shard_count < 1) ex.tile_shard_count = 1;
    if (ex.tile_shard_rank < 0 || ex.tile_shard_rank >= ex.tile_shard_count) { pa::set_error("bad tile shard"); return PA_ERR_BAD_ARG; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) {
        cudaGetLastError();
        pa::set_error("no CUDA device (this path has no CPU fallback)");
        return PA_ERR_NO_DEVICE;
    }
    PA_CK(cudaSetDevice(ex.device));
    if (pa_kernels_selfcheck() != 0) {
        cudaGetLastError();
        pa::set_error("no sm_100a kernel image for this device (library is built for B200 only)");
        return PA_ERR_NO_DEVICE;
    }
    pa_session *s = new pa_session();
    const int rc = session_create_impl(s, traj, jobs, njobs, ex, max_frames);
    if (rc) { pa_session_destroy(s); return rc; }
    *out = s;
    return 0;
}

static int session_create_impl(pa_session *s, const pa_traj *traj, const pa_job *jobs, int32_t njobs, const pa_exec &ex, int32_t max_frames)
{
    s->device = ex.device; s->exec = ex;
    s->force_general = (ex.flags & PA_EXEC_FORCE_GENERAL) != 0;
    s->verify = (ex.flags & 0x100) != 0;
    if (ex.flags & 0x200) s->exc_cap = 2048;          // test hook: tiny exception list to exercise the overflow fallback
    cudaDeviceProp prop;
    PA_CK(cudaGetDeviceProperties(&prop, ex.device));
    s->sm_count = prop.multiProcessorCount;
    for (int k = 0; k < 3; ++k) {
        s->box.lo[k] = traj->box_lo[k]; s->box.hi[k] = traj->box_hi[k];
        s->box.L[k] = traj->box_hi[k] - traj->box_lo[k];
    }
    s->box.max_dist_sq = traj->max_dist_sq;
    s->box.wrap_trajectory = traj->wrap_trajectory; s->box.com_boost = traj->com_boost;
    // validate jobs, decide which types are needed
    s->types.resize(traj->ntypes);
    int64_t off = 0;
    for (int t = 0; t < traj->ntypes; ++t) {
        const pa_type &ty = traj->types[t];
        if (ty.natoms < 1 || ty.nmol < 1) { pa::set_error("molecule type with natoms/nmol < 1"); return PA_ERR_BAD_ARG; }
        TypeMeta &m = s->types[t];
        m.charge = ty.charge; m.natoms = ty.natoms; m.nmol = ty.nmol; m.offset = (int32_t)off;
        m.mass = ty.mass; m.realcharge = (double)ty.realcharge; m.needed = false;
        off += ty.nmol;
    }
    s->M = off;
    for (int j = 0; j < njobs; ++j) {
        const pa_job &jb = jobs[j];
        if (jb.ref_type < 1 || jb.ref_type > traj->ntypes || jb.obs_type > traj->ntypes) {
            pa::set_error("molecule type index out of range (error 33)"); return PA_E_TYPE_INVALID;
        }
        if (jb.bin_a < 1 || jb.bin_a > 1000 || jb.bin_b < 1 || jb.bin_b > 1000 || jb.sampling_interval < 1 ||
            (jb.mode != PA_MODE_PDF && jb.mode != PA_MODE_CDF && jb.mode != PA_MODE_CHARGE_ARM)) {
            pa::set_error("invalid job (bins 1..1000, sampling_interval >= 1, mode cdf|pdf|charge_arm)"); return PA_ERR_BAD_ARG;
        }
        s->types[jb.ref_type - 1].needed = true;
        if (jb.mode == PA_MODE_CHARGE_ARM) continue;                // intramolecular: no observed types
        if (jb.obs_type < 1) for (auto &m : s->types) m.needed = true;
        else s->types[jb.obs_type - 1].needed = true;
        if (jb.use_coc) s->need_coc = true;
    }
    // frames per batch: slots must stay addressable with 32 bits (exception list entries)
    const int64_t cap_frames = std::max<int64_t>(1, (int64_t)0x7fffffff / std::max<int64_t>(1, s->M));
    s->max_frames = (int)std::min<int64_t>(max_frames, cap_frames);
    PA_CK(cudaStreamCreateWithFlags(&s->s_compute, cudaStreamNonBlocking));
    PA_CK(cudaStreamCreateWithFlags(&s->s_copy, cudaStreamNonBlocking));
    for (int b = 0; b < 2; ++b) {
        PA_CK(cudaEventCreateWithFlags(&s->ev_h2d[b], cudaEventDisableTiming));
        PA_CK(cudaEventCreateWithFlags(&s->ev_done[b], cudaEventDisableTiming));
    }
    // per-type constants and raw buffers
    size_t pinned = 0;
    for (int t = 0; t < traj->ntypes; ++t) {
        TypeMeta &m = s->types[t];
        const pa_type &ty = traj->types[t];
        PA_CK(pool_malloc(&m.d_mass, sizeof(double) * m.natoms));
        PA_CK(pool_malloc(&m.d_charge, sizeof(double) * m.natoms));
        std::vector<double> zeros(m.natoms, 0.0);
        PA_CK(cudaMemcpy(m.d_mass, ty.atom_mass ? ty.atom_mass : zeros.data(), sizeof(double) * m.natoms, cudaMemcpyHostToDevice));
        PA_CK(cudaMemcpy(m.d_charge, ty.atom_charge ? ty.atom_charge : zeros.data(), sizeof(double) * m.natoms, cudaMemcpyHostToDevice));
        if (m.needed) pinned += m.frame_floats() * (size_t)s->max_frames;
    }
    s->pinned_floats = pinned;
    for (int b = 0; b < 2; ++b) {
        s->d_raw[b].assign(traj->ntypes, nullptr);
        for (int t = 0; t < traj->ntypes; ++t)
            if (s->types[t].needed)
                PA_CK(pool_malloc(&s->d_raw[b][t], sizeof(float) * s->types[t].frame_floats() * (size_t)s->max_frames));
        s->h_pinned_bytes[b] = sizeof(float) * std::max<size_t>(pinned, 1);
        s->h_pinned[b] = static_cast<float *>(pinned_pool_take(s->h_pinned_bytes[b]));
        if (!s->h_pinned[b]) PA_CK(cudaMallocHost(&s->h_pinned[b], s->h_pinned_bytes[b]));
        PA_CK(pool_malloc(&s->d_types[b], sizeof(PaDevType) * traj->ntypes));
        int rc = upload_types_table(s, b);
        if (rc) return rc;
    }
    // points SoA
    const size_t nslots = (size_t)s->max_frames * (size_t)s->M;
    const int narr = s->need_coc ? 12 : 9;
    PA_CK(pool_malloc(&s->d_pts, sizeof(double) * nslots * narr));
    double *b0 = s->d_pts;
    s->pts.px = b0; s->pts.py = b0 + nslots; s->pts.pz = b0 + 2 * nslots;
    s->pts.wx = b0 + 3 * nslots; s->pts.wy = b0 + 4 * nslots; s->pts.wz = b0 + 5 * nslots;
    s->pts.cx = b0 + 6 * nslots; s->pts.cy = b0 + 7 * nslots; s->pts.cz = b0 + 8 * nslots;
    if (s->need_coc) { s->pts.qx = b0 + 9 * nslots; s->pts.qy = b0 + 10 * nslots; s->pts.qz = b0 + 11 * nslots; }
    else s->pts.qx = s->pts.qy = s->pts.qz = nullptr;
    s->pts.M = s->M;
    // jobs: tables + accumulators
    s->jobs.resize(njobs);
    size_t words = 0;
    for (int j = 0; j < njobs; ++j) {
        JobState &js = s->jobs[j];
        js.job = jobs[j];
        js.nbins = (size_t)jobs[j].bin_a * jobs[j].bin_b;
        js.acc_off = words;
        words += js.nbins + 2;
    }
    s->acc_words = words;
    PA_CK(pool_malloc(&s->d_acc, sizeof(unsigned long long) * words));
    PA_CK(cudaMemset(s->d_acc, 0, sizeof(unsigned long long) * words));
    PA_CK(pool_malloc(&s->d_exc, sizeof(uint2) * kExcCap));
    PA_CK(pool_malloc(&s->d_exc_count, sizeof(unsigned int) * 8));
    PA_CK(cudaMemset(s->d_exc_count, 0, sizeof(unsigned int) * 8));
    PA_CK(pool_malloc(&s->d_kstats, sizeof(unsigned long long) * 8));
    PA_CK(cudaMemset(s->d_kstats, 0, sizeof(unsigned long long) * 8));
    for (int j = 0; j < njobs; ++j) {
        JobState &js = s->jobs[j];
        const pa_job &jb = js.job;
        std::vector<unsigned long long> thr_a;
        std::vector<double> thr_b;
        if (jb.mode == PA_MODE_CHARGE_ARM) pa::build_thr_charge_arm(jb, thr_a, js.tiny_bits);
        else pa::build_thr_a(jb, traj->max_dist_sq, thr_a);
        PA_CK(pool_malloc(&js.d_thr_a, sizeof(unsigned long long) * thr_a.size()));
        PA_CK(cudaMemcpy(js.d_thr_a, thr_a.data(), sizeof(unsigned long long) * thr_a.size(), cudaMemcpyHostToDevice));
        if (jb.mode != PA_MODE_CDF) {
            pa::build_thr_b(jb, thr_b);
            PA_CK(pool_malloc(&js.d_thr_b, sizeof(double) * thr_b.size()));
            PA_CK(cudaMemcpy(js.d_thr_b, thr_b.data(), sizeof(double) * thr_b.size(), cudaMemcpyHostToDevice));
        }
        PaDevJob &d = js.dj;
        memset(&d, 0, sizeof d);
        d.mode = jb.mode; d.bin_a = jb.bin_a; d.bin_b = jb.bin_b; d.use_coc = jb.use_coc;
        for (int k = 0; k < 3; ++k) d.ref_vec[k] = jb.ref_vec[k];
        for (int k = 0; k < 9; ++k) d.rot[k] = jb.rot[k];
        d.maxdist = jb.maxdist; d.step_a = jb.step_a; d.step_b = jb.step_b; d.shift = jb.shift;
        d.maxdist_sq_f = jb.maxdist * jb.maxdist;
        d.thr_a = js.d_thr_a; d.thr_b = js.d_thr_b;
        d.cut_bits = thr_a[(size_t)(jb.mode == PA_MODE_PDF ? jb.bin_a : 0) + 1];
        d.guess_scale = 1.0f / jb.step_a;
        d.z_sign = jb.ref_vec[2] > 0 ? 1 : -1;
        d.acc = s->d_acc + js.acc_off;
        js.fast = !s->force_general && pa::job_is_rdf_fast(jb, *traj);   // false for cdf and charge_arm
        {
            // variant 0 needs the cutoff within 3e-7 (relative) of the boundary of bin na (REAL(4) step_a * bin_a is
            // maxdist up to a few float32 roundings): the class brackets are wider than the error budget by the actual
            // distance (delta_top), so every pair within that distance of the boundary -- binned in the last bin, "inside the cutoff with a-bin out
            // of range", or beyond the cutoff -- is decided by the exact fp64 path, whichever side the cutoff lies on.
            js.cells_variant = 1;
            js.delta_top = 0.0;
            if (jb.mode == PA_MODE_PDF && thr_a[(size_t)jb.bin_a + 1] != ~0ull) {
                double cutd, nad; memcpy(&cutd, &thr_a[(size_t)jb.bin_a + 1], 8); memcpy(&nad, &thr_a[(size_t)jb.bin_a], 8);
                const double xcut = std::sqrt(cutd) / (double)jb.step_a, xna = std::sqrt(nad) / (double)jb.step_a;   // thr_a[na] <= cut
                const double delta = std::max(std::fabs(xcut - jb.bin_a), std::fabs(xna - jb.bin_a)) / jb.bin_a;
                if (delta <= 3e-7) { js.cells_variant = 0; js.delta_top = delta; }
            }
            js.f32_cnt = 0;
            if (js.cells_variant == 0 && !(ex.flags & PA_EXEC_NO_F32)) {
                // classes a float32 segment can produce: any pair of a single-image run is closer than the box
                // diagonal; padded observed molecules sit at maxdist*1.001 + 96 (+32 for the tile radius)
                const double diag = std::sqrt(s->box.L[0] * s->box.L[0] + s->box.L[1] * s->box.L[1] + s->box.L[2] * s->box.L[2]);
                const double far = std::max(diag, (double)jb.maxdist * 1.001 + 128.0);
                const double ncl = far * (double)d.guess_scale * 1.00001 + 4.0;
                bool small = true;                                  // queue entries of the float32 launch hold sorted positions in 24 bits
                for (int t = 0; t < traj->ntypes; ++t) if (traj->types[t].nmol >= (1 << 24)) small = false;
                if (ncl + jb.bin_a + 8 <= 12288.0 && small) js.f32_cnt = (int)ncl;
            }
        }
    }
    // cell-sorted path: both partners of a pass need >= kCellsMinMolecules molecules (the grid never has fewer than 5 x 5 rows
    // of 20 cells; below ~2000 molecules a warp tile of 64 spans most of a row and nothing is pruned)
    s->grids.resize(traj->ntypes);
    // (with wrap_trajectory the positions are used as they are: the grid then needs every centre of mass inside the box,
    //  which pa_cell_count_kernel checks -- otherwise the sticky flag sends the call to the brute-force kernels)
    s->use_cells = !(ex.flags & PA_EXEC_NO_CELLS);
    if (s->use_cells) {
        for (const JobState &js : s->jobs) {
            const int r = js.job.ref_type - 1;
            if (js.job.mode == PA_MODE_CHARGE_ARM) continue;
            for (int o = 0; o < traj->ntypes; ++o) {
                if (js.job.obs_type >= 1 && o != js.job.obs_type - 1) continue;
                // ... and a pass below ~2e7 pairs per frame is finished by the brute-force tiles before the sort of a frame is
                if (s->types[r].nmol >= kCellsMinMolecules && s->types[o].nmol >= kCellsMinMolecules &&
                    (double)s->types[r].nmol * (double)s->types[o].nmol >= kCellsMinPairs) { s->grids[r].enabled = true; s->grids[o].enabled = true; }
            }
        }
        for (int t = 0; t < traj->ntypes; ++t) {
            TypeGrid &tg = s->grids[t];
            if (!tg.enabled) continue;
            PaCellGrid &g = tg.g;
            pa_cell_grid_dims(s->types[t].nmol, g.g);
            g.gmax = std::max(g.g[0], std::max(g.g[1], g.g[2]));
            g.ncell = g.g[0] * g.g[1] * g.g[2]; g.nmol = s->types[t].nmol; g.type_off = s->types[t].offset;
            for (int k = 0; k < 3; ++k) g.inv_w[k] = (double)g.g[k] / s->box.L[k];
            const size_t nf = (size_t)s->max_frames, nm = (size_t)g.nmol;
            PA_CK(pool_malloc(&g.cell_start, sizeof(int) * nf * (g.ncell + 1)));
            PA_CK(pool_malloc(&g.cell_fill, sizeof(int) * nf * g.ncell));
            PA_CK(pool_malloc(&g.cid, sizeof(int) * nf * nm));
            PA_CK(pool_malloc(&g.sx, sizeof(double) * nf * nm));
            PA_CK(pool_malloc(&g.sy, sizeof(double) * nf * nm));
            PA_CK(pool_malloc(&g.sz, sizeof(double) * nf * nm));
            PA_CK(pool_malloc(&g.sorig, sizeof(int) * nf * nm));
            PA_CK(pool_malloc(&g.sflag, nf * nm));
            PA_CK(pool_malloc(&g.lay_lo, sizeof(unsigned long long) * nf * 3 * g.gmax));
            PA_CK(pool_malloc(&g.lay_hi, sizeof(unsigned long long) * nf * 3 * g.gmax));
            // every cell yields at most one tile plus one extra per 256 molecules
            tg.tile_cap = (unsigned int)std::min<size_t>(nf * ((size_t)g.ncell + nm / 256 + 1), 0x7fffffffu);
            PA_CK(pool_malloc(&tg.tiles, sizeof(PaTile) * tg.tile_cap));
            // warp tiles: every row of cells yields at most one partial piece plus one per 64 molecules
            tg.wtile_cap = (unsigned int)std::min<size_t>(nf * ((size_t)g.ncell + nm / 64 + 1), 0x7fffffffu);
            PA_CK(pool_malloc(&tg.wtiles, sizeof(PaTile) * tg.wtile_cap));
            PA_CK(pool_malloc(&tg.row_off, sizeof(int) * (nf * (size_t)g.g[1] * g.g[2] + 1)));
            PA_CK(pool_malloc(&tg.counters, sizeof(unsigned int) * 4));
            PA_CK(cudaMemset(tg.counters, 0, sizeof(unsigned int) * 4));
        }
    }
    // the table uploads and memsets above ran on the legacy default stream, which does not synchronise with the
    // session's non-blocking streams: make sure all of them have landed before the first kernel can start
    PA_CK(cudaDeviceSynchronize());
    return 0;
}

int pa_session_upload(pa_session *s, const pa_traj *traj, int32_t first_step, int32_t nframes, int32_t step_stride)
{
    if (!s || !traj || nframes < 1 || nframes > s->max_frames || step_stride < 1 || first_step < 0 ||
        (int64_t)first_step + (int64_t)(nframes - 1) * step_stride >= traj->nsteps || traj->ntypes != (int)s->types.size()) {
        pa::set_error("pa_session_upload: bad frame range"); return PA_ERR_BAD_ARG;
    }
    PA_CK(cudaSetDevice(s->device));
    const int slot = s->cur ^ 1;
    // the pinned buffer of this slot may still be read by its previous H2D copy, and the device
    // buffer by the kernels of the batch that used it
    if (s->slot_used[slot]) {
        PA_CK(cudaEventSynchronize(s->ev_h2d[slot]));
        PA_CK(cudaStreamWaitEvent(s->s_copy, s->ev_done[slot], 0));
    }
    // Exactness certificate for the symmetric passes of the cell kernel: one-atom molecules (centre of
    // mass = the float32 coordinate), every coordinate 0 or 2^-17 <= |x| < 2^10, box bounds likewise and
    // float32-valued.  Then all wrapped positions and image shifts are multiples of 2^-40 below 2^11, every
    // (p2+s)-p1 needs at most 52 bits and is computed exactly, hence d2(i,j) == d2(j,i) bit for bit.
    bool exact = s->use_cells && s->box.com_boost;
    for (int k = 0; k < 3 && exact; ++k) {
        for (double v : { s->box.lo[k], s->box.hi[k] })
            if (!((double)(float)v == v && (v == 0.0 || (std::fabs(v) >= 7.62939453125e-06 && std::fabs(v) < 1024.0)))) exact = false;
    }
    // copy into the page-locked staging buffer (and scan for the exactness certificate) frame by frame; batches of tens of
    // MB are split over a few host threads -- one thread moves ~10 GB/s, the PCIe link takes five times that
    struct Seg { float *dst; const float *src; size_t n; };
    std::vector<Seg> segs;
    size_t poff = 0, total_floats = 0;
    for (size_t t = 0; t < s->types.size(); ++t) {
        const TypeMeta &m = s->types[t];
        if (!m.needed) continue;
        const pa_type &ty = traj->types[t];
        if (!ty.xyz || ty.natoms != m.natoms || ty.nmol != m.nmol) { pa::set_error("pa_session_upload: trajectory does not match the session"); return PA_ERR_BAD_ARG; }
        const size_t ff = m.frame_floats();
        float *dst = s->h_pinned[slot] + poff;
        for (int f = 0; f < nframes; ++f)
            segs.push_back({ dst + (size_t)f * ff, ty.xyz + (size_t)(first_step + (int64_t)f * step_stride) * ff, ff });
        total_floats += ff * (size_t)nframes;
        poff += ff * (size_t)s->max_frames;
    }
    {
        const bool scan = exact;
        auto work = [&segs, scan](size_t b, size_t e) -> bool {
            unsigned int bad = 0;
            for (size_t k = b; k < e; ++k) {
                memcpy(segs[k].dst, segs[k].src, segs[k].n * sizeof(float));
                if (scan) {
                    const uint32_t *u = reinterpret_cast<const uint32_t *>(segs[k].src);
                    for (size_t i = 0; i < segs[k].n; ++i) {
                        const uint32_t a = u[i] & 0x7fffffffu;                     // |x| as bits: 2^-17 = 0x37000000, 2^10 = 0x44800000
                        bad |= (unsigned)((a != 0u) & ((a < 0x37000000u) | (a >= 0x44800000u)));
                    }
                }
            }
            return bad != 0;
        };
        const size_t nthreads = std::min<size_t>({ (size_t)8, std::max<size_t>(1, std::thread::hardware_concurrency() / 2), segs.size(),
                                                   std::max<size_t>(1, total_floats * sizeof(float) >> 23) });     // >= 8 MB per thread
        bool bad = false;
        if (nthreads <= 1) bad = work(0, segs.size());
        else {
            std::vector<char> flags(nthreads, 0);
            std::vector<std::thread> th;
            for (size_t k = 0; k < nthreads; ++k)
                th.emplace_back([&, k] { flags[k] = work(segs.size() * k / nthreads, segs.size() * (k + 1) / nthreads) ? 1 : 0; });
            for (auto &t : th) t.join();
            for (char f : flags) bad = bad || f;
        }
        if (bad) exact = false;
    }
    poff = 0;
    for (size_t t = 0; t < s->types.size(); ++t) {
        const TypeMeta &m = s->types[t];
        if (!m.needed) continue;
        const size_t ff = m.frame_floats();
        PA_CK(cudaMemcpyAsync(s->d_raw[slot][t], s->h_pinned[slot] + poff, sizeof(float) * ff * nframes, cudaMemcpyHostToDevice, s->s_copy));
        s->stats.h2d_bytes += (int64_t)(sizeof(float) * ff * nframes);
        poff += ff * (size_t)s->max_frames;
    }
    PA_CK(cudaEventRecord(s->ev_h2d[slot], s->s_copy));
    s->slot_used[slot] = true;
    s->cur = slot;
    s->nframes = nframes;
    s->exact_ok = exact;
    return 0;
}

int pa_session_run(pa_session *s)
{
    if (!s) return PA_ERR_BAD_ARG;
    return pa_session_run_frames(s, 0, s->nframes);
}

int pa_session_run_frames(pa_session *s, int32_t first, int32_t count)
{
    if (!s || s->nframes < 1 || first < 0 || count < 1 || first + count > s->nframes) {
        pa::set_error("pa_session_run: frame range not resident"); return PA_ERR_BAD_ARG;
    }
    PA_CK(cudaSetDevice(s->device));
    const int slot = s->cur;
    PA_CK(cudaStreamWaitEvent(s->s_compute, s->ev_h2d[slot], 0));
    cudaEvent_t e0, e1;
    PA_CK(cudaEventCreate(&e0)); PA_CK(cudaEventCreate(&e1));
    PA_CK(cudaEventRecord(e0, s->s_compute));
    PaPrepareArgs pa{};
    pa.types = s->d_types[slot]; pa.ntypes = (int)s->types.size(); pa.nframes = count; pa.frame0 = first;
    pa.pts = s->pts; pa.box = s->box; pa.need_coc = s->need_coc ? 1 : 0;
    PA_CK((cudaError_t)pa_launch_prepare(pa, s->s_compute));
    s->stats.kernel_launches += 1;
    for (TypeGrid &tg : s->grids) {
        if (!tg.enabled) continue;
        PA_CK((cudaError_t)pa_launch_cell_sort(tg.g, s->pts, s->box, first, count, s->d_exc_count + 1, s->s_compute));
        tg.tiles_ready = tg.wtiles_ready = false;
        s->stats.kernel_launches += 5;
    }
    const long long target_items = (long long)s->sm_count * 16;
    // ---- plan: one pass per (job, observed type), Distribution:727-735 ---------------------------------
    struct Plan { int job, r, o; bool cells; bool sym; int partner; bool skip; };
    std::vector<Plan> plan;
    for (size_t j = 0; j < s->jobs.size(); ++j) {
        const pa_job &jb = s->jobs[j].job;
        const int r = jb.ref_type - 1;
        if (jb.mode == PA_MODE_CHARGE_ARM) {                        // fill_histogram_charge_arm, Distribution:847-895
            if (s->exec.tile_shard_rank != 0) continue;             // O(N) per frame: the first tile shard does it
            PaChargeArmArgs ca{};
            ca.types = s->d_types[slot]; ca.type = r; ca.frame0 = first; ca.nframes = count;
            ca.normalise_clm = jb.normalise_clm; ca.com_boost = s->box.com_boost;
            ca.job = s->jobs[j].dj; ca.tiny_bits = s->jobs[j].tiny_bits;
            PA_CK((cudaError_t)pa_launch_charge_arm(ca, s->s_compute));
            s->stats.kernel_launches += 1;
            s->stats.pairs_evaluated += (int64_t)count * s->types[r].nmol;   // molecules, not pairs, in this mode
            continue;
        }
        const int o_first = jb.obs_type < 1 ? 0 : jb.obs_type - 1;
        const int o_last = jb.obs_type < 1 ? (int)s->types.size() - 1 : jb.obs_type - 1;
        for (int o = o_first; o <= o_last; ++o) {
            Plan pl{ (int)j, r, o, false, false, -1, false };
            pl.cells = s->grids[r].enabled && s->grids[o].enabled;
            plan.push_back(pl);
        }
    }
    // When every coordinate difference is exact in fp64 (certificate computed at upload time),
    // d2(i,j) == d2(j,i) bit for bit: a same-type pass evaluates every unordered pair once, and the pass
    // (A around B) of one job doubles as the pass (B around A) of another job with the same bins.
    if (s->exact_ok && !(s->exec.flags & PA_EXEC_NO_SYMMETRY)) {
        for (size_t a = 0; a < plan.size(); ++a) {
            Plan &pa_ = plan[a];
            if (!pa_.cells || pa_.skip || !s->jobs[pa_.job].fast) continue;
            if (pa_.r == pa_.o) { pa_.sym = true; continue; }
            if (pa_.partner >= 0) continue;
            for (size_t b2 = a + 1; b2 < plan.size(); ++b2) {
                Plan &pb = plan[b2];
                if (!pb.cells || pb.skip || pb.partner >= 0 || pb.r != pa_.o || pb.o != pa_.r || !s->jobs[pb.job].fast) continue;
                const pa_job &ja = s->jobs[pa_.job].job, &jb2 = s->jobs[pb.job].job;
                if (ja.bin_a != jb2.bin_a || ja.maxdist != jb2.maxdist || ja.step_a != jb2.step_a ||
                    s->jobs[pa_.job].cells_variant != s->jobs[pb.job].cells_variant) continue;
                pa_.partner = (int)b2; pb.skip = true;
                break;
            }
        }
    }
    auto fill_pass = [&](PaPass &p, const JobState &js, int r, int o) {
        const pa_job &jb = js.job;
        p.pts = s->pts; p.box = s->box; p.job = js.dj;
        p.nframes = count; p.frame0 = first;
        p.ref_off = s->types[r].offset; p.nref = s->types[r].nmol;
        p.obs_off = s->types[o].offset; p.nobs = s->types[o].nmol;
        p.same_type = (r == o) ? 1 : 0;
        p.weight = jb.weigh_charge ? (long long)s->types[o].charge : 1ll;   // Distribution:833-838
        p.exc_list = s->d_exc; p.exc_count = s->d_exc_count; p.exc_cap = s->exc_cap / 2; p.flags = s->d_exc_count;
        p.exc_list2 = s->d_exc + s->exc_cap / 2; p.exc_count2 = s->d_exc_count + 4;
        p.verify_fail = s->verify ? s->d_exc_count + 2 : nullptr;
        p.f32_cnt = js.f32_cnt;
        {
            // brackets of the float32 segments: the largest float <= (1-eps)/step_a and the smallest float >= (1+eps)/step_a,
            // computed in double so that the scale factor itself carries no rounding towards the inside; eps = error budget
            // 5.4e-7 (pa_cells.cu) + the distance of the cutoff from the boundary of the last bin + 0.1e-7
            const double eps = 5.5e-7 + js.delta_top;
            const double inv = 1.0 / (double)jb.step_a;
            float lo = (float)(inv * (1.0 - eps)), hi = (float)(inv * (1.0 + eps));
            if ((double)lo > inv * (1.0 - eps)) lo = std::nextafterf(lo, 0.0f);
            if ((double)hi < inv * (1.0 + eps)) hi = std::nextafterf(hi, INFINITY);
            // fp64-difference segments: truncation of d2 to 24 bits 0.6e-7, sqrt.approx 1.2e-7, reference chain 1.5e-7,
            // three float32 roundings inside the in-kernel scale factors 1.8e-7 = 5.1e-7
            p.eps64 = (float)(5.5e-7 + js.delta_top);
            p.f32_lo = lo; p.f32_hi = hi;
            // approximate distance from which floor(r~ * f32_lo) > bin_a even after the 2^-22 relative error of sqrt.approx
            const double dead = ((double)jb.bin_a + 1.0) / (double)lo * (1.0 + 1.0e-6);
            p.dead_r = (float)dead;
            if ((double)p.dead_r < dead) p.dead_r = std::nextafterf(p.dead_r, INFINITY);
        }
        p.tshard_rank = s->exec.tile_shard_rank; p.tshard_count = s->exec.tile_shard_count;
        p.kstats = s->d_kstats;
    };
    for (const Plan &pl : plan) {
        const JobState &js = s->jobs[pl.job];
        const int r = pl.r, o = pl.o;
        PaPass p{};
        fill_pass(p, js, r, o);
        if (s->exec.tile_shard_rank == 0)                           // the statistics of all tile shards add up to the whole
            s->stats.pairs_evaluated += (int64_t)p.nframes * ((int64_t)p.nref * p.nobs - (p.same_type ? p.nref : 0));
        if (pl.skip) continue;                                     // counted by its partner's launch
        const int ri = js.fast ? pa_rdf_ri(p.nref) : pa_general_ri(p.nref);
        const int ti = PA_TILE_THREADS * ri;
        p.n_itiles = (p.nref + ti - 1) / ti;
        const long long items0 = (long long)p.nframes * p.n_itiles;
        const int max_splits = (p.nobs + PA_J_CHUNK - 1) / PA_J_CHUNK;
        long long want = items0 >= target_items ? 1 : (target_items + items0 - 1) / items0;
        if (want > max_splits) want = max_splits;
        int len = (int)((p.nobs + want - 1) / want);
        len = ((len + PA_J_CHUNK - 1) / PA_J_CHUNK) * PA_J_CHUNK;
        p.jsplit_len = len;
        p.n_jsplits = (p.nobs + len - 1) / len;
        double cut_d;
        memcpy(&cut_d, &js.dj.cut_bits, 8);
        const bool pdf_f32 = pl.cells && !js.fast && !s->force_general && js.job.mode == PA_MODE_PDF && !js.job.use_coc &&
                             js.cells_variant == 0 && !(s->exec.flags & PA_EXEC_NO_F32) && cut_d < s->box.max_dist_sq &&
                             s->types[r].nmol < (1 << 24) && s->types[o].nmol < (1 << 24);
        if (pdf_f32) {
            // 2-D pdf job on warp tiles, both classes bracketed in float32 (pa_pdf_cells_kernel)
            TypeGrid &gr = s->grids[r];
            if (!gr.wtiles_ready) {
                PA_CK((cudaError_t)pa_launch_build_tiles(gr.g, count, true, gr.wtiles, gr.counters + 2, gr.row_off, gr.wtile_cap, s->s_compute));
                gr.wtiles_ready = true;
                s->stats.kernel_launches += 3;
            }
            const long long tiles_est = std::max<long long>(1, (long long)count * (p.nref / 60 + (long long)gr.g.g[1] * gr.g.g[2]));
            long long nzg = (40ll * s->sm_count * 16 + tiles_est - 1) / tiles_est;
            nzg = std::max<long long>(1, std::min<long long>(nzg, s->grids[o].g.g[2]));
            PA_CK((cudaError_t)pa_launch_pdf_cells(p, gr.g, s->grids[o].g, gr.wtiles, gr.counters + 2, gr.wtile_cap, gr.counters + 1,
                                                   (int)nzg, s->sm_count, s->s_compute));
            s->stats.kernel_launches += 1;
        } else if (pl.cells && !js.fast) {
            TypeGrid &gr = s->grids[r];
            if (!gr.tiles_ready) {
                PA_CK((cudaError_t)pa_launch_build_tiles(gr.g, count, false, gr.tiles, gr.counters, gr.row_off, gr.tile_cap, s->s_compute));
                gr.tiles_ready = true;
                s->stats.kernel_launches += 3;
            }
            const long long tiles_est = std::max<long long>(1, (long long)count * (p.nref / 230 + 1));
            long long nzg = (20ll * s->sm_count * 5 + tiles_est - 1) / tiles_est;
            nzg = std::max<long long>(1, std::min<long long>(nzg, s->grids[o].g.g[2]));
            PA_CK((cudaError_t)pa_launch_general_cells(p, gr.g, s->grids[o].g, gr.tiles, gr.counters, gr.tile_cap, gr.counters + 1,
                                                       (int)nzg, s->sm_count, s->s_compute));
            s->stats.kernel_launches += 1;
        } else if (pl.cells) {
            TypeGrid &gr = s->grids[r];
            if (!gr.wtiles_ready) {
                PA_CK((cudaError_t)pa_launch_build_tiles(gr.g, count, true, gr.wtiles, gr.counters + 2, gr.row_off, gr.wtile_cap, s->s_compute));
                gr.wtiles_ready = true;
                s->stats.kernel_launches += 3;
            }
            PaPass p2{};
            p.sym = pl.sym ? 1 : 0;
            if (pl.partner >= 0) {
                const Plan &pb = plan[pl.partner];
                fill_pass(p2, s->jobs[pb.job], pb.r, pb.o);
                p.acc2 = p2.job.acc; p.weight2 = p2.weight;
                p2.exc_list = p.exc_list2; p2.exc_count = p.exc_count2;     // reversed pairs go through the partner's job
            }
            // work items = (warp tile, slab of observed z-layers): enough of them for ~40 per resident warp
            const long long tiles_est = std::max<long long>(1, (long long)count * (p.nref / 60 + (long long)gr.g.g[1] * gr.g.g[2]));
            long long nzg = (40ll * s->sm_count * 24 + tiles_est - 1) / tiles_est;
            nzg = std::max<long long>(1, std::min<long long>(nzg, s->grids[o].g.g[2]));
            cudaEvent_t f0 = nullptr, f1 = nullptr;
            if (js.cells_variant == 0 && js.f32_cnt > 0) {
                PA_CK(cudaEventCreate(&f0)); PA_CK(cudaEventCreate(&f1));
                s->f32_events.emplace_back(f0, f1);
            }
            PA_CK((cudaError_t)pa_launch_rdf_cells(p, gr.g, s->grids[o].g, gr.wtiles, gr.counters + 2, gr.wtile_cap, gr.counters + 1,
                                                   (int)nzg, js.cells_variant, s->sm_count, s->s_compute, f0, f1));
            PA_CK((cudaError_t)pa_launch_exceptions(p, 0, s->s_compute));
            s->stats.kernel_launches += 3 + ((js.cells_variant == 0 && js.f32_cnt > 0) ? 1 : 0);   // float32 runs and the rest are two launches
            if (pl.sym) {                                           // reversed pairs of the same job
                PaPass ps = p; ps.exc_list = p.exc_list2; ps.exc_count = p.exc_count2;
                PA_CK((cudaError_t)pa_launch_exceptions(ps, 0, s->s_compute));
                s->stats.kernel_launches += 2;
            } else if (pl.partner >= 0) {
                PA_CK((cudaError_t)pa_launch_exceptions(p2, 0, s->s_compute));
                s->stats.kernel_launches += 2;
            }
        } else if (js.fast) {
            PA_CK((cudaError_t)pa_launch_rdf_fast(p, s->sm_count, s->s_compute));
            PA_CK((cudaError_t)pa_launch_exceptions(p, 0, s->s_compute));
            s->stats.kernel_launches += 3;
        } else {
            PA_CK((cudaError_t)pa_launch_general(p, s->sm_count, s->s_compute));
            s->stats.kernel_launches += 1;
        }
    }
    PA_CK(cudaEventRecord(e1, s->s_compute));
    PA_CK(cudaEventRecord(s->ev_done[slot], s->s_compute));
    s->kernel_events.emplace_back(e0, e1);
    s->stats.frames_done += count;
    return 0;
}

int pa_session_stats(pa_session *s, pa_stats *st)
{
    if (!s || !st) return PA_ERR_BAD_ARG;
    PA_CK(cudaSetDevice(s->device));
    PA_CK(cudaStreamSynchronize(s->s_compute));
    for (auto &ev : s->kernel_events) {
        float ms = 0.f;
        PA_CK(cudaEventElapsedTime(&ms, ev.first, ev.second));
        s->stats.kernel_ms += ms;
        cudaEventDestroy(ev.first); cudaEventDestroy(ev.second);
    }
    s->kernel_events.clear();
    for (auto &ev : s->f32_events) {
        float ms = 0.f;
        PA_CK(cudaEventElapsedTime(&ms, ev.first, ev.second));
        s->stats.dominant_ms += ms; s->stats.dominant_launches += 1;
        cudaEventDestroy(ev.first); cudaEventDestroy(ev.second);
    }
    s->f32_events.clear();
    unsigned long long ks[8];
    PA_CK(cudaMemcpy(ks, s->d_kstats, sizeof ks, cudaMemcpyDeviceToHost));
    s->stats.cell_pair_evals = (int64_t)(ks[0] + ks[4]); s->stats.smem_atomics = (int64_t)(ks[1] + ks[5]);
    s->stats.unique_pairs_counted = (int64_t)ks[2];
    s->stats.dominant_pair_evals = (int64_t)ks[0]; s->stats.dominant_atomics = (int64_t)ks[1];
    *st = s->stats;
    return 0;
}

int pa_session_download(pa_session *s, int64_t *const *hists, int64_t *within_bin, int64_t *out_of_bounds, int32_t reset)
{
    if (!s) return PA_ERR_BAD_ARG;
    PA_CK(cudaSetDevice(s->device));
    PA_CK(cudaStreamSynchronize(s->s_compute));
    std::vector<unsigned long long> h(s->acc_words);
    unsigned int flags[4];
    PA_CK(cudaMemcpy(h.data(), s->d_acc, sizeof(unsigned long long) * s->acc_words, cudaMemcpyDeviceToHost));
    PA_CK(cudaMemcpy(flags, s->d_exc_count, sizeof flags, cudaMemcpyDeviceToHost));
    s->stats.d2h_bytes += (int64_t)(sizeof(unsigned long long) * s->acc_words + sizeof flags);
    s->stats.exception_pairs = (int64_t)flags[3];
    int64_t binned = 0;
    for (size_t j = 0; j < s->jobs.size(); ++j) {
        const JobState &js = s->jobs[j];
        if (hists && hists[j]) memcpy(hists[j], h.data() + js.acc_off, sizeof(int64_t) * js.nbins);
        if (within_bin) within_bin[j] = (int64_t)h[js.acc_off + js.nbins];
        if (out_of_bounds) out_of_bounds[j] = (int64_t)h[js.acc_off + js.nbins + 1];
        binned += (int64_t)h[js.acc_off + js.nbins];
    }
    s->stats.pairs_binned = binned;
    if (reset) {
        // stream-ordered with the kernels of the next run (a plain cudaMemset runs on the legacy default stream, which
        // the non-blocking compute stream does not wait for)
        PA_CK(cudaMemsetAsync(s->d_acc, 0, sizeof(unsigned long long) * s->acc_words, s->s_compute));
        PA_CK(cudaMemsetAsync(s->d_exc_count, 0, sizeof(unsigned int) * 8, s->s_compute));
    }
    if (flags[1]) {
        pa::set_error("exception list overflow: too many pairs (anti)parallel to the reference axis; rerun with PA_EXEC_FORCE_GENERAL");
        return PA_ERR_UNSUPPORTED;
    }
    if (flags[2]) {
        pa::set_error("internal: class-seed invariant violated " + std::to_string(flags[2]) + " times");
        return PA_ERR_CUDA;
    }
    return 0;
}

int pa_session_device_accumulators(pa_session *s, void **dev_ptr, int64_t *nwords)
{
    if (!s || !dev_ptr || !nwords) return PA_ERR_BAD_ARG;
    PA_CK(cudaSetDevice(s->device));
    PA_CK(cudaStreamSynchronize(s->s_compute));
    *dev_ptr = s->d_acc; *nwords = (int64_t)s->acc_words;
    return 0;
}

void pa_session_destroy(pa_session *s)
{
    if (!s) return;
    cudaSetDevice(s->device);
    if (s->s_compute) cudaStreamSynchronize(s->s_compute);
    if (s->s_copy) cudaStreamSynchronize(s->s_copy);
    for (auto &ev : s->kernel_events) { cudaEventDestroy(ev.first); cudaEventDestroy(ev.second); }
    for (auto &ev : s->f32_events) { cudaEventDestroy(ev.first); cudaEventDestroy(ev.second); }
    for (auto &m : s->types) { pool_free(m.d_mass); pool_free(m.d_charge); }
    for (int b = 0; b < 2; ++b) {
        for (float *p : s->d_raw[b]) pool_free(p);
        if (s->h_pinned[b]) pinned_pool_give(s->h_pinned[b], s->h_pinned_bytes[b]);
        pool_free(s->d_types[b]);
        if (s->ev_h2d[b]) cudaEventDestroy(s->ev_h2d[b]);
        if (s->ev_done[b]) cudaEventDestroy(s->ev_done[b]);
    }
    for (auto &js : s->jobs) { pool_free(js.d_thr_a); pool_free(js.d_thr_b); }
    for (auto &tg : s->grids) {
        if (!tg.enabled) continue;
        pool_free(tg.g.cell_start); pool_free(tg.g.cell_fill); pool_free(tg.g.cid); pool_free(tg.g.sx); pool_free(tg.g.sy); pool_free(tg.g.sz);
        pool_free(tg.g.sorig); pool_free(tg.g.sflag); pool_free(tg.g.lay_lo); pool_free(tg.g.lay_hi); pool_free(tg.tiles); pool_free(tg.wtiles); pool_free(tg.row_off); pool_free(tg.counters);
    }
    pool_free(s->d_pts); pool_free(s->d_acc); pool_free(s->d_exc); pool_free(s->d_exc_count); pool_free(s->d_kstats);
    if (s->s_compute) cudaStreamDestroy(s->s_compute);
    if (s->s_copy) cudaStreamDestroy(s->s_copy);
    delete s;
}

// ---------------------------------------------------------------------------
// one-shot entry points

static int choose_batch(const pa_traj *traj, const pa_exec *ex, int nsamp_local)
{
    if (ex && ex->frames_per_batch > 0) return std::min(ex->frames_per_batch, std::max(nsamp_local, 1));
    int64_t atoms = 0, M = 0;
    for (int t = 0; t < traj->ntypes; ++t) { atoms += (int64_t)traj->types[t].natoms * traj->types[t].nmol; M += traj->types[t].nmol; }
    // ~2^22 atoms of raw coordinates (50 MB of page-locked staging per buffer) and ~2^22 molecule slots (72-96 B each) per
    // batch: small enough that page-locking the staging buffers and the first host copy do not delay the first kernel by
    // more than a few ms, and that the copy of batch k+1 hides behind the kernels of batch k; large enough (>= 8 batches
    // when the trajectory allows) that the ~40 launches per batch do not matter
    int64_t b = std::min<int64_t>((int64_t(1) << 22) / std::max<int64_t>(atoms, 1), (int64_t(1) << 22) / std::max<int64_t>(M, 1));
    b = std::max<int64_t>(1, std::min<int64_t>(b, nsamp_local));
    return (int)b;
}

// This process's share of the work on ONE device (ex is already sharded): create a session, stream the frames through
// it and leave the results in its device accumulators.  A sticky device flag (exception list overflow / boxes too wide
// for the cell path / non-finite coordinates) sends the share through the kernel ladder cells -> brute force -> general
// kernel; every rung gives the same integers, so different devices may end on different rungs.
static int run_share_on_device(const pa_traj *traj, const pa_job *jobs, int32_t njobs, pa_exec ex, pa_session **out, float *total_ms)
{
    const int dt = jobs[0].sampling_interval;
    const int nsamp = (traj->nsteps + dt - 1) / dt;
    const int nlocal = nsamp > ex.shard_rank ? (nsamp - ex.shard_rank + ex.shard_count - 1) / ex.shard_count : 0;
    for (int attempt = 0; attempt < 3; ++attempt) {
        const int batch = choose_batch(traj, &ex, nlocal);
        pa_session *s = nullptr;
        const bool dbg_t = getenv("PA_DEBUG_TIMING") != nullptr;
        const auto w0 = std::chrono::steady_clock::now();
        auto lap = [&](const char *what) {
            if (dbg_t) fprintf(stderr, "[prealpha_b200] dev %d %-10s %8.2f ms\n", ex.device, what, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - w0).count());
        };
        int rc = pa_session_create(traj, jobs, njobs, &ex, batch, &s);
        if (rc) return rc;
        lap("create");
        cudaEvent_t t0, t1;
        cudaEventCreate(&t0); cudaEventCreate(&t1);
        cudaEventRecord(t0, s->s_compute);
        const int first = ex.shard_rank * dt, stride = ex.shard_count * dt;
        for (int done = 0; done < nlocal && rc == 0; done += s->max_frames) {
            const int n = std::min(s->max_frames, nlocal - done);
            rc = pa_session_upload(s, traj, first + done * stride, n, stride);
            if (rc == 0) rc = pa_session_run(s);
        }
        lap("enqueued");
        if (rc == 0) rc = pa_session_download(s, nullptr, nullptr, nullptr, 0);      // waits, and reads the sticky flags
        cudaEventRecord(t1, s->s_compute);
        cudaEventSynchronize(t1);
        lap("finished");
        if (total_ms) { *total_ms = 0.f; cudaEventElapsedTime(total_ms, t0, t1); }
        cudaEventDestroy(t0); cudaEventDestroy(t1);
        if (rc == 0) { *out = s; return 0; }
        pa_session_destroy(s);
        if (rc == PA_ERR_UNSUPPORTED && getenv("PA_DEBUG")) fprintf(stderr, "[prealpha_b200] attempt %d: %s -> retrying with a more general kernel\n", attempt, pa_last_error());
        if (rc == PA_ERR_UNSUPPORTED && !(ex.flags & PA_EXEC_NO_CELLS)) { ex.flags |= PA_EXEC_NO_CELLS; continue; }
        if (rc == PA_ERR_UNSUPPORTED && !(ex.flags & PA_EXEC_FORCE_GENERAL)) { ex.flags |= PA_EXEC_FORCE_GENERAL; continue; }
        return rc;
    }
    return PA_ERR_UNSUPPORTED;
}

int pa_distribution_histogram_multi(const pa_traj *traj, const pa_job *jobs, int32_t njobs, const pa_exec *exec,
                                    int64_t *const *hists, int64_t *within_bin, int64_t *out_of_bounds, pa_stats *stats)
{
    if (!traj || !jobs || njobs < 1) { pa::set_error("bad argument"); return PA_ERR_BAD_ARG; }
    pa_exec ex{};
    if (exec) ex = *exec;
    if (ex.shard_count < 1) ex.shard_count = 1;
    if (ex.tile_shard_count < 1) ex.tile_shard_count = 1;
    if (ex.shard_rank < 0 || ex.shard_rank >= ex.shard_count || ex.tile_shard_rank < 0 || ex.tile_shard_rank >= ex.tile_shard_count) {
        pa::set_error("bad shard"); return PA_ERR_BAD_ARG;
    }
    const int dt = jobs[0].sampling_interval;
    for (int j = 1; j < njobs; ++j)
        if (jobs[j].sampling_interval != dt) { pa::set_error("jobs of one call must share sampling_interval"); return PA_ERR_BAD_ARG; }
    if (dt < 1) { pa::set_error("sampling_interval < 1"); return PA_ERR_BAD_ARG; }
    if (ex.ndevices > 1 || ex.ndevices < 0)
        return pa_multi_device_histogram(traj, jobs, njobs, ex, hists, within_bin, out_of_bounds, stats);
    pa_session *s = nullptr;
    float total_ms = 0.f;
    int rc = run_share_on_device(traj, jobs, njobs, ex, &s, &total_ms);
    if (rc) return rc;
    rc = pa_session_download(s, hists, within_bin, out_of_bounds, 0);
    if (stats) { pa_session_stats(s, stats); stats->total_ms = total_ms; }
    pa_session_destroy(s);
    return rc;
}

}  // extern "C"

// ---------------------------------------------------------------------------
// several devices inside one call (pa_exec.ndevices): the parallel region of iterate_timesteps_parallelised
// (Distribution:704-753) with one host thread per GPU instead of one per core, and the CRITICAL-section merge of the
// private histograms (Distribution:745-751) as ONE sum of the int64 accumulators on the first device over NVLink (peer
// copies + an add kernel, or ncclReduce for large histograms: nccl_min_bytes).
namespace {

__global__ void pa_acc_add_kernel(unsigned long long *dst, const unsigned long long *src, size_t n)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] += src[i];
}

// NCCL is loaded at run time (libnccl.so.2 is in the image; a process that already uses NCCL, e.g. through
// torch.distributed, shares its copy), so that single-GPU users need no NCCL at all.
struct NcclApi {
    void *lib = nullptr;
    int (*CommInitAll)(void **, int, const int *) = nullptr;
    int (*CommDestroy)(void *) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*Reduce)(const void *, void *, size_t, int, int, int, void *, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    bool ok = false;
};
std::mutex g_nccl_mutex;
NcclApi g_nccl;
bool g_nccl_tried = false;
struct CommSet { std::vector<int> devs; std::vector<void *> comms; };
std::vector<CommSet> g_comm_cache;

const NcclApi &nccl_api()
{
    std::lock_guard<std::mutex> lock(g_nccl_mutex);
    if (g_nccl_tried) return g_nccl;
    g_nccl_tried = true;
    if (getenv("PA_NO_NCCL")) return g_nccl;                 // test hook: exercise the peer-copy path
    for (const char *name : { "libnccl.so.2", "libnccl.so" }) {
        g_nccl.lib = dlopen(name, RTLD_NOW | RTLD_LOCAL);
        if (g_nccl.lib) break;
    }
    if (!g_nccl.lib) return g_nccl;
    g_nccl.CommInitAll = (int (*)(void **, int, const int *))dlsym(g_nccl.lib, "ncclCommInitAll");
    g_nccl.CommDestroy = (int (*)(void *))dlsym(g_nccl.lib, "ncclCommDestroy");
    g_nccl.GroupStart = (int (*)())dlsym(g_nccl.lib, "ncclGroupStart");
    g_nccl.GroupEnd = (int (*)())dlsym(g_nccl.lib, "ncclGroupEnd");
    g_nccl.Reduce = (int (*)(const void *, void *, size_t, int, int, int, void *, cudaStream_t))dlsym(g_nccl.lib, "ncclReduce");
    g_nccl.GetErrorString = (const char *(*)(int))dlsym(g_nccl.lib, "ncclGetErrorString");
    g_nccl.ok = g_nccl.CommInitAll && g_nccl.CommDestroy && g_nccl.GroupStart && g_nccl.GroupEnd && g_nccl.Reduce;
    return g_nccl;
}

// The accumulators of a run are a few KB to a few MB.  Below this size the first device fetches them from its peers
// (copies over NVLink + an add kernel, microseconds); from this size on the sum is an ncclReduce.  A communicator set for
// 8 GPUs costs ~1 s to create, connects its transports on the first collective and has to be torn down at exit, which a
// 16 KB sum never earns back.  PA_NCCL_MIN_BYTES overrides the threshold (0: always NCCL -- the tests use it).
size_t nccl_min_bytes()
{
    if (const char *e = getenv("PA_NCCL_MIN_BYTES")) return (size_t)strtoull(e, nullptr, 10);
    return (size_t)4 << 20;
}

// communicators are expensive to create (hundreds of ms): one set per device list, kept until pa_trim()
std::vector<void *> nccl_comms_for(const std::vector<int> &devs)
{
    const NcclApi &api = nccl_api();
    if (!api.ok) return {};
    std::lock_guard<std::mutex> lock(g_nccl_mutex);
    for (const CommSet &c : g_comm_cache) if (c.devs == devs) return c.comms;
    CommSet c;
    c.devs = devs; c.comms.assign(devs.size(), nullptr);
    if (api.CommInitAll(c.comms.data(), (int)devs.size(), devs.data()) != 0) return {};
    g_comm_cache.push_back(c);
    return c.comms;
}

}  // namespace

// accumulator bytes of a session for these jobs (upper bound: bins + the within / out-of-bounds words of every job)
static size_t acc_bytes_estimate(const pa_job *jobs, int njobs)
{
    size_t words = 0;
    for (int j = 0; j < njobs; ++j) words += (size_t)std::max(jobs[j].bin_a, 1) * (size_t)std::max(jobs[j].bin_b, 1) + 8;
    return words * sizeof(unsigned long long);
}

// creates (and caches) the NCCL communicators of devices first..first+n-1 ahead of the reduce -- when the reduce of these
// jobs is going to use them (see nccl_min_bytes)
void pa_multi_device_prepare(int first, int n, const pa_job *jobs, int njobs)
{
    if (acc_bytes_estimate(jobs, njobs) < nccl_min_bytes()) return;
    std::vector<int> devs(n);
    for (int d = 0; d < n; ++d) devs[d] = first + d;
    nccl_comms_for(devs);
}

void pa_multi_device_trim()
{
    std::lock_guard<std::mutex> lock(g_nccl_mutex);
    if (g_nccl.ok)
        for (CommSet &c : g_comm_cache)
            for (void *comm : c.comms) if (comm) g_nccl.CommDestroy(comm);
    g_comm_cache.clear();
}

// Adds the accumulators of sessions 1..n-1 (one per device) to those of session 0: peer copies over NVLink and an add
// kernel on the first device for the usual few KB, one ncclReduce(sum) of the int64 words for large histograms (see
// nccl_min_bytes; also peer copies when libnccl cannot be loaded).
int pa_sessions_reduce_to_first(pa_session *const *sess, int n, const std::vector<void *> *comms_in)
{
    if (n < 2) return 0;
    int rc = 0;
    std::vector<int> devs(n);
    for (int d = 0; d < n; ++d) devs[d] = sess[d]->device;
    const size_t words = sess[0]->acc_words;
    for (int d = 1; d < n; ++d) if (sess[d]->acc_words != words) { pa::set_error("sessions with different accumulators"); return PA_ERR_BAD_ARG; }
    std::vector<void *> comms;
    if (words * sizeof(unsigned long long) >= nccl_min_bytes()) comms = (comms_in && !comms_in->empty()) ? *comms_in : nccl_comms_for(devs);
    bool reduced = false;
    if (!comms.empty()) {
        const NcclApi &api = nccl_api();
        int nrc = api.GroupStart();
        for (int d = 0; d < n && nrc == 0; ++d) {
            cudaSetDevice(devs[d]);
            nrc = api.Reduce(sess[d]->d_acc, sess[d]->d_acc, words, /*ncclUint64*/ 5, /*ncclSum*/ 0, 0, comms[d], sess[d]->s_compute);
        }
        const int erc = api.GroupEnd();
        if (nrc == 0) nrc = erc;
        for (int d = 0; d < n; ++d) { cudaSetDevice(devs[d]); cudaStreamSynchronize(sess[d]->s_compute); }
        if (nrc == 0) reduced = true;
        else if (getenv("PA_DEBUG")) fprintf(stderr, "[prealpha_b200] ncclReduce failed (%s): falling back to peer copies\n", api.GetErrorString ? api.GetErrorString(nrc) : "?");
    }
    if (!reduced) {
        // device d's accumulators are copied to a scratch buffer on the first device and added there; a few KB per job,
        // so the serial order does not matter
        for (int d = 1; d < n; ++d) { cudaSetDevice(devs[d]); cudaStreamSynchronize(sess[d]->s_compute); }
        cudaSetDevice(devs[0]);
        unsigned long long *scratch = nullptr;
        if (cudaMalloc(&scratch, sizeof(unsigned long long) * words) != cudaSuccess) { pa::set_error("cudaMalloc (reduce scratch)"); rc = PA_ERR_CUDA; }
        for (int d = 1; d < n && rc == 0; ++d) {
            if (cudaMemcpyPeerAsync(scratch, devs[0], sess[d]->d_acc, devs[d], sizeof(unsigned long long) * words, sess[0]->s_compute) != cudaSuccess) {
                pa::set_error("cudaMemcpyPeerAsync"); rc = PA_ERR_CUDA; break;
            }
            pa_acc_add_kernel<<<(unsigned)std::min<size_t>((words + 255) / 256, 1184), 256, 0, sess[0]->s_compute>>>(sess[0]->d_acc, scratch, words);
        }
        if (rc == 0 && cudaStreamSynchronize(sess[0]->s_compute) != cudaSuccess) { pa::set_error("peer reduce"); rc = PA_ERR_CUDA; }
        cudaFree(scratch);
    }
    cudaSetDevice(devs[0]);
    return rc;
}

void pa_sessions_merge_stats(pa_session *const *sess, int n, bool by_frames, pa_stats *out)
{
    pa_stats sum{};
    for (int d = 0; d < n; ++d) {
        pa_stats st{};
        pa_session_stats(sess[d], &st);
        sum.pairs_evaluated += st.pairs_evaluated;
        sum.frames_done = by_frames ? sum.frames_done + st.frames_done : st.frames_done;
        sum.kernel_launches += st.kernel_launches; sum.exception_pairs += st.exception_pairs;
        sum.h2d_bytes += st.h2d_bytes; sum.d2h_bytes += st.d2h_bytes;
        sum.kernel_ms = std::max(sum.kernel_ms, st.kernel_ms);
        sum.cell_pair_evals += st.cell_pair_evals; sum.smem_atomics += st.smem_atomics; sum.unique_pairs_counted += st.unique_pairs_counted;
        sum.dominant_launches += st.dominant_launches; sum.dominant_ms = std::max(sum.dominant_ms, st.dominant_ms);
        sum.dominant_pair_evals += st.dominant_pair_evals; sum.dominant_atomics += st.dominant_atomics;
    }
    *out = sum;
}

int pa_multi_device_histogram(const pa_traj *traj, const pa_job *jobs, int32_t njobs, pa_exec ex,
                              int64_t *const *hists, int64_t *within_bin, int64_t *out_of_bounds, pa_stats *stats)
{
    const int usable = pa_device_count();
    int n = ex.ndevices < 0 ? usable - ex.device : ex.ndevices;
    if (usable < 1) { pa::set_error("no CUDA device (this path has no CPU fallback)"); return PA_ERR_NO_DEVICE; }
    if (n < 1 || ex.device < 0 || ex.device + n > usable) { pa::set_error("pa_exec.device/ndevices: not that many usable devices"); return PA_ERR_BAD_ARG; }
    ex.ndevices = 1;
    if (n == 1) return pa_distribution_histogram_multi(traj, jobs, njobs, &ex, hists, within_bin, out_of_bounds, stats);
    const int dt = jobs[0].sampling_interval;
    const int nsamp = (traj->nsteps + dt - 1) / dt;
    const int nlocal = nsamp > ex.shard_rank ? (nsamp - ex.shard_rank + ex.shard_count - 1) / ex.shard_count : 0;
    // enough frames: every device takes whole frames (no replicated per-frame work); otherwise the frames are few and
    // huge, and the devices split the work items of each frame
    const bool by_frames = nlocal >= n;
    std::vector<int> devs(n);
    for (int d = 0; d < n; ++d) devs[d] = ex.device + d;
    // communicators are created while the kernels run
    std::vector<void *> comms;
    std::thread comm_thread([&] { if (acc_bytes_estimate(jobs, njobs) >= nccl_min_bytes()) comms = nccl_comms_for(devs); });
    std::vector<pa_session *> sess(n, nullptr);
    std::vector<int> rcs(n, 0);
    std::vector<float> total_ms(n, 0.f);
    std::vector<std::string> errs(n);
    std::vector<std::thread> workers;
    for (int d = 0; d < n; ++d) {
        workers.emplace_back([&, d] {
            pa_exec e = ex;
            e.device = devs[d];
            if (by_frames) { e.shard_rank = ex.shard_rank + ex.shard_count * d; e.shard_count = ex.shard_count * n; }
            else { e.tile_shard_rank = ex.tile_shard_rank + ex.tile_shard_count * d; e.tile_shard_count = ex.tile_shard_count * n; }
            rcs[d] = run_share_on_device(traj, jobs, njobs, e, &sess[d], &total_ms[d]);
            if (rcs[d]) errs[d] = pa_last_error();
        });
    }
    for (auto &w : workers) w.join();
    comm_thread.join();
    int rc = 0;
    for (int d = 0; d < n && rc == 0; ++d) if (rcs[d]) { rc = rcs[d]; pa::set_error("device " + std::to_string(devs[d]) + ": " + errs[d]); }
    if (rc == 0) rc = pa_sessions_reduce_to_first(sess.data(), n, &comms);
    if (rc == 0) rc = pa_session_download(sess[0], hists, within_bin, out_of_bounds, 0);
    if (rc == 0 && stats) {
        pa_sessions_merge_stats(sess.data(), n, by_frames, stats);
        for (int d = 0; d < n; ++d) stats->total_ms = std::max(stats->total_ms, (double)total_ms[d]);
        stats->pairs_binned = 0;
        for (int j = 0; j < njobs; ++j) if (within_bin) stats->pairs_binned += within_bin[j];
    }
    for (int d = 0; d < n; ++d) if (sess[d]) pa_session_destroy(sess[d]);
    return rc;
}

extern "C" {

int pa_distribution_histogram(const pa_traj *traj, const pa_job *job, int64_t *hist, int64_t *within_bin, int64_t *out_of_bounds)
{
    int64_t *hp[1] = { hist };
    return pa_distribution_histogram_multi(traj, job, 1, nullptr, hp, within_bin, out_of_bounds, nullptr);
}

}  // extern "C"
