// Internal declarations shared by the CUDA kernels (pa_kernels.cu), the session /
// C-ABI layer (pa_session.cu) and the host-side float32-exact logic (pa_host.cpp).
#ifndef PA_INTERNAL_H
#define PA_INTERNAL_H

#include <cstdint>
#include <cstddef>
#include <string>
#include <vector>
#include <vector_types.h>
#include "../../include/prealpha_b200.h"

// ---------------------------------------------------------------------------
// Device-side views

// Per molecule type, resident for a batch of frames.
struct PaDevType {
    const float  *xyz;          // [nframes][nmol][natoms][3] raw float32 (batch-local)
    const double *atom_mass;    // [natoms]
    const double *atom_charge;  // [natoms]
    double mass;                // %mass
    double realcharge;          // DBLE(%realcharge)
    int32_t natoms, nmol, charge;
    int32_t offset;             // first slot of this type inside a frame's slot range
};

// SoA of per-molecule points for a batch: slot = frame*M + type.offset + mol.
// p = centre of mass after wrap_vector (Molecular:1560); ws = its wrap shift;
// c = unwrapped centre of mass; q = centre of charge (only when some job uses it).
struct PaPoints {
    double *px, *py, *pz;
    double *wx, *wy, *wz;
    double *cx, *cy, *cz;
    double *qx, *qy, *qz;       // may be null
    int64_t M;                  // slots per frame (sum of nmol over all types)
};

struct PaBox {
    double lo[3], hi[3], L[3];  // L = hi - lo in fp64 (box_size, Molecular:937)
    double max_dist_sq;
    int32_t wrap_trajectory;    // 1: positions are NOT wrapped by the distance routine
    int32_t com_boost;
};

// Per job constants the kernels need.
struct PaDevJob {
    int32_t mode;               // PA_MODE_*
    int32_t bin_a, bin_b;
    int32_t use_coc;
    double  ref_vec[3];
    double  rot[9];
    float   maxdist, step_a, step_b, shift;
    float   maxdist_sq_f;       // REAL4 maxdist**2
    // class tables (device pointers, uint64 bit patterns of non-negative doubles):
    //   thr_a[k], k=0..bin_a+2: smallest d2 whose a-class is >= k (thr_a[0]=0).
    //   class bin_a = "inside cutoff but a-bin out of bounds", class bin_a+1 = beyond cutoff,
    //   thr_a[bin_a+2] = never.
    const unsigned long long *thr_a;
    //   thr_b[k], k=1..bin_b+1 (doubles): largest dot product whose b-class INT(acos/step_b) is >= k.
    const double *thr_b;
    unsigned long long cut_bits;   // = thr_a[bin_a+1]
    float   guess_scale;        // 1/step_a
    int32_t z_sign;             // fast path only: +1 / -1 for reference vector along +z / -z
    unsigned long long *acc;    // device accumulators: [bin_a*bin_b hist][within][oob]
};

// One launch = one job x one observed type over all frames of the batch.
struct PaPass {
    PaPoints pts;
    PaBox box;
    PaDevJob job;
    int32_t nframes;
    int32_t frame0;             // first resident frame this launch works on
    int32_t ref_off, nref;      // slot range of the reference type inside a frame
    int32_t obs_off, nobs;
    int32_t same_type;
    int32_t n_itiles, n_jsplits, jsplit_len;
    long long weight;           // +1 or charge of the observed type (weigh_charge)
    // exception list (pairs that need the literal per-pair evaluation)
    uint2 *exc_list;            // (reference slot, observed slot) of the pairs to evaluate literally
    unsigned int *exc_count;    // entries appended to exc_list by this launch
    unsigned int exc_cap;
    unsigned int *flags;        // [1] sticky error/overflow flag, [2] verify failures, [3] total exception pairs
    // symmetric / merged passes of the cell kernel (see pa_cells.cu):
    //   sym:  reference type == observed type and all arithmetic is exact -> every unordered pair is
    //         evaluated once and counted twice;
    //   acc2: a second job whose (observed, reference) types are this pass's (reference, observed) types
    //         with the same bins: d2(i,j) == d2(j,i) bitwise, so the counts are added to both.
    int32_t sym;
    unsigned long long *acc2;
    long long weight2;
    uint2 *exc_list2;           // reversed pairs (j,i) for the second job (or the same job when sym)
    unsigned int *exc_count2;
    unsigned int *verify_fail;  // debug counter for the class-guess invariant (may be null)
    int32_t f32_cnt;            // > 0: float32 segments enabled in the cell kernel, counters needed for any reachable class
    float eps64;                // half-width of the class brackets of the fp64-difference segments (variant 0)
    float f32_lo, f32_hi;       // (1 -+ eps)/step_a rounded outwards (eps: pa_session.cu); adjacent so that they load as one packed pair
    int32_t tshard_rank, tshard_count;   // tile shard of this device (pa_exec.tile_shard_*): work item w is taken when w % count == rank
    int32_t pad0_;
    float dead_r;               // (bin_a + 1)/f32_lo rounded up: from this approximate distance on, the lower class bracket is beyond the cutoff
    unsigned long long *kstats; // device counters of the cell kernels (8 words, may be null): [0] pair evaluations after cell pruning
                                // and [1] shared-memory atomics issued by the float32 launches, [4] / [5] the same for the other
                                // launches, [2] pairs counted into bins (each evaluated once)
};

// Uniform grid of one molecule type for a batch of frames (pa_cells.cu).  Arrays are
// batch-local: frame index f counts from the first frame the sort was launched on.
struct PaCellGrid {
    int32_t g[3], gmax;              // cells per component (x finer than y,z), max of the three
    int32_t ncell, nmol, type_off, pad_;
    double inv_w[3];                 // g[k] / L[k]
    int *cell_start;                 // [nframes][ncell+1] offsets into the sorted arrays
    int *cell_fill;                  // [nframes][ncell]   scatter cursors
    int *cid;                        // [nframes][nmol]    cell of every molecule
    double *sx, *sy, *sz;            // [nframes][nmol]    wrapped positions in cell order
    int *sorig;                      // [nframes][nmol]    original molecule index
    unsigned char *sflag;            // [nframes][nmol]    1: the centre of mass was moved by wrap_vector (w != 0)
    unsigned long long *lay_lo, *lay_hi;   // [nframes][3][gmax] exact bounds per cell layer (ordered keys)
};
struct PaTile { int32_t frame, begin, count, cx0, cx1, cy, cz, pad; };

// kernel launch wrappers (pa_kernels.cu).  All return cudaError_t as int.
struct PaPrepareArgs {
    const PaDevType *types;     // device array
    int32_t ntypes;
    int32_t nframes;
    int32_t frame0;
    PaPoints pts;
    PaBox box;
    int32_t need_coc;
};

int pa_launch_prepare(const PaPrepareArgs &a, void *stream);
int pa_launch_rdf_fast(const PaPass &p, int sm_count, void *stream);
int pa_launch_general(const PaPass &p, int sm_count, void *stream);
int pa_launch_exceptions(const PaPass &p, unsigned int count, void *stream);
int pa_launch_cell_sort(const PaCellGrid &g, const PaPoints &pts, const PaBox &box, int frame0, int nframes, unsigned int *err, void *stream);
// chunks = false: cell-aligned tiles of <= 256 molecules (general kernel); true: 64-molecule pieces of the rows (RDF kernel,
// one per warp); row_off: scratch of nframes * g[1] * g[2] ints
int pa_launch_build_tiles(const PaCellGrid &g, int nframes, bool chunks, PaTile *tiles, unsigned int *ntiles, int *row_off, unsigned int cap, void *stream);
int pa_launch_rdf_cells(const PaPass &p, const PaCellGrid &gi, const PaCellGrid &go, const PaTile *tiles, const unsigned int *ntiles,
                        unsigned int tile_cap, unsigned int *work_counter, int nzg, int variant, int sm_count, void *stream,
                        void *ev_f32_begin, void *ev_f32_end);   // optional CUDA events around the float32 launch (the dominant kernel)
// 2-D pdf jobs (distance x polar angle) on warp tiles with both classes bracketed in float32 (see pa_cells.cu)
int pa_launch_pdf_cells(const PaPass &p, const PaCellGrid &gi, const PaCellGrid &go, const PaTile *tiles, const unsigned int *ntiles,
                        unsigned int tile_cap, unsigned int *work_counter, int nzg, int sm_count, void *stream);
int pa_launch_general_cells(const PaPass &p, const PaCellGrid &gi, const PaCellGrid &go, const PaTile *tiles, const unsigned int *ntiles,
                            unsigned int tile_cap, unsigned int *work_counter, int nzg, int sm_count, void *stream);
void pa_cell_grid_dims(int nmol, int g[3]);
// charge_arm mode (O(N) per frame): one thread per reference molecule and frame
struct PaChargeArmArgs {
    const PaDevType *types;     // device array (raw coordinates of the resident batch)
    int32_t type;               // reference type, 0-based
    int32_t frame0, nframes;
    int32_t normalise_clm, com_boost;
    PaDevJob job;               // thr_a = classes of the squared arm length, thr_b as for pdf
    unsigned long long tiny_bits;   // smallest squared length whose REAL(4) length is >= 1E-8
};
int pa_launch_charge_arm(const PaChargeArmArgs &a, void *stream);
int pa_kernels_selfcheck();    // returns 0 when a kernel image for the current device exists
int pa_rdf_ri(int nref);       // reference molecules per thread the fast kernel will use (tile = 128*RI)
int pa_general_ri(int nref);
#define PA_TILE_THREADS 128
#define PA_J_CHUNK 128

// several sessions, one per device, working on shards of one job (pa_session.cu): combine the int64 accumulators on the
// first session's device (ncclReduce over NVLink, or peer copies + add) / merge their statistics
int pa_sessions_reduce_to_first(pa_session *const *sess, int n, const std::vector<void *> *comms_in = nullptr);
void pa_sessions_merge_stats(pa_session *const *sess, int n, bool by_frames, pa_stats *out);
void pa_multi_device_prepare(int first_device, int ndevices, const pa_job *jobs, int njobs);   // NCCL communicators ahead of time (cached), if the reduce will use them

// ---------------------------------------------------------------------------
// Host-side float32-exact logic (pa_host.cpp)

namespace pa {

// class tables, see PaDevJob
void build_thr_a(const pa_job &job, double max_dist_sq, std::vector<unsigned long long> &thr);
void build_thr_b(const pa_job &job, std::vector<double> &thr);
void build_thr_charge_arm(const pa_job &job, std::vector<unsigned long long> &thr, unsigned long long &tiny_bits);
int  a_class_of(const pa_job &job, double max_dist_sq, double d2);   // reference arithmetic
int  b_class_of(const pa_job &job, double dot);
bool job_is_rdf_fast(const pa_job &job, const pa_traj &traj);
void set_error(const std::string &msg);

}  // namespace pa

#endif
