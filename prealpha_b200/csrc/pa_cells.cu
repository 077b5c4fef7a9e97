// pa_cells.cu -- cell-sorted RDF kernel (the fast path for large systems).
//
// The brute-force tile kernel (pa_rdf_fast_kernel) must take the minimum over three periodic
// images per component for every pair because it knows nothing about where the two molecules
// are.  Here every frame is first sorted into a uniform grid (counting sort, hand-written:
// count -> scan -> scatter; the grid is finer along x so that runs of x-cells give tiles of
// almost exactly 256 molecules).  A tile of reference molecules (<= 256, held in registers) is
// then paired with runs of observed cells, and for each (tile, cell-run) the image shift per
// component is decided ONCE from the exact bounding boxes:
//   * "single": one shift s in {-L,0,+L} is strictly closest for every pair of the two boxes
//     (by a margin far above rounding) -> the pair costs 1 DSUB + 1 DMUL for that component;
//   * "double": two shifts are possible (boxes straddle L/2) -> 2 DSUB, min of |d|, 1 DMUL;
//   * cell runs whose closest possible distance is beyond the cutoff are skipped.
// Most runs are then classified in packed float32 from tile-local coordinates (pc_segment_f32: rigorous
// class brackets 5.5e-7 + delta wide, exact fp64 decision for the pairs the brackets leave open), which removes the
// FP64 pipe from the inner loop without changing a count.
// Nothing about the fp64 arithmetic changes: a pair evaluates exactly the candidates
// fl(fl(p2+s)-p1) that can be the minimum of the reference's 27-image loop (Molecular:1451-1467),
// so the fp64 result is bit-identical; skipped pairs are pairs the reference evaluates and then
// discards at `IF (current_distance_squared<maxdist_squared)` (Distribution:783).
#include <cuda_runtime.h>
#include <cstdint>
#include <algorithm>
#include "pa_internal.h"
#include "pa_pair_exact.cuh"

// RDF kernel: every WARP is autonomous -- it owns a tile of 32*PC_RI reference molecules (a compact slab of one row of
// cells), takes its own work items, classifies the observed cell layers against ITS bounding box, stages observed
// molecules into its own slice of shared memory and never meets a CTA-wide barrier inside the work loop; only the
// class counters (and the threshold table) are shared by the warps of a CTA.
#define PC_WARPS 4
#ifndef PC_PREFETCH
#define PC_PREFETCH 0               // float32 launch: load the candidates of the next staging step one step ahead (measured:
                                    // 212 ms instead of 203 ms per 1M-atom frame -- the six extra registers cost more than the latency)
#endif
#ifndef PDF_MINB
#define PDF_MINB 4                  // resident CTAs per SM the float32 2-D launch is compiled for
#endif
#ifndef PC_MINB
#define PC_MINB 6                   // resident CTAs per SM the float32 launch is compiled for (80 registers per thread)
#endif
#define PC_THREADS (32 * PC_WARPS)
#define PC_RI 2
#define PC_TILE (32 * PC_RI)
#define PC_JC 32                    // observed molecules staged per chunk: one per lane
#define PC_MAXG 256
// general (2-D) kernel: CTA-wide tiles
#define PG_THREADS 128
#define PG_RI 2
#define PC_GTILE (PG_THREADS * PG_RI)
#define PG_JC 128

// Reference molecule r (0..PC_RI-1) of this lane inside its warp's tile (consecutive in the row's sorted order).
static __device__ __forceinline__ int pc_kidx(int r) { return r * 32 + (int)(threadIdx.x & 31); }

static __device__ __forceinline__ unsigned long long dbits(double x) { return (unsigned long long)__double_as_longlong(x); }
// order-preserving map double -> uint64 (for atomicMin/atomicMax on coordinates)
static __device__ __forceinline__ unsigned long long okey(double x)
{
    const unsigned long long b = dbits(x);
    return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}
static __device__ __forceinline__ double okey_inv(unsigned long long k)
{
    return __longlong_as_double((long long)((k & 0x8000000000000000ull) ? (k & 0x7fffffffffffffffull) : ~k));
}

// ---------------------------------------------------------------------------
// K1: cell index per molecule, cell population, per-layer exact bounding boxes
__global__ void __launch_bounds__(256) pa_cell_count_kernel(PaCellGrid g, PaPoints pts, PaBox box, int frame0, int nframes, unsigned int *err)
{
    const long long total = (long long)nframes * g.nmol;
    for (long long id = blockIdx.x * (long long)blockDim.x + threadIdx.x; id < total; id += (long long)gridDim.x * blockDim.x) {
        const int f = (int)(id / g.nmol), m = (int)(id % g.nmol);
        const long long slot = (long long)(frame0 + f) * pts.M + g.type_off + m;
        const double p[3] = { pts.px[slot], pts.py[slot], pts.pz[slot] };
        int c[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            if (!(p[k] >= box.lo[k] && p[k] <= box.hi[k])) { atomicOr(err, 1u); c[k] = 0; continue; }   // NaN / not wrapped
            int ck = (int)((p[k] - box.lo[k]) * g.inv_w[k]);
            c[k] = min(g.g[k] - 1, max(0, ck));
            const long long lay = ((long long)f * 3 + k) * g.gmax + c[k];
            atomicMin(&g.lay_lo[lay], okey(p[k]));
            atomicMax(&g.lay_hi[lay], okey(p[k]));
        }
        const int cell = (c[2] * g.g[1] + c[1]) * g.g[0] + c[0];
        g.cid[(long long)f * g.nmol + m] = cell;
        atomicAdd(&g.cell_start[(long long)f * (g.ncell + 1) + cell + 1], 1);
    }
}

// K2: in-place inclusive scan of the populations (cell_start[f][c+1] held counts) -> offsets
__global__ void __launch_bounds__(1024) pa_cell_scan_kernel(PaCellGrid g)
{
    __shared__ int s_part[1024];
    int *cs = g.cell_start + (long long)blockIdx.x * (g.ncell + 1);
    const int n = g.ncell, per = (n + 1023) / 1024;
    const int b = min(n, (int)threadIdx.x * per), e = min(n, b + per);
    int sum = 0;
    for (int i = b; i < e; ++i) sum += cs[i + 1];
    s_part[threadIdx.x] = sum;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {
        int v = threadIdx.x >= off ? s_part[threadIdx.x - off] : 0;
        __syncthreads();
        s_part[threadIdx.x] += v;
        __syncthreads();
    }
    int run = threadIdx.x ? s_part[threadIdx.x - 1] : 0;
    for (int i = b; i < e; ++i) { run += cs[i + 1]; cs[i + 1] = run; }
    if (threadIdx.x == 0) cs[0] = 0;
}

// K3a: scatter the molecule indices into cell order.  The position inside a cell comes from an atomic cursor, i.e. it
// differs from run to run (and from GPU to GPU) ...
__global__ void __launch_bounds__(256) pa_cell_scatter_kernel(PaCellGrid g, int nframes)
{
    const long long total = (long long)nframes * g.nmol;
    for (long long id = blockIdx.x * (long long)blockDim.x + threadIdx.x; id < total; id += (long long)gridDim.x * blockDim.x) {
        const int f = (int)(id / g.nmol), m = (int)(id % g.nmol);
        const int cell = g.cid[id];
        const int pos = g.cell_start[(long long)f * (g.ncell + 1) + cell] + atomicAdd(&g.cell_fill[(long long)f * g.ncell + cell], 1);
        g.sorig[(long long)f * g.nmol + pos] = m;
    }
}

// K3b: ... so every cell is put into ascending molecule index (one warp per cell, rank sort: the rank of an entry is the
// number of smaller entries).  The sorted order of a frame is then a function of the frame alone, which is what lets
// tile shards on different GPUs -- whose warp tiles are pieces of this order -- agree on which molecules a tile holds.
// Output goes to g.cid (free after K3a).  A cell with more than PC_MAX_CELL molecules (the grid is useless for such a
// frame) raises the sticky flag: the host redoes the call without cells.
#define PC_MAX_CELL 4096
__global__ void __launch_bounds__(256) pa_cell_order_kernel(PaCellGrid g, int nframes, unsigned int *err)
{
    const int lane = threadIdx.x & 31;
    const long long ncells = (long long)nframes * g.ncell;
    for (long long c = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5; c < ncells; c += ((long long)gridDim.x * blockDim.x) >> 5) {
        const int f = (int)(c / g.ncell), cell = (int)(c % g.ncell);
        const int *cs = g.cell_start + (long long)f * (g.ncell + 1);
        const int beg = cs[cell], n = cs[cell + 1] - beg;
        const int *src = g.sorig + (long long)f * g.nmol + beg;
        int *dst = g.cid + (long long)f * g.nmol + beg;
        if (n <= 32) {
            const int v = lane < n ? src[lane] : 0x7fffffff;
            int rank = 0;
            for (int k = 0; k < n; ++k) rank += (__shfl_sync(0xffffffffu, v, k) < v) ? 1 : 0;
            if (lane < n) dst[rank] = v;
        } else if (n <= PC_MAX_CELL) {
            for (int i = lane; i < n; i += 32) {
                const int v = src[i];
                int rank = 0;
                for (int k = 0; k < n; ++k) rank += (src[k] < v) ? 1 : 0;
                dst[rank] = v;
            }
        } else {
            if (lane == 0) atomicOr(err, 1u);
            for (int i = lane; i < n; i += 32) dst[i] = src[i];
        }
    }
}

// K3c: gather the positions into the sorted order
__global__ void __launch_bounds__(256) pa_cell_gather_kernel(PaCellGrid g, PaPoints pts, PaBox box, int frame0, int nframes)
{
    const long long total = (long long)nframes * g.nmol;
    for (long long id = blockIdx.x * (long long)blockDim.x + threadIdx.x; id < total; id += (long long)gridDim.x * blockDim.x) {
        const int f = (int)(id / g.nmol);
        const int m = g.cid[id];
        const long long slot = (long long)(frame0 + f) * pts.M + g.type_off + m;
        // A position outside the box (not wrapped: wrap_trajectory, or beyond 1000 wrap moves) or NaN has already raised
        // the sticky flag in pa_cell_count_kernel and the whole call will be redone without cells; the pair kernels still
        // run, so such a position is clamped into the box here to keep every distance below the box diagonal (the
        // float32 segments size their counter window by that bound).  fmax / fmin return the non-NaN operand.
        g.sx[id] = fmin(fmax(pts.px[slot], box.lo[0]), box.hi[0]);
        g.sy[id] = fmin(fmax(pts.py[slot], box.lo[1]), box.hi[1]);
        g.sz[id] = fmin(fmax(pts.pz[slot], box.lo[2]), box.hi[2]);
        g.sorig[id] = m;
        g.sflag[id] = (pts.wx[slot] != 0.0 || pts.wy[slot] != 0.0 || pts.wz[slot] != 0.0) ? 1 : 0;
    }
}

// K4: reference tiles.  Two kinds, both numbered deterministically (count per row -> exclusive scan -> fill), because
// tile shards on different GPUs (pa_exec.tile_shard_*) must agree on which tile is which:
//   CHUNKS = false: greedy runs of x-neighbouring cells of one (y,z) row, <= PC_GTILE molecules and at most max_run cells
//                   wide (so that tile width + cell width stays below L/2) -- the general kernel's tiles;
//   CHUNKS = true:  consecutive PC_TILE-molecule pieces of a row's sorted order (x-cell ascending), cut without regard
//                   to cell boundaries -- one warp's worth of reference molecules of the RDF kernel, a compact slab.
template <bool CHUNKS, bool FILL>
__global__ void __launch_bounds__(128) pa_build_tiles_kernel(PaCellGrid g, int nframes, int max_run, PaTile *tiles, int *row_off, unsigned int cap)
{
    const int rows = g.g[1] * g.g[2], gx = g.g[0];
    const int id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= nframes * rows) return;
    const int f = id / rows, row = id % rows;
    const int *cs = g.cell_start + (long long)f * (g.ncell + 1) + (long long)row * gx;
    unsigned int t = FILL ? (unsigned int)row_off[id] : 0u;
    int n = 0;
    if (CHUNKS) {
        // a piece ends after PC_TILE molecules or at the end of the max_run-th cell it touches (sparse types: the tile's
        // x extent plus one observed cell must stay below L/2), whichever comes first
        const int row_end = cs[gx];
        int pos = cs[0], cx = 0;
        while (pos < row_end) {
            while (cs[cx + 1] <= pos) ++cx;                       // cell that holds `pos`
            const int end = min(min(pos + PC_TILE, cs[min(cx + max_run, gx)]), row_end);
            if (FILL && t + n < cap) {
                PaTile tl;
                tl.frame = f; tl.begin = pos; tl.count = end - pos;
                tl.cx0 = cx; tl.cx1 = min(cx + max_run, gx) - 1; tl.cy = row % g.g[1]; tl.cz = row / g.g[1]; tl.pad = 0;
                tiles[t + n] = tl;
            }
            ++n;
            pos = end;
        }
    } else {
        int cx = 0;
        while (cx < gx) {
            int beg = cs[cx], c1 = cx, cnt = cs[cx + 1] - beg;
            while (c1 + 1 < gx && (c1 + 1 - cx) < max_run && (cs[c1 + 2] - beg) <= PC_GTILE) { ++c1; cnt = cs[c1 + 1] - beg; }
            // a single over-full cell is split into several tiles
            for (int off = 0; off < cnt; off += PC_GTILE, ++n) {
                if (FILL && t + n < cap) {
                    PaTile tl;
                    tl.frame = f; tl.begin = beg + off; tl.count = min(PC_GTILE, cnt - off);
                    tl.cx0 = cx; tl.cx1 = c1; tl.cy = row % g.g[1]; tl.cz = row / g.g[1]; tl.pad = 0;
                    tiles[t + n] = tl;
                }
            }
            cx = c1 + 1;
        }
    }
    if (!FILL) row_off[id] = n;
}

// exclusive scan of row_off[0..n) in place (one block), total -> *ntiles
__global__ void __launch_bounds__(1024) pa_tile_scan_kernel(int *row_off, int n, unsigned int *ntiles)
{
    __shared__ int s_part[1024];
    const int per = (n + 1023) / 1024;
    const int b = min(n, (int)threadIdx.x * per), e = min(n, b + per);
    int sum = 0;
    for (int i = b; i < e; ++i) sum += row_off[i];
    s_part[threadIdx.x] = sum;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {
        int v = threadIdx.x >= off ? s_part[threadIdx.x - off] : 0;
        __syncthreads();
        s_part[threadIdx.x] += v;
        __syncthreads();
    }
    int run = threadIdx.x ? s_part[threadIdx.x - 1] : 0;
    for (int i = b; i < e; ++i) { const int c = row_off[i]; row_off[i] = run; run += c; }
    if (threadIdx.x == 1023) *ntiles = (unsigned int)s_part[1023];
}

// ---------------------------------------------------------------------------
// per (tile, observed layer, component) image classification
struct PcLayer { float dmin2; signed char mode, ca, cb, pad; };    // ca/cb: shift code 0,1,2 = -L,0,+L

static __device__ PcLayer pc_classify(double ilo, double ihi, double jlo, double jhi, double L)
{
    PcLayer r;
    const double a = jlo - ihi, b = jhi - ilo;          // p2 - p1 lies in [a, b]
    const double margin = 1e-9 * L + 1e-12;             // >> rounding of (p2+s)-p1, << any physical scale
    double mn[3], mx[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const double s = (double)(k - 1) * L;
        const double lo = a + s, hi = b + s;
        mn[k] = lo > 0.0 ? lo : (hi < 0.0 ? -hi : 0.0);
        mx[k] = fmax(fabs(lo), fabs(hi));
    }
    int best = 0;
    if (mx[1] < mx[best]) best = 1;
    if (mx[2] < (best == 0 ? mx[0] : mx[1])) best = 2;
    const double lim = (best == 0 ? mx[0] : (best == 1 ? mx[1] : mx[2])) + margin;
    int n = 0, c0 = best, c1 = best;
#pragma unroll
    for (int k = 0; k < 3; ++k)
        if (mn[k] <= lim) { if (n == 0) c0 = k; else c1 = k; ++n; }
    r.mode = (signed char)n;                             // 1, 2 or 3 (3 = boxes too wide: caller raises the error flag)
    r.ca = (signed char)c0; r.cb = (signed char)(n >= 2 ? c1 : c0); r.pad = 0;
    const double d = fmax(0.0, fmin(mn[0], fmin(mn[1], mn[2])) - margin);
    r.dmin2 = (float)fmin(d * d * (1.0 - 1e-6), 3.0e38);  // rounded down: still a lower bound
    return r;
}

template <int MODE>
static __device__ __forceinline__ double pc_comp_sq(double ja, double jb, double p1)
{
    if (MODE == 1) {
        const double d = __dsub_rn(ja, p1);
        return __dmul_rn(d, d);
    } else {
        // min over the two candidate images of fl(d^2) = fl((the d of smaller magnitude)^2); the
        // select works on the signed values so that |.| stays an operand modifier of DSETP
        const double da = __dsub_rn(ja, p1);
        const double db = __dsub_rn(jb, p1);
        const double m = fabs(db) < fabs(da) ? db : da;
        return __dmul_rn(m, m);
    }
}

struct PcShared {           // general kernel
    PcLayer lay[3][PC_MAXG];
    int work_item;
};
// Shared-memory class tables: thresholds thr[k] (8-byte stride, k = 0..na+2) and counters cnt[k]
// (4-byte stride, so that the 32 lanes of an ATOMS spread over all 32 banks).
static __device__ __forceinline__ double pc_lds_f64(unsigned int addr)
{
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}
static __device__ __forceinline__ unsigned int pc_lds_u32(unsigned int addr)
{
    unsigned int v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
static __device__ __forceinline__ void pc_sts_u32(unsigned int addr, unsigned int v)
{
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
// shared-memory increment (ATOMS.POPC.INC: equal addresses inside the warp are aggregated by the hardware)
static __device__ __forceinline__ void pc_red(unsigned int addr)
{
    asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(addr) : "memory");
}
// (A variant predicated on "counter is read at flush time" was measured: 402 ms instead of 314 ms per 1M-atom frame --
// partially active ATOMS.POPC.INC are slower than full warps that hit ignored counters.)

struct PcTables {
    unsigned int thr;       // shared address of thr[0]
    unsigned int cnt;       // shared address of cnt[0]
    float s_lo, s_hi;       // (1/step_a)(1 -+ eps): x~ brackets the reference's float32 quotient
    float r_max;            // clamp of the approximate distance: keeps the class index <= na+1
    int na;
    unsigned long long cut;
    unsigned int special;   // shared address of the counter for "inside the cutoff, a-bin out of range" decided by an exact path (VARIANT 0)
    // float32 segments (see pc_segment_f32)
    float pad_x;            // local x of padded observed molecules: beyond the cutoff for every tile molecule
    unsigned int cnt_pad;   // byte offset of the counter window used by padded reference slots
};

// VARIANT 0: the class of a pair is floor(sqrt(d2)/step_a) evaluated approximately (top 52 bits of d2,
//            MUFU.SQRT, two round-down FMAs with the scale shrunk / stretched by eps = 5.5e-7 + delta_top).  When both
//            brackets give the same integer the reference's exact float32 chain cannot give another one
//            (its relative distance to x~ is < 5.1e-7), so no table is touched.  Otherwise (~2e-3 of the
//            pairs) the pair takes the rare branch and is decided by the exact fp64 thresholds.  Pairs
//            beyond the cutoff land in counters >= na, which are ignored.  Valid when the cutoff sits
//            within delta_top <= 3e-7 of the boundary of bin na (pa_session.cu): the brackets are wider than
//            the error budget by delta_top, so every pair that close to the boundary reaches the exact branch, which
//            tells the last bin, "inside the cutoff, a-bin out of range" (counter T.special, counted
//            out of bounds at flush time) and "beyond the cutoff" apart.
// VARIANT 1: every pair is decided by the exact threshold table and an explicit cutoff test.
// DIAG: the observed run overlaps the tile itself (symmetric pass): only pairs whose observed
//       molecule comes later in the sorted order than the reference molecule are counted.
template <int MX, int MY, int MZ, int VARIANT, bool VERIFY, bool DIAG>
static __device__ __forceinline__ void pc_segment(const PaPass &p, const PaCellGrid &go, const PaCellGrid &gi, const PaTile &tl,
                                                  const double (&p1x)[PC_RI], const double (&p1y)[PC_RI], const double (&p1z)[PC_RI],
                                                  long long jbase, int jb, int je,
                                                  double sxa, double sxb, double sya, double syb, double sza, double szb,
                                                  double *s_j, const PcTables &T, unsigned int &it_all, unsigned int &it_live)
{
    const int hi_flag = 0x3E100000;                     // hi word of 2^-30
    const int na = T.na;
    const unsigned int cnt_m = T.cnt - 0x4B000000u * 4u;     // counter base minus the float "magic" offset
    const unsigned int thr_m = T.thr - 0x4B000000u * 8u;
    const unsigned int trash = T.cnt + (unsigned)(na + 2) * 4u;   // counter nobody reads
    constexpr int NP = 2 * PC_RI;
    constexpr int ND = MX + MY + MZ;                    // doubles per staged observed molecule
    for (int jc = jb; jc < je; jc += PC_JC) {
        __syncwarp();
        {
            const int t = threadIdx.x & 31;             // PC_JC == 32: one staged molecule per lane, in the warp's own buffer
            const int j = jc + t;
            double x = 1.0e15, y = 1.0e15, z = 1.0e15;  // padding: far away, finite in float range
            if (j < je) { x = go.sx[jbase + j]; y = go.sy[jbase + j]; z = go.sz[jbase + j]; }
            double *rec = s_j + t * ND;
            int o = 0;
            rec[o++] = __dadd_rn(x, sxa); if (MX == 2) rec[o++] = __dadd_rn(x, sxb);
            rec[o++] = __dadd_rn(y, sya); if (MY == 2) rec[o++] = __dadd_rn(y, syb);
            rec[o++] = __dadd_rn(z, sza); if (MZ == 2) rec[o++] = __dadd_rn(z, szb);
        }
        __syncwarp();
        const int cnt = min(PC_JC, je - jc);
        it_all += (unsigned)((cnt + 1) / 2); it_live += (unsigned)((cnt + 1) / 2);   // statistics: no iteration is skipped here
        for (int jj = 0; jj < cnt; jj += 2) {
            // two staged molecules = 2*ND doubles, 16-byte aligned because jj is even
            double v[2 * ND];
            {
                const double2 *rec = reinterpret_cast<const double2 *>(s_j + jj * ND);
#pragma unroll
                for (int k = 0; k < ND; ++k) { const double2 w = rec[k]; v[2 * k] = w.x; v[2 * k + 1] = w.y; }
            }
            double d2[NP];
            int s1hi[NP];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const double *w = v + u * ND;
                const double xa = w[0], xb = w[MX - 1];
                const double ya = w[MX], yb = w[MX + MY - 1];
                const double za = w[MX + MY], zb = w[MX + MY + MZ - 1];
#pragma unroll
                for (int r = 0; r < PC_RI; ++r) {
                    const double sx = pc_comp_sq<MX>(xa, xb, p1x[r]);
                    const double sy = pc_comp_sq<MY>(ya, yb, p1y[r]);
                    const double sz = pc_comp_sq<MZ>(za, zb, p1z[r]);
                    const double s1 = __dadd_rn(sx, sy);
                    s1hi[u * PC_RI + r] = __double2hiint(s1);
                    d2[u * PC_RI + r] = __dadd_rn(s1, sz);
                }
            }
            unsigned int ca[NP];      // shared address of the counter to increment
            unsigned int tlo[NP];     // float bits of MAGIC + floor(x~ (1-eps))
            int slow = 0;
#pragma unroll
            for (int q = 0; q < NP; ++q) {
                const int hi = __double2hiint(d2[q]);
                const unsigned int lo = (unsigned)__double2loint(d2[q]);
                // float32 with the top 24 significant bits of d2 (truncated): (hi:lo << 3) - bias
                const float f = __uint_as_float(__funnelshift_l(lo, (unsigned)hi, 3) + 0x40000000u);
                float rt;
                asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(rt) : "f"(f));
                rt = fminf(rt, T.r_max);
                const float t0 = __fmaf_rd(rt, T.s_lo, 8388608.0f);
                tlo[q] = (unsigned)__float_as_int(t0);
                if (VARIANT == 0) {
                    const float t1 = __fmaf_rd(rt, T.s_hi, 8388608.0f);
                    ca[q] = tlo[q] * 4u + cnt_m;
                    slow |= (s1hi[q] < hi_flag || t0 != t1) ? (1 << q) : 0;
                } else {
                    const double thr = pc_lds_f64(tlo[q] * 8u + thr_m + 8u);                   // thr[n+1]
                    ca[q] = tlo[q] * 4u + cnt_m + (d2[q] >= thr ? 4u : 0u);
                    if (!(dbits(d2[q]) < T.cut)) ca[q] = trash;
                    slow |= (s1hi[q] < hi_flag) ? (1 << q) : 0;
                }
            }
            if (DIAG) {
#pragma unroll
                for (int q = 0; q < NP; ++q) {
                    const int ipos = tl.begin + pc_kidx(q % PC_RI);
                    const int jpos = jc + jj + q / PC_RI;
                    if (!(ipos < jpos)) { ca[q] = trash; slow &= ~(1 << q); }
                }
            }
            if (slow) {
#pragma unroll
                for (int q = 0; q < NP; ++q) {
                    if (!((slow >> q) & 1)) continue;
                    if (s1hi[q] < hi_flag) {
                        // |link_xy|^2 < 2^-30: (anti)parallel to the reference axis, coincident, or the molecule
                        // itself -> literal evaluation by pa_exception_kernel, not binned here
                        ca[q] = trash;
                        const int k = pc_kidx(q % PC_RI);
                        const int j = jc + jj + q / PC_RI;
                        if (k < tl.count && j < je) {
                            const int io = gi.sorig[(long long)tl.frame * gi.nmol + tl.begin + k];
                            const int jo = go.sorig[jbase + j];
                            if (!(p.same_type && io == jo)) {
                                const unsigned int slot = atomicAdd(p.exc_count, 1u);
                                const long long fb = (long long)(p.frame0 + tl.frame) * p.pts.M;
                                if (slot < p.exc_cap)
                                    p.exc_list[slot] = make_uint2((unsigned int)(fb + p.ref_off + io), (unsigned int)(fb + p.obs_off + jo));
                                if (p.sym || p.acc2) {
                                    // the reversed ordered pair (reference j, observed i) is decided literally as well
                                    const unsigned int slot2 = atomicAdd(p.exc_count2, 1u);
                                    if (slot2 < p.exc_cap)
                                        p.exc_list2[slot2] = make_uint2((unsigned int)(fb + p.obs_off + jo), (unsigned int)(fb + p.ref_off + io));
                                }
                            }
                        }
                    } else if (VARIANT == 0) {
                        // the two brackets disagree: decide with the exact fp64 threshold thr[n+1]; next to the cutoff the
                        // outcome can also be "inside the cutoff, a-bin out of range" (class na) or "beyond the cutoff"
                        const double thr = pc_lds_f64(tlo[q] * 8u + thr_m + 8u);
                        const unsigned int n = tlo[q] + (d2[q] >= thr ? 1u : 0u);
                        ca[q] = n * 4u + cnt_m;
                        if (n == 0x4B000000u + (unsigned)na) ca[q] = dbits(d2[q]) < T.cut ? T.special : trash;
                    }
                }
            }
            if (VERIFY) {
#pragma unroll
                for (int q = 0; q < NP; ++q) {
                    if (s1hi[q] < hi_flag || !(d2[q] < 1e29) || ca[q] == trash) continue;
                    // exact class by binary search in the table
                    int lo_k = 0, hi_k = na + 1;
                    while (lo_k < hi_k) { const int mid = (lo_k + hi_k + 1) >> 1; if (d2[q] >= pc_lds_f64(T.thr + 8u * (unsigned)mid)) lo_k = mid; else hi_k = mid - 1; }
                    const int got = (int)((ca[q] - T.cnt) / 4u);
                    const bool ok = lo_k < na ? (got == lo_k) : (VARIANT == 0 && lo_k == na ? ca[q] == T.special : (got >= na && ca[q] != T.special));
                    if (!ok) atomicAdd(p.verify_fail, 1u);
                }
            }
#pragma unroll
            for (int q = 0; q < NP; ++q) pc_red(ca[q]);
        }
    }
}


// ---------------------------------------------------------------------------
// float32 segments (VARIANT 0 only, all three components in "single image" mode).
//
// The approximate class of a pair is computed entirely in float32 from LOCAL coordinates about the tile
// origin O (T = largest distance of a tile molecule from O, measured per tile):
//     xi~ = fl32(xi - O),  xj~ = fl32(fl64(xj + fl64(s - O))),  d~ = fl32(xj~ - xi~),
//     d2~ = fma(dz~,dz~, fma(dy~,dy~, fl32(dx~*dx~))),   r~ = sqrt.approx(d2~),
// two reference molecules per instruction (sm_100 packed FADD2 / FMUL2 / FFMA2).
// With u = 2^-24: |d~_k - d_k| <= u(|xi_k-O_k| + |xj~_k| + |d~_k|) <= u(2|d_k| + 2T_k)(1+o(1)), and
// | |d~| - |d| | <= |d~ - d| <= 2u(r + |T|): a relative error of 2u = 1.2e-7 in r plus the absolute term
// a = 2u|T|.  On top come the three roundings of d2~ (1.5u = 0.9e-7 in r), sqrt.approx (2^-23 = 1.2e-7) and the
// reference's own float32 chain fl(fl(sqrtf(fl(d2)))/step_a), which deviates from sqrt(d2)/step_a by
// 2.5u = 1.5e-7.  The bracket factors (1 -+ eps)/step_a are computed by the host in double and rounded
// outwards (PaPass::f32_lo/f32_hi), and the FFMA.RM is a single rounding, so the scale adds nothing.
//   * far runs (boxes >= 2|T| apart): a <= u r = 0.6e-7, total 5.4e-7 relative; eps = 5.5e-7 + delta_top because the
//     cutoff may sit delta_top <= 3e-7 off the boundary of the last bin (either side): every pair that close to the boundary
//     must reach the exact path, which also knows the class "inside the cutoff, a-bin out of range";
//   * NEAR runs: the brackets are evaluated at r~ -+ a (one more FADD2, whose rounding adds u): 5.4e-7.
// A pair whose brackets disagree, or whose |link_xy|^2 may be below 2^-30, is re-evaluated from the staged
// fp64 values exactly as pc_segment does, so results stay bit-identical.  Padded slots (tile tail, chunk
// tail) need no clamp: padded observed molecules sit at local (pad_x,0,0), beyond the cutoff for every
// tile molecule, and padded reference slots count into a shifted window of the counter array that is
// never read.
static __device__ __forceinline__ unsigned long long pc_pack(float lo, float hi)
{
    unsigned long long v;
    asm("mov.b64 %0, {%1, %2};" : "=l"(v) : "f"(lo), "f"(hi));
    return v;
}
static __device__ __forceinline__ float pc_lo(unsigned long long v) { return __uint_as_float((unsigned int)v); }
static __device__ __forceinline__ float pc_hi(unsigned long long v) { return __uint_as_float((unsigned int)(v >> 32)); }
static __device__ __forceinline__ unsigned long long pc_sub2(unsigned long long a, unsigned long long b)
{
    unsigned long long d;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
static __device__ __forceinline__ unsigned long long pc_add2(unsigned long long a, unsigned long long b)
{
    unsigned long long d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
static __device__ __forceinline__ unsigned long long pc_mul2(unsigned long long a, unsigned long long b)
{
    unsigned long long d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
static __device__ __forceinline__ unsigned long long pc_fma2(unsigned long long a, unsigned long long b, unsigned long long c)
{
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
static __device__ __forceinline__ unsigned long long pc_fma2_rm(unsigned long long a, unsigned long long b, unsigned long long c)
{
    unsigned long long d;
    asm("fma.rm.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}

static __device__ __forceinline__ float4 pc_lds_f32x4(unsigned int addr)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}

// Exact decisions of the float32 segments are deferred: a lane with a flagged pair (~1e-3 of the pairs) appends ONE
// entry for its iteration to a 32-entry queue of its warp, and the queue is drained with one entry per lane, so the fp64
// re-evaluation runs at full lane occupancy instead of stalling 31 idle lanes on every hit.  An entry is self-contained
// (sorted positions of the two observed molecules of the iteration, source lane, which of the 4 pairs), the drain reads
// the coordinates from global memory and decides the class from the fp64 threshold table alone, so entries survive
// refills of the staging buffer and a drain only happens when the queue is full or the run ends.
//   entry: bits 0-23 sorted position of observed molecule u = 0, 24-47 of u = 1, 48-52 source lane,
//          53-56 mask of the pairs to decide (bit u * PC_RI + r)
#define PC_QCAP 32
#define PC_FBUF 128                 // staged (compacted) observed molecules per warp
template <int MX, int MY, int MZ>
static __device__ __forceinline__ void pc_exact_pair(const PaCellGrid &go, const PaCellGrid &gi, long long jslot, long long islot,
                                                     double sxa, double sxb, double sya, double syb, double sza, double szb,
                                                     double &e1, double &e2)
{
    const double gx0 = go.sx[jslot], gy0 = go.sy[jslot], gz0 = go.sz[jslot];
    const double ex = pc_comp_sq<MX>(__dadd_rn(gx0, sxa), __dadd_rn(gx0, sxb), gi.sx[islot]);
    const double ey = pc_comp_sq<MY>(__dadd_rn(gy0, sya), __dadd_rn(gy0, syb), gi.sy[islot]);
    const double ez = pc_comp_sq<MZ>(__dadd_rn(gz0, sza), __dadd_rn(gz0, szb), gi.sz[islot]);
    e1 = __dadd_rn(ex, ey);
    e2 = __dadd_rn(e1, ez);
}
// exact class of d2 from the fp64 threshold table (global memory: the float32 launch keeps no copy in shared memory):
// the largest k <= na + 1 with thr_a[k] <= d2 (class na = inside the cutoff with the a-bin out of range, na + 1 = beyond)
static __device__ __forceinline__ int pc_exact_class(const PaPass &p, double d2, int na)
{
    int n = (int)fmin(sqrt(d2) * (double)p.job.guess_scale, (double)(na + 1));
    while (n > 0 && !(d2 >= __longlong_as_double((long long)__ldg(p.job.thr_a + n)))) --n;
    while (n <= na && d2 >= __longlong_as_double((long long)__ldg(p.job.thr_a + n + 1))) ++n;
    return n;
}
template <int MX, int MY, int MZ, bool ZCOL>
static __device__ __noinline__ void pc_f32_drain(const PaPass &p, const PaCellGrid &go, const PaCellGrid &gi, const PaTile &tl,
                                                 const unsigned long long *queue, int count, long long jbase,
                                                 double sxa, double sxb, double sya, double syb, double sza, double szb,
                                                 unsigned int cnt0, int na, unsigned long long cut, unsigned int special)
{
    __syncwarp();
    const int lane = threadIdx.x & 31;
    if (lane < count) {
        const unsigned long long ent = queue[lane];
        const int src = (int)((ent >> 48) & 31u);
        unsigned int mask = (unsigned int)((ent >> 53) & 15u);
        while (mask) {
            const int q = __ffs((int)mask) - 1;
            mask &= mask - 1u;
            const int r = q % PC_RI, u = q / PC_RI;
            const long long jpos = (long long)((ent >> (24 * u)) & 0xffffffull);
            const long long iref = (long long)tl.frame * gi.nmol + tl.begin + r * 32 + src;
            double e1, e2;
            pc_exact_pair<MX, MY, MZ>(go, gi, jbase + jpos, iref, sxa, sxb, sya, syb, sza, szb, e1, e2);
            if (ZCOL && __double2hiint(e1) < 0x3E100000) {
                // |link_xy|^2 < 2^-30: literal evaluation by pa_exception_kernel, not binned here
                const int io = gi.sorig[iref];
                const int jo = go.sorig[jbase + jpos];
                if (!(p.same_type && io == jo)) {
                    const unsigned int slot = atomicAdd(p.exc_count, 1u);
                    const long long fb = (long long)(p.frame0 + tl.frame) * p.pts.M;
                    if (slot < p.exc_cap)
                        p.exc_list[slot] = make_uint2((unsigned int)(fb + p.ref_off + io), (unsigned int)(fb + p.obs_off + jo));
                    if (p.sym || p.acc2) {
                        const unsigned int slot2 = atomicAdd(p.exc_count2, 1u);
                        if (slot2 < p.exc_cap)
                            p.exc_list2[slot2] = make_uint2((unsigned int)(fb + p.obs_off + jo), (unsigned int)(fb + p.ref_off + io));
                    }
                }
            } else {
                const int cls = pc_exact_class(p, e2, na);
                if (cls < na) pc_red(cnt0 + (unsigned)cls * 4u);
                else if (cls == na && dbits(e2) < cut) pc_red(special);          // inside the cutoff, a-bin out of range
                // else: beyond the cutoff, not counted
            }
        }
    }
    __syncwarp();
}

// per-half minimum of two packed pairs
static __device__ __forceinline__ unsigned long long pc_min2(unsigned long long a, unsigned long long b)
{
    return pc_pack(fminf(pc_lo(a), pc_lo(b)), fminf(pc_hi(a), pc_hi(b)));
}

// MX/MY/MZ: 1 = single image, 2 = two candidate images for that component (at most one component; its second
// candidate travels in the .w lane of the staged float4, and the squared component is the smaller of the two squares).
//
// Staging compacts: a lane converts one observed molecule of the run to local float32 coordinates and keeps it only if
// it can reach the tile at all -- with h = the tile's half extents about its origin (exact maxima of the |local float
// coordinates| of its molecules), every tile molecule i has |xj~ - xi~| >= g_x := max(0, |xj~| - h_x) (rounded down), and
// because rounding is monotone the d2~ the loop would compute for (i, j) is >= the same chain of operations evaluated
// on (g_x, g_y, g_z).  If that bound is >= dead_d2 (lower class bracket beyond the cutoff even after sqrt.approx's error
// and the NEAR term) no pair of j with this tile is ever counted or queued, and j is dropped.  The survivors are
// appended to the warp's buffer (ballot + popc), so the pair loop only ever sees molecules with something to count:
// for a cutoff of L/2 this removes the ~15 % of the evaluated pairs (and their atomics) that the cell-level pruning
// must leave in.
template <int MX, int MY, int MZ, bool ZCOL, bool NEAR, bool VERIFY>
static __device__ __forceinline__ void pc_segment_f32(const PaPass &p, const PaCellGrid &go, const PaCellGrid &gi, const PaTile &tl,
                                                      const float (&xf)[PC_RI], const float (&yf)[PC_RI], const float (&zf)[PC_RI],
                                                      const unsigned int (&cb)[PC_RI],
                                                      float hx, float hy, float hz,
                                                      double ox, double oy, double oz,
                                                      long long jbase, int jb, int je,
                                                      double sxa, double sxb, double sya, double syb, double sza, double szb,
                                                      float4 *s_jf, int *s_jidx, const PcTables &T, float near_a, unsigned long long *queue,
                                                      float dead_d2, unsigned int &it_all, unsigned int &it_live)
{
    static_assert(PC_RI == 2, "the packed float32 loop pairs the two reference molecules of a thread");
    static_assert(MX + MY + MZ <= 4, "at most one component with two candidate images");
    int qn = 0;                                         // entries in this warp's queue (warp-uniform)
    const int lane = threadIdx.x & 31;
    const unsigned int lane_lt = (1u << lane) - 1u;
    const unsigned long long X01 = pc_pack(xf[0], xf[1]), Y01 = pc_pack(yf[0], yf[1]), Z01 = pc_pack(zf[0], zf[1]);
    const unsigned long long F2 = pc_pack(p.f32_lo, p.f32_hi), M2 = pc_pack(8388608.0f, 8388608.0f), A2 = pc_pack(-near_a, near_a);
    const float zthr = 2.0e-9f;                         // 2^-30 plus the float32 error of dx~^2 + dy~^2 near the axis
    const int na = T.na;
    const unsigned int trash = T.cnt + (unsigned)(na + 2) * 4u;
    const unsigned int a_jf = (unsigned int)__cvta_generic_to_shared(s_jf);
    const long long ibase = (long long)tl.frame * gi.nmol + tl.begin;
    constexpr int NU = 2;                               // staged molecules per iteration (4 measured slower)
    static_assert(NU == 2, "queue entries and statistics assume two staged molecules per iteration");
    constexpr int NP = NU * PC_RI;
    const unsigned int tlim = 0x4B000000u + (unsigned)na;
    // pairs (bit u * PC_RI + r) whose reference slot r holds a real molecule of the tile
    const unsigned int kmask = (pc_kidx(0) < tl.count ? 0x5u : 0u) | (pc_kidx(1) < tl.count ? 0xAu : 0u);
    int fill = 0;                                       // molecules in the warp's buffer (warp-uniform)
    // local coordinate of an observed molecule: fl32(fl64(xj + c)) with c = fl64(s - O), image shift and tile origin folded
    // into one constant per run (the two fp64 roundings, < 1.3e-13 A together, are far inside the float32 budget)
    const double cxa = __dsub_rn(sxa, ox), cya = __dsub_rn(sya, oy), cza = __dsub_rn(sza, oz);
    const double cb2 = MX == 2 ? __dsub_rn(sxb, ox) : (MY == 2 ? __dsub_rn(syb, oy) : __dsub_rn(szb, oz));   // second candidate of the two-image component
    // the candidates of a staging step are loaded one step ahead (the global-memory latency overlaps the compaction
    // and the pair loop of the step before)
#if PC_PREFETCH
    double nx = 0.0, ny = 0.0, nz = 0.0;
    if (jb + lane < je) { nx = go.sx[jbase + jb + lane]; ny = go.sy[jbase + jb + lane]; nz = go.sz[jbase + jb + lane]; }
#endif
    for (int jc = jb; jc < je; jc += 32) {
        // ---- stage: one candidate per lane, keep the ones that can reach the tile -------------------------------
        {
            const int j = jc + lane;
            bool live = false;
            float4 f = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
#if PC_PREFETCH
            const double gx0 = nx, gy0 = ny, gz0 = nz;
            if (j + 32 < je) { nx = go.sx[jbase + j + 32]; ny = go.sy[jbase + j + 32]; nz = go.sz[jbase + j + 32]; }
#else
            double gx0 = 0.0, gy0 = 0.0, gz0 = 0.0;
            if (j < je) { gx0 = go.sx[jbase + j]; gy0 = go.sy[jbase + j]; gz0 = go.sz[jbase + j]; }
#endif
            if (j < je) {
                f.x = __double2float_rn(__dadd_rn(gx0, cxa));
                f.y = __double2float_rn(__dadd_rn(gy0, cya));
                f.z = __double2float_rn(__dadd_rn(gz0, cza));
                if (MX == 2 || MY == 2 || MZ == 2) f.w = __double2float_rn(__dadd_rn(MX == 2 ? gx0 : (MY == 2 ? gy0 : gz0), cb2));
                // lower bound of the d2~ of any pair (tile molecule, j): same operations as the pair loop below
                float g1 = fmaxf(__fsub_rd(fabsf(f.x), hx), 0.0f), g2 = fmaxf(__fsub_rd(fabsf(f.y), hy), 0.0f), g3 = fmaxf(__fsub_rd(fabsf(f.z), hz), 0.0f);
                float bound;
                if (MX == 2) {
                    g1 = fminf(g1, fmaxf(__fsub_rd(fabsf(f.w), hx), 0.0f));
                    bound = __fmaf_rn(g3, g3, __fmaf_rn(g2, g2, __fmul_rn(g1, g1)));
                } else if (MY == 2) {
                    g2 = fminf(g2, fmaxf(__fsub_rd(fabsf(f.w), hy), 0.0f));
                    bound = __fmaf_rn(g3, g3, __fadd_rn(__fmul_rn(g1, g1), __fmul_rn(g2, g2)));
                } else if (MZ == 2) {
                    g3 = fminf(g3, fmaxf(__fsub_rd(fabsf(f.w), hz), 0.0f));
                    bound = __fadd_rn(__fmaf_rn(g2, g2, __fmul_rn(g1, g1)), __fmul_rn(g3, g3));
                } else {
                    bound = __fmaf_rn(g3, g3, __fmaf_rn(g2, g2, __fmul_rn(g1, g1)));
                }
                live = !(bound >= dead_d2);
                if (VERIFY && !live) {
                    // a dropped molecule must be beyond the cutoff (class na + 1) for every molecule of the tile
#pragma unroll 1
                    for (int k = 0; k < tl.count; ++k) {
                        double e1, e2;
                        pc_exact_pair<MX, MY, MZ>(go, gi, jbase + j, ibase + k, sxa, sxb, sya, syb, sza, szb, e1, e2);
                        if (!(e2 >= __longlong_as_double((long long)__ldg(p.job.thr_a + (na + 1))))) atomicAdd(p.verify_fail, 1u);
                    }
                }
            }
            const unsigned int m = __ballot_sync(0xffffffffu, live);
            if (live) {
                const int pos = fill + __popc(m & lane_lt);
                s_jf[pos] = f;
                s_jidx[pos] = j;
            }
            fill += __popc(m);
        }
        if (!(fill > PC_FBUF - 32 || (jc + 32 >= je && fill > 0))) continue;
        // ---- evaluate the buffer ---------------------------------------------------------------------------------
        if ((fill & 1) && lane == 0) {                   // odd count: one padded slot, beyond the cutoff for every tile molecule
            s_jf[fill] = make_float4(T.pad_x, 0.0f, 0.0f, MX == 2 ? T.pad_x : 0.0f);
            s_jidx[fill] = 0;
        }
        __syncwarp();
        it_all += (unsigned)((fill + NU - 1) / NU); it_live += (unsigned)((fill + NU - 1) / NU);
        const unsigned int a_end = a_jf + (unsigned)fill * 16u;
        for (unsigned int aj = a_jf; aj < a_end; aj += 16u * NU) {
            float4 wv[NU];
#pragma unroll
            for (int u = 0; u < NU; ++u) wv[u] = pc_lds_f32x4(aj + 16u * u);
            unsigned int ca[NP], tlo[NP], thi[NP];
            float s1[NP];
            bool slow = false;
#pragma unroll
            for (int u = 0; u < NU; ++u) {
                // both reference molecules of the thread at once (packed float32 pairs, sm_100 FADD2 / FMUL2 / FFMA2)
                const float4 w = wv[u];
                const unsigned long long dx = pc_sub2(pc_pack(w.x, w.x), X01), dy = pc_sub2(pc_pack(w.y, w.y), Y01), dz = pc_sub2(pc_pack(w.z, w.z), Z01);
                unsigned long long s1p, d2p;
                if (MX == 2) {
                    const unsigned long long db = pc_sub2(pc_pack(w.w, w.w), X01);
                    s1p = pc_fma2(dy, dy, pc_min2(pc_mul2(dx, dx), pc_mul2(db, db)));
                    d2p = pc_fma2(dz, dz, s1p);
                } else if (MY == 2) {
                    const unsigned long long db = pc_sub2(pc_pack(w.w, w.w), Y01);
                    s1p = pc_add2(pc_mul2(dx, dx), pc_min2(pc_mul2(dy, dy), pc_mul2(db, db)));
                    d2p = pc_fma2(dz, dz, s1p);
                } else if (MZ == 2) {
                    const unsigned long long db = pc_sub2(pc_pack(w.w, w.w), Z01);
                    s1p = pc_fma2(dy, dy, pc_mul2(dx, dx));
                    d2p = pc_add2(s1p, pc_min2(pc_mul2(dz, dz), pc_mul2(db, db)));
                } else {
                    s1p = pc_fma2(dy, dy, pc_mul2(dx, dx));
                    d2p = pc_fma2(dz, dz, s1p);
                }
#pragma unroll
                for (int r = 0; r < PC_RI; ++r) {
                    const int q = u * PC_RI + r;
                    float rt;
                    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(rt) : "f"(r ? pc_hi(d2p) : pc_lo(d2p)));
                    unsigned long long rr = pc_pack(rt, rt);
                    if (NEAR) rr = pc_add2(rr, A2);                               // (r~ - a, r~ + a)
                    const unsigned long long tt = pc_fma2_rm(rr, F2, M2);        // (MAGIC + floor(lo bracket), MAGIC + floor(hi bracket))
                    tlo[q] = __float_as_uint(pc_lo(tt)); thi[q] = __float_as_uint(pc_hi(tt));
                    s1[q] = r ? pc_hi(s1p) : pc_lo(s1p);
                    slow |= tlo[q] != thi[q];
                    if (ZCOL) slow |= s1[q] < zthr;
                    ca[q] = tlo[q] * 4u + cb[r];
                }
            }
            if (__any_sync(0xffffffffu, slow)) {                                  // warp-uniform: ~10% of the iterations
                const int jj = (int)((aj - a_jf) >> 4);
                unsigned int need = 0u;
#pragma unroll
                for (int q = 0; q < NP; ++q) {
                    // not when even the lower bracket is beyond the cutoff (class > na: never counted; this also covers the
                    // padded observed slot, which sits beyond the cutoff), and not for padded reference slots (kmask)
                    if ((tlo[q] != thi[q] || (ZCOL && s1[q] < zthr)) && tlo[q] <= tlim) need |= 1u << q;
                }
                need &= kmask;
#pragma unroll
                for (int q = 0; q < NP; ++q) if (need & (1u << q)) ca[q] = trash;   // counted by the drain
                const unsigned int m = __ballot_sync(0xffffffffu, need != 0u);
                if (m != 0u) {
                    const int add = __popc(m);
                    if (qn + add > PC_QCAP) {
                        pc_f32_drain<MX, MY, MZ, ZCOL>(p, go, gi, tl, queue, qn, jbase, sxa, sxb, sya, syb, sza, szb, T.cnt, na, T.cut, T.special);
                        qn = 0;
                    }
                    if (need)
                        queue[qn + __popc(m & lane_lt)] = (unsigned long long)(unsigned int)s_jidx[jj] | ((unsigned long long)(unsigned int)s_jidx[jj + 1] << 24) |
                                                          ((unsigned long long)lane << 48) | ((unsigned long long)need << 53);
                    qn += add;
                }
            }
            if (VERIFY) {
                const int jj = (int)((aj - a_jf) >> 4);
#pragma unroll
                for (int q = 0; q < NP; ++q) {
                    const int r = q % PC_RI, u = q / PC_RI;
                    if (pc_kidx(r) >= tl.count || jj + u >= fill || ca[q] == trash) continue;
                    double e1, e2;
                    pc_exact_pair<MX, MY, MZ>(go, gi, jbase + s_jidx[jj + u], ibase + pc_kidx(r), sxa, sxb, sya, syb, sza, szb, e1, e2);
                    int lo_k = 0, hi_k = na + 1;
                    while (lo_k < hi_k) { const int mid = (lo_k + hi_k + 1) >> 1; if (e2 >= __longlong_as_double((long long)__ldg(p.job.thr_a + mid))) lo_k = mid; else hi_k = mid - 1; }
                    const int got = (int)((ca[q] - T.cnt) / 4u);
                    const bool ok = (__double2hiint(e1) >= 0x3E100000) && (lo_k < na ? (got == lo_k) : (got >= na && lo_k != na));   // class na is always queued
                    if (!ok) atomicAdd(p.verify_fail, 1u);
                }
            }
#pragma unroll
            for (int q = 0; q < NP; ++q) pc_red(ca[q]);
        }
        fill = 0;
        __syncwarp();                                    // the buffer is refilled by the next staging step
    }
    if (qn > 0) pc_f32_drain<MX, MY, MZ, ZCOL>(p, go, gi, tl, queue, qn, jbase, sxa, sxb, sya, syb, sza, szb, T.cnt, na, T.cut, T.special);   // the shifts belong to this run
}

template <int VARIANT>
static __device__ void pc_flush(const PaPass &p, const PcTables &T, bool reset)
{
    const int na = T.na;
    const long long mult = p.sym ? 2 : 1;                // symmetric pass: every unordered pair was seen once
    unsigned long long within = 0;
    for (int k = threadIdx.x; k < na + 3; k += blockDim.x) {
        const unsigned int c = pc_lds_u32(T.cnt + (unsigned)k * 4u);
        if (c) {
            if (k < na) {
                atomicAdd(&p.job.acc[k], (unsigned long long)((long long)c * p.weight * mult));
                if (p.acc2) atomicAdd(&p.acc2[k], (unsigned long long)((long long)c * p.weight2));
                within += c;
            } else if (VARIANT == 1 && k == na) {                                                   // inside the cutoff, a-bin out of range
                atomicAdd(&p.job.acc[na + 1], (unsigned long long)((long long)c * mult));
                if (p.acc2) atomicAdd(&p.acc2[na + 1], (unsigned long long)c);
            }
            if (reset) pc_sts_u32(T.cnt + (unsigned)k * 4u, 0u);
        }
    }
    if (within) {
        atomicAdd(&p.job.acc[na], within * (unsigned long long)mult);
        if (p.acc2) atomicAdd(&p.acc2[na], within);
        if (p.kstats) atomicAdd(p.kstats + 2, within);               // pairs counted, each evaluated once
    }
    if (VARIANT == 0 && threadIdx.x == 0) {                                                          // class na found by an exact path
        const unsigned int c = pc_lds_u32(T.special);
        if (c) {
            atomicAdd(&p.job.acc[na + 1], (unsigned long long)((long long)c * mult));
            if (p.acc2) atomicAdd(&p.acc2[na + 1], (unsigned long long)c);
            if (reset) pc_sts_u32(T.special, 0u);
        }
    }
}

// The same flush by ONE warp while the other warps of the CTA keep counting: every counter is fetched and zeroed with
// an atomic exchange, so no increment is lost or seen twice.  Used before the 32-bit counters could wrap.
static __device__ __forceinline__ unsigned int pc_exch_u32(unsigned int addr)
{
    unsigned int v;
    asm volatile("atom.shared.exch.b32 %0, [%1], 0;" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
template <int VARIANT>
static __device__ __noinline__ void pc_flush_warp(const PaPass &p, const PcTables &T)
{
    const int na = T.na, lane = threadIdx.x & 31;
    const long long mult = p.sym ? 2 : 1;
    unsigned long long within = 0;
    for (int k = lane; k < na + 3; k += 32) {
        if (k >= na && !(VARIANT == 1 && k == na)) continue;                                        // ignored classes
        const unsigned int c = pc_exch_u32(T.cnt + (unsigned)k * 4u);
        if (!c) continue;
        if (k < na) {
            atomicAdd(&p.job.acc[k], (unsigned long long)((long long)c * p.weight * mult));
            if (p.acc2) atomicAdd(&p.acc2[k], (unsigned long long)((long long)c * p.weight2));
            within += c;
        } else {
            atomicAdd(&p.job.acc[na + 1], (unsigned long long)((long long)c * mult));
            if (p.acc2) atomicAdd(&p.acc2[na + 1], (unsigned long long)c);
        }
    }
    if (within) {
        atomicAdd(&p.job.acc[na], within * (unsigned long long)mult);
        if (p.acc2) atomicAdd(&p.acc2[na], within);
        if (p.kstats) atomicAdd(p.kstats + 2, within);
    }
    if (VARIANT == 0 && lane == 0) {
        const unsigned int c = pc_exch_u32(T.special);
        if (c) {
            atomicAdd(&p.job.acc[na + 1], (unsigned long long)((long long)c * mult));
            if (p.acc2) atomicAdd(&p.acc2[na + 1], (unsigned long long)c);
        }
    }
    __syncwarp();
}

// warp-wide minimum / maximum of a double (all lanes receive the result)
static __device__ __forceinline__ double pc_warp_min(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
static __device__ __forceinline__ double pc_warp_max(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// PART 0: every run in one launch.  PART 1 / PART 2: the float32-eligible runs (single image or one two-image
// component) and all other runs as two launches over the same work items, so that each gets its own register allocation.
static __host__ __device__ __forceinline__ size_t pc_warp_bytes(bool part1, int nlay)
{
    // per warp (16-byte aligned slices): PART 1: staged local float4 (PC_FBUF) + their sorted positions + queue of deferred
    // exact decisions; other parts: staged doubles (PC_JC * 6); both: layer classification of the warp's tile
    const size_t stage = part1 ? (size_t)PC_FBUF * 16 + (size_t)PC_FBUF * 4 + (size_t)PC_QCAP * 8 : (size_t)PC_JC * 6 * 8;
    return stage + (((size_t)nlay * sizeof(PcLayer) + 15) & ~(size_t)15);
}
template <int VARIANT, bool VERIFY, bool SYM, int PART>
__global__ void __launch_bounds__(PC_THREADS, PART == 1 ? PC_MINB : 0) pa_rdf_cells_kernel(const __grid_constant__ PaPass p, const __grid_constant__ PaCellGrid gi, const __grid_constant__ PaCellGrid go, const PaTile *tiles,
                                                                   const unsigned int *ntiles_ptr, unsigned int tile_cap,
                                                                   unsigned int *work_counter, int nzg)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int na = p.job.bin_a;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool f32_on = VARIANT == 0 && PART != 0 && p.f32_cnt > 0;       // PART 0 is launched only when float32 segments are off
    const int ncnt = ((f32_on && PART == 1) ? p.f32_cnt + na + 8 : na + 3) + 1;                          // counters (see pc_segment_f32) + the class-na counter
    const int nlay = go.g[0] + go.g[1] + go.g[2];
    // CTA-wide: threshold table (PART 1 reads it from global memory when needed) and the class counters
    unsigned long long *s_thr = reinterpret_cast<unsigned long long *>(smem_raw);          // na+3 (PART 1: none)
    unsigned int *s_cnt = reinterpret_cast<unsigned int *>(s_thr + (PART == 1 ? 0 : na + 3));   // ncnt (+1 pad)
    unsigned char *wbase = smem_raw + ((((size_t)(PART == 1 ? 0 : na + 3) * 8 + (size_t)(ncnt + 1) * 4) + 15) & ~(size_t)15) +
                           (size_t)warp * pc_warp_bytes(PART == 1, nlay);
    double *s_j = reinterpret_cast<double *>(wbase);                                       // PART 0 / 2
    float4 *s_jf = reinterpret_cast<float4 *>(wbase);                                      // PART 1
    int *s_jidx = reinterpret_cast<int *>(s_jf + PC_FBUF);
    unsigned long long *queue = reinterpret_cast<unsigned long long *>(s_jidx + PC_FBUF);
    PcLayer *lay = reinterpret_cast<PcLayer *>(wbase + (PART == 1 ? (size_t)PC_FBUF * 20 + (size_t)PC_QCAP * 8 : (size_t)PC_JC * 6 * 8));

    for (int k = threadIdx.x; k < (PART == 1 ? 0 : na + 3); k += blockDim.x)
        s_thr[k] = p.job.thr_a[k];       // "never" (all ones) is a NaN as a double: every `d2 >= NaN` is false, as required
    for (int k = threadIdx.x; k < ncnt; k += blockDim.x) s_cnt[k] = 0u;
    __syncthreads();

    PcTables T;
    T.thr = (unsigned int)__cvta_generic_to_shared(s_thr);
    T.cnt = (unsigned int)__cvta_generic_to_shared(s_cnt);
    T.na = na;
    T.cut = p.job.cut_bits;
    if (VARIANT == 0) {
        T.s_lo = p.job.guess_scale * (1.0f - p.eps64);
        T.s_hi = p.job.guess_scale * (1.0f + p.eps64);
    } else {
        T.s_lo = T.s_hi = p.job.guess_scale * (1.0f - 9.5367431640625e-07f);             // floor(x~) <= true class
    }
    T.r_max = ((float)na + 1.5f) / p.job.guess_scale;
    T.special = T.cnt + (unsigned)(ncnt - 1) * 4u;
    const float t2max = 32.0f;                                                            // largest |T| a float32 tile may have
    T.pad_x = p.job.maxdist * 1.001f + 3.0f * t2max;
    T.cnt_pad = (unsigned)(na + 8) * 4u;
    const double cut_d2 = __longlong_as_double((long long)T.cut) * (1.0 + 1e-12);
    const float cutf = (float)fmin(cut_d2 * (1.0 + 1e-6), 3.0e38);                        // NaN ("never") -> 3e38
    const unsigned int ntiles = min(*ntiles_ptr, tile_cap);
    if (*ntiles_ptr > tile_cap && threadIdx.x == 0 && blockIdx.x == 0) atomicOr(p.flags + 1, 1u);
    const long long n_items = (long long)ntiles * nzg;
    unsigned long long since_flush = 0;                       // this warp's increments since its own last flush
    const int gx = go.g[0], gy = go.g[1], gz = go.g[2];
    // statistics (warp-uniform): loop iterations (4 pair evaluations per lane each) and those that issued their atomics
    unsigned int it_all = 0, it_live = 0;
    unsigned long long st_all = 0, st_live = 0;

    for (;;) {
        st_all += it_all; st_live += it_live; it_all = it_live = 0;
        __syncwarp();
        unsigned int w = 0;
        if (lane == 0) w = atomicAdd(work_counter, 1u);
        w = __shfl_sync(0xffffffffu, w, 0);
        const long long item = (long long)w * p.tshard_count + p.tshard_rank;   // tile shard: items w with w % count == rank
        if (item >= n_items) break;
        const PaTile tl = tiles[item / nzg];
        const int zg = (int)(item % nzg);
        const int z0 = (int)((long long)gz * zg / nzg), z1 = (int)((long long)gz * (zg + 1) / nzg);
        if (z0 == z1) continue;
        // reference molecules of the tile -> registers; padded slots are a copy of the tile's last molecule for the
        // float32 coordinates (so that they never decide a warp-uniform test) and far away for the fp64 segments
        double p1x[PC_RI], p1y[PC_RI], p1z[PC_RI];
        double q1x[PC_RI], q1y[PC_RI], q1z[PC_RI];
        const long long ibase = (long long)tl.frame * gi.nmol;
#pragma unroll
        for (int r = 0; r < PC_RI; ++r) {
            const int k = pc_kidx(r), kk = min(k, tl.count - 1);
            q1x[r] = gi.sx[ibase + tl.begin + kk]; q1y[r] = gi.sy[ibase + tl.begin + kk]; q1z[r] = gi.sz[ibase + tl.begin + kk];
            if (k < tl.count) { p1x[r] = q1x[r]; p1y[r] = q1y[r]; p1z[r] = q1z[r]; }
            else p1x[r] = p1y[r] = p1z[r] = -1.0e15;   // padding: far away, finite in float range
        }
        // exact bounding box of the tile (padded slots repeat a real molecule, so they do not widen it)
        double blox = q1x[0], bloy = q1y[0], bloz = q1z[0], bhix = q1x[0], bhiy = q1y[0], bhiz = q1z[0];
#pragma unroll
        for (int r = 1; r < PC_RI; ++r) {
            blox = fmin(blox, q1x[r]); bhix = fmax(bhix, q1x[r]);
            bloy = fmin(bloy, q1y[r]); bhiy = fmax(bhiy, q1y[r]);
            bloz = fmin(bloz, q1z[r]); bhiz = fmax(bhiz, q1z[r]);
        }
        blox = pc_warp_min(blox); bloy = pc_warp_min(bloy); bloz = pc_warp_min(bloz);
        bhix = pc_warp_max(bhix); bhiy = pc_warp_max(bhiy); bhiz = pc_warp_max(bhiz);
        // float32 segments: local coordinates about the centre of the tile's bounding box
        float xf[PC_RI], yf[PC_RI], zf[PC_RI];
        unsigned int cb[PC_RI];
        const double ox = 0.5 * (blox + bhix), oy = 0.5 * (bloy + bhiy), oz = 0.5 * (bloz + bhiz);
        float tile_t2 = 0.0f, hx = 0.0f, hy = 0.0f, hz = 0.0f;
        if (f32_on) {
#pragma unroll
            for (int r = 0; r < PC_RI; ++r) {
                const bool live = pc_kidx(r) < tl.count;
                xf[r] = __double2float_rn(__dsub_rn(q1x[r], ox));
                yf[r] = __double2float_rn(__dsub_rn(q1y[r], oy));
                zf[r] = __double2float_rn(__dsub_rn(q1z[r], oz));
                cb[r] = T.cnt - 0x4B000000u * 4u + (live ? 0u : T.cnt_pad);      // padded slots count into a window nobody reads
                tile_t2 = fmaxf(tile_t2, __fmaf_ru(zf[r], zf[r], __fmaf_ru(yf[r], yf[r], __fmul_ru(xf[r], xf[r]))));
                hx = fmaxf(hx, fabsf(xf[r])); hy = fmaxf(hy, fabsf(yf[r])); hz = fmaxf(hz, fabsf(zf[r]));
            }
            tile_t2 = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(tile_t2)));
            // half extents of the tile in local float32 coordinates (non-negative floats order like their bit patterns)
            hx = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(hx)));
            hy = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(hy)));
            hz = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(hz)));
        } else {
#pragma unroll
            for (int r = 0; r < PC_RI; ++r) { xf[r] = yf[r] = zf[r] = 0.0f; cb[r] = 0u; }
        }
        // classification of every observed layer against the tile's bounding box
        __syncwarp();
        for (int t = lane; t < nlay; t += 32) {
            const int c = t < gx ? 0 : (t < gx + gy ? 1 : 2);
            const int k = c == 0 ? t : (c == 1 ? t - gx : t - gx - gy);
            const long long lj = ((long long)tl.frame * 3 + c) * go.gmax;
            PcLayer cl;
            if (go.lay_lo[lj + k] > go.lay_hi[lj + k]) { cl.mode = 0; cl.ca = cl.cb = 1; cl.pad = 0; cl.dmin2 = 3.0e38f; }   // empty layer
            else cl = pc_classify(c == 0 ? blox : (c == 1 ? bloy : bloz), c == 0 ? bhix : (c == 1 ? bhiy : bhiz),
                                  okey_inv(go.lay_lo[lj + k]), okey_inv(go.lay_hi[lj + k]), c == 0 ? p.box.L[0] : (c == 1 ? p.box.L[1] : p.box.L[2]));
            if (cl.mode == 3) atomicOr(p.flags + 1, 1u);                       // boxes too wide: host redoes the call without cells
            lay[t] = cl;
        }
        __syncwarp();
        // 32-bit shared counters: the CTA's warps together add at most PC_WARPS * 2^28 between two flushes of any one of them
        const unsigned long long item_pairs = (unsigned long long)PC_TILE * (unsigned long long)go.nmol;
        if (since_flush + item_pairs > (0x7fffffffull / PC_WARPS)) {
            pc_flush_warp<VARIANT>(p, T);
            since_flush = 0;
        }
        since_flush += item_pairs;
        // float32 eligibility of this tile: |T| <= t2max; runs must then be at least 2|T| away (rounded up)
        const bool tile_f32 = f32_on && tile_t2 <= t2max * t2max;
        if (PART == 1 && !tile_f32) continue;
        const float rmin2 = 4.0f * tile_t2 * 1.0001f + 1.0e-3f;
        const float near_a = 1.25e-7f * sqrtf(tile_t2) + 1.0e-12f;       // 2u|T| (u = 2^-24), rounded up
        // d2~ from which the lower bracket floor((sqrt.approx(d2~) - a) * f32_lo) is certainly > na (sqrt.approx: 2^-22 relative)
        const float dead_r = __fadd_ru(p.dead_r, near_a);
        const float dead_d2 = __fmul_ru(__fmul_ru(dead_r, dead_r), 1.000001f);

        const long long jbase = (long long)tl.frame * go.nmol;
        const int *cs = go.cell_start + (long long)tl.frame * (go.ncell + 1);
        // symmetric pass: rows (cy,cz) that come before the tile's own row in the sorted order are skipped,
        // they see this tile as "after" when the roles are reversed
        const int tile_row = tl.cz * gy + tl.cy, tile_end = tl.begin + tl.count;
        for (int cz = z0; cz < z1; ++cz) {
            const PcLayer lz = lay[gx + gy + cz];
            if (!(lz.dmin2 < cutf)) continue;
            if (SYM && cz < tl.cz) continue;
            for (int cy = 0; cy < gy; ++cy) {
                if (SYM && cz * gy + cy < tile_row) continue;
                const PcLayer ly = lay[gx + cy];
                const float dyz2 = lz.dmin2 + ly.dmin2;
                if (!(dyz2 < cutf)) continue;
                // PART 2 has nothing to do in a row whose y and z layers are single-image, at least 2|T| away and not in the
                // tile's own xy column: every run there is float32-eligible whatever its x class (nearr and zcol are false)
                if (PART == 2 && tile_f32 && ly.mode == 1 && lz.mode == 1 && dyz2 >= rmin2 && ly.dmin2 > 1.0e-6f &&
                    !(SYM && cz * gy + cy == tile_row)) continue;
                const int *row = cs + ((long long)cz * gy + cy) * gx;
                int cx = 0;
                while (cx < gx) {
                    const PcLayer lx = lay[cx];
                    if (!(dyz2 + lx.dmin2 < cutf)) { ++cx; continue; }            // cannot reach the cutoff
                    // float32 segment: all components single-image, boxes >= 2|T| apart; zcol: the link vector of a
                    // pair of this run can be (anti)parallel to z
                    const bool far = tile_f32;                                     // every single-image run of a float32 tile
                    const bool nearr = !(dyz2 + lx.dmin2 >= rmin2);                // closer than 2|T|: brackets get the absolute term
                    const bool zcol = !(lx.dmin2 + ly.dmin2 > 1.0e-6f);
                    int ce = cx;                                                   // extend while the x classification is identical
                    while (ce + 1 < gx) {
                        const PcLayer ln = lay[ce + 1];
                        if (ln.mode != lx.mode || ln.ca != lx.ca || ln.cb != lx.cb || !(dyz2 + ln.dmin2 < cutf)) break;
                        if (tile_f32 && (!(dyz2 + ln.dmin2 >= rmin2) != nearr || !(ln.dmin2 + ly.dmin2 > 1.0e-6f) != zcol)) break;
                        ++ce;
                    }
                    int jb = row[cx], je = row[ce + 1];
                    cx = ce + 1;
                    // float32-eligible runs belong to the PART 1 launch -- except the part of a run that overlaps the tile
                    // itself in a symmetric pass (db..de below), which always takes the fp64 segment of PART 0 / 2
                    // (all components single image, or -- far runs only -- exactly one component with two candidates)
                    const int dcode = (lx.mode >= 2 ? 1 : 0) | (ly.mode >= 2 ? 2 : 0) | (lz.mode >= 2 ? 4 : 0);
                    const bool f32run = far && lx.mode <= 2 && ly.mode <= 2 && lz.mode <= 2 &&
                                        (dcode == 0 || (!nearr && (dcode == 4 || ((dcode == 1 || dcode == 2) && !zcol))));
                    if (PART == 1 && !f32run) continue;
                    int db = 0, de = 0;                                              // part of the run that overlaps the tile itself
                    if (SYM && cz * gy + cy == tile_row) {
                        if (je <= tl.begin) continue;                               // entirely before the tile
                        db = max(jb, tl.begin); de = min(je, tile_end);
                        jb = max(jb, tile_end);
                    }
                    if (jb >= je && db >= de) continue;
                    const double Lx = p.box.L[0], Ly = p.box.L[1], Lz = p.box.L[2];
                    const double sxa = (double)(lx.ca - 1) * Lx, sxb = (double)(lx.cb - 1) * Lx;
                    const double sya = (double)(ly.ca - 1) * Ly, syb = (double)(ly.cb - 1) * Ly;
                    const double sza = (double)(lz.ca - 1) * Lz, szb = (double)(lz.cb - 1) * Lz;
                    const int code = (lx.mode >= 2 ? 1 : 0) | (ly.mode >= 2 ? 2 : 0) | (lz.mode >= 2 ? 4 : 0);
                    if (PART != 1 && SYM && db < de) {
                        // staging needs an even start so that pairs of records stay 16-byte aligned: start at db & ~1,
                        // the extra molecule (if any) lies before the tile or inside it and is masked by ipos < jpos
                        pc_segment<2, 2, 2, VARIANT, VERIFY, true>(p, go, gi, tl, p1x, p1y, p1z, jbase, db, de, sxa, sxb, sya, syb, sza, szb, s_j, T, it_all, it_live);
                    }
                    if (jb >= je) continue;
                    if (PART == 2 && f32run) continue;                              // counted by the PART 1 launch
                    if (VARIANT == 0 && PART == 1 && f32run) {
#define PC_F32(MX_, MY_, MZ_, Z_, N_) pc_segment_f32<MX_, MY_, MZ_, Z_, N_, VERIFY>(p, go, gi, tl, xf, yf, zf, cb, hx, hy, hz, ox, oy, oz, jbase, jb, je, \
                                                                                    sxa, sxb, sya, syb, sza, szb, s_jf, s_jidx, T, near_a, queue, dead_d2, it_all, it_live)
                        if (dcode == 0) {
                            if (nearr) { if (zcol) PC_F32(1, 1, 1, true, true); else PC_F32(1, 1, 1, false, true); }
                            else { if (zcol) PC_F32(1, 1, 1, true, false); else PC_F32(1, 1, 1, false, false); }
                        } else if (dcode == 1) PC_F32(2, 1, 1, false, false);
                        else if (dcode == 2) PC_F32(1, 2, 1, false, false);
                        else { if (zcol) PC_F32(1, 1, 2, true, false); else PC_F32(1, 1, 2, false, false); }
#undef PC_F32
                        continue;
                    }
                    if (PART == 1) continue;
#define PC_SEG(MX_, MY_, MZ_)                                                                                              \
    pc_segment<MX_, MY_, MZ_, VARIANT, VERIFY, false>(p, go, gi, tl, p1x, p1y, p1z, jbase, jb, je, sxa, sxb, sya, syb, sza, szb,  \
                                               s_j, T, it_all, it_live)
                    switch (code) {
                    case 0: PC_SEG(1, 1, 1); break;
                    case 1: PC_SEG(2, 1, 1); break;
                    case 2: PC_SEG(1, 2, 1); break;
                    case 3: PC_SEG(2, 2, 1); break;
                    case 4: PC_SEG(1, 1, 2); break;
                    case 5: PC_SEG(2, 1, 2); break;
                    case 6: PC_SEG(1, 2, 2); break;
                    default: PC_SEG(2, 2, 2); break;
                    }
#undef PC_SEG
                }
            }
        }
    }
    if (p.kstats && lane == 0) {
        st_all += it_all; st_live += it_live;
        unsigned long long *ks = p.kstats + (PART == 1 ? 0 : 4);    // the float32 launch is accounted separately
        if (st_all) atomicAdd(ks + 0, st_all * 128ull);             // pair evaluations (lanes x 4 per iteration)
        if (st_live) atomicAdd(ks + 1, st_live * 128ull);           // shared-memory atomics issued (one per evaluated pair of a live iteration)
    }
    __syncthreads();
    pc_flush<VARIANT>(p, T, false);
}

// ---------------------------------------------------------------------------
// Cell-sorted 2-D (distance x polar angle) kernel for `pdf` jobs with bin_count_b > 1 or any reference vector: the
// warp-autonomous tiles, layer classification, staging filter and compaction of the RDF kernel, with a pair loop that
// brackets BOTH classes in float32:
//   a-class  as in pc_segment_f32 (MUFU.SQRT, two round-down FMAs; brackets must agree);
//   b-class  = #{k >= 1 : dot <= thr_b[k]}, dot = link.v / |link| (Distribution:812-827).  In local float32
//            coordinates: N~ = fma(dz~,vz~, fma(dy~,vy~, dx~ vx~)), dot~ = N~ * rcp.approx(r~).  With u = 2^-24,
//            |d~ - d| <= 2u(r + |T|), |v~ - v| <= u:  |N~ - N| <= u(8.2 r + 2|T|);  r~ = sqrt.approx(d2~) is within
//            5.5u r + 2u|T| of r and rcp.approx adds 2u, so 1/r~ = (1/r)(1 + rho), |rho| <= 7.5u + 2u|T|/r; the product
//            rounds once more:   |dot~ - dot| <= u(16.7 + 4|T|/r)   (first order; pairs with a bound above 1e-3, i.e.
//            r within ~1e-4 |T| of zero, are never decided here).  E = 1.2e-6 + 2.6e-7 |T| / r~ covers that with 5 %
//            to spare plus 1e-7 for the float32 rounding of the threshold table and of dot~ +- E; the reference's own
//            fp64 normalisation and the wrap-shift bookkeeping of its link vector differ from the exact quotient by
//            < 1e-12.  A class c seeded by an approximate acos is accepted when the whole interval
//            [dot~ - E, dot~ + E] lies inside (thr_b[c+1], thr_b[c]] of the float threshold table.
// A pair whose a- or b-class is not certain, whose b-class would be bin_count_b (polar angle pi: out of bounds), or
// that may be the molecule itself goes to the warp's queue and is decided by pa_eval_pair_exact from the original
// slots (the literal evaluation every other kernel uses), so the result is bit-identical by construction.  Runs that
// are not float32-eligible (two candidate images for a component, tiles wider than 32 A) are evaluated pair by pair in
// fp64 and every pair inside the cutoff takes the literal path.  Needs the same job properties as the float32 RDF
// launch (cutoff within 3e-7 of the last bin boundary), no centre of charge, and a cutoff below
// maximum_distance_squared; other jobs use pa_general_cells_kernel below.
struct PdfConsts {
    float vx, vy, vz;          // reference vector, float32
    float e_a, e_b;            // E = e_a + e_b / r~
    unsigned int tab;          // shared address of the float threshold table T[0..nb+1] (T[0] = +inf, T[nb+1] = -inf)
    int nb;                    // bin_count_b
    float inv_step_b;
    unsigned int hist;         // shared address of the CTA's 32-bit histogram [bin_a * bin_b | out of bounds], 0: global atomics
};

// candidate b-class of a dot product: floor(acos~(x) / step_b) from a 4-term approximation of acos (|error| < 7e-5 rad,
// Abramowitz & Stegun 4.4.45), moved by at most one class towards the interval of the float threshold table that holds
// x (T decreasing, T[0] = +inf, T[nb+1] = -inf).  Only a seed: the caller accepts it when [x - E, x + E] lies inside
// (T[c+1], T[c]] and sends the pair to the literal evaluation otherwise, so a wrong seed costs time, never a count.
static __device__ __forceinline__ int pdf_class(const PdfConsts &C, float x, float &t_hi, float &t_lo)
{
    const float ax = fminf(fabsf(x), 1.0f);
    float s;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(s) : "f"(1.0f - ax));
    float th = s * __fmaf_rn(__fmaf_rn(__fmaf_rn(-0.0187293f, ax, 0.0742610f), ax, -0.2121144f), ax, 1.5707288f);
    if (x < 0.0f) th = 3.14159265f - th;
    int c = min(max((int)(th * C.inv_step_b), 0), C.nb - 1);
    float t0, t1;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(t0) : "r"(C.tab + 4u * (unsigned)c));
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(t1) : "r"(C.tab + 4u * (unsigned)(c + 1)));
    c += (x <= t1) ? 1 : ((x > t0 && c > 0) ? -1 : 0);
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(t_hi) : "r"(C.tab + 4u * (unsigned)c));
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(t_lo) : "r"(C.tab + 4u * (unsigned)(c + 1)));
    return c;
}

template <bool VERIFY>
static __device__ __noinline__ void pdf_drain(const PaPass &p, const PaCellGrid &go, const PaCellGrid &gi, const PaTile &tl,
                                              const unsigned long long *queue, int count, long long jbase, unsigned int hist,
                                              unsigned long long &g_within, unsigned long long &g_oob)
{
    __syncwarp();
    const int lane = threadIdx.x & 31;
    if (lane < count) {
        const unsigned long long ent = queue[lane];
        const int src = (int)((ent >> 48) & 31u);
        unsigned int mask = (unsigned int)((ent >> 53) & 15u);
        const long long fb = (long long)(p.frame0 + tl.frame) * p.pts.M;
        while (mask) {
            const int q = __ffs((int)mask) - 1;
            mask &= mask - 1u;
            const int r = q % PC_RI, u = q / PC_RI;
            const long long jpos = (long long)((ent >> (24 * u)) & 0xffffffull);
            const long long iref = (long long)tl.frame * gi.nmol + tl.begin + r * 32 + src;
            const int io = gi.sorig[iref], jo = go.sorig[jbase + jpos];
            if (p.same_type && io == jo) continue;                       // the molecule itself (Distribution:778)
            int bin = 0;
            const int st = pa_eval_pair_exact(p, fb + p.ref_off + io, fb + p.obs_off + jo, &bin);
            if (st == 1) {
                if (hist) pc_red(hist + 4u * (unsigned)bin);
                else { atomicAdd(&p.job.acc[bin], (unsigned long long)p.weight); ++g_within; }
            } else if (st == 2) {
                if (hist) pc_red(hist + 4u * (unsigned)(p.job.bin_a * p.job.bin_b));
                else ++g_oob;
            }
        }
    }
    __syncwarp();
}

// Pair loop in two phases, because only about half of the pairs that survive the staging filter are inside the cutoff
// and the classification costs ~60 instructions: phase 1 evaluates d2~ for the 4 pairs of a lane's iteration (packed)
// and appends the pairs that may be inside the cutoff -- (dx~, dy~, dz~, reference index | observed position) -- to a
// stack of the warp (ballot + popc); whenever the stack holds 64 entries or more, phase 2 pops two entries per lane and
// classifies them at full lane occupancy.  A popped pair that is not certain is evaluated literally on the spot (~1e-3
// of the pairs).
#define PDF_STACK 192               // 63 left over + 4 x 32 appended by one iteration
// classification of one popped pair: straight-line arithmetic (so that two pairs of a lane interleave), then the decision
struct PdfDecision { int bin; bool alive, sure; };
template <bool NEAR>
static __device__ __forceinline__ PdfDecision pdf_decide(const PaPass &p, const float4 e, const PcTables &T, const PdfConsts &C, float near_a)
{
    const int na = T.na;
    const unsigned int tlim = 0x4B000000u + (unsigned)na;
    const float d2 = __fmaf_rn(e.z, e.z, __fmaf_rn(e.y, e.y, __fmul_rn(e.x, e.x)));
    float rt, rinv;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(rt) : "f"(d2));
    unsigned long long rr = pc_pack(rt, rt);
    if (NEAR) rr = pc_add2(rr, pc_pack(-near_a, near_a));
    const unsigned long long tt = pc_fma2_rm(rr, pc_pack(p.f32_lo, p.f32_hi), pc_pack(8388608.0f, 8388608.0f));
    const unsigned int tlo = __float_as_uint(pc_lo(tt)), thi = __float_as_uint(pc_hi(tt));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rinv) : "f"(rt));
    const float dot = __fmaf_rn(e.z, C.vz, __fmaf_rn(e.y, C.vy, __fmul_rn(e.x, C.vx))) * rinv;
    const float E = __fmaf_rn(C.e_b, rinv, C.e_a);
    float t_hi, t_lo;                                                         // thr_b[c] >= dot > thr_b[c+1] must hold with margin E
    const int c = pdf_class(C, dot, t_hi, t_lo);
    PdfDecision d;
    // beyond the cutoff: lower bracket above class na, or both brackets ON class na -- the job's cutoff sits within
    // delta_top of the boundary of bin na and the brackets are wider than that, so a pair of true class na ("inside the
    // cutoff, a-bin out of range") always has disagreeing brackets (same argument as VARIANT 0 of the RDF kernel)
    d.alive = !(tlo > tlim || (tlo == tlim && thi == tlim));
    // (|dot| + E <= 1: next to +-1 the reference's fp64 dot product can exceed 1 by an ulp -> ACOS = NaN -> out of bounds)
    d.sure = tlo == thi && tlo < tlim && c < C.nb && E < 1.0e-3f && (dot + E <= t_hi) && (dot - E > t_lo) && (fabsf(dot) + E <= 1.0f);
    d.bin = c * na + (int)(tlo - 0x4B000000u);
    return d;
}

template <bool VERIFY>
static __device__ __forceinline__ void pdf_commit(const PaPass &p, const PaCellGrid &go, const PaCellGrid &gi, const PaTile &tl, const float4 e,
                                                  const PdfDecision d, const PdfConsts &C, long long jbase,
                                                  unsigned long long &g_within, unsigned long long &g_oob)
{
    if (!d.alive) return;
    if (d.sure && !VERIFY) {
        if (C.hist) pc_red(C.hist + 4u * (unsigned)d.bin);
        else { atomicAdd(&p.job.acc[d.bin], (unsigned long long)p.weight); ++g_within; }
        return;
    }
    // literal evaluation from the original slots
    const unsigned int id = __float_as_uint(e.w);
    const long long iref = (long long)tl.frame * gi.nmol + tl.begin + (int)(id & 63u);
    const int io = gi.sorig[iref], jo = go.sorig[jbase + (long long)(id >> 6)];
    const bool self = p.same_type && io == jo;                                // the molecule itself (Distribution:778)
    const long long fb = (long long)(p.frame0 + tl.frame) * p.pts.M;
    int bin = -1;
    const int st = self ? 0 : pa_eval_pair_exact(p, fb + p.ref_off + io, fb + p.obs_off + jo, &bin);
    if (VERIFY && d.sure && !(st == 1 && bin == d.bin)) atomicAdd(p.verify_fail, 1u);
    if (st == 1) {
        if (C.hist) pc_red(C.hist + 4u * (unsigned)bin);
        else { atomicAdd(&p.job.acc[bin], (unsigned long long)p.weight); ++g_within; }
    } else if (st == 2) {
        if (C.hist) pc_red(C.hist + 4u * (unsigned)(p.job.bin_a * C.nb));
        else ++g_oob;
    }
}

template <bool NEAR, bool VERIFY>
static __device__ __forceinline__ void pc_segment_pdf(const PaPass &p, const PaCellGrid &go, const PaCellGrid &gi, const PaTile &tl,
                                                      const float (&xf)[PC_RI], const float (&yf)[PC_RI], const float (&zf)[PC_RI],
                                                      float hx, float hy, float hz, double ox, double oy, double oz,
                                                      long long jbase, int jb, int je, double sxa, double sya, double sza,
                                                      float4 *s_jf, int *s_jidx, const PcTables &T, const PdfConsts &C, float near_a,
                                                      float4 *stack, float dead_d2,
                                                      unsigned int &it_all, unsigned long long &g_within, unsigned long long &g_oob)
{
    int sn = 0;                                          // entries on the warp's stack (warp-uniform)
    const int lane = threadIdx.x & 31;
    const unsigned int lane_lt = (1u << lane) - 1u;
    const unsigned long long X01 = pc_pack(xf[0], xf[1]), Y01 = pc_pack(yf[0], yf[1]), Z01 = pc_pack(zf[0], zf[1]);
    const unsigned int kmask = (pc_kidx(0) < tl.count ? 0x5u : 0u) | (pc_kidx(1) < tl.count ? 0xAu : 0u);
    const unsigned int a_jf = (unsigned int)__cvta_generic_to_shared(s_jf);
    const double cxa = __dsub_rn(sxa, ox), cya = __dsub_rn(sya, oy), cza = __dsub_rn(sza, oz);
    int fill = 0;
    for (int jc = jb; jc < je; jc += 32) {
        {
            const int j = jc + lane;
            bool live = false;
            float4 f = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            if (j < je) {
                f.x = __double2float_rn(__dadd_rn(go.sx[jbase + j], cxa));
                f.y = __double2float_rn(__dadd_rn(go.sy[jbase + j], cya));
                f.z = __double2float_rn(__dadd_rn(go.sz[jbase + j], cza));
                const float g1 = fmaxf(__fsub_rd(fabsf(f.x), hx), 0.0f), g2 = fmaxf(__fsub_rd(fabsf(f.y), hy), 0.0f), g3 = fmaxf(__fsub_rd(fabsf(f.z), hz), 0.0f);
                live = !(__fmaf_rn(g3, g3, __fmaf_rn(g2, g2, __fmul_rn(g1, g1))) >= dead_d2);
            }
            const unsigned int m = __ballot_sync(0xffffffffu, live);
            if (live) { const int pos = fill + __popc(m & lane_lt); s_jf[pos] = f; s_jidx[pos] = j; }
            fill += __popc(m);
        }
        if (!(fill > PC_FBUF - 32 || (jc + 32 >= je && fill > 0))) continue;
        if ((fill & 1) && lane == 0) { s_jf[fill] = make_float4(T.pad_x, 0.0f, 0.0f, 0.0f); s_jidx[fill] = 0; }
        __syncwarp();
        it_all += (unsigned)((fill + 1) / 2);
        const unsigned int a_end = a_jf + (unsigned)fill * 16u;
        for (unsigned int aj = a_jf; aj < a_end; aj += 32u) {
            const int jj = (int)((aj - a_jf) >> 4);
            const unsigned int valid = kmask & (jj + 1 < fill ? 0xFu : 0x3u);  // real reference slots, not the padded observed slot
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const float4 w = pc_lds_f32x4(aj + 16u * u);
                const unsigned long long dx = pc_sub2(pc_pack(w.x, w.x), X01), dy = pc_sub2(pc_pack(w.y, w.y), Y01), dz = pc_sub2(pc_pack(w.z, w.z), Z01);
                const unsigned long long d2p = pc_fma2(dz, dz, pc_fma2(dy, dy, pc_mul2(dx, dx)));
                const unsigned int jid = (unsigned int)s_jidx[jj + u] << 6;
#pragma unroll
                for (int r = 0; r < PC_RI; ++r) {
                    const int q = u * PC_RI + r;
                    // d2~ >= dead_d2: the lower class bracket is beyond the cutoff, the pair can never be counted
                    const bool in = ((valid >> q) & 1u) && (r ? pc_hi(d2p) : pc_lo(d2p)) < dead_d2;
                    const unsigned int m = __ballot_sync(0xffffffffu, in);
                    if (in)
                        stack[sn + __popc(m & lane_lt)] = make_float4(r ? pc_hi(dx) : pc_lo(dx), r ? pc_hi(dy) : pc_lo(dy), r ? pc_hi(dz) : pc_lo(dz),
                                                                      __uint_as_float(jid | (unsigned int)pc_kidx(r)));
                    sn += __popc(m);
                }
            }
            while (sn >= 64) {                                                    // two entries per lane: their arithmetic interleaves
                __syncwarp();
                sn -= 64;
                const float4 e0 = stack[sn + lane], e1 = stack[sn + 32 + lane];
                __syncwarp();
                const PdfDecision d0 = pdf_decide<NEAR>(p, e0, T, C, near_a), d1 = pdf_decide<NEAR>(p, e1, T, C, near_a);
                pdf_commit<VERIFY>(p, go, gi, tl, e0, d0, C, jbase, g_within, g_oob);
                pdf_commit<VERIFY>(p, go, gi, tl, e1, d1, C, jbase, g_within, g_oob);
            }
        }
        fill = 0;
        __syncwarp();
    }
    // the brackets of the entries depend on this run (NEAR): empty the stack before the next run
    __syncwarp();
    for (int b = 0; b < sn; b += 32)
        if (b + lane < sn) {
            const float4 e = stack[b + lane];
            pdf_commit<VERIFY>(p, go, gi, tl, e, pdf_decide<NEAR>(p, e, T, C, near_a), C, jbase, g_within, g_oob);
        }
    __syncwarp();
}

// runs that are not float32-eligible: every pair in fp64 (two candidate images per component), pairs inside the cutoff
// go to the queue and are evaluated literally
template <bool VERIFY>
static __device__ __forceinline__ void pc_segment_pdf64(const PaPass &p, const PaCellGrid &go, const PaCellGrid &gi, const PaTile &tl,
                                                        const double (&p1x)[PC_RI], const double (&p1y)[PC_RI], const double (&p1z)[PC_RI],
                                                        long long jbase, int jb, int je, double sxa, double sxb, double sya, double syb, double sza, double szb,
                                                        unsigned long long *queue, unsigned int hist, unsigned int &it_all, unsigned long long &g_within, unsigned long long &g_oob)
{
    int qn = 0;
    const int lane = threadIdx.x & 31;
    const unsigned int lane_lt = (1u << lane) - 1u;
    const unsigned long long cut = p.job.cut_bits;
    for (int j = jb; j < je; j += 2) {                                           // two observed molecules per step (uniform loads)
        unsigned int need = 0u;
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            if (j + u >= je) continue;
            const double gx0 = go.sx[jbase + j + u], gy0 = go.sy[jbase + j + u], gz0 = go.sz[jbase + j + u];
#pragma unroll
            for (int r = 0; r < PC_RI; ++r) {
                if (pc_kidx(r) >= tl.count) continue;
                const double ex = pc_comp_sq<2>(__dadd_rn(gx0, sxa), __dadd_rn(gx0, sxb), p1x[r]);
                const double ey = pc_comp_sq<2>(__dadd_rn(gy0, sya), __dadd_rn(gy0, syb), p1y[r]);
                const double ez = pc_comp_sq<2>(__dadd_rn(gz0, sza), __dadd_rn(gz0, szb), p1z[r]);
                if (dbits(__dadd_rn(__dadd_rn(ex, ey), ez)) < cut) need |= 1u << (u * PC_RI + r);
            }
        }
        ++it_all;
        const unsigned int m = __ballot_sync(0xffffffffu, need != 0u);
        if (m == 0u) continue;
        const int add = __popc(m);
        if (qn + add > PC_QCAP) { pdf_drain<VERIFY>(p, go, gi, tl, queue, qn, jbase, hist, g_within, g_oob); qn = 0; }
        if (need)
            queue[qn + __popc(m & lane_lt)] = (unsigned long long)(unsigned int)j | ((unsigned long long)(unsigned int)(j + 1 < je ? j + 1 : j) << 24) |
                                              ((unsigned long long)lane << 48) | ((unsigned long long)need << 53);
        qn += add;
    }
    if (qn > 0) pdf_drain<VERIFY>(p, go, gi, tl, queue, qn, jbase, hist, g_within, g_oob);
}

// flush of the CTA's shared histogram by ONE warp while the others keep counting (atomic exchange), and at the end
static __device__ void pdf_flush(const PaPass &p, unsigned int hist, int nbins, int first, int stride)
{
    unsigned long long within = 0;
    for (int k = first; k < nbins + 1; k += stride) {
        const unsigned int c = pc_exch_u32(hist + 4u * (unsigned)k);
        if (!c) continue;
        if (k < nbins) { atomicAdd(&p.job.acc[k], (unsigned long long)((long long)c * p.weight)); within += c; }
        else atomicAdd(&p.job.acc[nbins + 1], (unsigned long long)c);
    }
    if (within) atomicAdd(&p.job.acc[nbins], within);
}

// SMEMH: one CTA of 16 warps per SM sharing a 32-bit histogram of all bins in shared memory (global 64-bit atomics of
// 148 x 16 warps into the few thousand words of a 2-D histogram run at ~1e11/s and were the limit of the first version);
// otherwise 4-warp CTAs and global atomics (histograms above ~40 000 bins).
template <bool VERIFY, bool SMEMH>
__global__ void __launch_bounds__(SMEMH ? 512 : PC_THREADS, SMEMH ? 1 : PDF_MINB) pa_pdf_cells_kernel(const __grid_constant__ PaPass p, const __grid_constant__ PaCellGrid gi, const __grid_constant__ PaCellGrid go, const PaTile *tiles,
                                                                     const unsigned int *ntiles_ptr, unsigned int tile_cap, unsigned int *work_counter, int nzg)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int na = p.job.bin_a, nb = p.job.bin_b;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nlay = go.g[0] + go.g[1] + go.g[2];
    float *s_tab = reinterpret_cast<float *>(smem_raw);                                   // nb + 2 floats
    const int nbins = na * nb;
    const size_t warp_bytes = pc_warp_bytes(true, nlay) + (size_t)PDF_STACK * 16;
    unsigned int *s_hist = reinterpret_cast<unsigned int *>(smem_raw + ((((size_t)(nb + 2) * 4) + 15) & ~(size_t)15));   // nbins + 1 (SMEMH)
    unsigned char *wbase = reinterpret_cast<unsigned char *>(s_hist) + (SMEMH ? ((((size_t)(nbins + 1) * 4) + 15) & ~(size_t)15) : 0) + (size_t)warp * warp_bytes;
    float4 *s_jf = reinterpret_cast<float4 *>(wbase);
    float4 *stack = s_jf + PC_FBUF;                                                        // PDF_STACK entries
    int *s_jidx = reinterpret_cast<int *>(stack + PDF_STACK);
    unsigned long long *queue = reinterpret_cast<unsigned long long *>(s_jidx + PC_FBUF);  // fp64 segments (pc_segment_pdf64)
    PcLayer *lay = reinterpret_cast<PcLayer *>(wbase + (size_t)PC_FBUF * 20 + (size_t)PDF_STACK * 16 + (size_t)PC_QCAP * 8);
    for (int k = threadIdx.x; k < nb + 2; k += blockDim.x)
        s_tab[k] = k == 0 ? __int_as_float(0x7f800000) : (k <= nb ? __double2float_rn(p.job.thr_b[k]) : __int_as_float(0xff800000));
    if (SMEMH) for (int k = threadIdx.x; k < nbins + 1; k += blockDim.x) s_hist[k] = 0u;
    __syncthreads();
    PcTables T{};
    T.na = na; T.cut = p.job.cut_bits;
    const float t2max = 32.0f;
    T.pad_x = p.job.maxdist * 1.001f + 3.0f * t2max;
    PdfConsts C;
    C.vx = __double2float_rn(p.job.ref_vec[0]); C.vy = __double2float_rn(p.job.ref_vec[1]); C.vz = __double2float_rn(p.job.ref_vec[2]);
    C.tab = (unsigned int)__cvta_generic_to_shared(s_tab);
    C.nb = nb; C.inv_step_b = 1.0f / p.job.step_b;
    C.hist = SMEMH ? (unsigned int)__cvta_generic_to_shared(s_hist) : 0u;
    C.e_a = 1.2e-6f;
    const int nwarps = (int)(blockDim.x >> 5);
    unsigned long long since_flush = 0;
    const double cut_d2 = __longlong_as_double((long long)T.cut) * (1.0 + 1e-12);
    const float cutf = (float)fmin(cut_d2 * (1.0 + 1e-6), 3.0e38);
    const unsigned int ntiles = min(*ntiles_ptr, tile_cap);
    if (*ntiles_ptr > tile_cap && threadIdx.x == 0 && blockIdx.x == 0) atomicOr(p.flags + 1, 1u);
    const long long n_items = (long long)ntiles * nzg;
    const int gx = go.g[0], gy = go.g[1], gz = go.g[2];
    unsigned int it_all = 0;
    unsigned long long st_all = 0, g_within = 0, g_oob = 0;
    for (;;) {
        st_all += it_all; it_all = 0;
        __syncwarp();
        unsigned int w = 0;
        if (lane == 0) w = atomicAdd(work_counter, 1u);
        w = __shfl_sync(0xffffffffu, w, 0);
        const long long item = (long long)w * p.tshard_count + p.tshard_rank;
        if (item >= n_items) break;
        const PaTile tl = tiles[item / nzg];
        const int zg = (int)(item % nzg);
        const int z0 = (int)((long long)gz * zg / nzg), z1 = (int)((long long)gz * (zg + 1) / nzg);
        if (z0 == z1) continue;
        double p1x[PC_RI], p1y[PC_RI], p1z[PC_RI], q1x[PC_RI], q1y[PC_RI], q1z[PC_RI];
        const long long ibase = (long long)tl.frame * gi.nmol;
#pragma unroll
        for (int r = 0; r < PC_RI; ++r) {
            const int k = pc_kidx(r), kk = min(k, tl.count - 1);
            q1x[r] = gi.sx[ibase + tl.begin + kk]; q1y[r] = gi.sy[ibase + tl.begin + kk]; q1z[r] = gi.sz[ibase + tl.begin + kk];
            p1x[r] = q1x[r]; p1y[r] = q1y[r]; p1z[r] = q1z[r];
        }
        double blox = fmin(q1x[0], q1x[1]), bloy = fmin(q1y[0], q1y[1]), bloz = fmin(q1z[0], q1z[1]);
        double bhix = fmax(q1x[0], q1x[1]), bhiy = fmax(q1y[0], q1y[1]), bhiz = fmax(q1z[0], q1z[1]);
        blox = pc_warp_min(blox); bloy = pc_warp_min(bloy); bloz = pc_warp_min(bloz);
        bhix = pc_warp_max(bhix); bhiy = pc_warp_max(bhiy); bhiz = pc_warp_max(bhiz);
        const double ox = 0.5 * (blox + bhix), oy = 0.5 * (bloy + bhiy), oz = 0.5 * (bloz + bhiz);
        float xf[PC_RI], yf[PC_RI], zf[PC_RI];
        float tile_t2 = 0.0f, hx = 0.0f, hy = 0.0f, hz = 0.0f;
#pragma unroll
        for (int r = 0; r < PC_RI; ++r) {
            xf[r] = __double2float_rn(__dsub_rn(q1x[r], ox)); yf[r] = __double2float_rn(__dsub_rn(q1y[r], oy)); zf[r] = __double2float_rn(__dsub_rn(q1z[r], oz));
            tile_t2 = fmaxf(tile_t2, __fmaf_ru(zf[r], zf[r], __fmaf_ru(yf[r], yf[r], __fmul_ru(xf[r], xf[r]))));
            hx = fmaxf(hx, fabsf(xf[r])); hy = fmaxf(hy, fabsf(yf[r])); hz = fmaxf(hz, fabsf(zf[r]));
        }
        tile_t2 = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(tile_t2)));
        hx = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(hx)));
        hy = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(hy)));
        hz = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(hz)));
        __syncwarp();
        for (int t = lane; t < nlay; t += 32) {
            const int c = t < gx ? 0 : (t < gx + gy ? 1 : 2);
            const int k = c == 0 ? t : (c == 1 ? t - gx : t - gx - gy);
            const long long lj = ((long long)tl.frame * 3 + c) * go.gmax;
            PcLayer cl;
            if (go.lay_lo[lj + k] > go.lay_hi[lj + k]) { cl.mode = 0; cl.ca = cl.cb = 1; cl.pad = 0; cl.dmin2 = 3.0e38f; }
            else cl = pc_classify(c == 0 ? blox : (c == 1 ? bloy : bloz), c == 0 ? bhix : (c == 1 ? bhiy : bhiz),
                                  okey_inv(go.lay_lo[lj + k]), okey_inv(go.lay_hi[lj + k]), c == 0 ? p.box.L[0] : (c == 1 ? p.box.L[1] : p.box.L[2]));
            if (cl.mode == 3) atomicOr(p.flags + 1, 1u);
            lay[t] = cl;
        }
        __syncwarp();
        if (SMEMH) {
            // 32-bit shared counters: the CTA's warps together add less than 2^31 between two flushes of any one of them
            const unsigned long long item_pairs = (unsigned long long)PC_TILE * (unsigned long long)go.nmol;
            if (since_flush + item_pairs > (0x7fffffffull / (unsigned)nwarps)) { pdf_flush(p, C.hist, nbins, lane, 32); since_flush = 0; }
            since_flush += item_pairs;
        }
        const bool tile_f32 = tile_t2 <= t2max * t2max;
        const float rmin2 = 4.0f * tile_t2 * 1.0001f + 1.0e-3f;
        const float near_a = 1.25e-7f * sqrtf(tile_t2) + 1.0e-12f;
        const float dead_r = __fadd_ru(p.dead_r, near_a);
        const float dead_d2 = __fmul_ru(__fmul_ru(dead_r, dead_r), 1.000001f);
        C.e_b = 2.6e-7f * __fsqrt_ru(tile_t2) + 1.0e-12f;
        const long long jbase = (long long)tl.frame * go.nmol;
        const int *cs = go.cell_start + (long long)tl.frame * (go.ncell + 1);
        for (int cz = z0; cz < z1; ++cz) {
            const PcLayer lz = lay[gx + gy + cz];
            if (!(lz.dmin2 < cutf)) continue;
            for (int cy = 0; cy < gy; ++cy) {
                const PcLayer ly = lay[gx + cy];
                const float dyz2 = lz.dmin2 + ly.dmin2;
                if (!(dyz2 < cutf)) continue;
                const int *row = cs + ((long long)cz * gy + cy) * gx;
                int cx = 0;
                while (cx < gx) {
                    const PcLayer lx = lay[cx];
                    if (!(dyz2 + lx.dmin2 < cutf)) { ++cx; continue; }
                    const bool nearr = !(dyz2 + lx.dmin2 >= rmin2);
                    int ce = cx;
                    while (ce + 1 < gx) {
                        const PcLayer ln = lay[ce + 1];
                        if (ln.mode != lx.mode || ln.ca != lx.ca || ln.cb != lx.cb || !(dyz2 + ln.dmin2 < cutf)) break;
                        if (tile_f32 && (!(dyz2 + ln.dmin2 >= rmin2) != nearr)) break;
                        ++ce;
                    }
                    const int jb = row[cx], je = row[ce + 1];
                    cx = ce + 1;
                    if (jb >= je) continue;
                    const double Lx = p.box.L[0], Ly = p.box.L[1], Lz = p.box.L[2];
                    const double sxa = (double)(lx.ca - 1) * Lx, sxb = (double)(lx.cb - 1) * Lx;
                    const double sya = (double)(ly.ca - 1) * Ly, syb = (double)(ly.cb - 1) * Ly;
                    const double sza = (double)(lz.ca - 1) * Lz, szb = (double)(lz.cb - 1) * Lz;
                    const bool single = lx.mode <= 1 && ly.mode <= 1 && lz.mode <= 1;
                    if (tile_f32 && single) {
                        if (nearr) pc_segment_pdf<true, VERIFY>(p, go, gi, tl, xf, yf, zf, hx, hy, hz, ox, oy, oz, jbase, jb, je, sxa, sya, sza, s_jf, s_jidx, T, C, near_a, stack, dead_d2, it_all, g_within, g_oob);
                        else pc_segment_pdf<false, VERIFY>(p, go, gi, tl, xf, yf, zf, hx, hy, hz, ox, oy, oz, jbase, jb, je, sxa, sya, sza, s_jf, s_jidx, T, C, near_a, stack, dead_d2, it_all, g_within, g_oob);
                    } else {
                        pc_segment_pdf64<VERIFY>(p, go, gi, tl, p1x, p1y, p1z, jbase, jb, je, sxa, sxb, sya, syb, sza, szb, queue, C.hist, it_all, g_within, g_oob);
                    }
                }
            }
        }
    }
    if (g_within) atomicAdd(&p.job.acc[nbins], g_within);
    if (g_oob) atomicAdd(&p.job.acc[nbins + 1], g_oob);
    if (SMEMH) {
        __syncthreads();
        pdf_flush(p, C.hist, nbins, (int)threadIdx.x, (int)blockDim.x);
    }
    if (p.kstats && lane == 0) { st_all += it_all; if (st_all) atomicAdd(p.kstats + 4, st_all * 128ull); }
}

// ---------------------------------------------------------------------------
// Cell-sorted GENERAL kernel (2-D pdf, cdf, centre of charge, any reference vector).  Same tiles,
// same per-run image classification and cutoff pruning as above; a pair inside the cutoff gets the
// literal evaluation (pa_eval_pair_exact) through its original slots.  With cutoffs well below L/2
// (the usual 2-D case) almost all cell runs are skipped, which is where the speed-up comes from.
// Pairs inside the cutoff are not evaluated where they are found (a quarter of the lanes at a time) but appended
// to a queue of their warp -- entry = (reference index inside the tile, observed index inside the staged chunk) --
// and drained with one entry per lane.  An entry whose two molecules were not moved by wrap_vector, in a run with
// one possible image per component, needs no global data: translation + wrapshift is the run's image shift, so
// the reference's link vector (translation + com_obs) - com_ref (Distribution:786-792) is bit for bit the
// difference fl(p2+s) - p1 already staged, and the bin follows from pa_bin_from_link.  Everything else (centre of
// charge, wrapped molecules, two candidate images, m >= maximum_distance_squared) takes pa_eval_pair_exact.
#define PG_QCAP 64
template <int MODE, bool SMEM_HIST>
static __device__ __noinline__ void pg_drain(const PaPass &p, const PaCellGrid &go, const PaCellGrid &gi, const PaTile &tl,
                                             const unsigned int *queue, int count, const double *s_j, const int *s_jo,
                                             long long jbase, int jc, unsigned int *s_hist, int nbins,
                                             const unsigned long long *thr_a, const double *thr_b,
                                             unsigned long long &g_within, unsigned long long &g_oob)
{
    __syncwarp();
    const int lane = threadIdx.x & 31;
    const long long fb = (long long)(p.frame0 + tl.frame) * p.pts.M;
    for (int base = 0; base < count; base += 32) {
        if (base + lane >= count) continue;
        const unsigned int ent = queue[base + lane];
        const int k = (int)(ent & 0xffffu), jl = (int)(ent >> 16);
        const long long iref = (long long)tl.frame * gi.nmol + tl.begin + k;
        const int io = gi.sorig[iref], jo = s_jo[jl];
        int bin = 0, st;
        bool fast = false;
        if (MODE == 1 && !p.job.use_coc && gi.sflag[iref] == 0 && go.sflag[jbase + jc + jl] == 0) {
            const double *rec = s_j + jl * 6;
            const double link[3] = { __dsub_rn(rec[0], gi.sx[iref]), __dsub_rn(rec[2], gi.sy[iref]), __dsub_rn(rec[4], gi.sz[iref]) };
            const double m = __dadd_rn(__dadd_rn(__dmul_rn(link[0], link[0]), __dmul_rn(link[1], link[1])), __dmul_rn(link[2], link[2]));
            if (m < p.box.max_dist_sq) { st = pa_bin_from_link(p, m, link, &bin, thr_a, thr_b); fast = true; }
        }
        if (!fast) st = pa_eval_pair_exact(p, fb + p.ref_off + io, fb + p.obs_off + jo, &bin);
        if (SMEM_HIST) {
            if (st == 1) atomicAdd(&s_hist[bin], 1u);
            else if (st == 2) atomicAdd(&s_hist[nbins], 1u);
        } else {
            if (st == 1) { atomicAdd(&p.job.acc[bin], (unsigned long long)p.weight); ++g_within; }
            else if (st == 2) ++g_oob;
        }
    }
    __syncwarp();
}

template <int MODE, bool SMEM_HIST>
static __device__ __forceinline__ void pg_segment(const PaPass &p, const PaCellGrid &go, const PaCellGrid &gi, const PaTile &tl,
                                                  const double (&p1x)[PG_RI], const double (&p1y)[PG_RI], const double (&p1z)[PG_RI],
                                                  const int (&iorig)[PG_RI], long long jbase, int jb, int je,
                                                  double sxa, double sxb, double sya, double syb, double sza, double szb,
                                                  double *s_j, int *s_jo, unsigned int *s_hist, int nbins, unsigned int *queue,
                                                  const unsigned long long *thr_a, const double *thr_b, int &parity,
                                                  unsigned long long &g_within, unsigned long long &g_oob)
{
    // The staging area is double buffered (s_j0 / s_jo0 hold two chunks): chunk g goes to buffer g & 1 and there is ONE
    // barrier per chunk, after the staging.  A warp restages a buffer only after the barrier of the chunk in between,
    // which every warp passes only after it has finished (evaluated and drained) the chunk that used the buffer before;
    // so warps may be one chunk apart and a long drain in one warp no longer stalls the other three.
    double *const s_j0 = s_j;
    int *const s_jo0 = s_jo;
    const unsigned long long cut = p.job.cut_bits;
    const unsigned int lane_lt = (1u << (threadIdx.x & 31)) - 1u;
    int qn = 0;                                          // entries in this warp's queue (warp-uniform)
    for (int jc = jb; jc < je; jc += PG_JC) {
        s_j = s_j0 + (parity & 1) * (PG_JC * 6);
        s_jo = s_jo0 + (parity & 1) * PG_JC;
        ++parity;
        {
            const int t = threadIdx.x;
            const int j = jc + t;
            double x = 1.0e15, y = 1.0e15, z = 1.0e15;
            int o = -1;
            if (j < je) { x = go.sx[jbase + j]; y = go.sy[jbase + j]; z = go.sz[jbase + j]; o = go.sorig[jbase + j]; }
            double *rec = s_j + t * 6;
            rec[0] = __dadd_rn(x, sxa); rec[1] = __dadd_rn(x, sxb);
            rec[2] = __dadd_rn(y, sya); rec[3] = __dadd_rn(y, syb);
            rec[4] = __dadd_rn(z, sza); rec[5] = __dadd_rn(z, szb);
            s_jo[t] = o;
        }
        __syncthreads();
        const int cnt = min(PG_JC, je - jc);
        for (int jj = 0; jj < cnt; ++jj) {
            const double2 *rec = reinterpret_cast<const double2 *>(s_j + jj * 6);
            const double2 rx = rec[0], ry = rec[1], rz = rec[2];
            const int jo = s_jo[jj];
#pragma unroll
            for (int r = 0; r < PG_RI; ++r) {
                const double sx = pc_comp_sq<MODE>(rx.x, rx.y, p1x[r]);
                const double sy = pc_comp_sq<MODE>(ry.x, ry.y, p1y[r]);
                const double sz = pc_comp_sq<MODE>(rz.x, rz.y, p1z[r]);
                const double d2 = __dadd_rn(__dadd_rn(sx, sy), sz);
                // inside the cutoff, not a padding lane, not the molecule itself (Distribution:778)
                const bool need = dbits(d2) < cut && iorig[r] >= 0 && !(p.same_type && iorig[r] == jo);
                const unsigned int m = __ballot_sync(0xffffffffu, need);
                if (m == 0u) continue;
                const int add = __popc(m);
                if (qn + add > PG_QCAP) { pg_drain<MODE, SMEM_HIST>(p, go, gi, tl, queue, qn, s_j, s_jo, jbase, jc, s_hist, nbins, thr_a, thr_b, g_within, g_oob); qn = 0; }
                if (need) queue[qn + __popc(m & lane_lt)] = ((unsigned int)jj << 16) | (unsigned int)(r * PG_THREADS + threadIdx.x);
                qn += add;
            }
        }
        if (qn > 0) { pg_drain<MODE, SMEM_HIST>(p, go, gi, tl, queue, qn, s_j, s_jo, jbase, jc, s_hist, nbins, thr_a, thr_b, g_within, g_oob); qn = 0; }   // before the chunk is overwritten
    }
}

template <bool SMEM_HIST>
static __device__ void pg_flush(const PaPass &p, unsigned int *s_hist, int nbins, bool reset)
{
    unsigned long long within = 0;
    for (int k = threadIdx.x; k < nbins + 1; k += blockDim.x) {
        const unsigned int c = s_hist[k];
        if (c) {
            if (k < nbins) { atomicAdd(&p.job.acc[k], (unsigned long long)((long long)c * p.weight)); within += c; }
            else atomicAdd(&p.job.acc[nbins + 1], (unsigned long long)c);
            if (reset) s_hist[k] = 0u;
        }
    }
    if (within) atomicAdd(&p.job.acc[nbins], within);
}

template <bool SMEM_HIST>
__global__ void __launch_bounds__(PG_THREADS, 5) pa_general_cells_kernel(const __grid_constant__ PaPass p, PaCellGrid gi, PaCellGrid go, const PaTile *tiles,
                                                                       const unsigned int *ntiles_ptr, unsigned int tile_cap,
                                                                       unsigned int *work_counter, int nzg)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int nbins = p.job.bin_a * p.job.bin_b;
    double *s_j = reinterpret_cast<double *>(smem_raw);                         // 2 x PG_JC*6 doubles (double buffered)
    int *s_jo = reinterpret_cast<int *>(s_j + 2 * PG_JC * 6);                   // 2 x PG_JC original indices
    PcShared *sh = reinterpret_cast<PcShared *>(s_jo + 2 * PG_JC);
    int parity = 0;
    unsigned int *queue = reinterpret_cast<unsigned int *>(sh + 1) + (threadIdx.x >> 5) * PG_QCAP;   // per warp
    // class tables of the job in shared memory (pdf: thr_a[bin_a+3], thr_b[bin_b+1]; <= 16 KB), searched by every drained pair
    unsigned long long *s_thr_a = reinterpret_cast<unsigned long long *>(
        (reinterpret_cast<uintptr_t>(reinterpret_cast<unsigned int *>(sh + 1) + (PG_THREADS / 32) * PG_QCAP) + 7) & ~uintptr_t(7));
    const int n_thr_a = p.job.mode == PA_MODE_PDF ? p.job.bin_a + 3 : 0, n_thr_b = p.job.mode == PA_MODE_PDF ? p.job.bin_b + 1 : 0;
    double *s_thr_b = reinterpret_cast<double *>(s_thr_a + n_thr_a);
    for (int k = threadIdx.x; k < n_thr_a; k += blockDim.x) s_thr_a[k] = p.job.thr_a[k];
    for (int k = threadIdx.x; k < n_thr_b; k += blockDim.x) s_thr_b[k] = p.job.thr_b[k];
    const unsigned long long *thr_a = n_thr_a ? s_thr_a : p.job.thr_a;
    const double *thr_b = n_thr_b ? s_thr_b : p.job.thr_b;
    unsigned int *s_hist = reinterpret_cast<unsigned int *>(s_thr_b + n_thr_b);   // nbins+1 (SMEM_HIST)
    if (SMEM_HIST) for (int k = threadIdx.x; k < nbins + 1; k += blockDim.x) s_hist[k] = 0u;
    __syncthreads();
    const double cut_d2 = __longlong_as_double((long long)p.job.cut_bits) * (1.0 + 1e-12);
    const float cutf = (float)fmin(cut_d2 * (1.0 + 1e-6), 3.0e38);
    const unsigned int ntiles = min(*ntiles_ptr, tile_cap);
    if (*ntiles_ptr > tile_cap && threadIdx.x == 0 && blockIdx.x == 0) atomicOr(p.flags + 1, 1u);
    const long long n_items = (long long)ntiles * nzg;
    const int gx = go.g[0], gy = go.g[1], gz = go.g[2];
    unsigned long long g_within = 0, g_oob = 0, since_flush = 0;

    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) { sh->work_item = (int)atomicAdd(work_counter, 1u); }
        __syncthreads();
        const long long item = (long long)sh->work_item * p.tshard_count + p.tshard_rank;   // tile shard: items w with w % count == rank
        if (item >= n_items) break;
        const PaTile tl = tiles[item / nzg];
        const int zg = (int)(item % nzg);
        const int z0 = (int)((long long)gz * zg / nzg), z1 = (int)((long long)gz * (zg + 1) / nzg);
        if (z0 == z1) continue;
        double p1x[PG_RI], p1y[PG_RI], p1z[PG_RI];
        int iorig[PG_RI];
        const long long ibase = (long long)tl.frame * gi.nmol;
#pragma unroll
        for (int r = 0; r < PG_RI; ++r) {
            const int k = r * PG_THREADS + threadIdx.x;
            if (k < tl.count) {
                p1x[r] = gi.sx[ibase + tl.begin + k]; p1y[r] = gi.sy[ibase + tl.begin + k]; p1z[r] = gi.sz[ibase + tl.begin + k];
                iorig[r] = gi.sorig[ibase + tl.begin + k];
            } else { p1x[r] = p1y[r] = p1z[r] = -1.0e15; iorig[r] = -1; }
        }
        for (int t = threadIdx.x; t < gx + gy + gz; t += blockDim.x) {
            const int c = t < gx ? 0 : (t < gx + gy ? 1 : 2);
            const int k = c == 0 ? t : (c == 1 ? t - gx : t - gx - gy);
            const long long li = ((long long)tl.frame * 3 + c) * gi.gmax, lj = ((long long)tl.frame * 3 + c) * go.gmax;
            double ilo, ihi;
            if (c == 0) {
                ilo = 1e300; ihi = -1e300;
                for (int q = tl.cx0; q <= tl.cx1; ++q)
                    if (gi.lay_lo[li + q] <= gi.lay_hi[li + q]) { ilo = fmin(ilo, okey_inv(gi.lay_lo[li + q])); ihi = fmax(ihi, okey_inv(gi.lay_hi[li + q])); }
            } else {
                const int q = c == 1 ? tl.cy : tl.cz;
                ilo = okey_inv(gi.lay_lo[li + q]); ihi = okey_inv(gi.lay_hi[li + q]);
            }
            PcLayer cl;
            if (go.lay_lo[lj + k] > go.lay_hi[lj + k]) { cl.mode = 0; cl.ca = cl.cb = 1; cl.pad = 0; cl.dmin2 = 3.0e38f; }
            else cl = pc_classify(ilo, ihi, okey_inv(go.lay_lo[lj + k]), okey_inv(go.lay_hi[lj + k]), p.box.L[c]);
            if (cl.mode == 3) atomicOr(p.flags + 1, 1u);
            sh->lay[c][k] = cl;
        }
        __syncthreads();
        if (SMEM_HIST) {
            const unsigned long long item_pairs = (unsigned long long)PC_GTILE * (unsigned long long)go.nmol;
            if (since_flush + item_pairs > 0x7fffffffull) { __syncthreads(); pg_flush<SMEM_HIST>(p, s_hist, nbins, true); since_flush = 0; }
            since_flush += item_pairs;
        }
        const long long jbase = (long long)tl.frame * go.nmol;
        const int *cs = go.cell_start + (long long)tl.frame * (go.ncell + 1);
        for (int cz = z0; cz < z1; ++cz) {
            const PcLayer lz = sh->lay[2][cz];
            if (!(lz.dmin2 < cutf)) continue;
            for (int cy = 0; cy < gy; ++cy) {
                const PcLayer ly = sh->lay[1][cy];
                const float dyz2 = lz.dmin2 + ly.dmin2;
                if (!(dyz2 < cutf)) continue;
                const int *row = cs + ((long long)cz * gy + cy) * gx;
                int cx = 0;
                while (cx < gx) {
                    const PcLayer lx = sh->lay[0][cx];
                    if (!(dyz2 + lx.dmin2 < cutf)) { ++cx; continue; }
                    int ce = cx;
                    while (ce + 1 < gx) {
                        const PcLayer ln = sh->lay[0][ce + 1];
                        if (ln.mode != lx.mode || ln.ca != lx.ca || ln.cb != lx.cb || !(dyz2 + ln.dmin2 < cutf)) break;
                        ++ce;
                    }
                    const int jb = row[cx], je = row[ce + 1];
                    cx = ce + 1;
                    if (jb == je) continue;
                    const double Lx = p.box.L[0], Ly = p.box.L[1], Lz = p.box.L[2];
                    const double sxa = (double)(lx.ca - 1) * Lx, sxb = (double)(lx.cb - 1) * Lx;
                    const double sya = (double)(ly.ca - 1) * Ly, syb = (double)(ly.cb - 1) * Ly;
                    const double sza = (double)(lz.ca - 1) * Lz, szb = (double)(lz.cb - 1) * Lz;
                    if (lx.mode < 2 && ly.mode < 2 && lz.mode < 2)
                        pg_segment<1, SMEM_HIST>(p, go, gi, tl, p1x, p1y, p1z, iorig, jbase, jb, je, sxa, sxb, sya, syb, sza, szb, s_j, s_jo, s_hist, nbins, queue, thr_a, thr_b, parity, g_within, g_oob);
                    else
                        pg_segment<2, SMEM_HIST>(p, go, gi, tl, p1x, p1y, p1z, iorig, jbase, jb, je, sxa, sxb, sya, syb, sza, szb, s_j, s_jo, s_hist, nbins, queue, thr_a, thr_b, parity, g_within, g_oob);
                }
            }
        }
    }
    __syncthreads();
    if (SMEM_HIST) pg_flush<SMEM_HIST>(p, s_hist, nbins, false);
    else {
        if (g_within) atomicAdd(&p.job.acc[nbins], g_within);
        if (g_oob) atomicAdd(&p.job.acc[nbins + 1], g_oob);
    }
}

// ---------------------------------------------------------------------------
// launch wrappers

int pa_launch_cell_sort(const PaCellGrid &g, const PaPoints &pts, const PaBox &box, int frame0, int nframes, unsigned int *err, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    const long long total = (long long)nframes * g.nmol;
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    cudaMemsetAsync(g.cell_start, 0, sizeof(int) * (size_t)nframes * (g.ncell + 1), st);
    cudaMemsetAsync(g.cell_fill, 0, sizeof(int) * (size_t)nframes * g.ncell, st);
    cudaMemsetAsync(g.lay_lo, 0xff, sizeof(unsigned long long) * (size_t)nframes * 3 * g.gmax, st);
    cudaMemsetAsync(g.lay_hi, 0x00, sizeof(unsigned long long) * (size_t)nframes * 3 * g.gmax, st);
    pa_cell_count_kernel<<<blocks, 256, 0, st>>>(g, pts, box, frame0, nframes, err);
    pa_cell_scan_kernel<<<nframes, 1024, 0, st>>>(g);
    pa_cell_scatter_kernel<<<blocks, 256, 0, st>>>(g, nframes);
    const long long order_threads = (long long)nframes * g.ncell * 32;
    pa_cell_order_kernel<<<(int)std::min<long long>((order_threads + 255) / 256, 148 * 16), 256, 0, st>>>(g, nframes, err);
    pa_cell_gather_kernel<<<blocks, 256, 0, st>>>(g, pts, box, frame0, nframes);
    return (int)cudaGetLastError();
}

int pa_launch_build_tiles(const PaCellGrid &g, int nframes, bool chunks, PaTile *tiles, unsigned int *ntiles, int *row_off, unsigned int cap, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    const int rows = nframes * g.g[1] * g.g[2];
    // tile width (max_run cells) + one observed cell must stay below L/2: (max_run + 1.1) / gx < 0.5
    int max_run = (int)(0.45 * g.g[0]) - 1;
    if (max_run < 1) max_run = 1;
    const int blocks = (rows + 127) / 128;
    if (chunks) {
        pa_build_tiles_kernel<true, false><<<blocks, 128, 0, st>>>(g, nframes, max_run, tiles, row_off, cap);
        pa_tile_scan_kernel<<<1, 1024, 0, st>>>(row_off, rows, ntiles);
        pa_build_tiles_kernel<true, true><<<blocks, 128, 0, st>>>(g, nframes, max_run, tiles, row_off, cap);
    } else {
        pa_build_tiles_kernel<false, false><<<blocks, 128, 0, st>>>(g, nframes, max_run, tiles, row_off, cap);
        pa_tile_scan_kernel<<<1, 1024, 0, st>>>(row_off, rows, ntiles);
        pa_build_tiles_kernel<false, true><<<blocks, 128, 0, st>>>(g, nframes, max_run, tiles, row_off, cap);
    }
    return (int)cudaGetLastError();
}

void pa_cell_grid_dims(int nmol, int g[3])
{
    // ~100 molecules per (y,z)-isotropic cell, split 4x finer along x: tiles are runs of x-cells
    int G = (int)cbrt((double)nmol / 100.0);
    if (G < 5) G = 5;
    if (G > 64) G = 64;
    g[1] = g[2] = G;
    g[0] = 4 * G;
    if (g[0] > PC_MAXG) g[0] = PC_MAXG;
}

size_t pa_cells_smem_bytes(int na, int f32_cnt, int nlayers)   // f32_cnt > 0: the PART 1 launch
{
    const int ncnt = (f32_cnt > 0 ? f32_cnt + na + 8 : na + 3) + 1;
    const size_t cta = ((f32_cnt > 0 ? 0 : (size_t)(na + 3) * 8) + (size_t)(ncnt + 1) * 4 + 15) & ~(size_t)15;
    return cta + (size_t)PC_WARPS * pc_warp_bytes(f32_cnt > 0, nlayers) + 16;
}

int pa_launch_rdf_cells(const PaPass &p, const PaCellGrid &gi, const PaCellGrid &go, const PaTile *tiles, const unsigned int *ntiles,
                        unsigned int tile_cap, unsigned int *work_counter, int nzg, int variant, int sm_count, void *stream,
                        void *ev_f32_begin, void *ev_f32_end)
{
    cudaStream_t st = (cudaStream_t)stream;
    const size_t smem = pa_cells_smem_bytes(p.job.bin_a, variant == 0 ? p.f32_cnt : 0, go.g[0] + go.g[1] + go.g[2]);
    const size_t smem0 = pa_cells_smem_bytes(p.job.bin_a, 0, go.g[0] + go.g[1] + go.g[2]);
    const bool verify = p.verify_fail != nullptr;
#define PC_LAUNCH3(O_, V_, S_, P_, SMEM_)                                                                        \
    do {                                                                                                        \
        int occ = 1;                                                                                            \
        cudaMemsetAsync(work_counter, 0, sizeof(unsigned int), st);                                             \
        cudaFuncSetAttribute(pa_rdf_cells_kernel<O_, V_, S_, P_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(SMEM_)); \
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, pa_rdf_cells_kernel<O_, V_, S_, P_>, PC_THREADS, (SMEM_)) != cudaSuccess || occ < 1) occ = 1; \
        pa_rdf_cells_kernel<O_, V_, S_, P_><<<sm_count * occ, PC_THREADS, (SMEM_), st>>>(p, gi, go, tiles, ntiles, tile_cap, work_counter, nzg); \
    } while (0)
#define PC_LAUNCH2(O_, V_, S_)                                                                                  \
    do {                                                                                                        \
        if (O_ == 0 && p.f32_cnt > 0) {                                                                         \
            if (ev_f32_begin) cudaEventRecord((cudaEvent_t)ev_f32_begin, st);                                   \
            PC_LAUNCH3(O_, V_, S_, 1, smem);                                                                    \
            if (ev_f32_end) cudaEventRecord((cudaEvent_t)ev_f32_end, st);                                       \
            PC_LAUNCH3(O_, V_, S_, 2, smem0);                                                                   \
        }                                                                                                       \
        else PC_LAUNCH3(O_, V_, S_, 0, smem0);                                                                  \
    } while (0)
#define PC_LAUNCH(O_, V_) do { if (p.sym) PC_LAUNCH2(O_, V_, true); else PC_LAUNCH2(O_, V_, false); } while (0)
    if (variant == 1) { if (verify) PC_LAUNCH(1, true); else PC_LAUNCH(1, false); }
    else { if (verify) PC_LAUNCH(0, true); else PC_LAUNCH(0, false); }
#undef PC_LAUNCH3
#undef PC_LAUNCH2
#undef PC_LAUNCH
    return (int)cudaGetLastError();
}

int pa_launch_pdf_cells(const PaPass &p, const PaCellGrid &gi, const PaCellGrid &go, const PaTile *tiles, const unsigned int *ntiles,
                        unsigned int tile_cap, unsigned int *work_counter, int nzg, int sm_count, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    const int nbins = p.job.bin_a * p.job.bin_b;
    const size_t warp_bytes = pc_warp_bytes(true, go.g[0] + go.g[1] + go.g[2]) + (size_t)PDF_STACK * 16;
    const size_t tab = (((size_t)(p.job.bin_b + 2) * 4) + 15) & ~(size_t)15;
    const size_t smem_h = tab + ((((size_t)(nbins + 1) * 4) + 15) & ~(size_t)15) + 16 * warp_bytes + 16;     // 16 warps, shared histogram
    const size_t smem_g = tab + (size_t)PC_WARPS * warp_bytes + 16;
    const bool smemh = smem_h <= 200 * 1024;
    cudaMemsetAsync(work_counter, 0, sizeof(unsigned int), st);
#define PDF_LAUNCH(V_, H_, THREADS_, SMEM_)                                                                     \
    do {                                                                                                        \
        int occ = 1;                                                                                            \
        cudaFuncSetAttribute(pa_pdf_cells_kernel<V_, H_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(SMEM_)); \
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, pa_pdf_cells_kernel<V_, H_>, (THREADS_), (SMEM_)) != cudaSuccess || occ < 1) occ = 1; \
        pa_pdf_cells_kernel<V_, H_><<<sm_count * occ, (THREADS_), (SMEM_), st>>>(p, gi, go, tiles, ntiles, tile_cap, work_counter, nzg); \
    } while (0)
    if (smemh) { if (p.verify_fail) PDF_LAUNCH(true, true, 512, smem_h); else PDF_LAUNCH(false, true, 512, smem_h); }
    else { if (p.verify_fail) PDF_LAUNCH(true, false, PC_THREADS, smem_g); else PDF_LAUNCH(false, false, PC_THREADS, smem_g); }
#undef PDF_LAUNCH
    return (int)cudaGetLastError();
}

int pa_launch_general_cells(const PaPass &p, const PaCellGrid &gi, const PaCellGrid &go, const PaTile *tiles, const unsigned int *ntiles,
                            unsigned int tile_cap, unsigned int *work_counter, int nzg, int sm_count, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    const int nbins = p.job.bin_a * p.job.bin_b;
    // a shared-memory histogram only while it leaves room for >= 6 CTAs per SM (occupancy matters more to this
    // latency-bound kernel than the atomics: larger histograms take 64-bit atomics in L2)
    const bool smem_hist = nbins <= 6000;
    const size_t smem = 2 * ((size_t)PG_JC * 6 * 8 + (size_t)PG_JC * 4) + sizeof(PcShared) + (size_t)(PG_THREADS / 32) * PG_QCAP * 4 +
                        (p.job.mode == PA_MODE_PDF ? (size_t)(p.job.bin_a + 3 + p.job.bin_b + 1) * 8 : 0) +
                        (smem_hist ? (size_t)(nbins + 1) * 4 : 0) + 24;
    cudaMemsetAsync(work_counter, 0, sizeof(unsigned int), st);
#define PG_LAUNCH(S_)                                                                                           \
    do {                                                                                                        \
        int occ = 1;                                                                                            \
        cudaFuncSetAttribute(pa_general_cells_kernel<S_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, pa_general_cells_kernel<S_>, PG_THREADS, smem) != cudaSuccess || occ < 1) occ = 1; \
        pa_general_cells_kernel<S_><<<sm_count * occ, PG_THREADS, smem, st>>>(p, gi, go, tiles, ntiles, tile_cap, work_counter, nzg); \
    } while (0)
    if (smem_hist) PG_LAUNCH(true); else PG_LAUNCH(false);
#undef PG_LAUNCH
    return (int)cudaGetLastError();
}
