// pa_microbench2.cu -- float32 / packed float32 (f32x2) instruction rates on sm_100a, for the float32
// segments of the cell-sorted RDF kernel (pa_cells.cu): does FFMA2 halve the issue slots of an FMA-heavy
// loop, and at what rate do FFMA2.RM / FADD2 / MUFU / integer instructions co-issue with it?
// Build: make -C prealpha_b200/csrc ../pa_microbench2 ; run on a B200: prealpha_b200/pa_microbench2
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)
constexpr int ITERS = 4096;
typedef unsigned long long u64;

__device__ __forceinline__ u64 pk(float a, float b) { return ((u64)__float_as_uint(b) << 32) | __float_as_uint(a); }

__global__ void k_ffma(float *out, float a, float b)
{
    float x[16];
    for (int i = 0; i < 16; ++i) x[i] = a + threadIdx.x * 1e-6f + i;
    for (int it = 0; it < ITERS; ++it)
#pragma unroll
        for (int i = 0; i < 16; ++i) x[i] = __fmaf_rn(x[i], b, a);
    float s = 0; for (int i = 0; i < 16; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_ffma_rm(float *out, float a, float b)
{
    float x[16];
    for (int i = 0; i < 16; ++i) x[i] = a + threadIdx.x * 1e-6f + i;
    for (int it = 0; it < ITERS; ++it)
#pragma unroll
        for (int i = 0; i < 16; ++i) x[i] = __fmaf_rd(x[i], b, a);
    float s = 0; for (int i = 0; i < 16; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE>   // 0: fma.rn.f32x2, 1: fma.rm.f32x2, 2: add.rn.f32x2, 3: fma2 + 1 LOP3 each, 4: fma2 + 2 LOP3 each, 5: fma2 + 1 IMAD each
__global__ void k_f32x2(float *out, float a, float b)
{
    u64 x[16];
    unsigned int z[16];
    const u64 vb = pk(b, b), va = pk(a, a);
    for (int i = 0; i < 16; ++i) { x[i] = pk(a + threadIdx.x * 1e-6f + i, a + i); z[i] = threadIdx.x + i; }
    for (int it = 0; it < ITERS; ++it)
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (MODE == 1) asm volatile("fma.rm.f32x2 %0, %0, %1, %2;" : "+l"(x[i]) : "l"(vb), "l"(va));
            else if (MODE == 2) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(x[i]) : "l"(va));
            else asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x[i]) : "l"(vb), "l"(va));
            if (MODE == 3 || MODE == 4) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(z[i]) : "r"(it), "r"(z[(i + 1) & 15]));
            if (MODE == 4) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(z[(i + 5) & 15]) : "r"(it), "r"(z[(i + 7) & 15]));
            if (MODE == 5) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(z[i]) : "r"(it | 1), "r"(z[(i + 1) & 15]));
        }
    float s = 0;
    for (int i = 0; i < 16; ++i) s += __uint_as_float((unsigned)x[i]) + __uint_as_float((unsigned)(x[i] >> 32)) + (float)z[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// the inner loop's mix per 4 pairs, scalar: 12 FADD, 4 FMUL, 8 FFMA, 4 MUFU, 8 FFMA.RM, 4 IMAD, 4 ISETP-ish LOP3
__global__ void k_mix_scalar(float *out, float a, float b)
{
    float x[4], y[4], zc[4]; unsigned int acc = 0;
    for (int i = 0; i < 4; ++i) { x[i] = a + threadIdx.x * 1e-3f + i; y[i] = b + i; zc[i] = a - i; }
    for (int it = 0; it < ITERS; ++it) {
        const float w = __int_as_float(0x3f800000 | (it & 0xff));
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float dx = __fsub_rn(w, x[i]), dy = __fsub_rn(w, y[i]), dz = __fsub_rn(w, zc[i]);
            const float d2 = __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
            float rt; asm volatile("sqrt.approx.ftz.f32 %0, %1;" : "=f"(rt) : "f"(d2));
            const float t0 = __fmaf_rd(rt, a, 8388608.0f), t1 = __fmaf_rd(rt, b, 8388608.0f);
            acc += __float_as_uint(t0) * 4u + (__float_as_uint(t0) ^ __float_as_uint(t1));
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = __uint_as_float(acc);
}
// same arithmetic, packed: 6 FADD2, 2 FMUL2, 4 FFMA2, 4 MUFU, 4 FFMA2.RM, + the same integer tail
__global__ void k_mix_packed(float *out, float a, float b)
{
    u64 x[2], y[2], zc[2]; unsigned int acc = 0;
    for (int i = 0; i < 2; ++i) { x[i] = pk(a + threadIdx.x * 1e-3f + i, a + i + 2); y[i] = pk(b + i, b + i + 2); zc[i] = pk(a - i, a - i - 2); }
    const u64 va = pk(a, a), vb = pk(b, b), mg = pk(8388608.0f, 8388608.0f);
    for (int it = 0; it < ITERS; ++it) {
        const float wf = __int_as_float(0x3f800000 | (it & 0xff));
        const u64 w = pk(wf, wf);
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            u64 dx, dy, dz, s, t0, t1;
            asm volatile("sub.rn.f32x2 %0, %1, %2;" : "=l"(dx) : "l"(w), "l"(x[i]));
            asm volatile("sub.rn.f32x2 %0, %1, %2;" : "=l"(dy) : "l"(w), "l"(y[i]));
            asm volatile("sub.rn.f32x2 %0, %1, %2;" : "=l"(dz) : "l"(w), "l"(zc[i]));
            asm volatile("mul.rn.f32x2 %0, %1, %1;" : "=l"(s) : "l"(dx));
            asm volatile("fma.rn.f32x2 %0, %1, %1, %0;" : "+l"(s) : "l"(dy));
            asm volatile("fma.rn.f32x2 %0, %1, %1, %0;" : "+l"(s) : "l"(dz));
            float r0, r1;
            asm volatile("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(__uint_as_float((unsigned)s)));
            asm volatile("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r1) : "f"(__uint_as_float((unsigned)(s >> 32))));
            const u64 rr = pk(r0, r1);
            asm volatile("fma.rm.f32x2 %0, %1, %2, %3;" : "=l"(t0) : "l"(rr), "l"(va), "l"(mg));
            asm volatile("fma.rm.f32x2 %0, %1, %2, %3;" : "=l"(t1) : "l"(rr), "l"(vb), "l"(mg));
            acc += (unsigned)t0 * 4u + ((unsigned)t0 ^ (unsigned)t1);
            acc += (unsigned)(t0 >> 32) * 4u + ((unsigned)(t0 >> 32) ^ (unsigned)(t1 >> 32));
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = __uint_as_float(acc);
}


// Shared-memory atomics of the pair kernels: the 32 lanes of a warp hit ~32 different counters inside a window of 128
// classes (one observed molecule against 32 reference molecules).  LAYOUT 0: counters at a 4-byte stride as in
// pa_cells.cu (different classes collide in a bank: ~3 wavefronts per ATOMS); LAYOUT 1: one counter per class AND lane
// ([class][32], 128 KB for 1000 classes): every lane its own bank, one wavefront per instruction.  Same index stream.
constexpr int ITERS_A = 4096;
template <int LAYOUT>
__global__ void k_atoms_layout(unsigned int *out, int nbins, int window, unsigned int seed)
{
    extern __shared__ unsigned int h[];
    const int nwords = LAYOUT ? nbins * 32 : nbins;
    for (int k = threadIdx.x; k < nwords; k += blockDim.x) h[k] = 0;
    __syncthreads();
    unsigned int s = seed ^ (blockIdx.x * 9781u + threadIdx.x * 6271u + 1u);
    unsigned int sw = seed ^ (blockIdx.x * 7919u + (threadIdx.x >> 5) * 104729u + 7u);      // warp-uniform stream: window position
    const int lane = threadIdx.x & 31;
    unsigned int acc = 0;
    for (int it = 0; it < ITERS_A; ++it) {
        sw = sw * 1664525u + 1013904223u;
        const int base = (int)(((unsigned long long)(sw >> 8) * (unsigned)(nbins - window)) >> 24);   // warp-uniform window position (cheap)
        s = s * 1664525u + 1013904223u;
        const int bin = base + (int)((s >> 8) & (unsigned)(window - 1));                              // window must be a power of two
        const unsigned int addr = LAYOUT ? (unsigned)(bin * 32 + lane) : (unsigned)bin;
        asm volatile("red.shared.add.u32 [%0], 1;" ::"r"((unsigned)__cvta_generic_to_shared(h + addr)) : "memory");
        acc += bin;
    }
    __syncthreads();
    unsigned int t = acc;
    for (int k = threadIdx.x; k < nwords; k += blockDim.x) t += h[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = t;
}

template <typename F>
static double time_ms(F launch, int reps = 5)
{
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch(); launch();
    CK(cudaDeviceSynchronize());
    std::vector<float> t;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); t.push_back(ms);
    }
    CK(cudaGetLastError());
    std::sort(t.begin(), t.end());
    return t[0];
}

int main(int argc, char **argv)
{
    const int dev = argc > 1 ? atoi(argv[1]) : 0;
    CK(cudaSetDevice(dev));
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, dev));
    const int sms = p.multiProcessorCount, threads = 256, blocks = sms * 8;
    float *fout; CK(cudaMalloc(&fout, sizeof(float) * threads * blocks));
    const double n = (double)threads * blocks * ITERS * 16.0;      // thread-level instructions of the measured kind
    const double t_ffma = time_ms([&] { k_ffma<<<blocks, threads>>>(fout, 1.0f, 0.999f); });
    const double t_frm = time_ms([&] { k_ffma_rm<<<blocks, threads>>>(fout, 1.0f, 0.999f); });
    const double t_p0 = time_ms([&] { k_f32x2<0><<<blocks, threads>>>(fout, 1.0f, 0.999f); });
    const double t_p1 = time_ms([&] { k_f32x2<1><<<blocks, threads>>>(fout, 1.0f, 0.999f); });
    const double t_p2 = time_ms([&] { k_f32x2<2><<<blocks, threads>>>(fout, 1.0f, 0.999f); });
    const double t_p3 = time_ms([&] { k_f32x2<3><<<blocks, threads>>>(fout, 1.0f, 0.999f); });
    const double t_p4 = time_ms([&] { k_f32x2<4><<<blocks, threads>>>(fout, 1.0f, 0.999f); });
    const double t_p5 = time_ms([&] { k_f32x2<5><<<blocks, threads>>>(fout, 1.0f, 0.999f); });
    const double npairs = (double)threads * blocks * ITERS * 4.0;
    const double t_ms = time_ms([&] { k_mix_scalar<<<blocks, threads>>>(fout, 7.37f, 7.38f); });
    const double t_mp = time_ms([&] { k_mix_packed<<<blocks, threads>>>(fout, 7.37f, 7.38f); });
    // shared atomics: 128-thread CTAs; layout 0 at the occupancy of the pair kernel (6 CTAs/SM), layout 1 with one 128 KB
    // histogram per SM shared by a 1024-thread CTA
    unsigned int *uout; CK(cudaMalloc(&uout, sizeof(unsigned int) * 1024 * sms * 8));
    CK(cudaFuncSetAttribute(k_atoms_layout<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 1000 * 32 * 4));
    const double t_a0 = time_ms([&] { k_atoms_layout<0><<<sms * 6, 128, 1000 * 4>>>(uout, 1000, 128, 12345u); });
    const double t_a1 = time_ms([&] { k_atoms_layout<1><<<sms, 768, 1000 * 32 * 4>>>(uout, 1000, 128, 12345u); });
    const double n_a0 = (double)sms * 6 * 128 * ITERS_A, n_a1 = (double)sms * 768 * ITERS_A;
    printf("{\"gpu\": \"%s\", \"sms\": %d,\n", p.name, sms);
    printf(" \"smem_atomics_4byte_stride_per_s\": %.4e, \"smem_atomics_lane_private_banks_per_s\": %.4e,\n", n_a0 / (t_a0 * 1e-3), n_a1 / (t_a1 * 1e-3));
    printf(" \"ffma_inst_per_s\": %.4e, \"ffma_rm_inst_per_s\": %.4e,\n", n / (t_ffma * 1e-3), n / (t_frm * 1e-3));
    printf(" \"ffma2_inst_per_s\": %.4e, \"ffma2_rm_inst_per_s\": %.4e, \"fadd2_inst_per_s\": %.4e,\n", n / (t_p0 * 1e-3), n / (t_p1 * 1e-3), n / (t_p2 * 1e-3));
    printf(" \"ffma2_with_1_lop3_inst_per_s\": %.4e, \"ffma2_with_2_lop3_inst_per_s\": %.4e, \"ffma2_with_1_imad_inst_per_s\": %.4e,\n",
           n / (t_p3 * 1e-3), n / (t_p4 * 1e-3), n / (t_p5 * 1e-3));
    printf(" \"pair_mix_scalar_pairs_per_s\": %.4e, \"pair_mix_packed_pairs_per_s\": %.4e}\n", npairs / (t_ms * 1e-3), npairs / (t_mp * 1e-3));
    return 0;
}
