// pa_microbench.cu -- measures the denominators of the roofline this path is bound by on the
// GPU it runs on: FP64-pipe instruction issue rate without FMA credit (DADD/DMUL/DSETP), the
// shared-memory atomic rate with an RDF-like bin distribution, and MUFU.SQRT.  MEASURED_PEAKS.json
// (driver-written) only has HBM and bf16 numbers, neither of which bounds this kernel.
// Prints one JSON object.  Usage: pa_microbench [device]
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

constexpr int ITERS = 4096;

// 8 independent chains per thread, 2 instructions per chain per iteration
__global__ void k_dadd(double *out, double a, double b)
{
    double x[8];
    for (int i = 0; i < 8; ++i) x[i] = a + threadIdx.x * 1e-9 + i;
    for (int it = 0; it < ITERS; ++it)
#pragma unroll
        for (int i = 0; i < 8; ++i) { x[i] = __dadd_rn(x[i], b); x[i] = __dadd_rn(x[i], -a); }
    double s = 0; for (int i = 0; i < 8; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_dmul(double *out, double a, double b)
{
    double x[8];
    for (int i = 0; i < 8; ++i) x[i] = a + threadIdx.x * 1e-9 + i;
    for (int it = 0; it < ITERS; ++it)
#pragma unroll
        for (int i = 0; i < 8; ++i) { x[i] = __dmul_rn(x[i], b); x[i] = __dmul_rn(x[i], a); }
    double s = 0; for (int i = 0; i < 8; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_dmix(double *out, double a, double b)      // the kernel's mix: sub, mul
{
    double x[8];
    for (int i = 0; i < 8; ++i) x[i] = a + threadIdx.x * 1e-9 + i;
    for (int it = 0; it < ITERS; ++it)
#pragma unroll
        for (int i = 0; i < 8; ++i) { x[i] = __dsub_rn(x[i], b); x[i] = __dmul_rn(x[i], a); }
    double s = 0; for (int i = 0; i < 8; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_dfma(double *out, double a, double b)
{
    double x[8];
    for (int i = 0; i < 8; ++i) x[i] = a + threadIdx.x * 1e-9 + i;
    for (int it = 0; it < ITERS; ++it)
#pragma unroll
        for (int i = 0; i < 8; ++i) { x[i] = __fma_rn(x[i], b, a); x[i] = __fma_rn(x[i], a, b); }
    double s = 0; for (int i = 0; i < 8; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_dmin(double *out, double a, double b)      // DSETP + 2 FSEL per min
{
    double x[8], y[8];
    for (int i = 0; i < 8; ++i) { x[i] = a + threadIdx.x * 1e-9 + i; y[i] = b + i; }
    for (int it = 0; it < ITERS; ++it)
#pragma unroll
        for (int i = 0; i < 8; ++i) { x[i] = x[i] < y[i] ? x[i] : y[i]; y[i] = y[i] < x[i] ? x[i] : y[i]; y[i] = __longlong_as_double(__double_as_longlong(y[i]) ^ it); }
    double s = 0; for (int i = 0; i < 8; ++i) s += x[i] + y[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_mufu(float *out, float a)
{
    float x[8];
    for (int i = 0; i < 8; ++i) x[i] = a + threadIdx.x + i;
    for (int it = 0; it < ITERS; ++it)
#pragma unroll
        for (int i = 0; i < 8; ++i) { asm volatile("sqrt.approx.ftz.f32 %0, %0;" : "+f"(x[i])); asm volatile("sqrt.approx.ftz.f32 %0, %0;" : "+f"(x[i])); }
    float s = 0; for (int i = 0; i < 8; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_f2f(float *out, double a)
{
    double x[8]; float f[8];
    for (int i = 0; i < 8; ++i) { x[i] = a + threadIdx.x * 1e-3 + i; f[i] = 0; }
    for (int it = 0; it < ITERS; ++it)
#pragma unroll
        for (int i = 0; i < 8; ++i) { float t = __double2float_rn(x[i]); f[i] += t; x[i] = __longlong_as_double(__double_as_longlong(x[i]) + __float_as_int(t)); float u = __double2float_rn(x[i]); f[i] += u; }
    float s = 0; for (int i = 0; i < 8; ++i) s += f[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// shared-memory atomics: every thread does ITERS_A increments into an nbins histogram with bin
// probability ~ r^2 (an RDF out to L/2); `active` of 32 lanes take part (cutoff sphere = 52%).
constexpr int ITERS_A = 2048;
__global__ void k_atoms(unsigned int *out, int nbins, int stride_words, int active_pct, unsigned int seed)
{
    extern __shared__ unsigned int h[];
    for (int k = threadIdx.x; k < nbins * stride_words; k += blockDim.x) h[k] = 0;
    __syncthreads();
    unsigned int s = seed ^ (blockIdx.x * 9781u + threadIdx.x * 6271u + 1u);
    unsigned int acc = 0;
    for (int it = 0; it < ITERS_A; ++it) {
        s = s * 1664525u + 1013904223u;
        const float u = (float)(s >> 8) * (1.0f / 16777216.0f);
        const int bin = min(nbins - 1, (int)(cbrtf(u) * nbins));          // p(bin) ~ bin^2
        s = s * 1664525u + 1013904223u;
        const bool on = (int)((s >> 8) % 100u) < active_pct;
        if (on) atomicAdd(&h[bin * stride_words], 1u);
        acc += bin;
    }
    __syncthreads();
    unsigned int t = acc;
    for (int k = threadIdx.x; k < nbins; k += blockDim.x) t += h[k * stride_words];
    out[blockIdx.x * blockDim.x + threadIdx.x] = t;
}
// same loop without the atomic: the cost of generating the bins, to subtract
__global__ void k_atoms_ref(unsigned int *out, int nbins, int active_pct, unsigned int seed)
{
    unsigned int s = seed ^ (blockIdx.x * 9781u + threadIdx.x * 6271u + 1u);
    unsigned int acc = 0;
    for (int it = 0; it < ITERS_A; ++it) {
        s = s * 1664525u + 1013904223u;
        const float u = (float)(s >> 8) * (1.0f / 16777216.0f);
        const int bin = min(nbins - 1, (int)(cbrtf(u) * nbins));
        s = s * 1664525u + 1013904223u;
        const bool on = (int)((s >> 8) % 100u) < active_pct;
        acc += on ? bin : 1;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
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
    const int sms = p.multiProcessorCount;
    const int threads = 256, blocks = sms * 8;
    double *dout; float *fout; unsigned int *uout;
    CK(cudaMalloc(&dout, sizeof(double) * threads * blocks));
    CK(cudaMalloc(&fout, sizeof(float) * threads * blocks));
    CK(cudaMalloc(&uout, sizeof(unsigned int) * threads * blocks));
    const double n64 = (double)threads * blocks * ITERS * 16.0;   // instructions (thread-level)
    const double t_add = time_ms([&] { k_dadd<<<blocks, threads>>>(dout, 1.0, 1e-7); });
    const double t_mul = time_ms([&] { k_dmul<<<blocks, threads>>>(dout, 1.0000001, 0.9999999); });
    const double t_mix = time_ms([&] { k_dmix<<<blocks, threads>>>(dout, 1.0000001, 1e-7); });
    const double t_fma = time_ms([&] { k_dfma<<<blocks, threads>>>(dout, 1.0000001, 1e-7); });
    const double t_min = time_ms([&] { k_dmin<<<blocks, threads>>>(dout, 1.0, 2.0); });
    const double t_mufu = time_ms([&] { k_mufu<<<blocks, threads>>>(fout, 2.0f); });
    const double t_f2f = time_ms([&] { k_f2f<<<blocks, threads>>>(fout, 2.0); });
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz_max\": %d,\n", p.name, sms, p.clockRate);
    printf(" \"dadd_inst_per_s\": %.4e, \"dmul_inst_per_s\": %.4e, \"dsub_dmul_inst_per_s\": %.4e, \"dfma_inst_per_s\": %.4e,\n",
           n64 / (t_add * 1e-3), n64 / (t_mul * 1e-3), n64 / (t_mix * 1e-3), n64 / (t_fma * 1e-3));
    printf(" \"dmin_per_s\": %.4e, \"mufu_sqrt_per_s\": %.4e, \"f2f_f64_f32_per_s\": %.4e,\n",
           n64 / (t_min * 1e-3), n64 / (t_mufu * 1e-3), n64 / (t_f2f * 1e-3));
    // shared atomics
    const int at = 128, ab = sms * 8;
    printf(" \"smem_atomics\": [");
    bool first = true;
    for (int nbins : { 100, 300, 1000 })
        for (int stride : { 1, 2 })
            for (int pct : { 52, 100 }) {
                const size_t smem = (size_t)nbins * stride * 4;
                const double ta = time_ms([&] { k_atoms<<<ab, at, smem>>>(uout, nbins, stride, pct, 12345u); });
                const double tr = time_ms([&] { k_atoms_ref<<<ab, at>>>(uout, nbins, pct, 12345u); });
                const double n = (double)at * ab * ITERS_A * pct / 100.0;
                printf("%s\n  {\"bins\": %d, \"stride_words\": %d, \"active_pct\": %d, \"atomics_per_s\": %.4e, \"atomics_per_s_net\": %.4e}",
                       first ? "" : ",", nbins, stride, pct, n / (ta * 1e-3), n / (std::max(ta - tr, 1e-6) * 1e-3));
                first = false;
            }
    printf("]}\n");
    return 0;
}
