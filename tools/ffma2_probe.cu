// Microbenchmark: does sm_100a's packed FP32 arithmetic (fma.rn.f32x2 -> FFMA2) free issue slots in a kernel that is
// bound by instruction issue?  Measures warp-instructions per clock per SM sub-partition for
//   A: scalar FFMA chains            B: packed FFMA2 chains (same number of FMAs)
//   C: scalar FFMA + as many LOP3    D: packed FFMA2 + the same LOP3 count (same FMAs as C)
//   E / F: the same with half as many LOP3 (so that the half-rate ALU pipe is not what binds)
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/ffma2_probe.cu -o build/ffma2_probe   (run it on the GPU box)
// Measured on a B200 (1965 MHz): A 70.5 TFLOP/s, B 72.8, C 35.5, D 36.1 (the half-rate ALU pipe binds C and D), E 47.0, F 69.4:
// with issue slots as the limit, the packed form is 1.48x faster for the same arithmetic.
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 4096
#define NCH 8      // independent chains per thread

__device__ __forceinline__ unsigned long long pk(float a, float b)
{
    return ((unsigned long long)__float_as_uint(b) << 32) | __float_as_uint(a);
}
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c)
{
    unsigned long long r;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ float fma1(float a, float b, float c)
{
    float r;
    asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}
__device__ __forceinline__ unsigned lop(unsigned a, unsigned b, unsigned c)
{
    unsigned r;
    asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, float a, float b, unsigned m)
{
    float x[NCH];
    unsigned long long y[NCH / 2];
    unsigned u[NCH];
#pragma unroll
    for (int i = 0; i < NCH; i++) { x[i] = a + i + threadIdx.x; u[i] = m + i + threadIdx.x; }
#pragma unroll
    for (int i = 0; i < NCH / 2; i++) y[i] = pk(x[2 * i], x[2 * i + 1]);
    const unsigned long long ab = pk(a, a), bb = pk(b, b);
    for (int it = 0; it < ITERS; it++) {
        if (MODE == 0 || MODE == 2 || MODE == 4) {
#pragma unroll
            for (int i = 0; i < NCH; i++) x[i] = fma1(x[i], a, b);
        } else {
#pragma unroll
            for (int i = 0; i < NCH / 2; i++) y[i] = fma2(y[i], ab, bb);
        }
        if (MODE == 2 || MODE == 3) {
#pragma unroll
            for (int i = 0; i < NCH; i++) u[i] = lop(u[i], m, u[(i + 1) % NCH]);
        }
        if (MODE == 4 || MODE == 5) {     // half as many ALU-pipe instructions: the 16-lane ALU pipe is not the limit
#pragma unroll
            for (int i = 0; i < NCH / 2; i++) u[i] = lop(u[i], m, u[(i + 1) % (NCH / 2)]);
        }
    }
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < NCH; i++) acc += x[i] + __uint_as_float(u[i]);
#pragma unroll
    for (int i = 0; i < NCH / 2; i++) acc += __uint_as_float((unsigned)y[i]) + __uint_as_float((unsigned)(y[i] >> 32));
    if (acc == 12345.678f) out[0] = acc;
}

template <int MODE>
static void run(const char* name, int sms, double fmas_per_iter, double instr_per_iter)
{
    float* out;
    cudaMalloc(&out, 4);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int blocks = sms * 8;
    k<MODE><<<blocks, 256>>>(out, 1.0001f, 0.5f, 0x1234u);
    cudaDeviceSynchronize();
    float best = 1e9f;
    for (int r = 0; r < 5; r++) {
        cudaEventRecord(e0);
        k<MODE><<<blocks, 256>>>(out, 1.0001f, 0.5f, 0x1234u);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    const double warps = (double)blocks * 8;
    const double tflops = warps * 32 * ITERS * fmas_per_iter * 2 / (best * 1e-3) / 1e12;
    const double ginstr = warps * ITERS * instr_per_iter / (best * 1e-3) / 1e9;
    printf("%-28s %8.3f ms  %7.2f TFLOP/s  %8.1f G warp-instr/s  (per SM sub-partition: %.3f warp-instr/ns)\n", name, best,
           tflops, ginstr, ginstr / (sms * 4));
}

int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    printf("%s, %d SMs, %d MHz max\n", p.name, p.multiProcessorCount, p.clockRate / 1000);
    run<0>("A scalar FFMA", p.multiProcessorCount, NCH, NCH);
    run<1>("B packed FFMA2", p.multiProcessorCount, NCH, NCH / 2);
    run<2>("C scalar FFMA + LOP3", p.multiProcessorCount, NCH, 2 * NCH);
    run<3>("D packed FFMA2 + LOP3", p.multiProcessorCount, NCH, NCH / 2 + NCH);
    run<4>("E scalar FFMA + LOP3/2", p.multiProcessorCount, NCH, NCH + NCH / 2);
    run<5>("F packed FFMA2 + LOP3/2", p.multiProcessorCount, NCH, NCH / 2 + NCH / 2);
    return 0;
}
