// Microbenchmark: issue rate of packed FFMA2/FADD2/FMUL2 (fma.rn.f32x2 ...) vs scalar FFMA on sm_100a.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a --fmad=false -O3 -o ffma2 ffma2.cu ; run on a B200.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c)
{ unsigned long long d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b)
{ unsigned long long d; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ unsigned long long mul2(unsigned long long a, unsigned long long b)
{ unsigned long long d; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ float fma1(float a, float b, float c)
{ float d; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
__device__ __forceinline__ float add1(float a, float b)
{ float d; asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b)); return d; }

constexpr int ITERS = 4096;
constexpr int ILP = 8;

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, float s)
{
    float r[2 * ILP];
    unsigned long long p[ILP];
    for (int i = 0; i < 2 * ILP; i++) r[i] = s + threadIdx.x + i;
    for (int i = 0; i < ILP; i++) p[i] = ((unsigned long long)__float_as_uint(r[2*i]) << 32) | __float_as_uint(r[2*i+1]);
    unsigned long long cs = ((unsigned long long)__float_as_uint(s) << 32) | __float_as_uint(s * 0.5f);
    for (int it = 0; it < ITERS; it++) {
        if (MODE == 0) {          // 2*ILP scalar FFMA
#pragma unroll
            for (int i = 0; i < 2 * ILP; i++) r[i] = fma1(r[i], s, r[i]);
        } else if (MODE == 1) {   // ILP packed FFMA2 (same flops as mode 0)
#pragma unroll
            for (int i = 0; i < ILP; i++) p[i] = fma2(p[i], cs, p[i]);
        } else if (MODE == 2) {   // 2*ILP scalar FADD
#pragma unroll
            for (int i = 0; i < 2 * ILP; i++) r[i] = add1(r[i], s);
        } else if (MODE == 3) {   // ILP packed FADD2
#pragma unroll
            for (int i = 0; i < ILP; i++) p[i] = add2(p[i], cs);
        } else if (MODE == 4) {   // ILP packed FMUL2
#pragma unroll
            for (int i = 0; i < ILP; i++) p[i] = mul2(p[i], cs);
        } else if (MODE == 5) {   // mixed: ILP FFMA2 + ILP LOP3 (alu pipe) -- does FFMA2 leave issue slots free?
#pragma unroll
            for (int i = 0; i < ILP; i++) { p[i] = fma2(p[i], cs, p[i]); r[i] = __uint_as_float(__float_as_uint(r[i]) ^ (__float_as_uint(r[i + ILP]) >> 3)); }
        } else if (MODE == 6) {   // mixed: 2*ILP FFMA + ILP LOP3
#pragma unroll
            for (int i = 0; i < ILP; i++) { r[i] = fma1(r[i], s, r[i]); r[i+ILP] = fma1(r[i+ILP], s, r[i+ILP]); }
#pragma unroll
            for (int i = 0; i < ILP; i++) { p[i] ^= (p[i] >> 3); }
        }
    }
    float acc = 0;
    for (int i = 0; i < 2 * ILP; i++) acc += r[i];
    for (int i = 0; i < ILP; i++) acc += __uint_as_float((uint32_t)p[i]) + __uint_as_float((uint32_t)(p[i] >> 32));
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int MODE>
void run(const char* name, float* out, double flops_per_iter_thread)
{
    const int blocks = 148 * 8, threads = 256;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k<MODE><<<blocks, threads>>>(out, 1.0001f);
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    for (int i = 0; i < 5; i++) k<MODE><<<blocks, threads>>>(out, 1.0001f);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); ms /= 5;
    double ops = (double)blocks * threads * ITERS * flops_per_iter_thread;
    printf("%-28s %8.3f ms  %8.2f G lane-ops/s  (%.2f lane-ops/clk/SM at 1.965 GHz)\n", name, ms, ops / ms * 1e-6,
           ops / (ms * 1e-3) / 148 / 1.965e9);
}

int main()
{
    float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
    run<0>("16x FFMA (scalar)", out, 16);
    run<1>("8x FFMA2 (packed)", out, 16);
    run<2>("16x FADD (scalar)", out, 16);
    run<3>("8x FADD2 (packed)", out, 16);
    run<4>("8x FMUL2 (packed)", out, 16);
    run<5>("8x FFMA2 + 8x LOP3", out, 16);
    run<6>("16x FFMA + 8x LOP", out, 16);
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
