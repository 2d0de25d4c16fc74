// Does a tiny kernel on a high-priority side stream start while a machine-filling kernel runs on the main stream?
// Pattern of the overlapped strip exchange: main: A (long) -> T (fills the GPU); side: wait(A done) -> B (tiny).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(256) k_big(float* p, size_t n, int reps)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        float v = p[i];
        for (int r = 0; r < reps; ++r) v = v * 1.0001f + 0.5f;
        p[i] = v;
    }
}
__global__ void __launch_bounds__(128) k_stream(const float4* __restrict__ in, float4* __restrict__ out, size_t n4)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;     // one float4 per thread: many short CTAs, like k_trail_rows
    if (i < n4) { float4 v = in[i]; v.x += 1.f; out[i] = v; }
}
__global__ void k_tiny(int* flag) { if (threadIdx.x == 0) atomicAdd(flag, 1); }

int main()
{
    const size_t n = 64ull << 20;   // 256 MB
    float *a, *b, *c; int* flag;
    cudaMalloc(&a, n * 4); cudaMalloc(&b, n * 4); cudaMalloc(&c, n * 4); cudaMalloc(&flag, 4);
    cudaMemset(a, 0, n * 4); cudaMemset(b, 0, n * 4); cudaMemset(flag, 0, 4);
    int lo, hi; cudaDeviceGetStreamPriorityRange(&lo, &hi);
    printf("priority range: least %d greatest %d\n", lo, hi);
    cudaStream_t mainS, sideS[2];
    cudaStreamCreateWithFlags(&mainS, cudaStreamNonBlocking);
    cudaStreamCreateWithPriority(&sideS[0], cudaStreamNonBlocking, hi);
    cudaStreamCreateWithPriority(&sideS[1], cudaStreamNonBlocking, lo);
    cudaEvent_t fork, e0, e1, e2, e3, e4;
    cudaEventCreateWithFlags(&fork, cudaEventDisableTiming);
    cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&e2); cudaEventCreate(&e3); cudaEventCreate(&e4);
    for (int variant = 0; variant < 6; ++variant) {
        const int prio = variant & 1;            // 0: high-priority side stream, 1: low
        const int order = variant >> 1;          // 0: T launched first, 1: B launched first, 2: B x4 chained, launched first
        cudaStream_t side = sideS[prio];
        float tB = 0, tT = 0;
        for (int it = 0; it < 6; ++it) {
            k_big<<<148 * 8, 256, 0, mainS>>>(a, n / 4, 40);                   // "A"
            cudaEventRecord(fork, mainS);
            cudaEventRecord(e0, mainS);
            cudaStreamWaitEvent(side, fork, 0);
            auto launchT = [&] { k_stream<<<(unsigned)((n / 4 + 127) / 128), 128, 0, mainS>>>((const float4*)b, (float4*)c, n / 4); cudaEventRecord(e1, mainS); };
            auto launchB = [&] { cudaEventRecord(e2, side); for (int k = 0; k < (order == 2 ? 4 : 1); ++k) k_tiny<<<1, 128, 0, side>>>(flag); cudaEventRecord(e3, side); };
            if (order == 0) { launchT(); launchB(); } else { launchB(); launchT(); }
            cudaStreamWaitEvent(mainS, e3, 0);
            cudaEventRecord(e4, mainS);
            cudaDeviceSynchronize();
            if (it >= 2) { float x; cudaEventElapsedTime(&x, e0, e3); tB += x; cudaEventElapsedTime(&x, e0, e1); tT += x; }
        }
        printf("side prio %s, %s: tiny kernel(s) done %.1f us after A, T done %.1f us after A\n", prio ? "LOW " : "HIGH",
               order == 0 ? "T launched first      " : order == 1 ? "B launched first      " : "4 chained B, launched first", tB / 4 * 1e3, tT / 4 * 1e3);
    }
    printf("status %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
