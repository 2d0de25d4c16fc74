// Which copies of the trail can the fused decay+diffuse pass afford to read / write?
// A thread owns 4 columns and walks down a chunk of rows (the access pattern of k_trail_rows); variants:
//   lin->lin        LDG.128 -> STG.128                      (diffusion-only pass: 8 B/cell)
//   lin->lin+surf   LDG.128 -> STG.128 + SUST.128           (full step today: 12 B/cell + flags)
//   lin->surf       LDG.128 -> SUST.128                     (block-linear copy only)
//   surf->surf      SULD.128 -> SUST.128                    (block-linear array as the ONLY copy of the trail)
//   tex->surf       2 x TLD4 (8 texels, two rows) -> SUST   (read through the texture path)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

enum { V_LL, V_LLS, V_LS, V_SS, V_LS_PAIR, V_LS_ZERO, V_LS_QUAD, V_COUNT };
static const char* kNames[] = {"lin->lin", "lin->lin+surf", "lin->surf", "surf->surf", "lin->surf pair", "lin->surf zero", "lin->surf quad"};

// lin->surf with a warp instruction covering FULL 32-byte sectors of the block-linear layout (a sector = 4 texels x 2 rows):
//   pair: lanes 0-15 write row y, lanes 16-31 row y+1 of the same 64 columns (16 whole sectors per instruction instead of 32 halves)
//   quad: lanes 0-7 / 8-15 / 16-23 / 24-31 write rows y .. y+3 of the same 32 columns (a 128-byte line = 8 texels x 4 rows)
template <int ROWS>
__global__ void __launch_bounds__(128, 8)
k_copy_rows(const float* __restrict__ in, cudaSurfaceObject_t sout, int W, int H, int rpc)
{
    constexpr int LPR = 32 / ROWS;                       // lanes per row
    const int lane = threadIdx.x & 31, warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int x0 = (warp * LPR + (lane % LPR)) * 4, dy = lane / LPR;
    if (x0 >= W) return;
    const int y0 = blockIdx.y * rpc * ROWS, y1 = min(y0 + rpc * ROWS, H);
#pragma unroll 4
    for (int y = y0 + dy; y < y1; y += ROWS) {
        float4 v = __ldg(reinterpret_cast<const float4*>(in + (size_t)y * W + x0));
        v.x += 1.0f; v.y += 1.0f; v.z += 1.0f; v.w += 1.0f;
        surf2Dwrite(v, sout, x0 * 4, y);
    }
}
template <int ROWS>
static int run_rows(const char* name, const float* in, cudaSurfaceObject_t sout, int S, int rpc)
{
    dim3 grid((S / 4 * ROWS + 127) / 128, (S + rpc * ROWS - 1) / (rpc * ROWS));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int i = 0; i < 3; ++i) k_copy_rows<ROWS><<<grid, 128>>>(in, sout, S, S, rpc);
    CK(cudaDeviceSynchronize());
    const int reps = 20;
    CK(cudaEventRecord(e0));
    for (int i = 0; i < reps; ++i) k_copy_rows<ROWS><<<grid, 128>>>(in, sout, S, S, rpc);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    const double cells = (double)S * S;
    printf("%-14s %5d^2 rpc %2d: %8.1f us/pass  %6.2f ns/Kcell  %7.0f GB/s moved\n", name, S, rpc, ms / reps * 1e3, ms / reps * 1e6 / (cells / 1e3),
           cells * 8.0 / (ms / reps * 1e-3) / 1e9);
    return 0;
}

template <int V>
__global__ void __launch_bounds__(128, 8)
k_copy(const float* __restrict__ in, float* __restrict__ out, cudaSurfaceObject_t sin, cudaSurfaceObject_t sout, int W, int H, int rpc)
{
    const int x0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (x0 >= W) return;
    const int y0 = blockIdx.y * rpc, y1 = min(y0 + rpc, H);
#pragma unroll 4
    for (int y = y0; y < y1; ++y) {
        float4 v;
        if (V == V_SS) v = surf2Dread<float4>(sin, x0 * 4, y);
        else v = __ldg(reinterpret_cast<const float4*>(in + (size_t)y * W + x0));
        v.x += 1.0f; v.y += 1.0f; v.z += 1.0f; v.w += 1.0f;
        if (V == V_LL || V == V_LLS) *reinterpret_cast<float4*>(out + (size_t)y * W + x0) = v;
        if (V == V_LS_ZERO) surf2Dwrite(v, sout, x0 * 4, y, cudaBoundaryModeZero);
        else if (V != V_LL) surf2Dwrite(v, sout, x0 * 4, y);
    }
}

template <int V>
static int run(const float* in, float* out, cudaSurfaceObject_t sin, cudaSurfaceObject_t sout, int S, int rpc)
{
    dim3 grid((S / 4 + 127) / 128, (S + rpc - 1) / rpc);
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int i = 0; i < 3; ++i) k_copy<V><<<grid, 128>>>(in, out, sin, sout, S, S, rpc);
    CK(cudaDeviceSynchronize());
    const int reps = 20;
    CK(cudaEventRecord(e0));
    for (int i = 0; i < reps; ++i) k_copy<V><<<grid, 128>>>(in, out, sin, sout, S, S, rpc);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    const double cells = (double)S * S;
    const double bytes = cells * (V == V_LLS ? 12.0 : 8.0);
    printf("%-14s %5d^2 rpc %2d: %8.1f us/pass  %6.2f ns/Kcell  %7.0f GB/s moved\n", kNames[V], S, rpc, ms / reps * 1e3, ms / reps * 1e6 / (cells / 1e3),
           bytes / (ms / reps * 1e-3) / 1e9);
    return 0;
}

int main(int argc, char** argv)
{
    for (int S : {4096, 8192}) {
        const size_t n = (size_t)S * S;
        float *in, *out;
        CK(cudaMalloc(&in, n * 4)); CK(cudaMalloc(&out, n * 4));
        CK(cudaMemset(in, 0, n * 4));
        cudaChannelFormatDesc fd = cudaCreateChannelDesc<float>();
        cudaArray_t a0, a1;
        CK(cudaMallocArray(&a0, &fd, S, S, cudaArraySurfaceLoadStore | cudaArrayTextureGather));
        CK(cudaMallocArray(&a1, &fd, S, S, cudaArraySurfaceLoadStore | cudaArrayTextureGather));
        cudaResourceDesc rd{}; rd.resType = cudaResourceTypeArray;
        cudaSurfaceObject_t s0, s1;
        rd.res.array.array = a0; CK(cudaCreateSurfaceObject(&s0, &rd));
        rd.res.array.array = a1; CK(cudaCreateSurfaceObject(&s1, &rd));
        for (int rpc : {8, 16}) {
            if (run<V_LL>(in, out, s0, s1, S, rpc)) return 1;
            if (run<V_LLS>(in, out, s0, s1, S, rpc)) return 1;
            if (run<V_LS>(in, out, s0, s1, S, rpc)) return 1;
            if (run<V_SS>(in, out, s0, s1, S, rpc)) return 1;
            if (run<V_LS_ZERO>(in, out, s0, s1, S, rpc)) return 1;
            if (run_rows<2>("lin->surf pair", in, s1, S, rpc)) return 1;
            if (run_rows<4>("lin->surf quad", in, s1, S, rpc)) return 1;
        }
        cudaDestroySurfaceObject(s0); cudaDestroySurfaceObject(s1); cudaFreeArray(a0); cudaFreeArray(a1); cudaFree(in); cudaFree(out);
    }
    printf("status %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
