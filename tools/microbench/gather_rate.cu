// How many incoherent 2x2 footprints per second can one B200 fetch, and what limits it?
// (round 2: k_agents at sensor distance 225 runs with l1tex__m_l1tex2xbar_req_cycles_active at 84 % -- one L1-miss request
//  per cycle and SM -- and 1.41 requests per tap; this measures the candidates for fewer requests per tap.)
//
// Every thread fetches 3 independent random footprints per iteration (like the three sensors) from a region that fits L2.
//   tld4_any      tex2Dgather on a block-linear f32 array, footprint anywhere
//   tld4_line     ... footprint inside one 8x4-texel block (one 128-byte line)
//   tld4_sector   ... footprint inside one 4x2-texel block (one 32-byte sector)
//   tld4_ypad     ... rows padded 3 -> 4 (footprint never straddles a 4-row group): x anywhere
//   texq          tex2D<float4> point fetch from a float4 array holding the whole footprint per cell ("quad layout")
//   ldg128q       LDG.128 from a linear float4 quad layout
//   ldg64x2       2 x LDG.64 from a linear "row pair" layout (v[y][x], v[y+1][x]) per cell
//   ldg4          4 x LDG.32 from the row-major field (the LDG sampler)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__device__ __forceinline__ uint32_t mix32(uint32_t x)
{
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

enum { M_TLD4_ANY, M_TLD4_LINE, M_TLD4_SECTOR, M_TLD4_YPAD, M_TEXQ, M_LDG128Q, M_LDG64X2, M_LDG4, M_COUNT };
static const char* kNames[] = {"tld4_any", "tld4_line", "tld4_sector", "tld4_ypad", "texq", "ldg128q", "ldg64x2", "ldg4"};

template <int MODE>
__global__ void __launch_bounds__(256, 5)
k_gather(cudaTextureObject_t tex, cudaTextureObject_t texq, const float4* __restrict__ quad, const float2* __restrict__ pair,
         const float* __restrict__ lin, int S, int iters, float* __restrict__ out)
{
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    float acc = 0.0f;
    uint32_t h = mix32(tid * 2654435761u + 12345u);
    const uint32_t mask = (uint32_t)S - 1u;        // S power of two; footprints in [0, S-2]
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
        float v[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            h = mix32(h + 0x9e3779b9u);
            uint32_t x = h & mask, y = (h >> 16) & mask;
            if (x > mask - 1u) x = mask - 1u;
            if (y > mask - 1u) y = mask - 1u;
            if (MODE == M_TLD4_LINE) { if ((x & 7u) == 7u) x -= 1u; if ((y & 3u) == 3u) y -= 1u; }
            if (MODE == M_TLD4_SECTOR) { if ((x & 3u) == 3u) x -= 1u; y &= ~1u; }
            if (MODE == M_TLD4_YPAD) { y = (y * 3u) >> 2; y = y + y / 3u; }      // padded row index: never 3 mod 4 ... (y%3 in 0..2 within a 4-row group)
            if (MODE <= M_TLD4_YPAD) {
                const float4 g = tex2Dgather<float4>(tex, (float)x + 1.0f, (float)y + 1.0f, 0);
                v[k] = (g.x + g.y) + (g.z + g.w);
            } else if (MODE == M_TEXQ) {
                const float4 g = tex2D<float4>(texq, (float)x + 0.5f, (float)y + 0.5f);
                v[k] = (g.x + g.y) + (g.z + g.w);
            } else if (MODE == M_LDG128Q) {
                const float4 g = __ldg(quad + (size_t)y * S + x);
                v[k] = (g.x + g.y) + (g.z + g.w);
            } else if (MODE == M_LDG64X2) {
                const float2 a = __ldg(pair + (size_t)y * S + x), b = __ldg(pair + (size_t)y * S + x + 1);
                v[k] = (a.x + a.y) + (b.x + b.y);
            } else {
                const float* r0 = lin + (size_t)y * S + x;
                v[k] = (__ldg(r0) + __ldg(r0 + 1)) + (__ldg(r0 + S) + __ldg(r0 + S + 1));
            }
        }
        acc += v[0] + v[1] + v[2];
    }
    if (acc == 123.456f) out[tid] = acc;
}

template <int MODE>
static int run(cudaTextureObject_t tex, cudaTextureObject_t texq, const float4* quad, const float2* pair, const float* lin, int S, float* out,
               int sm_count, double mhz)
{
    const int iters = 64, blocks = sm_count * 5 * 8;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    k_gather<MODE><<<blocks, 256>>>(tex, texq, quad, pair, lin, S, 4, out);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    k_gather<MODE><<<blocks, 256>>>(tex, texq, quad, pair, lin, S, iters, out);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms = 0; CK(cudaEventElapsedTime(&ms, e0, e1));
    const double taps = 3.0 * iters * (double)blocks * 256;
    printf("%-12s region %5d^2: %8.3f ms  %7.2f Gtaps/s  %5.2f SM-cycles per tap (at %.0f MHz)\n", kNames[MODE], S, ms, taps / ms * 1e-6,
           ms * 1e-3 * mhz * 1e6 * sm_count / taps, mhz);
    return 0;
}

int main(int argc, char** argv)
{
    const int S = argc > 1 ? atoi(argv[1]) : 2048;          // 2048^2 f32 = 16 MB (quad layout 64 MB): L2-resident like the sweep band of config 3
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    int clk = 0; CK(cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0));
    const double mhz = clk / 1000.0;
    const size_t n = (size_t)S * S;
    float* h = (float*)malloc(n * 4);
    for (size_t i = 0; i < n; ++i) h[i] = (float)(i % 977) * 1e-3f;
    cudaChannelFormatDesc fd = cudaCreateChannelDesc<float>();
    cudaArray_t arr; CK(cudaMallocArray(&arr, &fd, S, S + S / 3 + 4, cudaArrayTextureGather));
    CK(cudaMemcpy2DToArray(arr, 0, 0, h, (size_t)S * 4, (size_t)S * 4, S, cudaMemcpyHostToDevice));
    CK(cudaMemcpy2DToArray(arr, 0, S, h, (size_t)S * 4, (size_t)S * 4, S / 3, cudaMemcpyHostToDevice));
    cudaResourceDesc rd{}; rd.resType = cudaResourceTypeArray; rd.res.array.array = arr;
    cudaTextureDesc td{}; td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp; td.filterMode = cudaFilterModePoint;
    td.readMode = cudaReadModeElementType; td.normalizedCoords = 0;
    cudaTextureObject_t tex; CK(cudaCreateTextureObject(&tex, &rd, &td, nullptr));
    cudaChannelFormatDesc fq = cudaCreateChannelDesc<float4>();
    cudaArray_t arrq; CK(cudaMallocArray(&arrq, &fq, S, S));
    float4* hq = (float4*)malloc(n * 16);
    for (size_t i = 0; i < n; ++i) hq[i] = make_float4(h[i], h[(i + 1) % n], h[(i + S) % n], h[(i + S + 1) % n]);
    CK(cudaMemcpy2DToArray(arrq, 0, 0, hq, (size_t)S * 16, (size_t)S * 16, S, cudaMemcpyHostToDevice));
    cudaResourceDesc rq{}; rq.resType = cudaResourceTypeArray; rq.res.array.array = arrq;
    cudaTextureObject_t texq; CK(cudaCreateTextureObject(&texq, &rq, &td, nullptr));
    float4* quad; float2* pair; float* lin; float* out;
    CK(cudaMalloc(&quad, n * 16 + 64)); CK(cudaMalloc(&pair, n * 8 + 64)); CK(cudaMalloc(&lin, n * 4 + (size_t)S * 8)); CK(cudaMalloc(&out, (size_t)prop.multiProcessorCount * 5 * 8 * 256 * 4));
    CK(cudaMemcpy(quad, hq, n * 16, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(pair, hq, n * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(lin, h, n * 4, cudaMemcpyHostToDevice));
    printf("%s, %d SMs, %.0f MHz; 3 taps in flight per thread, 256 x 5 CTAs per SM\n", prop.name, prop.multiProcessorCount, mhz);
    const int sms = prop.multiProcessorCount;
    if (run<M_TLD4_ANY>(tex, texq, quad, pair, lin, S, out, sms, mhz)) return 1;
    if (run<M_TLD4_LINE>(tex, texq, quad, pair, lin, S, out, sms, mhz)) return 1;
    if (run<M_TLD4_SECTOR>(tex, texq, quad, pair, lin, S, out, sms, mhz)) return 1;
    if (run<M_TLD4_YPAD>(tex, texq, quad, pair, lin, S, out, sms, mhz)) return 1;
    if (run<M_TEXQ>(tex, texq, quad, pair, lin, S, out, sms, mhz)) return 1;
    if (run<M_LDG128Q>(tex, texq, quad, pair, lin, S, out, sms, mhz)) return 1;
    if (run<M_LDG64X2>(tex, texq, quad, pair, lin, S, out, sms, mhz)) return 1;
    if (run<M_LDG4>(tex, texq, quad, pair, lin, S, out, sms, mhz)) return 1;
    printf("status %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
