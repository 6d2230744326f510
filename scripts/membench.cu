// membench.cu -- characterises B200's random-access memory path for the lookup
// kernels (run on the GPU box: nvcc -arch=sm_100a -O3 membench.cu -o membench).
// Each thread issues `iters` rounds of ILP independent random "touches"; one
// touch = V consecutive 16-byte loads starting at a random aligned chunk of
// V*16 bytes.  Reports touches/s and the implied bytes/s.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint64_t mix(uint64_t x) { x ^= x >> 33; x *= 0xFF51AFD7ED558CCDull; x ^= x >> 33; x *= 0xC4CEB9FE1A85EC53ull; return x ^ (x >> 33); }

template <int V, int ILP, int W>   // V 16-byte loads per touch; W: width of a single load in bytes (4, 8, 16)
__global__ void __launch_bounds__(256) touch_kernel(const uint8_t *buf, uint64_t n_chunks, uint32_t iters, uint64_t seed, unsigned long long *sink) {
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t x = mix(t + seed);
    uint64_t acc = 0;
    for (uint32_t it = 0; it < iters; ++it) {
        uint4 v[ILP][V];
#pragma unroll
        for (int j = 0; j < ILP; ++j) {
            x = mix(x + j + 1);
            const uint8_t *p = buf + __umul64hi(x, n_chunks) * (uint64_t)(V * 16);
#pragma unroll
            for (int k = 0; k < V; ++k) {
                if (W == 16) v[j][k] = __ldg(reinterpret_cast<const uint4 *>(p) + k);
                else if (W == 8) { uint2 u = __ldg(reinterpret_cast<const uint2 *>(p + 16 * k)); v[j][k] = make_uint4(u.x, u.y, 0, 0); }
                else { uint32_t u = __ldg(reinterpret_cast<const uint32_t *>(p + 16 * k)); v[j][k] = make_uint4(u, 0, 0, 0); }
            }
        }
#pragma unroll
        for (int j = 0; j < ILP; ++j)
#pragma unroll
            for (int k = 0; k < V; ++k) acc += v[j][k].x ^ v[j][k].y ^ v[j][k].z ^ v[j][k].w;
    }
    if (acc == 0x1234567ull) atomicAdd(sink, 1ull);
}

template <int V, int ILP, int W>
static int run(const char *name, const uint8_t *buf, uint64_t ws, unsigned long long *sink, int blocks_per_sm) {
    uint64_t n_chunks = ws / (V * 16);
    uint32_t iters = 64 / ILP;
    int blocks = 148 * blocks_per_sm * 16;
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    float best = 1e30f;
    const int reps = getenv("MEMBENCH_REPS") ? atoi(getenv("MEMBENCH_REPS")) : 4;
    for (int rep = 0; rep < reps; ++rep) {
        CK(cudaEventRecord(a));
        touch_kernel<V, ILP, W><<<blocks, 256>>>(buf, n_chunks, iters, 1234 + rep, sink);
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        float ms; CK(cudaEventElapsedTime(&ms, a, b));
        if ((rep || reps == 1) && ms < best) best = ms;
    }
    double touches = (double)blocks * 256 * iters * ILP;
    printf("%-34s ws=%5.1f GB  %7.2f G touches/s  %8.1f GB/s requested  (%.2f ms)\n", name, ws / 1e9, touches / best / 1e6,
           touches * V * (W == 16 ? 16 : W) / best / 1e6, best);
    return 0;
}

int main() {
    size_t ws_big = (size_t)8 << 30, ws_small = (size_t)1 << 30;
    uint8_t *buf; unsigned long long *sink;
    CK(cudaMalloc(&buf, ws_big)); CK(cudaMalloc(&sink, 8));
    CK(cudaMemset(buf, 1, ws_big)); CK(cudaMemset(sink, 0, 8));
    for (int pass = 0; pass < 2; ++pass) {
        size_t lim = 0;
        if (pass == 1) { cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, 32); }
        cudaDeviceGetLimit(&lim, cudaLimitMaxL2FetchGranularity);
        printf("---- L2 fetch granularity limit = %zu\n", lim);
        for (size_t ws : {ws_big, ws_small}) {
            run<1, 8, 4>("1 x LDG.32  per 16B chunk, ILP8", buf, ws, sink, 8);
            run<1, 8, 8>("1 x LDG.64  per 16B chunk, ILP8", buf, ws, sink, 8);
            run<1, 8, 16>("1 x LDG.128 per 16B chunk, ILP8", buf, ws, sink, 8);
            run<2, 8, 16>("2 x LDG.128 per 32B sector, ILP8", buf, ws, sink, 8);
            run<4, 4, 16>("4 x LDG.128 per 64B chunk, ILP4", buf, ws, sink, 8);
            run<8, 2, 16>("8 x LDG.128 per 128B line, ILP2", buf, ws, sink, 8);
            run<1, 1, 16>("1 x LDG.128 per 16B chunk, ILP1", buf, ws, sink, 8);
            run<2, 1, 16>("2 x LDG.128 per 32B sector, ILP1", buf, ws, sink, 8);
            run<4, 1, 16>("4 x LDG.128 per 64B chunk, ILP1", buf, ws, sink, 8);
            run<1, 2, 16>("1 x LDG.128 per 16B chunk, ILP2", buf, ws, sink, 8);
            run<1, 4, 16>("1 x LDG.128 per 16B chunk, ILP4", buf, ws, sink, 8);
        }
    }
    return 0;
}
