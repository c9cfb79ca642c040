// mem_probe.cu — what HBM3e / L2 give for the access patterns the identify path can choose between (design input, not product):
//   gather  : independent random reads of W bytes (W = 8, 16, 32, 64, 128; aligned to W) from a region of S bytes
//   chain   : two dependent random reads (prefix table entry -> index entries), W bytes each
//   scatter : random writes of W bytes (W = 4, 8, 16, 32) into a region of S bytes (full 32-byte sectors vs partial sectors)
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/mem_probe.cu -o gpurun_out/mem_probe
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

__host__ __device__ __forceinline__ uint64_t mix(uint64_t x)
{
    uint64_t z = x + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

template <int W>
__device__ __forceinline__ uint64_t load_w(const char *p)
{
    uint64_t acc = 0;
    if (W == 8) {
        acc = *reinterpret_cast<const uint64_t *>(p);
    } else if (W == 16) {
        uint64_t a, b;
        asm volatile("ld.global.nc.L1::no_allocate.v2.b64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p));
        acc = a ^ b;
    } else {
#pragma unroll
        for (int o = 0; o < W; o += 32) {
            uint64_t a, b, c, d;
            asm volatile("ld.global.nc.L1::no_allocate.v4.b64 {%0, %1, %2, %3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p + o));
            acc ^= a ^ b ^ c ^ d;
        }
    }
    return acc;
}

template <int W, int U>
__global__ void __launch_bounds__(256) gather_kernel(const char *base, uint64_t slots, uint64_t iters, uint64_t *sink)
{
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t acc = 0;
    for (uint64_t it = 0; it < iters; it += U) {
        uint64_t v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint64_t s = mix(tid * 0x10001 + it + u) % slots;
            v[u] = load_w<W>(base + s * W);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) acc ^= v[u];
    }
    if (acc == 0x1234567) sink[0] = acc;
}

// table (8-byte entries) -> entries: the second address depends on the first load's value
template <int W, int U>
__global__ void __launch_bounds__(256) chain_kernel(const uint64_t *table, uint64_t tslots, const char *base, uint64_t slots, uint64_t iters,
                                                    uint64_t *sink)
{
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t acc = 0;
    for (uint64_t it = 0; it < iters; it += U) {
        uint64_t t[U], v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) t[u] = __ldg(table + mix(tid * 0x10001 + it + u) % tslots);
#pragma unroll
        for (int u = 0; u < U; ++u) v[u] = load_w<W>(base + (t[u] % slots) * W);
#pragma unroll
        for (int u = 0; u < U; ++u) acc ^= v[u];
    }
    if (acc == 0x1234567) sink[0] = acc;
}

template <int W>
__global__ void __launch_bounds__(256) scatter_kernel(char *base, uint64_t slots, uint64_t iters)
{
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (uint64_t it = 0; it < iters; ++it) {
        const uint64_t h = mix(tid * 0x10001 + it);
        char *p = base + (h % slots) * W;
        if (W == 4) *reinterpret_cast<uint32_t *>(p) = (uint32_t)h;
        else if (W == 8) *reinterpret_cast<uint64_t *>(p) = h;
        else if (W == 16) asm volatile("st.global.v2.b64 [%0], {%1, %2};" ::"l"(p), "l"(h), "l"(h) : "memory");
        else asm volatile("st.global.v4.b64 [%0], {%1, %2, %3, %4};" ::"l"(p), "l"(h), "l"(h), "l"(h), "l"(h) : "memory");
    }
}

__global__ void fill_kernel(uint64_t *p, uint64_t n)
{
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) p[i] = mix(i);
}

template <class F>
float time_ms(F f)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    f();                                   // warm-up
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    f();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

int main(int argc, char **argv)
{
    const uint64_t big = (argc > 1 ? strtoull(argv[1], 0, 10) : 20ull) << 30;       // DRAM-resident region
    const uint64_t small = 48ull << 20;                                            // L2-resident region
    char *buf;
    uint64_t *sink, *table;
    const uint64_t tslots = 1ull << 30;                                            // 8 GB prefix table stand-in
    if (cudaMalloc(&buf, big) != cudaSuccess || cudaMalloc(&table, tslots * 8) != cudaSuccess) { printf("alloc failed\n"); return 1; }
    cudaMalloc(&sink, 8);
    fill_kernel<<<148 * 8, 256>>>(reinterpret_cast<uint64_t *>(buf), big / 8);
    fill_kernel<<<148 * 8, 256>>>(table, tslots);
    cudaDeviceSynchronize();
    const int grid = 148 * 8, block = 256;
    const uint64_t threads = (uint64_t)grid * block, iters = 1024;
    const double accesses = (double)threads * iters;
    auto report = [&](const char *what, int w, const char *where, float ms, double n) {
        printf("{\"pattern\": \"%s\", \"bytes\": %d, \"region\": \"%s\", \"ms\": %.3f, \"G_per_s\": %.2f, \"GB_per_s\": %.1f}\n", what, w, where, ms,
               n / ms * 1e-6, n * w / ms * 1e-6);
        fflush(stdout);
    };
#define GATHER(W, U)                                                                                                                    \
    report("gather_u" #U, W, "dram", time_ms([&] { gather_kernel<W, U><<<grid, block>>>(buf, big / W, iters, sink); }), accesses);     \
    report("gather_u" #U, W, "l2", time_ms([&] { gather_kernel<W, U><<<grid, block>>>(buf, small / W, iters, sink); }), accesses);
    GATHER(8, 8) GATHER(16, 8) GATHER(32, 8) GATHER(64, 8) GATHER(128, 4)
    GATHER(8, 2) GATHER(32, 2)
#define CHAIN(W, U)                                                                                                                     \
    report("chain_u" #U, W, "dram", time_ms([&] { chain_kernel<W, U><<<grid, block>>>(table, tslots, buf, big / W, iters, sink); }), accesses); \
    report("chain_u" #U, W, "l2", time_ms([&] { chain_kernel<W, U><<<grid, block>>>(table, (16ull << 20) / 8, buf, small / W, iters, sink); }), accesses);
    CHAIN(32, 8) CHAIN(64, 8) CHAIN(128, 4) CHAIN(32, 2)
#define SCATTER(W)                                                                                                                      \
    report("scatter", W, "dram", time_ms([&] { scatter_kernel<W><<<grid, block>>>(buf, big / W, iters / 4); }), accesses / 4);          \
    report("scatter", W, "l2", time_ms([&] { scatter_kernel<W><<<grid, block>>>(buf, small / W, iters / 4); }), accesses / 4);
    SCATTER(4) SCATTER(8) SCATTER(16) SCATTER(32)
    printf("{\"error\": \"%s\"}\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
