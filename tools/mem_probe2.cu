// mem_probe2.cu — random scatter / gather rate as a function of the region size (L2 .. HBM), design input.
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
__global__ void __launch_bounds__(256) scatter_kernel(char *base, uint64_t slots, uint64_t iters)
{
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (uint64_t it = 0; it < iters; ++it) {
        const uint64_t h = mix(tid * 0x10001 + it);
        char *p = base + (h % slots) * W;
        if (W == 8) *reinterpret_cast<uint64_t *>(p) = h;
        else if (W == 16) asm volatile("st.global.v2.b64 [%0], {%1, %2};" ::"l"(p), "l"(h), "l"(h) : "memory");
        else asm volatile("st.global.v4.b64 [%0], {%1, %2, %3, %4};" ::"l"(p), "l"(h), "l"(h), "l"(h), "l"(h) : "memory");
    }
}
// a scatter that is a permutation inside windows: every slot of the region written exactly once, windows visited in order
// (what an L2-local "scatter inside a window" does): slot = window base + bit-reversal-ish hash permutation inside the window
template <int W>
__global__ void __launch_bounds__(256) window_scatter_kernel(char *base, uint64_t slots, uint64_t win_slots)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < slots; i += stride) {
        const uint64_t w = i / win_slots, j = i % win_slots;
        const uint64_t pj = (j * 0x9E3779B97F4A7C15ull) % win_slots;     // odd multiplier: a permutation when win_slots is a power of two
        char *p = base + (w * win_slots + pj) * W;
        const uint64_t h = i;
        if (W == 8) *reinterpret_cast<uint64_t *>(p) = h;
        else if (W == 16) asm volatile("st.global.v2.b64 [%0], {%1, %2};" ::"l"(p), "l"(h), "l"(h) : "memory");
        else asm volatile("st.global.v4.b64 [%0], {%1, %2, %3, %4};" ::"l"(p), "l"(h), "l"(h), "l"(h), "l"(h) : "memory");
    }
}
template <int W>
__global__ void __launch_bounds__(256) gather_kernel(const char *base, uint64_t slots, uint64_t iters, uint64_t *sink)
{
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t acc = 0;
    for (uint64_t it = 0; it < iters; it += 4) {
        uint64_t v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const uint64_t s = mix(tid * 0x10001 + it + u) % slots;
            uint64_t a, b, c, d;
            asm volatile("ld.global.nc.L1::no_allocate.v4.b64 {%0, %1, %2, %3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(base + s * W));
            v[u] = a ^ b ^ c ^ d;
        }
        acc ^= v[0] ^ v[1] ^ v[2] ^ v[3];
    }
    if (acc == 0x1234567) sink[0] = acc;
}
// warp-coalesced random 128-byte lines: lane l reads 4 bytes of the line chosen by the warp
__global__ void __launch_bounds__(256) gather_line_kernel(const char *base, uint64_t lines, uint64_t iters, uint64_t *sink)
{
    const uint64_t wid = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    uint32_t acc = 0;
    for (uint64_t it = 0; it < iters; it += 4) {
        uint32_t v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = __ldg(reinterpret_cast<const uint32_t *>(base + (mix(wid * 0x10001 + it + u) % lines) * 128) + lane);
        acc ^= v[0] ^ v[1] ^ v[2] ^ v[3];
    }
    if (acc == 0x1234567) sink[0] = acc;
}
template <class F>
float time_ms(F f)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    f();
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    f();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}
int main()
{
    const uint64_t cap = 64ull << 30;
    char *buf;
    uint64_t *sink;
    if (cudaMalloc(&buf, cap) != cudaSuccess) { printf("alloc failed\n"); return 1; }
    cudaMalloc(&sink, 8);
    cudaMemset(buf, 1, cap);
    const int grid = 148 * 8, block = 256;
    const uint64_t threads = (uint64_t)grid * block, iters = 512;
    const double n = (double)threads * iters;
    for (uint64_t mb : {64ull, 256ull, 1024ull, 4096ull, 16384ull, 56000ull}) {
        const uint64_t bytes = mb << 20;
        float s32 = time_ms([&] { scatter_kernel<32><<<grid, block>>>(buf, bytes / 32, iters); });
        float s16 = time_ms([&] { scatter_kernel<16><<<grid, block>>>(buf, bytes / 16, iters); });
        float s8 = time_ms([&] { scatter_kernel<8><<<grid, block>>>(buf, bytes / 8, iters); });
        float g32 = time_ms([&] { gather_kernel<32><<<grid, block>>>(buf, bytes / 32, iters, sink); });
        float gl = time_ms([&] { gather_line_kernel<<<grid, block>>>(buf, bytes / 128, iters, sink); });
        printf("{\"region_MB\": %llu, \"scatter32_G\": %.1f, \"scatter16_G\": %.1f, \"scatter8_G\": %.1f, \"gather32_G\": %.1f, \"gather_line128_G\": %.1f}\n",
               (unsigned long long)mb, n / s32 * 1e-6, n / s16 * 1e-6, n / s8 * 1e-6, n / g32 * 1e-6, n / 32 / gl * 1e-6);
        fflush(stdout);
    }
    // permutation scatter inside windows, whole 16 GB region written once
    for (uint64_t win_mb : {16ull, 32ull, 64ull, 128ull, 512ull}) {
        const uint64_t bytes = 16ull << 30;
        float w8 = time_ms([&] { window_scatter_kernel<8><<<grid, block>>>(buf, bytes / 8, (win_mb << 20) / 8); });
        float w16 = time_ms([&] { window_scatter_kernel<16><<<grid, block>>>(buf, bytes / 16, (win_mb << 20) / 16); });
        float w32 = time_ms([&] { window_scatter_kernel<32><<<grid, block>>>(buf, bytes / 32, (win_mb << 20) / 32); });
        printf("{\"window_MB\": %llu, \"perm8_G\": %.1f, \"perm8_GBs\": %.0f, \"perm16_G\": %.1f, \"perm16_GBs\": %.0f, \"perm32_G\": %.1f, \"perm32_GBs\": %.0f}\n", (unsigned long long)win_mb,
               bytes / 8 / w8 * 1e-6, bytes / w8 * 1e-6, bytes / 16 / w16 * 1e-6, bytes / w16 * 1e-6, bytes / 32 / w32 * 1e-6, bytes / w32 * 1e-6);
        fflush(stdout);
    }
    printf("{\"error\": \"%s\"}\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
