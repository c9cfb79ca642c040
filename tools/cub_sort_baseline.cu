// Library baseline for the radix sort: cub::DeviceRadixSort::SortPairs (CUDA 12.9's CUB, onesweep) on the identify sort's
// shape — 64-bit keys, 32-bit payload, bits [28, 60) — timed with CUDA events.  Not part of the product; DESIGN.md quotes it.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/cub_sort_baseline.cu -o gpurun_out/cub_sort_baseline && gpurun_out/cub_sort_baseline 1742000000
#include <cub/cub.cuh>
#include <cstdio>
#include <cstdlib>

__global__ void fill(unsigned long long *k, unsigned *v, size_t n)
{
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned long long z = i + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    k[i] = (z ^ (z >> 31)) & ((1ull << 60) - 1);
    v[i] = (unsigned)i;
}

int main(int argc, char **argv)
{
    size_t n = argc > 1 ? strtoull(argv[1], 0, 10) : 1742000000ull;
    int begin_bit = argc > 2 ? atoi(argv[2]) : 28, end_bit = argc > 3 ? atoi(argv[3]) : 60;
    unsigned long long *k0, *k1;
    unsigned *v0, *v1;
    cudaMalloc(&k0, n * 8); cudaMalloc(&k1, n * 8); cudaMalloc(&v0, n * 4); cudaMalloc(&v1, n * 4);
    cub::DoubleBuffer<unsigned long long> dk(k0, k1);
    cub::DoubleBuffer<unsigned> dv(v0, v1);
    size_t tmp_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, dk, dv, n, begin_bit, end_bit);
    void *tmp;
    cudaMalloc(&tmp, tmp_bytes);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int it = 0; it < 5; ++it) {
        fill<<<(unsigned)((n + 255) / 256), 256>>>(dk.Current(), dv.Current(), n);
        cudaEventRecord(e0);
        cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, dk, dv, n, begin_bit, end_bit);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (it && ms < best) best = ms;
    }
    cudaError_t err = cudaGetLastError();
    printf("{\"what\": \"cub::DeviceRadixSort::SortPairs u64 keys + u32 values, bits [%d, %d)\", \"n\": %zu, \"ms\": %.3f, \"temp_bytes\": %zu, \"error\": \"%s\"}\n",
           begin_bit, end_bit, n, best, tmp_bytes, cudaGetErrorString(err));
    return 0;
}
