// common.cuh — shared device/host helpers for libkasa_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <string>

namespace kasa {

constexpr int kSMs = 148;                 // B200: 2 dies x 74 SMs
constexpr uint32_t kFull = 0xffffffffu;

// kASA letter codes (kASA.hpp:621-667, letter & 31)
constexpr uint32_t kLetterCaret = 30;     // '^' codon contains X
constexpr uint32_t kLetterUnder = 31;     // '_' codon contains Z

struct Error {
    int code = 0;
    std::string msg;
};

#define KASA_CUDA_TRY(expr)                                                                        \
    do {                                                                                           \
        cudaError_t _e = (expr);                                                                   \
        if (_e != cudaSuccess) {                                                                   \
            char _b[512];                                                                          \
            snprintf(_b, sizeof(_b), "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            throw kasa::Error{-2, _b};                                                             \
        }                                                                                          \
    } while (0)

#define KASA_REQUIRE(cond, code, text)                                                             \
    do {                                                                                           \
        if (!(cond)) throw kasa::Error{(code), (text)};                                            \
    } while (0)

// identical to kasa_b200/synth.py::splitmix64
__host__ __device__ __forceinline__ uint64_t splitmix64(uint64_t x)
{
    uint64_t z = x + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }
__device__ __forceinline__ uint32_t lanemask_lt()
{
    uint32_t m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

// single-word status traffic for decoupled look-back: relaxed gpu-scope accesses bypass L1
__device__ __forceinline__ uint64_t ld_relaxed_gpu(const uint64_t *p)
{
    uint64_t v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_gpu(uint64_t *p, uint64_t v)
{
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// streaming 64/32-bit accesses: written once / read once data should not pollute L1
__device__ __forceinline__ uint64_t ldg_stream(const uint64_t *p)
{
    uint64_t v;
    asm volatile("ld.global.nc.L1::no_allocate.u64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ uint32_t ldg_stream(const uint32_t *p)
{
    uint32_t v;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}

template <typename T>
__host__ __device__ __forceinline__ T div_up(T a, T b) { return (a + b - 1) / b; }

}  // namespace kasa
