// Host stand-in for <cuda_runtime.h>: lets g++ compile the device hashers of seqhasher_b200/csrc/*.cuh as plain
// C++ so their arithmetic can be checked against the oracle WITHOUT a GPU (tests/test_emul_cpu.py).  TEST
// INFRASTRUCTURE ONLY — nothing in the product includes this directory; the product has no CPU path.
#pragma once
#include <cstdint>
#include <cstring>
#define __device__
#define __host__
#define __global__
#define __forceinline__ inline __attribute__((always_inline))
#define __noinline__ __attribute__((noinline))
#define __constant__
#define __shared__ static
#define __restrict__ __restrict
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))
struct uint2 { uint32_t x, y; };
struct uint4 { uint32_t x, y, z, w; };
struct ulonglong2 { unsigned long long x, y; };
static inline uint2 make_uint2(uint32_t x, uint32_t y) { return uint2{x, y}; }
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }
static inline uint32_t __funnelshift_l(uint32_t lo, uint32_t hi, uint32_t s) { s &= 31; return s ? (hi << s) | (lo >> (32 - s)) : hi; }
static inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t s) { s &= 31; return s ? (lo >> s) | (hi << (32 - s)) : lo; }
static inline uint32_t __funnelshift_lc(uint32_t lo, uint32_t hi, uint32_t s) { if (s > 32) s = 32; return s == 0 ? hi : s == 32 ? lo : (hi << s) | (lo >> (32 - s)); }
static inline uint32_t __funnelshift_rc(uint32_t lo, uint32_t hi, uint32_t s) { if (s > 32) s = 32; return s == 0 ? lo : s == 32 ? hi : (lo >> s) | (hi << (32 - s)); }
static inline uint32_t __byte_perm(uint32_t a, uint32_t b, uint32_t sel)
{
    const uint64_t v = ((uint64_t)b << 32) | a;
    uint32_t r = 0;
    for (int i = 0; i < 4; i++) r |= (uint32_t)((v >> (8 * ((sel >> (4 * i)) & 7))) & 0xff) << (8 * i);
    return r;
}
static inline uint64_t __umul64hi(uint64_t a, uint64_t b) { return (uint64_t)(((unsigned __int128)a * b) >> 64); }
static inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
template <class T> static inline T __ldg(const T *p) { return *p; }
static inline void __syncwarp(unsigned mask = 0xffffffffu) { (void)mask; }
static inline uint32_t __brev(uint32_t x)
{
    uint32_t r = 0;
    for (int i = 0; i < 32; i++) r |= ((x >> i) & 1u) << (31 - i);
    return r;
}
static inline int __popc(uint32_t x) { return __builtin_popcount(x); }
static inline int __clz(uint32_t x) { return x ? __builtin_clz(x) : 32; }
static inline int __ffs(uint32_t x) { return __builtin_ffs((int)x); }
static inline size_t __cvta_generic_to_shared(const void *p) { return (size_t)p; }
using std::min;
using std::max;
typedef int cudaError_t;
typedef void *cudaStream_t;
struct emul_dim3 { unsigned x = 0, y = 0, z = 0; };
static emul_dim3 threadIdx, blockIdx;
static emul_dim3 blockDim = {1, 1, 1}, gridDim = {1, 1, 1};
