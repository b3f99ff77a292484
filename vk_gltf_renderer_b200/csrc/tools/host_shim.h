// host_shim.h — lets g++ compile the DEVICE traversal code (../traverse.cuh, ../vec.cuh) for the host, one "lane" at a
// time, so that the exact source the sm_100a kernels run can be checked against brute force without a GPU
// (tools/host_traverse_check.cpp, tests/test_device_source_on_host.py).  Only what traverse.cuh uses is provided.
#pragma once
#include <cuda_runtime.h>  // vector types; __host__ / __device__ expand to nothing outside nvcc

#include <cmath>
#include <cstdint>
#include <cstring>

#ifndef __CUDACC__
#ifndef __forceinline__
#define __forceinline__ inline
#endif
#ifndef __noinline__
#define __noinline__ __attribute__((noinline))
#endif
static inline uint32_t __float_as_uint(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
static inline float    __uint_as_float(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }
static inline int      __float_as_int(float f) { int i; std::memcpy(&i, &f, 4); return i; }
static inline float    __int_as_float(int i) { float f; std::memcpy(&f, &i, 4); return f; }
static inline int      __clz(uint32_t x) { return x ? __builtin_clz(x) : 32; }
static inline int      __popc(uint32_t x) { return __builtin_popcount(x); }
template <typename T>
static inline T __ldg(const T* p) { return *p; }
// PRMT: result byte i = byte (selector nibble i) of the 8-byte pool {x bytes 0..3, y bytes 4..7}
static inline uint32_t __byte_perm(uint32_t x, uint32_t y, uint32_t s)
{
  const uint64_t pool = (uint64_t)x | ((uint64_t)y << 32);
  uint32_t       r = 0;
  for(int i = 0; i < 4; i++)
    r |= (uint32_t)((pool >> (8 * ((s >> (4 * i)) & 7))) & 0xff) << (8 * i);
  return r;
}
using std::isnan;
using std::isinf;
// sequential stand-ins for the atomics of the build kernels (the host harness runs the "threads" one after another)
template <typename T>
static inline T atomicAdd(T* p, T v) { const T o = *p; *p = o + v; return o; }
static inline int atomicMin(int* p, int v) { const int o = *p; if(v < o) *p = v; return o; }
static inline int atomicMax(int* p, int v) { const int o = *p; if(v > o) *p = v; return o; }
// one lane per "warp"
static inline unsigned __activemask() { return 1u; }
static inline unsigned __ballot_sync(unsigned, int pred) { return pred ? 1u : 0u; }
#endif
