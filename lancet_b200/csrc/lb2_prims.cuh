// lb2_prims.cuh -- the handful of CTA primitives the pipeline is written against.
//
// Product build: nvcc, sm_100a, one CTA of LB2_THREADS threads per window.
// LB2_HOSTSIM build (tests/hostsim only): the same source compiled by g++ with a ONE-thread
// "CTA" so the control flow can be debugged in a container without a GPU.  The simulation is
// never shipped, never loaded by the package and is not the oracle.
#ifndef LB2_PRIMS_CUH
#define LB2_PRIMS_CUH

#include <stdint.h>

#ifdef LB2_HOSTSIM
#include <string.h>
#include <math.h>
#define LB2_DEV static inline
#define LB2_DEVNI static
static inline unsigned lb2_tid()  { return 0; }
static inline unsigned lb2_nthr() { return 1; }
static inline void     lb2_sync() {}
static inline uint64_t lb2_cas64(uint64_t *p, uint64_t cmp, uint64_t val) { uint64_t o = *p; if (o == cmp) *p = val; return o; }
static inline uint32_t lb2_cas32(uint32_t *p, uint32_t cmp, uint32_t val) { uint32_t o = *p; if (o == cmp) *p = val; return o; }
static inline uint32_t lb2_add32(uint32_t *p, uint32_t v) { uint32_t o = *p; *p = o + v; return o; }
static inline uint32_t lb2_sub32(uint32_t *p, uint32_t v) { uint32_t o = *p; *p = o - v; return o; }
static inline uint32_t lb2_max32(uint32_t *p, uint32_t v) { uint32_t o = *p; if (v > o) *p = v; return o; }
static inline uint32_t lb2_min32(uint32_t *p, uint32_t v) { uint32_t o = *p; if (v < o) *p = v; return o; }
static inline uint32_t lb2_or32 (uint32_t *p, uint32_t v) { uint32_t o = *p; *p = o | v; return o; }
static inline uint64_t lb2_ld64(const uint64_t *p) { return *p; }
static inline uint32_t lb2_ld32(const uint32_t *p) { return *p; }
static inline int lb2_ctz64(uint64_t x) { return __builtin_ctzll(x); }
static inline int lb2_clz32(uint32_t x) { return __builtin_clz(x); }
static inline unsigned long long lb2_clock() { return 0; }
#else
#define LB2_DEV   __device__ __forceinline__
#define LB2_DEVNI __device__ __noinline__
LB2_DEV unsigned lb2_tid()  { return threadIdx.x; }
LB2_DEV unsigned lb2_nthr() { return blockDim.x; }
LB2_DEV void     lb2_sync() { __syncthreads(); }
LB2_DEV uint64_t lb2_cas64(uint64_t *p, uint64_t cmp, uint64_t val) { return atomicCAS((unsigned long long *)p, (unsigned long long)cmp, (unsigned long long)val); }
LB2_DEV uint32_t lb2_cas32(uint32_t *p, uint32_t cmp, uint32_t val) { return atomicCAS(p, cmp, val); }
LB2_DEV uint32_t lb2_add32(uint32_t *p, uint32_t v) { return atomicAdd(p, v); }
LB2_DEV uint32_t lb2_sub32(uint32_t *p, uint32_t v) { return atomicSub(p, v); }
LB2_DEV uint32_t lb2_max32(uint32_t *p, uint32_t v) { return atomicMax(p, v); }
LB2_DEV uint32_t lb2_min32(uint32_t *p, uint32_t v) { return atomicMin(p, v); }
LB2_DEV uint32_t lb2_or32 (uint32_t *p, uint32_t v) { return atomicOr(p, v); }
LB2_DEV uint64_t lb2_ld64(const uint64_t *p) { return *(const volatile uint64_t *)p; }
LB2_DEV uint32_t lb2_ld32(const uint32_t *p) { return *(const volatile uint32_t *)p; }
LB2_DEV int lb2_ctz64(uint64_t x) { return __ffsll((long long)x) - 1; }
LB2_DEV int lb2_clz32(uint32_t x) { return __clz((int)x); }
LB2_DEV unsigned long long lb2_clock() { return (unsigned long long)clock64(); }
#endif

#endif
