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
static inline uint32_t lb2_cas32(uint32_t *p, uint32_t cmp, uint32_t val) { uint32_t o = *p; if (o == cmp) *p = val; return o; }
static inline uint32_t lb2_add32(uint32_t *p, uint32_t v) { uint32_t o = *p; *p = o + v; return o; }
static inline void lb2_max32(uint32_t *p, uint32_t v) { if (v > *p) *p = v; }
static inline void lb2_min32(uint32_t *p, uint32_t v) { if (v < *p) *p = v; }
static inline void lb2_or32 (uint32_t *p, uint32_t v) { *p |= v; }
static inline uint32_t lb2g_add32(uint32_t *p, uint32_t v) { uint32_t o = *p; *p = o + v; return o; }
static inline void lb2g_red_add(uint32_t *p, uint32_t v) { *p += v; }
static inline void lb2g_red_max(uint32_t *p, uint32_t v) { if (v > *p) *p = v; }
static inline void lb2g_red_or(uint32_t *p, uint32_t v) { *p |= v; }
static inline uint32_t lb2g_min32(uint32_t *p, uint32_t v) { uint32_t o = *p; if (v < o) *p = v; return o; }
static inline uint32_t lb2_ld32(const uint32_t *p) { return *p; }
static inline uint32_t lb2_lds(const uint32_t *p) { return *p; }
static inline int lb2_ctz64(uint64_t x) { return __builtin_ctzll(x); }
static inline int lb2_clz32(uint32_t x) { return __builtin_clz(x); }
static inline int lb2_ctz32(uint32_t x) { return __builtin_ctz(x); }
static inline int lb2_popc32(uint32_t x) { return __builtin_popcount(x); }
// length-weighted coverage average of compressNode (src/Graph.cc:2631-2636): two products, a sum, a quotient, each rounded
// (the reference is x86-64 without FMA; the device must not contract the sum of products)
static inline float lb2_wavg(float a, int la, float b, int lb) { volatile float p = a * la; volatile float q = b * lb; volatile float sm = p + q; return sm / (la + lb); }
// the same average for a serial fold, with the reciprocal of the (integer) divisor supplied by the caller
static inline float lb2_rcp_int(uint32_t n) { return 1.0f / (float)n; }
static inline float lb2_wavg_rcp(float a, int la, float b, int lb, float r) { (void)r; return lb2_wavg(a, la, b, lb); }
// (simulation of the device's division, for tests/hostsim's divtest)
static inline float lb2_div_nr2(float s, float n, float r) { float q0 = s * r; float e0 = fmaf(-n, q0, s); float q1 = fmaf(e0, r, q0); float e1 = fmaf(-n, q1, s); return fmaf(e1, r, q1); }
static inline unsigned long long lb2_clock() { return 0; }
// CTA-wide exclusive prefix sum of one value per lane (sc: >= 34 words of shared scratch); every lane must call it
static inline uint32_t lb2_block_excl(uint32_t *sc, uint32_t v, uint32_t *total) { (void)sc; *total = v; return 0; }
// atomics on an address that may be shared OR global (scratch that falls back to the workspace slab)
static inline uint32_t lb2x_exch32(uint32_t *p, uint32_t v) { uint32_t o = *p; *p = v; return o; }
static inline uint32_t lb2x_cas32(uint32_t *p, uint32_t cmp, uint32_t v) { uint32_t o = *p; if (o == cmp) { *p = v; } return o; }
// ---- sub-warp groups for the read staging (8 lanes per read on the device, 1 in the simulation) ----
#define LB2_GS 1
// ---- shared-memory words addressed by a precomputed base (device: 32-bit shared-window address, no generic->shared
// ---- conversion per access; simulation: a plain pointer) ----
typedef uint32_t *lb2_sp;
static inline lb2_sp lb2_sp_of(const void *p) { return (uint32_t *)p; }
static inline lb2_sp lb2_sp_at(lb2_sp b, uint32_t i) { return b + i; }
static inline uint32_t lb2s_ldv(lb2_sp a) { return *a; }
static inline uint32_t lb2s_ld(lb2_sp a) { return *a; }
static inline uint32_t lb2s_cas(lb2_sp a, uint32_t cmp, uint32_t val) { uint32_t o = *a; if (o == cmp) *a = val; return o; }
static inline void lb2s_min(lb2_sp a, uint32_t v) { if (v < *a) *a = v; }
static inline void lb2s_or(lb2_sp a, uint32_t v) { *a |= v; }
static inline uint32_t lb2_fsr(uint32_t lo, uint32_t hi, uint32_t sh) { sh &= 31u; return sh ? ((lo >> sh) | (hi << (32u - sh))) : lo; }
// work items handed out to whole warps: every lane of the warp calls this together and gets its own item index
static inline uint32_t lb2_batch_next(uint32_t *ctr) { return (*ctr)++; }
#define LB2_WARP 1     /* lanes that run warp-cooperative code (one in the simulation) */
static inline uint32_t lb2_ballot(bool p) { return p ? 1u : 0u; }
static inline void lb2_warp_sync() {}
static inline uint32_t lb2_warp_max(uint32_t v) { return v; }
static inline unsigned lb2_lane() { return 0; }
static inline uint32_t lb2_match_any(uint32_t) { return 1u; }      // lanes of the warp holding the same value
static inline uint32_t lb2_warp_excl(uint32_t v, uint32_t *total) { *total = v; return 0; }      // warp-wide exclusive prefix sum
static inline uint32_t lb2_shfl(uint32_t v, uint32_t) { return v; }
static inline uint32_t lb2_shfl_up1(uint32_t v) { return v; }
#define LB2_FQ 1      /* lanes per chain in the coverage fold of the parallel compaction (one per channel on the device) */
static inline unsigned lb2_glane() { return 0; }
static inline unsigned lb2_group() { return 0; }
static inline unsigned lb2_ngroups() { return 1; }
static inline uint32_t lb2_gmin(uint32_t v) { return v; }
static inline uint32_t lb2_gmax(uint32_t v) { return v; }
static inline uint32_t lb2_gor(uint32_t v) { return v; }
// 16 bytes from an arbitrarily aligned address (the device version reads whole aligned words around them)
static inline void lb2_load16(const char *p, uint32_t o[4]) { memcpy(o, p, 16); }
// ---- bulk-async staging (device: cp.async.bulk global -> shared, completion counted on an mbarrier; simulation: memcpy) ----
typedef uint64_t lb2_mbar;
static inline void lb2_mbar_init(lb2_mbar *, uint32_t) {}
static inline void lb2_async_fence() {}
static inline void lb2_bulk_g2s(void *dst, const void *src, uint32_t bytes, lb2_mbar *) { memcpy(dst, src, bytes); }
static inline void lb2_mbar_arrive(lb2_mbar *) {}
static inline void lb2_mbar_wait(lb2_mbar *, uint32_t) {}
// per-byte compares: 0xFF in every byte lane where the predicate holds
static inline uint32_t lb2_eq4(uint32_t w, uint32_t c4) { uint32_t r = 0; for (int b = 0; b < 4; ++b) { if (((w >> (8 * b)) & 0xFF) == ((c4 >> (8 * b)) & 0xFF)) { r |= 0xFFu << (8 * b); } } return r; }
static inline uint32_t lb2_ltu4(uint32_t w, uint32_t c4) { uint32_t r = 0; for (int b = 0; b < 4; ++b) { if (((w >> (8 * b)) & 0xFF) < ((c4 >> (8 * b)) & 0xFF)) { r |= 0xFFu << (8 * b); } } return r; }
#else
#define LB2_DEV   __device__ __forceinline__
#define LB2_DEVNI __device__ __noinline__
LB2_DEV unsigned lb2_tid()  { return threadIdx.x; }
LB2_DEV unsigned lb2_nthr() { return blockDim.x; }
LB2_DEV void     lb2_sync() { __syncthreads(); }
// lb2_* atomics act on SHARED memory (the Mer->Node table, window state): explicit .shared PTX, because a generic
// atomic whose address happens to be shared takes the slow path.  lb2g_* are the few global-memory atomics.
LB2_DEV uint32_t lb2_saddr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
LB2_DEV uint32_t lb2_cas32(uint32_t *p, uint32_t cmp, uint32_t val) { uint32_t o; asm volatile("atom.shared.cas.b32 %0, [%1], %2, %3;" : "=r"(o) : "r"(lb2_saddr(p)), "r"(cmp), "r"(val) : "memory"); return o; }
LB2_DEV uint32_t lb2_add32(uint32_t *p, uint32_t v) { uint32_t o; asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(o) : "r"(lb2_saddr(p)), "r"(v) : "memory"); return o; }
LB2_DEV void     lb2_max32(uint32_t *p, uint32_t v) { asm volatile("red.shared.max.u32 [%0], %1;" :: "r"(lb2_saddr(p)), "r"(v) : "memory"); }
LB2_DEV void     lb2_min32(uint32_t *p, uint32_t v) { asm volatile("red.shared.min.u32 [%0], %1;" :: "r"(lb2_saddr(p)), "r"(v) : "memory"); }
LB2_DEV void     lb2_or32 (uint32_t *p, uint32_t v) { asm volatile("red.shared.or.b32 [%0], %1;" :: "r"(lb2_saddr(p)), "r"(v) : "memory"); }
LB2_DEV uint32_t lb2_ld32(const uint32_t *p) { uint32_t o; asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(o) : "r"(lb2_saddr(p)) : "memory"); return o; }
// read-only shared data (packed bases, quality mask): plain ld.shared, free to be scheduled/merged by the compiler
LB2_DEV uint32_t lb2_lds(const uint32_t *p) { uint32_t o; asm("ld.shared.u32 %0, [%1];" : "=r"(o) : "r"(lb2_saddr(p))); return o; }
LB2_DEV uint32_t lb2g_add32(uint32_t *p, uint32_t v) { return atomicAdd(p, v); }
// fire-and-forget global reductions (no return value => the lane does not wait for L2)
LB2_DEV void lb2g_red_add(uint32_t *p, uint32_t v) { asm volatile("red.global.add.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory"); }
LB2_DEV void lb2g_red_max(uint32_t *p, uint32_t v) { asm volatile("red.global.max.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory"); }
LB2_DEV void lb2g_red_or(uint32_t *p, uint32_t v) { asm volatile("red.global.or.b32 [%0], %1;" :: "l"(p), "r"(v) : "memory"); }
LB2_DEV uint32_t lb2g_min32(uint32_t *p, uint32_t v) { return atomicMin(p, v); }
LB2_DEV int lb2_ctz64(uint64_t x) { return __ffsll((long long)x) - 1; }
LB2_DEV int lb2_clz32(uint32_t x) { return __clz((int)x); }
LB2_DEV int lb2_ctz32(uint32_t x) { return __ffs((int)x) - 1; }
LB2_DEV int lb2_popc32(uint32_t x) { return __popc(x); }
// length-weighted coverage average of compressNode (src/Graph.cc:2631-2636): two products, a sum, a quotient, each rounded
// (the reference is x86-64 without FMA; the device must not contract the sum of products)
LB2_DEV float lb2_wavg(float a, int la, float b, int lb) { return __fdiv_rn(__fadd_rn(__fmul_rn(a, (float)la), __fmul_rn(b, (float)lb)), (float)(la + lb)); }
// the same average for a serial fold: the correctly rounded reciprocal r = RN(1/(la+lb)) of the (integer-valued) divisor
// comes from the caller (computed in parallel beforehand), and the division is q0 = s*r followed by two fused residual
// corrections.  With r = RN(1/n) and q1 faithful, the last correction returns RN(s/n) (Markstein); tests/hostsim
// "divtest" checks the sequence against IEEE division.  s is a finite non-negative normal (coverage x length sums),
// n an integer < 2^16.  Only product, sum and the corrections depend on a.
LB2_DEV float lb2_rcp_int(uint32_t n) { return __frcp_rn((float)n); }
LB2_DEV float lb2_wavg_rcp(float a, int la, float b, int lb, float r) {
	const float n = (float)(la + lb), q = __fmul_rn(b, (float)lb);
	const float s = __fadd_rn(__fmul_rn(a, (float)la), q);
	const float q0 = __fmul_rn(s, r), e0 = __fmaf_rn(-n, q0, s), q1 = __fmaf_rn(e0, r, q0), e1 = __fmaf_rn(-n, q1, s);
	return __fmaf_rn(e1, r, q1);
}
LB2_DEV unsigned long long lb2_clock() { return (unsigned long long)clock64(); }
// CTA-wide exclusive prefix sum of one value per lane (sc: >= 34 words of shared scratch); every lane must call it
LB2_DEV uint32_t lb2_block_excl(uint32_t *sc, uint32_t v, uint32_t *total) {
	const unsigned lane = threadIdx.x & 31u, wid = threadIdx.x >> 5, nwarp = (blockDim.x + 31u) >> 5;
	uint32_t x = v;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, o); if (lane >= (unsigned)o) { x += y; } }
	if (lane == 31u) { sc[wid] = x; }
	__syncthreads();
	uint32_t base = 0, tot = 0;
	for (unsigned w = 0; w < nwarp; ++w) { uint32_t t = sc[w]; if (w < wid) { base += t; } tot += t; }
	__syncthreads();
	*total = tot;
	return base + x - v;
}
// atomics on an address that may be shared OR global (scratch that falls back to the workspace slab)
LB2_DEV uint32_t lb2x_exch32(uint32_t *p, uint32_t v) { return atomicExch(p, v); }
LB2_DEV uint32_t lb2x_cas32(uint32_t *p, uint32_t cmp, uint32_t v) { return atomicCAS(p, cmp, v); }
// ---- sub-warp groups for the read staging (8 lanes per read) ----
#define LB2_GS 8
// ---- shared-memory words addressed by a precomputed 32-bit shared-window address (no generic->shared conversion per access) ----
typedef uint32_t lb2_sp;
LB2_DEV lb2_sp lb2_sp_of(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
LB2_DEV lb2_sp lb2_sp_at(lb2_sp b, uint32_t i) { return b + 4u * i; }
LB2_DEV uint32_t lb2s_ldv(lb2_sp a) { uint32_t o; asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(o) : "r"(a) : "memory"); return o; }
LB2_DEV uint32_t lb2s_ld(lb2_sp a) { uint32_t o; asm("ld.shared.u32 %0, [%1];" : "=r"(o) : "r"(a)); return o; }
LB2_DEV uint32_t lb2s_cas(lb2_sp a, uint32_t cmp, uint32_t val) { uint32_t o; asm volatile("atom.shared.cas.b32 %0, [%1], %2, %3;" : "=r"(o) : "r"(a), "r"(cmp), "r"(val) : "memory"); return o; }
LB2_DEV void lb2s_min(lb2_sp a, uint32_t v) { asm volatile("red.shared.min.u32 [%0], %1;" :: "r"(a), "r"(v) : "memory"); }
LB2_DEV void lb2s_or(lb2_sp a, uint32_t v) { asm volatile("red.shared.or.b32 [%0], %1;" :: "r"(a), "r"(v) : "memory"); }
LB2_DEV uint32_t lb2_fsr(uint32_t lo, uint32_t hi, uint32_t sh) { return __funnelshift_r(lo, hi, sh); }
// work items handed out to whole warps (32 consecutive items per fetch): every lane of the warp calls this together
LB2_DEV uint32_t lb2_batch_next(uint32_t *ctr) {
	uint32_t base = 0;
	__syncwarp();
	if ((threadIdx.x & 31u) == 0) { base = lb2_add32(ctr, 32u); }
	return __shfl_sync(0xFFFFFFFFu, base, 0) + (threadIdx.x & 31u);
}
#define LB2_WARP 32
LB2_DEV uint32_t lb2_ballot(bool p) { return __ballot_sync(0xFFFFFFFFu, p); }
LB2_DEV void lb2_warp_sync() { __syncwarp(); }
LB2_DEV uint32_t lb2_warp_max(uint32_t v) { return __reduce_max_sync(0xFFFFFFFFu, v); }
LB2_DEV unsigned lb2_lane() { return threadIdx.x & 31u; }
LB2_DEV uint32_t lb2_match_any(uint32_t v) { return __match_any_sync(0xFFFFFFFFu, v); }      // lanes of the warp holding the same value
LB2_DEV uint32_t lb2_warp_excl(uint32_t v, uint32_t *total) {      // warp-wide exclusive prefix sum (all 32 lanes call it)
	uint32_t x = v;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, o); if ((threadIdx.x & 31u) >= (unsigned)o) { x += y; } }
	*total = __shfl_sync(0xFFFFFFFFu, x, 31);
	return x - v;
}
LB2_DEV uint32_t lb2_shfl(uint32_t v, uint32_t src) { return __shfl_sync(0xFFFFFFFFu, v, (int)src); }
LB2_DEV uint32_t lb2_shfl_up1(uint32_t v) { return __shfl_up_sync(0xFFFFFFFFu, v, 1); }
#define LB2_FQ 4      /* lanes per chain in the coverage fold of the parallel compaction: one per channel */
LB2_DEV unsigned lb2_glane() { return threadIdx.x & 7u; }
LB2_DEV unsigned lb2_group() { return threadIdx.x >> 3; }
LB2_DEV unsigned lb2_ngroups() { return blockDim.x >> 3; }
LB2_DEV uint32_t lb2_gmin(uint32_t v) { for (int o = 1; o < 8; o <<= 1) { uint32_t y = __shfl_xor_sync(0xFFFFFFFFu, v, o); v = y < v ? y : v; } return v; }
LB2_DEV uint32_t lb2_gmax(uint32_t v) { for (int o = 1; o < 8; o <<= 1) { uint32_t y = __shfl_xor_sync(0xFFFFFFFFu, v, o); v = y > v ? y : v; } return v; }
LB2_DEV uint32_t lb2_gor(uint32_t v) { for (int o = 1; o < 8; o <<= 1) { v |= __shfl_xor_sync(0xFFFFFFFFu, v, o); } return v; }
// 16 bytes from an arbitrarily aligned address: five aligned words, funnel-shifted (reads up to 3 bytes before and 7 after)
LB2_DEV void lb2_load16(const char *p, uint32_t o[4]) {
	const uint32_t *a = (const uint32_t *)((uintptr_t)p & ~(uintptr_t)3); const uint32_t sh = (uint32_t)((uintptr_t)p & 3u) * 8u;
	const uint32_t w0 = __ldg(a), w1 = __ldg(a + 1), w2 = __ldg(a + 2), w3 = __ldg(a + 3), w4 = __ldg(a + 4);
	o[0] = __funnelshift_r(w0, w1, sh); o[1] = __funnelshift_r(w1, w2, sh); o[2] = __funnelshift_r(w2, w3, sh); o[3] = __funnelshift_r(w3, w4, sh);
}
// ---- bulk-async staging: 1-D cp.async.bulk (the TMA engine's untiled copy) from global to this CTA's shared memory, 16-byte
// aligned on both sides, size a multiple of 16.  Completion is counted in bytes on an mbarrier: every lane announces the
// bytes of its own copies (expect_tx, no arrival), issues them, and arrives once when it has issued all of them; the phase
// completes when all lanes have arrived and every announced byte has landed.  Data written by a bulk copy is visible to
// ordinary loads of the lanes that observed the phase flip.
typedef uint64_t lb2_mbar;
LB2_DEV void lb2_mbar_init(lb2_mbar *b, uint32_t count) {
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(lb2_saddr(b)), "r"(count) : "memory");
	asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// orders this lane's earlier ordinary shared-memory accesses before its later bulk copies (generic -> async proxy)
LB2_DEV void lb2_async_fence() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
LB2_DEV void lb2_bulk_g2s(void *dst, const void *src, uint32_t bytes, lb2_mbar *b) {
	const uint32_t ba = lb2_saddr(b);
	asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" :: "r"(ba), "r"(bytes) : "memory");
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
	             :: "r"(lb2_saddr(dst)), "l"(__cvta_generic_to_global(src)), "r"(bytes), "r"(ba) : "memory");
}
LB2_DEV void lb2_mbar_arrive(lb2_mbar *b) { asm volatile("{ .reg .b64 st; mbarrier.arrive.shared::cta.b64 st, [%0]; }" :: "r"(lb2_saddr(b)) : "memory"); }
LB2_DEV void lb2_mbar_wait(lb2_mbar *b, uint32_t parity) {
	const uint32_t ba = lb2_saddr(b); uint32_t done = 0;
	while (!done) {
		asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(ba), "r"(parity) : "memory");
	}
}
LB2_DEV uint32_t lb2_eq4(uint32_t w, uint32_t c4) { return __vcmpeq4(w, c4); }
LB2_DEV uint32_t lb2_ltu4(uint32_t w, uint32_t c4) { return __vcmpltu4(w, c4); }
#endif

#endif
