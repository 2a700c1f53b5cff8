// lb2_cuda.cu -- the sm_100a kernels and the extern "C" boundary declared in include/lancet_b200.h.
//
// lb2_pack_*_kernel: the read pool is classified once (trim, 2-bit bases, quality mask; lb2_pack.cuh) -- a streaming pass.
// lb2_window_kernel: one persistent CTA per resident slot; each CTA pulls window indices from a global counter and runs
// the whole micro-assembly of that window (lb2_process_window) out of its own workspace slab, with the window's reads
// staged from the packed pool into shared memory by bulk-async copies.  No CPU fallback exists in this file: every
// entry point fails with LB2_ERR_CUDA when there is no usable device.
#include <cuda_runtime.h>
#include <nccl.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>
#include <algorithm>
#include <numeric>
#include <atomic>
#include <thread>
#include <chrono>

#include "lb2_pipeline.cuh"

#define LB2_KERNEL_VERSION "r02-v3"

struct lb2_launch {
	lb2_params P; lb2_cfg C; lb2_dev_batch B; lb2_dev_out O;
	uint8_t *ws_base; size_t ws_stride; uint32_t *counter;
	uint32_t w_begin, w_end;      // first pass: the windows of this launch (one launch per upload segment)
	const uint32_t *win_list; const uint32_t *n_list;   // escalation pass: indices of the windows to redo (NULL = [w_begin, w_end))
	uint32_t *retry_list; uint32_t *retry_count;
	// compaction outputs
	uint32_t *var_off; uint32_t *str_off; uint32_t *totals; lb2_variant *cvars; char *cstr;
};

__global__ void __launch_bounds__(256, 3)
lb2_window_kernel(const __grid_constant__ lb2_launch L)
{
	extern __shared__ __align__(16) uint8_t smem[];
	__shared__ uint32_t s_next;
	// the window descriptor (some 150 pointers, identical for every lane) lives in SHARED memory: as a kernel-local
	// struct handed by reference to the pipeline's functions it sat in per-thread local memory, which with three 72 KB
	// CTAs per SM has next to no L1 behind it -- every pointer fetch was an L2 round trip
	__shared__ lb2_win sW;
	// ... and so do the parameter blocks it points to (read inside lane-0 loops)
	__shared__ lb2_params sP; __shared__ lb2_cfg sC; __shared__ lb2_dev_batch sB; __shared__ lb2_dev_out sO;
	lb2_win &W = sW;
	if (threadIdx.x == 0) {
		sP = L.P; sC = L.C; sB = L.B; sO = L.O;
		W.P = &sP; W.C = &sC; W.B = &sB; W.O = &sO; W.escal = (L.win_list != nullptr);
		lb2_ws_layout(L.C, L.ws_base + (size_t)blockIdx.x * L.ws_stride, &W.ws); W.ws0 = W.ws;
		W.sh = (lb2_sh *)smem;
		W.ref_raw = (char *)smem + ((sizeof(lb2_sh) + 15) & ~(size_t)15);
		W.bits = (uint32_t *)(W.ref_raw + LB2_MAX_REF);
		W.lowq = W.bits + (L.C.max_bp / 16 + 4);
		W.treg = smem + ((lb2_smem_fixed(L.C.max_bp) + 15) & ~(size_t)15);
		lb2_mbar_init(&W.sh->mbar, blockDim.x); W.sh->mbar_phase = 0;
	}
	__syncthreads();
	const uint32_t nwin = L.win_list ? *L.n_list : (L.w_end - L.w_begin);
	while (true) {
		if (threadIdx.x == 0) { s_next = atomicAdd(L.counter, 1u); }
		__syncthreads();
		uint32_t w = s_next;
		__syncthreads();
		if (w >= nwin) { break; }
		w = L.win_list ? L.win_list[w] : L.w_begin + w;
		lb2_process_window(W, w);
	}
}

// ---- the pre-pack pass over pool reads [r0, r1), r0 a multiple of LB2_PACK_BLOCK: one block per LB2_PACK_BLOCK reads.
// A block's reads start at packed word floor(pool offset of its first read / 16) + index of that read: an upper bound of
// the words all earlier reads take (a read of len bases takes ceil(len/16) <= len/16 + 1 words), so blocks never overlap
// and every block finds its place from the pool offsets alone -- stretches of the pool can be packed in any order.
// 128 threads and 32 registers on purpose: one such block fits into what three resident window CTAs leave
// of an SM (4096 registers, 6.9 KB shared memory), so the pass over the next upload segment runs beside the window
// kernel of the current one instead of waiting for its CTAs to drain.
#define LB2_PACK_BLOCK 512
#define LB2_PACK_THREADS 128
__global__ void __launch_bounds__(LB2_PACK_THREADS, 16) lb2_pack_kernel(const lb2_dev_batch B, lb2_pkread *pk, uint32_t *pk_bits, uint16_t *pk_lowq, uint32_t qtrim4, uint32_t qcall4,
                                                                        uint32_t r0, uint32_t r1)
{
	__shared__ uint32_t sc[40]; __shared__ uint32_t s_w[LB2_PACK_BLOCK]; __shared__ uint32_t s_o[LB2_PACK_BLOCK + 1]; __shared__ uint8_t s_f[LB2_PACK_BLOCK];      // (offsets relative to the block's first read: 4.7 KB in all, see above)
	constexpr uint32_t PER = LB2_PACK_BLOCK / LB2_PACK_THREADS;
	const uint32_t rb = r0 + blockIdx.x * LB2_PACK_BLOCK, t = threadIdx.x;
	// the block's pool offsets and flag bytes first (coalesced), so that the per-read work below starts with the base loads
	const uint64_t o_first = B.base_off[rb];
	for (uint32_t j = t; j <= LB2_PACK_BLOCK; j += LB2_PACK_THREADS) { const uint32_t r = rb + j; s_o[j] = (uint32_t)(B.base_off[r <= r1 ? r : r1] - o_first); if (j < LB2_PACK_BLOCK) { s_f[j] = (r < r1) ? B.flags[r] : (uint8_t)0; } }
	__syncthreads();
	uint32_t sum = 0;
	for (uint32_t q = 0; q < PER; ++q) { const uint32_t j = t * PER + q; const uint32_t c = (rb + j < r1) ? lb2_pack_nwords(s_o[j + 1] - s_o[j]) : 0u; s_w[j] = c; sum += c; }
	uint32_t total = 0, ex = lb2_block_excl(sc, sum, &total) + (uint32_t)(o_first >> 4) + rb;
	for (uint32_t q = 0; q < PER; ++q) { const uint32_t c = s_w[t * PER + q]; s_w[t * PER + q] = ex; ex += c; }
	__syncthreads();
	// a group of LB2_GS lanes per read, neighbouring groups on neighbouring reads (every lane of a warp runs every round)
	for (uint32_t j = lb2_group(); j < LB2_PACK_BLOCK; j += lb2_ngroups()) {
		const uint32_t r = rb + j;
		lb2_pack_read(B, pk, pk_bits, pk_lowq, qtrim4, qcall4, r < r1, r, s_w[j], o_first + s_o[j], (uint64_t)(s_o[j + 1] - s_o[j]), s_f[j]);
	}
}

// windows that ran out of a per-CTA capacity in the first pass are redone by the escalation pass
// (same kernel, larger table / arena / BFS queue / staging area, one CTA per SM)
__global__ void lb2_collect_kernel(const __grid_constant__ lb2_launch L)
{
	const uint32_t n = L.B.n_windows;
	for (uint32_t w = blockIdx.x * blockDim.x + threadIdx.x; w < n; w += gridDim.x * blockDim.x) {
		lb2_window_info wi = L.O.info[w];
		if (wi.status == LB2_WIN_OVERFLOW) {
			uint32_t d = wi.detail;
			if (d == LB2_D_HASH_FULL || d == LB2_D_NODES || d == LB2_D_ARENA || d == LB2_D_QUEUE || d == LB2_D_SMEM || d == LB2_D_BUCKETS ||
			    d == LB2_D_STACK || d == LB2_D_EDGES || d == LB2_D_SPECIAL || d == LB2_D_READS ||
			    d == LB2_D_VARIANTS || d == LB2_D_STRINGS) {      // (escalated windows emit into the large output slabs)
				L.retry_list[atomicAdd(L.retry_count, 1u)] = w;
			}
		}
	}
}

// exclusive scans of per-window variant counts and string bytes (one block)
__global__ void lb2_scan_kernel(const __grid_constant__ lb2_launch L)
{
	__shared__ uint32_t pv[1024], ps[1024];
	const uint32_t n = L.B.n_windows, t = threadIdx.x, nt = blockDim.x;
	const uint32_t chunk = (n + nt - 1) / nt, lo = min(n, t * chunk), hi = min(n, lo + chunk);
	uint32_t sv = 0, ss = 0;
	for (uint32_t w = lo; w < hi; ++w) { sv += L.O.info[w].n_variants; ss += (L.O.info[w].n_variants ? L.O.str_used[w] : 0); }
	pv[t] = sv; ps[t] = ss; __syncthreads();
	if (t == 0) {
		uint32_t av = 0, as = 0;
		for (uint32_t i = 0; i < nt; ++i) { uint32_t v = pv[i], s = ps[i]; pv[i] = av; ps[i] = as; av += v; as += s; }
		L.totals[0] = av; L.totals[1] = as;
	}
	__syncthreads();
	uint32_t av = pv[t], as = ps[t];
	for (uint32_t w = lo; w < hi; ++w) {
		L.var_off[w] = av; L.str_off[w] = as;
		uint32_t nv = L.O.info[w].n_variants; av += nv; as += (nv ? L.O.str_used[w] : 0);
	}
}

// gather the per-window slabs into dense arrays (one block per window)
__global__ void lb2_gather_kernel(const __grid_constant__ lb2_launch L)
{
	const uint32_t w = blockIdx.x; const uint32_t nv = L.O.info[w].n_variants;
	if (!nv) { return; }
	const uint32_t vo = L.var_off[w], so = L.str_off[w], sb = L.O.str_used[w];
	const uint32_t big = L.O.big_slot[w];
	const lb2_variant *vsrc = (big != 0xFFFFFFFFu) ? L.O.big_variants + (size_t)big * L.O.big_max_var : L.O.variants + (size_t)w * L.C.max_var;
	for (uint32_t i = threadIdx.x; i < nv; i += blockDim.x) {
		lb2_variant v = vsrc[i]; v.str_off += so; L.cvars[vo + i] = v;
	}
	const char *src = (big != 0xFFFFFFFFu) ? L.O.big_strings + (size_t)big * L.O.big_str_bytes : L.O.strings + (size_t)w * L.C.str_bytes;
	for (uint32_t i = threadIdx.x; i < sb; i += blockDim.x) { L.cstr[so + i] = src[i]; }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
#define LB2_MAX_SEG 16
// an upload segment: windows [w0, w1) and the stretches of the pool (whole pack blocks) that they use and that are not on the device yet
struct lb2_seg { uint32_t w0, w1, nr; uint32_t r0[4], r1[4]; };
struct lb2_ctx {
	int device = 0; int sm_count = 0; uint32_t threads = 256; size_t smem_optin = 0;
	cudaStream_t stream = nullptr, copy_stream = nullptr, pack_stream = nullptr, wstream[2] = { nullptr, nullptr };
	cudaEvent_t ev[6] = { nullptr, nullptr, nullptr, nullptr, nullptr, nullptr };      // run: start, packed, first pass, escalation, compaction
	cudaEvent_t ev_ready = nullptr, ev_seg[LB2_MAX_SEG] = {}, ev_cp[LB2_MAX_SEG] = {}, ev_w[2] = { nullptr, nullptr };
	lb2_params P; lb2_cfg C;
	std::string err;
	struct Buf { void *p = nullptr; size_t cap = 0; };
	// device buffers of the batch: the caller's arrays, the packed pool, the outputs
	Buf d_ref_off, d_ref_start, d_wr_off, d_wr_idx, d_base_off, d_flags, d_name_rank, d_ref_seq, d_seq, d_qual;
	Buf d_pk, d_pk_bits, d_pk_lowq;
	Buf d_info, d_vars, d_strs, d_str_used, d_var_off, d_str_off, d_cvars, d_cstr, d_ws, d_big_vars, d_big_strs, d_big_slot, d_ws2, d_wsm, d_retry;
	uint32_t *d_counters = nullptr;      // [LB2_MAX_SEG] window counters of the first-pass launches, [LB2_MAX_SEG] escalation, +1 retry count, +2 big count, +3 pack carry, +4.. totals
	unsigned long long *d_prof = nullptr;
	uint32_t big_cap = 256, big_max_var = 1024, big_str_bytes = 128u << 10;
	lb2_launch L, L2, Lm;
	uint32_t n_windows = 0, n_reads = 0; bool resident = false, ran = false, escalate = true;
	size_t ws_stride = 0; uint32_t ws_slots = 0, ws_sets = 0; size_t ws2_stride = 0; uint32_t ws2_slots = 0;
	lb2_cfg C2, Cm; size_t wsm_stride = 0; uint32_t wsm_slots = 0; bool mid = true;      // Cm: the middle pass (two CTAs per SM)
	// per window: the one or two stretches of the pool its reads lie in, [a0,a1) [b0,b1) (a window's list is tumour reads then
	// normal reads, each ascending: two stretches far apart in a pool that holds all tumour reads before all normal ones)
	std::vector<uint32_t> h_rng;
	uint32_t plan_need_bp = 0, plan_max_reads = 0, smem_cap = 0, bp1_cached = 0, bp1_slots = 0;      // results of the window plan; cached first-pass staging size
	uint64_t launches = 0;
	float kernel_ms = 0;
	// host result
	std::vector<lb2_window_info> h_info; std::vector<lb2_variant> h_vars; std::vector<char> h_str;
	uint64_t h2d_bytes = 0, d2h_bytes = 0;
	// record gather across ranks (one process per GPU)
	ncclComm_t comm = nullptr; int comm_rank = 0, comm_world = 1; Buf d_comm_send, d_comm_recv, d_comm_cnt;
	std::vector<lb2_variant> g_vars; std::vector<char> g_str;
};
enum { LB2_CTR_RETRY = 2 * LB2_MAX_SEG, LB2_CTR_BIG, LB2_CTR_CARRY, LB2_CTR_TOTALS, LB2_CTR_N = LB2_CTR_TOTALS + 2 };

#define LB2_CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_); return LB2_ERR_CUDA; } } while (0)

static int lb2_reserve(lb2_ctx *ctx, lb2_ctx::Buf &b, size_t bytes)
{
	if (bytes <= b.cap && b.p) { return LB2_OK; }
	if (b.p) { cudaFree(b.p); b.p = nullptr; b.cap = 0; }
	size_t cap = bytes + bytes / 4 + 256;
	LB2_CK(cudaMalloc(&b.p, cap));
	b.cap = cap;
	return LB2_OK;
}

extern "C" void lb2_default_params(lb2_params *p)
{
	// reference src/Lancet.hh:33-79
	p->min_k = 11; p->max_k = 101; p->min_qual_trim = 10 + 33; p->min_qual_call = 17 + 33; p->cov_threshold = 5;
	p->low_cov_threshold = 1; p->max_tip_len = 11; p->dfs_limit = 1000000; p->max_indel_len = 500; p->max_mismatch = 2;
	p->max_unit_len = 4; p->min_report_units = 3; p->min_report_len = 7; p->dist_from_str = 1; p->min_cov_ratio = 0.01;
}

extern "C" const char *lb2_strerror(const lb2_ctx *ctx, int code)
{
	switch (code) {
		case LB2_OK: return "ok";
		case LB2_ERR_CUDA: return (ctx && !ctx->err.empty()) ? ctx->err.c_str() : "CUDA error / no sm_100 device (this library has no CPU fallback)";
		case LB2_ERR_ARG: return "invalid argument";
		case LB2_ERR_NOMEM: return "out of memory";
		case LB2_ERR_STATE: return "call sequence error (upload -> run -> download)";
	}
	return "unknown error";
}
extern "C" const char *lb2_kernel_version(void) { return LB2_KERNEL_VERSION; }

static uint32_t env_u32(const char *name, uint32_t dflt) { const char *s = getenv(name); return s ? (uint32_t)strtoul(s, nullptr, 10) : dflt; }

static void lb2_comm_release(ncclComm_t c);
extern "C" void lb2_destroy(lb2_ctx *ctx)
{
	if (!ctx) { return; }
	cudaSetDevice(ctx->device);
	lb2_ctx::Buf *bufs[] = { &ctx->d_ref_off, &ctx->d_ref_start, &ctx->d_wr_off, &ctx->d_wr_idx, &ctx->d_base_off, &ctx->d_flags, &ctx->d_name_rank,
		&ctx->d_ref_seq, &ctx->d_seq, &ctx->d_qual, &ctx->d_pk, &ctx->d_pk_bits, &ctx->d_pk_lowq, &ctx->d_info, &ctx->d_vars, &ctx->d_strs, &ctx->d_str_used,
		&ctx->d_var_off, &ctx->d_str_off, &ctx->d_cvars, &ctx->d_cstr, &ctx->d_ws, &ctx->d_big_vars, &ctx->d_big_strs, &ctx->d_big_slot, &ctx->d_ws2, &ctx->d_wsm, &ctx->d_retry };
	for (auto b : bufs) { if (b->p) { cudaFree(b->p); } }
	if (ctx->comm) { lb2_comm_release(ctx->comm); }
	{ lb2_ctx::Buf *cb[] = { &ctx->d_comm_send, &ctx->d_comm_recv, &ctx->d_comm_cnt }; for (auto b : cb) { if (b->p) { cudaFree(b->p); } } }
	if (ctx->d_counters) { cudaFree(ctx->d_counters); } if (ctx->d_prof) { cudaFree(ctx->d_prof); }
	for (auto &e : ctx->ev) { if (e) { cudaEventDestroy(e); } } for (auto &e : ctx->ev_seg) { if (e) { cudaEventDestroy(e); } } for (auto &e : ctx->ev_cp) { if (e) { cudaEventDestroy(e); } } for (auto &e : ctx->ev_w) { if (e) { cudaEventDestroy(e); } }
	if (ctx->ev_ready) { cudaEventDestroy(ctx->ev_ready); }
	if (ctx->stream) { cudaStreamDestroy(ctx->stream); } if (ctx->copy_stream) { cudaStreamDestroy(ctx->copy_stream); } if (ctx->pack_stream) { cudaStreamDestroy(ctx->pack_stream); }
	for (auto &st : ctx->wstream) { if (st) { cudaStreamDestroy(st); } }
	cudaGetLastError();
	delete ctx;
}

extern "C" int lb2_create(lb2_ctx **out, const lb2_params *params, int device)
{
	if (!out || !params) { return LB2_ERR_ARG; }
	*out = nullptr;
	int ndev = 0;
	if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) { cudaGetLastError(); return LB2_ERR_CUDA; }
	lb2_ctx *ctx = new lb2_ctx();
	ctx->device = device; ctx->P = *params;
	auto fail = [&]() { lb2_destroy(ctx); return LB2_ERR_CUDA; };      // (frees whatever the half-built context holds)
	cudaDeviceProp prop;
	if (cudaSetDevice(device) != cudaSuccess || cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major < 10) { return fail(); }
	ctx->sm_count = prop.multiProcessorCount; ctx->smem_optin = prop.sharedMemPerBlockOptin;
	cudaStream_t *streams[] = { &ctx->stream, &ctx->copy_stream, &ctx->pack_stream, &ctx->wstream[0], &ctx->wstream[1] };
	for (auto st : streams) { if (cudaStreamCreateWithFlags(st, cudaStreamNonBlocking) != cudaSuccess) { return fail(); } }
	for (auto &e : ctx->ev) { if (cudaEventCreate(&e) != cudaSuccess) { return fail(); } }
	for (auto &e : ctx->ev_seg) { if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) { return fail(); } }
	for (auto &e : ctx->ev_cp) { if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) { return fail(); } }
	for (auto &e : ctx->ev_w) { if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) { return fail(); } }
	if (cudaEventCreateWithFlags(&ctx->ev_ready, cudaEventDisableTiming) != cudaSuccess) { return fail(); }
	lb2_cfg &C = ctx->C; memset(&C, 0, sizeof C);
	C.table_slots = env_u32("LB2_TABLE_SLOTS", 4096); C.max_nodes = C.table_slots - C.table_slots / 4;
	C.max_reads = 4096; C.max_bp = 0;
	C.arena_bytes = env_u32("LB2_ARENA_BYTES", 512u << 10); C.deficit_bytes = env_u32("LB2_DEFICIT_BYTES", 1u << 20);
	C.max_inst = env_u32("LB2_MAX_INST", 1u << 18);
	C.queue_cap = env_u32("LB2_QUEUE_CAP", 1u << 16); C.max_var = env_u32("LB2_MAX_VAR", 32); C.str_bytes = env_u32("LB2_STR_BYTES", 4096);
	C.debug_flags = env_u32("LB2_DEBUG_FLAGS", 0); C.max_special = env_u32("LB2_MAX_SPECIAL", 128); C.bucket_cap = 10273; C.max_k = 127; C.graph_bytes = env_u32("LB2_GRAPH_BYTES", 48u << 10);
	if (cudaMalloc(&ctx->d_prof, 24 * 8) != cudaSuccess || cudaMemset(ctx->d_prof, 0, 24 * 8) != cudaSuccess) { return fail(); }
	if (cudaMalloc(&ctx->d_counters, sizeof(uint32_t) * LB2_CTR_N) != cudaSuccess || cudaMemset(ctx->d_counters, 0, sizeof(uint32_t) * LB2_CTR_N) != cudaSuccess) { return fail(); }
	ctx->escalate = env_u32("LB2_ESCALATE", 1) != 0; ctx->mid = env_u32("LB2_MID", 1) != 0;
	ctx->big_cap = env_u32("LB2_BIG_SLABS", 256); ctx->big_max_var = env_u32("LB2_BIG_MAX_VAR", 1024); ctx->big_str_bytes = env_u32("LB2_BIG_STR_BYTES", 128u << 10);
	ctx->threads = env_u32("LB2_THREADS", 256); if (ctx->threads < 32 || ctx->threads > 256 || (ctx->threads & 31)) { ctx->threads = 256; }
	*out = ctx;
	return LB2_OK;
}

extern "C" uint64_t lb2_kernel_launches(const lb2_ctx *ctx) { return ctx ? ctx->launches : 0; }

// ---- planning.  lb2_prepare_static: launch configuration of the first pass, device buffers, launch descriptor -- from the
// batch header alone.  lb2_plan_windows: one pass over the windows' read lists (millions of entries for a 1 Mb region,
// split over host threads): the one or two stretches of the pool every window's reads lie in, the deepest window's staging
// need and read count (sizes of the escalation passes).  lb2_prepare_escalation: the configurations of the middle and the
// last pass.  lb2_process runs the window pass beside the upload of its first segment.
struct lb2_win_plan { uint32_t rng[4]; uint32_t runs; uint32_t nreads; bool bad; };
// (only the index list is walked: the staging need of a window is bounded from its read count, its runs of pool-consecutive
// reads and the longest read of the pool -- exact when all reads have the same length)
static inline lb2_win_plan lb2_plan_one(const lb2_batch *b, uint32_t w)
{
	lb2_win_plan p; p.bad = false; p.runs = 0; p.nreads = 0; p.rng[0] = p.rng[1] = p.rng[2] = p.rng[3] = 0;
	const uint32_t R = b->n_reads;
	uint32_t top = 0, low = 0xFFFFFFFFu, prev = 0xFFFFFFFEu, gap = 0, gap_lo = 0, gap_hi = 0; bool ascending = true;
	if (b->wr_off[w + 1] < b->wr_off[w] || b->wr_off[w + 1] > b->n_wr || b->ref_off[w + 1] < b->ref_off[w]) { p.bad = true; return p; }
	for (uint32_t x = b->wr_off[w]; x < b->wr_off[w + 1]; ++x) {
		const uint32_t r = b->wr_idx[x]; if (r >= R) { p.bad = true; return p; }
		if (r != prev + 1u) { ++p.runs; }
		if (prev != 0xFFFFFFFEu) { if (r < prev) { ascending = false; } else if (r - prev > gap) { gap = r - prev; gap_lo = prev + 1; gap_hi = r; } }
		prev = r;
		if (r >= top) { top = r + 1; } if (r < low) { low = r; }
	}
	// one stretch [low, top), or two around the largest jump of an ascending list when that jump is worth it (a window's list is
	// tumour reads then normal reads, each ascending: two stretches far apart in a pool that holds all tumour reads first)
	if (top) {
		if (ascending && gap > 8u * LB2_PACK_BLOCK) { p.rng[0] = low; p.rng[1] = gap_lo; p.rng[2] = gap_hi; p.rng[3] = top; }
		else { p.rng[0] = low; p.rng[1] = top; }
	}
	p.nreads = b->wr_off[w + 1] - b->wr_off[w];
	return p;
}

static int lb2_plan_windows(lb2_ctx *ctx, const lb2_batch *b)
{
	const uint32_t W = b->n_windows;
	uint32_t max_bp = 0, max_reads = 0;
	const unsigned hw = std::thread::hardware_concurrency();
	const unsigned share = std::max(1u, hw / (unsigned)std::max(1, ctx->comm_world));      // (several ranks on one host share its cores)
	const unsigned T = (W >= 2048 && hw > 1) ? std::min<unsigned>(std::min<unsigned>(share, 16u), std::max(1u, env_u32("LB2_HOST_THREADS", 16))) : 1u;
	std::vector<uint32_t> t_bp(T, 0), t_rd(T, 0); std::vector<int> t_bad(T, 0);
	std::vector<uint32_t> t_mw(T, 0);      // longest read of the pool, in 16-base words
	auto work = [&](unsigned t) {
		{ const uint32_t R_ = b->n_reads, r0 = (uint32_t)((uint64_t)R_ * t / T), r1 = (uint32_t)((uint64_t)R_ * (t + 1) / T); uint32_t mw = 0;
		  for (uint32_t r = r0; r < r1; ++r) { mw = std::max(mw, lb2_pack_nwords(b->base_off[r + 1] - b->base_off[r])); } t_mw[t] = mw; }
		const uint32_t w0 = (uint32_t)((uint64_t)W * t / T), w1 = (uint32_t)((uint64_t)W * (t + 1) / T);
		uint32_t mbp = 0, mrd = 0;      // (mbp: reads + 14 pad words per run, still without the per-read words)
		for (uint32_t w = w0; w < w1; ++w) {
			const lb2_win_plan p = lb2_plan_one(b, w);
			if (p.bad) { t_bad[t] = 1; return; }
			uint32_t *rg = ctx->h_rng.data() + 4 * (size_t)w; rg[0] = p.rng[0]; rg[1] = p.rng[1]; rg[2] = p.rng[2]; rg[3] = p.rng[3];
			mbp = std::max(mbp, (((b->ref_off[w + 1] - b->ref_off[w]) + 31u) & ~31u) + 128u + 224u * p.runs);
			mrd = std::max(mrd, p.nreads);
		}
		t_bp[t] = mbp; t_rd[t] = mrd;
	};
	if (T == 1) { work(0); }
	else { std::vector<std::thread> th; for (unsigned t = 0; t < T; ++t) { th.emplace_back(work, t); } for (auto &x : th) { x.join(); } }
	uint32_t maxw = 0;
	for (unsigned t = 0; t < T; ++t) { if (t_bad[t]) { return LB2_ERR_ARG; } max_bp = std::max(max_bp, t_bp[t]); max_reads = std::max(max_reads, t_rd[t]); maxw = std::max(maxw, t_mw[t]); }
	const uint64_t bound = (uint64_t)max_bp + (uint64_t)max_reads * maxw * 16u;      // (deepest window x longest read: an upper bound of every window's need)
	ctx->plan_need_bp = (uint32_t)std::min<uint64_t>((bound + 1023) & ~(uint64_t)1023, (1u << 20) - 1024);      // (the first-occurrence index in a table key has 20 bits)
	ctx->plan_max_reads = max_reads;
	return LB2_OK;
}

static int lb2_prepare_static(lb2_ctx *ctx, const lb2_batch *b, bool two_sets)
{
	if (!ctx || !b) { return LB2_ERR_ARG; }
	if (cudaSetDevice(ctx->device) != cudaSuccess) { return LB2_ERR_CUDA; }
	ctx->resident = false; ctx->ran = false;
	cudaStreamSynchronize(ctx->stream);      // (work of an earlier batch may still read the buffers re-planned below)
	const uint32_t W = b->n_windows, R = b->n_reads;
	if ((W && (!b->ref_off || !b->ref_start || !b->wr_off)) || (R && (!b->base_off || !b->flags || !b->name_rank)) || (b->n_wr && !b->wr_idx)) { return LB2_ERR_ARG; }
	if (b->n_base_bytes / 16 + R > 0xFFFFFF00ull) { ctx->err = "read pool too large for 32-bit word offsets: split the batch"; return LB2_ERR_ARG; }
	ctx->h_rng.resize(4 * (size_t)W);
	const uint32_t smem_cap = (uint32_t)std::min<size_t>(ctx->smem_optin ? ctx->smem_optin : (227u << 10), 227u << 10) - 2048u;      // (static shared memory of the kernel comes on top)
	ctx->smem_cap = smem_cap;
	lb2_cfg &C = ctx->C;
	C.table_slots = std::min<uint32_t>(env_u32("LB2_TABLE_SLOTS", 4096), 16384u);      // (occurrence words hold 14-bit slot numbers)
	// First pass: the same configuration for every batch -- the largest staging area that still lets LB2_MIN_CTAS_PER_SM CTAs
	// share an SM (it doubles as the scratch of the graph stage: parallel first compaction, alignment rows, BFS queue head
	// want 18-21 KB whatever the depth of the windows), room for 4096 reads.  A window that needs more goes to the
	// escalation passes, which are sized from the plan (one outlier must not cost every window its occupancy).
	uint32_t bp1 = env_u32("LB2_MIN_BP", 94208u);
	LB2_CK(cudaFuncSetAttribute(lb2_window_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cap));
	if (ctx->bp1_cached && ctx->bp1_slots == C.table_slots) { bp1 = ctx->bp1_cached; }
	else {
		const uint32_t want_occ = std::max(1u, env_u32("LB2_MIN_CTAS_PER_SM", 3));
		while (bp1 > 32768u) {
			int occ = 0; const size_t sm = lb2_smem_bytes(bp1, C.table_slots, C.graph_bytes);
			if (sm <= smem_cap && cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, lb2_window_kernel, (int)ctx->threads, sm) == cudaSuccess && (uint32_t)occ >= want_occ) { break; }
			bp1 -= 1024;
		}
		cudaGetLastError();
		while (C.table_slots > 1024 && lb2_smem_bytes(bp1, C.table_slots, C.graph_bytes) > smem_cap) { C.table_slots >>= 1; }
		ctx->bp1_cached = bp1; ctx->bp1_slots = C.table_slots;
	}
	C.max_nodes = C.table_slots - C.table_slots / 4;
	C.max_bp = bp1; C.smem_bytes = (uint32_t)lb2_smem_bytes(bp1, C.table_slots, C.graph_bytes); C.max_reads = env_u32("LB2_MAX_READS", 4096) + 2;
	int occ = 0;
	LB2_CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, lb2_window_kernel, (int)ctx->threads, C.smem_bytes));
	if (occ < 1) { ctx->err = "kernel does not fit on an SM"; return LB2_ERR_CUDA; }
	C.n_slots = (uint32_t)ctx->sm_count * std::min<uint32_t>((uint32_t)occ, env_u32("LB2_MAX_CTAS_PER_SM", 16));
	const size_t stride = lb2_ws_layout(C, nullptr, nullptr); const uint32_t sets = two_sets ? 2u : 1u;
	if (stride != ctx->ws_stride || C.n_slots > ctx->ws_slots || sets > ctx->ws_sets) {
		if (ctx->d_ws.p) { cudaFree(ctx->d_ws.p); ctx->d_ws.p = nullptr; ctx->d_ws.cap = 0; }
		const size_t bytes = stride * C.n_slots * sets;
		LB2_CK(cudaMalloc(&ctx->d_ws.p, bytes)); ctx->d_ws.cap = bytes;
		ctx->ws_stride = stride; ctx->ws_slots = C.n_slots; ctx->ws_sets = sets;
	}
	int rc;
#define LB2_RS(buf, bytes) do { if ((rc = lb2_reserve(ctx, ctx->buf, (bytes)))) { return rc; } } while (0)
	LB2_RS(d_ref_off, sizeof(uint32_t) * (size_t)(W + 1)); LB2_RS(d_ref_start, sizeof(int32_t) * (size_t)W + 4); LB2_RS(d_wr_off, sizeof(uint32_t) * (size_t)(W + 1));
	LB2_RS(d_wr_idx, sizeof(uint32_t) * (size_t)b->n_wr + 4); LB2_RS(d_base_off, sizeof(uint64_t) * (size_t)(R + 1)); LB2_RS(d_flags, (size_t)R + 4);
	LB2_RS(d_name_rank, sizeof(uint32_t) * (size_t)R + 4); LB2_RS(d_ref_seq, (size_t)b->n_ref_bytes + 4); LB2_RS(d_seq, (size_t)b->n_base_bytes + 64); LB2_RS(d_qual, (size_t)b->n_base_bytes + 64);
	const size_t pk_words = (size_t)(b->n_base_bytes / 16) + R + 64;
	LB2_RS(d_pk, sizeof(lb2_pkread) * ((size_t)R + 1)); LB2_RS(d_pk_bits, 4 * pk_words); LB2_RS(d_pk_lowq, 2 * pk_words);
	LB2_RS(d_info, sizeof(lb2_window_info) * (size_t)W + 16); LB2_RS(d_vars, sizeof(lb2_variant) * (size_t)W * C.max_var + 64); LB2_RS(d_strs, (size_t)W * C.str_bytes + 64);
	LB2_RS(d_str_used, sizeof(uint32_t) * (size_t)W + 4); LB2_RS(d_var_off, sizeof(uint32_t) * (size_t)W + 4); LB2_RS(d_str_off, sizeof(uint32_t) * (size_t)W + 4);
	const uint32_t nbig = ctx->escalate ? std::min<uint32_t>(ctx->big_cap, std::max(W, 1u)) : 0u;
	LB2_RS(d_cvars, sizeof(lb2_variant) * ((size_t)W * C.max_var + (size_t)nbig * ctx->big_max_var) + 64); LB2_RS(d_cstr, (size_t)W * C.str_bytes + (size_t)nbig * ctx->big_str_bytes + 64);
	LB2_RS(d_big_vars, sizeof(lb2_variant) * (size_t)nbig * ctx->big_max_var + 64); LB2_RS(d_big_strs, (size_t)nbig * ctx->big_str_bytes + 64);
	LB2_RS(d_big_slot, sizeof(uint32_t) * (size_t)(W + 1)); LB2_RS(d_retry, sizeof(uint32_t) * (size_t)(W + 1));
#undef LB2_RS
	lb2_launch &L = ctx->L; memset(&L, 0, sizeof L);
	L.P = ctx->P; L.C = C;
	L.B.n_windows = W; L.B.ref_off = (const uint32_t *)ctx->d_ref_off.p; L.B.ref_start = (const int32_t *)ctx->d_ref_start.p;
	L.B.wr_off = (const uint32_t *)ctx->d_wr_off.p; L.B.wr_idx = (const uint32_t *)ctx->d_wr_idx.p;
	L.B.base_off = (const uint64_t *)ctx->d_base_off.p; L.B.flags = (const uint8_t *)ctx->d_flags.p;
	L.B.name_rank = (const uint32_t *)ctx->d_name_rank.p; L.B.ref_seq = (const char *)ctx->d_ref_seq.p;
	L.B.seq = (const char *)ctx->d_seq.p; L.B.qual = (const char *)ctx->d_qual.p;
	L.B.pk = (const lb2_pkread *)ctx->d_pk.p; L.B.pk_bits = (const uint32_t *)ctx->d_pk_bits.p; L.B.pk_lowq = (const uint16_t *)ctx->d_pk_lowq.p;
	L.O.info = (lb2_window_info *)ctx->d_info.p; L.O.variants = (lb2_variant *)ctx->d_vars.p; L.O.strings = (char *)ctx->d_strs.p;
	L.O.str_used = (uint32_t *)ctx->d_str_used.p; L.O.prof = ctx->d_prof;
	L.O.big_variants = (lb2_variant *)ctx->d_big_vars.p; L.O.big_strings = (char *)ctx->d_big_strs.p; L.O.big_slot = (uint32_t *)ctx->d_big_slot.p;
	L.O.big_count = ctx->d_counters + LB2_CTR_BIG; L.O.big_cap = nbig; L.O.big_max_var = ctx->big_max_var; L.O.big_str_bytes = ctx->big_str_bytes;
	L.ws_base = (uint8_t *)ctx->d_ws.p; L.ws_stride = ctx->ws_stride; L.counter = ctx->d_counters;
	L.w_begin = 0; L.w_end = W; L.win_list = nullptr; L.n_list = nullptr;
	L.retry_list = (uint32_t *)ctx->d_retry.p; L.retry_count = ctx->d_counters + LB2_CTR_RETRY;
	L.var_off = (uint32_t *)ctx->d_var_off.p; L.str_off = (uint32_t *)ctx->d_str_off.p; L.totals = ctx->d_counters + LB2_CTR_TOTALS;
	L.cvars = (lb2_variant *)ctx->d_cvars.p; L.cstr = (char *)ctx->d_cstr.p;
	ctx->n_windows = W; ctx->n_reads = R;
	return LB2_OK;
}

static int lb2_prepare_escalation(lb2_ctx *ctx)
{
	const lb2_cfg &C = ctx->C; const lb2_launch &L = ctx->L; const uint32_t smem_cap = ctx->smem_cap, bp1 = C.max_bp, need_bp = ctx->plan_need_bp;
	if (ctx->escalate) {
		// escalation pass: one CTA per SM, the largest staging area / table / graph region that fit, big arena / BFS queue
		lb2_cfg &C2 = ctx->C2; C2 = C; C2.max_reads = std::max(C.max_reads, ctx->plan_max_reads + 2);
		C2.table_slots = std::min<uint32_t>(env_u32("LB2_TABLE_SLOTS2", 16384), 16384u); C2.graph_bytes = env_u32("LB2_GRAPH_BYTES2", 184u << 10);
		C2.max_bp = std::max(need_bp, bp1);
		while (C2.graph_bytes > C.graph_bytes && lb2_smem_bytes(C2.max_bp, C2.table_slots, C2.graph_bytes) > smem_cap) { C2.graph_bytes -= 4096; }
		while (C2.table_slots > C.table_slots && lb2_smem_bytes(C2.max_bp, C2.table_slots, C2.graph_bytes) > smem_cap) { C2.table_slots >>= 1; }
		while (C2.max_bp > bp1 && lb2_smem_bytes(C2.max_bp, C2.table_slots, C2.graph_bytes) > smem_cap) { C2.max_bp -= 1024; }      // windows beyond this report LB2_WIN_OVERFLOW
		C2.max_nodes = C2.table_slots - C2.table_slots / 4;
		C2.smem_bytes = (uint32_t)lb2_smem_bytes(C2.max_bp, C2.table_slots, C2.graph_bytes);
		C2.arena_bytes = env_u32("LB2_ARENA_BYTES2", 8u << 20); C2.deficit_bytes = env_u32("LB2_DEFICIT_BYTES2", 16u << 20);
		C2.queue_cap = env_u32("LB2_QUEUE_CAP2", 1u << 23);      /* DFS_LIMIT visits x a few children each */ C2.max_inst = env_u32("LB2_MAX_INST2", 1u << 20); C2.max_special = env_u32("LB2_MAX_SPECIAL2", 2048);
		C2.n_slots = std::min<uint32_t>((uint32_t)ctx->sm_count, env_u32("LB2_SLOTS2", 1024));
		// (the workspaces of the escalation passes -- gigabytes -- are allocated when a batch first needs them: lb2_enqueue_finish)
		const size_t stride2 = lb2_ws_layout(C2, nullptr, nullptr);
		if (stride2 != ctx->ws2_stride || C2.n_slots > ctx->ws2_slots) { if (ctx->d_ws2.p) { cudaFree(ctx->d_ws2.p); ctx->d_ws2.p = nullptr; ctx->d_ws2.cap = 0; } ctx->ws2_stride = stride2; ctx->ws2_slots = C2.n_slots; }
		lb2_launch &L2 = ctx->L2; L2 = L; L2.C = C2;
		L2.ws_base = nullptr; L2.ws_stride = stride2; L2.counter = ctx->d_counters + LB2_MAX_SEG;
		L2.win_list = (const uint32_t *)ctx->d_retry.p; L2.n_list = ctx->d_counters + LB2_CTR_RETRY;
		if (ctx->mid) {
			// middle pass: windows with a few thousand k-mers (sequencing errors, low-complexity sequence) outgrow the first
			// pass's table but do not need a whole SM -- twice the table, a larger graph region, two CTAs per SM
			lb2_cfg &Cm = ctx->Cm; Cm = C;
			Cm.table_slots = std::min<uint32_t>(2 * C.table_slots, 16384u); Cm.graph_bytes = env_u32("LB2_GRAPH_BYTES_MID", 80u << 10); Cm.max_bp = bp1;
			while (Cm.graph_bytes > C.graph_bytes) {
				int occ_m = 0; const size_t sm = lb2_smem_bytes(Cm.max_bp, Cm.table_slots, Cm.graph_bytes);
				if (sm <= smem_cap && cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_m, lb2_window_kernel, (int)ctx->threads, sm) == cudaSuccess && occ_m >= 2) { break; }
				Cm.graph_bytes -= 4096;
			}
			cudaGetLastError();
			Cm.max_nodes = Cm.table_slots - Cm.table_slots / 4; Cm.smem_bytes = (uint32_t)lb2_smem_bytes(Cm.max_bp, Cm.table_slots, Cm.graph_bytes);
			Cm.arena_bytes = 2u << 20; Cm.deficit_bytes = 4u << 20; Cm.queue_cap = 1u << 20; Cm.max_inst = 1u << 19; Cm.max_special = 2048;
			int occ_m = 1; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_m, lb2_window_kernel, (int)ctx->threads, Cm.smem_bytes); cudaGetLastError();
			Cm.n_slots = (uint32_t)ctx->sm_count * (uint32_t)std::max(1, std::min(occ_m, 2));
			const size_t stridem = lb2_ws_layout(Cm, nullptr, nullptr);
			if (stridem != ctx->wsm_stride || Cm.n_slots > ctx->wsm_slots) { if (ctx->d_wsm.p) { cudaFree(ctx->d_wsm.p); ctx->d_wsm.p = nullptr; ctx->d_wsm.cap = 0; } ctx->wsm_stride = stridem; ctx->wsm_slots = Cm.n_slots; }
			lb2_launch &Lm = ctx->Lm; Lm = L2; Lm.C = Cm;
			Lm.ws_base = nullptr; Lm.ws_stride = stridem; Lm.counter = ctx->d_counters + LB2_MAX_SEG + 1;
		}
	}
	return LB2_OK;
}

// enqueue the pre-pack pass over pool reads [r0, r1) on stream st
static int lb2_enqueue_pack(lb2_ctx *ctx, uint32_t r0, uint32_t r1, cudaStream_t st)
{
	if (r1 <= r0) { return LB2_OK; }
	if (r0 % LB2_PACK_BLOCK) { return LB2_ERR_ARG; }
	const uint32_t nblk = (r1 - r0 + LB2_PACK_BLOCK - 1) / LB2_PACK_BLOCK;
	const uint32_t qt = (uint32_t)ctx->P.min_qual_trim & 0xFFu, qc = (uint32_t)ctx->P.min_qual_call & 0xFFu;
	lb2_pack_kernel<<<nblk, LB2_PACK_THREADS, 0, st>>>(ctx->L.B, (lb2_pkread *)ctx->d_pk.p, (uint32_t *)ctx->d_pk_bits.p, (uint16_t *)ctx->d_pk_lowq.p, qt * 0x01010101u, qc * 0x01010101u, r0, r1);
	ctx->launches += 1;
	LB2_CK(cudaGetLastError());
	return LB2_OK;
}

// enqueue one first-pass launch over windows [w0, w1) on stream st, out of workspace set `set`, counted on counter `ci`
static int lb2_enqueue_windows(lb2_ctx *ctx, uint32_t w0, uint32_t w1, uint32_t set, uint32_t ci, cudaStream_t st)
{
	if (w1 <= w0) { return LB2_OK; }
	lb2_launch L = ctx->L; L.w_begin = w0; L.w_end = w1; L.counter = ctx->d_counters + ci;
	L.ws_base = (uint8_t *)ctx->d_ws.p + (size_t)set * ctx->ws_stride * ctx->C.n_slots;
	const uint32_t grid = std::min<uint32_t>(ctx->C.n_slots, w1 - w0);
	lb2_window_kernel<<<grid, ctx->threads, ctx->C.smem_bytes, st>>>(L);
	ctx->launches += 1;
	LB2_CK(cudaGetLastError());
	return LB2_OK;
}

// collect the windows the first pass could not hold, redo them, compact the record slabs (stream st)
static int lb2_enqueue_finish(lb2_ctx *ctx, cudaStream_t st, cudaEvent_t after_escalation)
{
	const uint32_t W = ctx->n_windows;
	if (ctx->escalate) {
		// how many windows the pass before could not hold (one 4-byte read-back per pass: the escalation workspaces are only
		// allocated, and the passes only launched, for batches that need them)
		auto pending = [&](uint32_t *n) -> int {
			lb2_collect_kernel<<<64, 256, 0, st>>>(ctx->L); ctx->launches += 1;
			LB2_CK(cudaMemcpyAsync(n, ctx->d_counters + LB2_CTR_RETRY, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
			LB2_CK(cudaStreamSynchronize(st));
			return LB2_OK;
		};
		auto ensure = [&](lb2_ctx::Buf &buf, size_t stride, uint32_t slots, lb2_launch &Lx) -> int {
			if (!buf.p) { LB2_CK(cudaMalloc(&buf.p, stride * slots)); buf.cap = stride * slots; }
			Lx.ws_base = (uint8_t *)buf.p;
			return LB2_OK;
		};
		uint32_t n_retry = 0; int rc = pending(&n_retry); if (rc) { return rc; }
		if (n_retry && ctx->mid) {      // middle pass first; what it cannot hold either is collected again for the last pass
			if ((rc = ensure(ctx->d_wsm, ctx->wsm_stride, ctx->wsm_slots, ctx->Lm))) { return rc; }
			lb2_window_kernel<<<std::min(ctx->Cm.n_slots, n_retry), ctx->threads, ctx->Cm.smem_bytes, st>>>(ctx->Lm); ctx->launches += 1;
			LB2_CK(cudaMemsetAsync(ctx->d_counters + LB2_CTR_RETRY, 0, sizeof(uint32_t), st));
			if ((rc = pending(&n_retry))) { return rc; }
		}
		if (n_retry) {
			if ((rc = ensure(ctx->d_ws2, ctx->ws2_stride, ctx->ws2_slots, ctx->L2))) { return rc; }
			lb2_window_kernel<<<std::min(ctx->C2.n_slots, n_retry), ctx->threads, ctx->C2.smem_bytes, st>>>(ctx->L2); ctx->launches += 1;
		}
	}
	if (after_escalation) { LB2_CK(cudaEventRecord(after_escalation, st)); }
	lb2_scan_kernel<<<1, 1024, 0, st>>>(ctx->L);
	lb2_gather_kernel<<<W, 64, 0, st>>>(ctx->L);
	ctx->launches += 2;
	LB2_CK(cudaGetLastError());
	return LB2_OK;
}

static int lb2_enqueue_reset(lb2_ctx *ctx, cudaStream_t st)
{
	LB2_CK(cudaMemsetAsync(ctx->d_counters, 0, sizeof(uint32_t) * LB2_CTR_N, st));
	LB2_CK(cudaMemsetAsync(ctx->d_big_slot.p, 0xFF, sizeof(uint32_t) * (size_t)(ctx->n_windows + 1), st));
	return LB2_OK;
}

extern "C" int lb2_upload(lb2_ctx *ctx, const lb2_batch *b)
{
	int rc = lb2_prepare_static(ctx, b, false); if (rc) { return rc; }
	if ((rc = lb2_plan_windows(ctx, b))) { return rc; }
	if ((rc = lb2_prepare_escalation(ctx))) { return rc; }
	const uint32_t W = b->n_windows, R = b->n_reads; uint64_t h2d = 0;
#define LB2_UP(buf, ptr, bytes) do { if (bytes) { LB2_CK(cudaMemcpyAsync(ctx->buf.p, (ptr), (bytes), cudaMemcpyHostToDevice, ctx->stream)); h2d += (bytes); } } while (0)
	LB2_UP(d_ref_off, b->ref_off, sizeof(uint32_t) * (size_t)(W + 1)); LB2_UP(d_ref_start, b->ref_start, sizeof(int32_t) * (size_t)W); LB2_UP(d_wr_off, b->wr_off, sizeof(uint32_t) * (size_t)(W + 1));
	LB2_UP(d_wr_idx, b->wr_idx, sizeof(uint32_t) * (size_t)b->n_wr); LB2_UP(d_base_off, b->base_off, sizeof(uint64_t) * (size_t)(R + 1)); LB2_UP(d_flags, b->flags, (size_t)R);
	LB2_UP(d_name_rank, b->name_rank, sizeof(uint32_t) * (size_t)R); LB2_UP(d_ref_seq, b->ref_seq, (size_t)b->n_ref_bytes); LB2_UP(d_seq, b->seq, (size_t)b->n_base_bytes); LB2_UP(d_qual, b->qual, (size_t)b->n_base_bytes);
#undef LB2_UP
	ctx->h2d_bytes = h2d;
	LB2_CK(cudaStreamSynchronize(ctx->stream));
	ctx->resident = true;
	return LB2_OK;
}

// the hot path over a resident batch: pre-pack pass, window pipeline, escalation pass, record compaction
extern "C" int lb2_run(lb2_ctx *ctx)
{
	if (!ctx) { return LB2_ERR_ARG; }
	if (!ctx->resident) { return LB2_ERR_STATE; }
	if (cudaSetDevice(ctx->device) != cudaSuccess) { return LB2_ERR_CUDA; }
	const uint32_t W = ctx->n_windows; int rc;
	LB2_CK(cudaFuncSetAttribute(lb2_window_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smem_cap));
	if ((rc = lb2_enqueue_reset(ctx, ctx->stream))) { return rc; }
	LB2_CK(cudaEventRecord(ctx->ev[0], ctx->stream));
	if (W) { if ((rc = lb2_enqueue_pack(ctx, 0, ctx->n_reads, ctx->stream))) { return rc; } }
	LB2_CK(cudaEventRecord(ctx->ev[1], ctx->stream));
	if (W) { if ((rc = lb2_enqueue_windows(ctx, 0, W, 0, 0, ctx->stream))) { return rc; } }
	LB2_CK(cudaEventRecord(ctx->ev[2], ctx->stream));
	if (W) { if ((rc = lb2_enqueue_finish(ctx, ctx->stream, ctx->ev[3]))) { return rc; } } else { LB2_CK(cudaEventRecord(ctx->ev[3], ctx->stream)); }
	LB2_CK(cudaEventRecord(ctx->ev[4], ctx->stream));
	ctx->ran = true;
	return LB2_OK;
}

extern "C" int lb2_download(lb2_ctx *ctx, lb2_result *res)
{
	if (!ctx || !res) { return LB2_ERR_ARG; }
	if (!ctx->ran) { return LB2_ERR_STATE; }
	if (cudaSetDevice(ctx->device) != cudaSuccess) { return LB2_ERR_CUDA; }
	const uint32_t W = ctx->n_windows;
	uint32_t totals[2] = { 0, 0 };
	ctx->h_info.resize(W);
	if (W) {
		LB2_CK(cudaMemcpyAsync(totals, ctx->d_counters + LB2_CTR_TOTALS, 8, cudaMemcpyDeviceToHost, ctx->stream));
		LB2_CK(cudaMemcpyAsync(ctx->h_info.data(), ctx->d_info.p, sizeof(lb2_window_info) * (size_t)W, cudaMemcpyDeviceToHost, ctx->stream));
	}
	LB2_CK(cudaStreamSynchronize(ctx->stream));
	ctx->h_vars.resize(totals[0]); ctx->h_str.resize(totals[1]);
	if (totals[0]) { LB2_CK(cudaMemcpyAsync(ctx->h_vars.data(), ctx->d_cvars.p, sizeof(lb2_variant) * (size_t)totals[0], cudaMemcpyDeviceToHost, ctx->stream)); }
	if (totals[1]) { LB2_CK(cudaMemcpyAsync(ctx->h_str.data(), ctx->d_cstr.p, totals[1], cudaMemcpyDeviceToHost, ctx->stream)); }
	LB2_CK(cudaStreamSynchronize(ctx->stream));
	float ms = 0; cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[4]); ctx->kernel_ms = ms;
	ctx->d2h_bytes = 8 + sizeof(lb2_window_info) * (size_t)W + sizeof(lb2_variant) * (size_t)totals[0] + totals[1];
	res->n_windows = W; res->n_variants = totals[0]; res->windows = ctx->h_info.data(); res->variants = ctx->h_vars.data();
	res->strings = ctx->h_str.data(); res->n_string_bytes = totals[1]; res->kernel_ms = ms;
	return LB2_OK;
}

// Host buffers in, host buffers out.  The batch is cut into a few segments of consecutive windows (short ones first);
// per segment the copy stream uploads what its windows read that is not on the device yet -- their reference bases,
// their read lists, the next stretch of the pool -- and runs the pre-pack pass over those reads; the segment's windows
// are then assembled by their own launch on one of two compute streams, while the copy stream is already busy with
// the next segment.  Plain stream/event ordering: no kernel ever waits for a copy that was issued after it.
// From pageable memory the copies do not overlap anything, so the batch is one segment.
extern "C" int lb2_process(lb2_ctx *ctx, const lb2_batch *batch, lb2_result *result)
{
	if (!ctx || !batch || !result) { return LB2_ERR_ARG; }
	if (cudaSetDevice(ctx->device) != cudaSuccess) { return LB2_ERR_CUDA; }
	const bool timing = env_u32("LB2_TIMING", 0) != 0; const auto t0 = std::chrono::steady_clock::now();
	bool pinned = batch->n_base_bytes > 0 && batch->n_windows > 0 && env_u32("LB2_STREAM", 1) != 0;
	{
		const void *arrs[] = { batch->seq, batch->qual, batch->base_off, batch->flags, batch->name_rank, batch->wr_idx, batch->ref_seq };
		const uint64_t sizes[] = { batch->n_base_bytes, batch->n_base_bytes, 1, batch->n_reads, batch->n_reads, batch->n_wr, batch->n_ref_bytes };
		for (int i = 0; i < 7 && pinned; ++i) {
			if (!sizes[i]) { continue; }
			cudaPointerAttributes at; if (cudaPointerGetAttributes(&at, arrs[i]) != cudaSuccess || at.type != cudaMemoryTypeHost) { pinned = false; }
		}
		cudaGetLastError();
	}
	int rc = lb2_prepare_static(ctx, batch, pinned); if (rc) { return rc; }
	const uint32_t W = ctx->n_windows, R = ctx->n_reads;
	// The pass over the windows' read lists runs on its own thread(s) beside the upload of the first segment, whose few
	// hundred windows are planned right here; everything the first-pass launches need is fixed by lb2_prepare_static.
	int plan_rc = LB2_OK; std::thread planner; struct lb2_joiner { std::thread &t; ~lb2_joiner() { if (t.joinable()) { t.join(); } } } joiner{ planner };
	bool planning = false;
	std::atomic<bool> plan_done(false);
	if (pinned && W >= 4096 && env_u32("LB2_PLAN_OVERLAP", 1)) { planner = std::thread([&]() { plan_rc = lb2_plan_windows(ctx, batch); plan_done.store(true, std::memory_order_release); }); planning = true; }
	else { if ((rc = lb2_plan_windows(ctx, batch))) { return rc; } if ((rc = lb2_prepare_escalation(ctx))) { return rc; } }
	const auto t1 = std::chrono::steady_clock::now();
	LB2_CK(cudaFuncSetAttribute(lb2_window_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smem_cap));
	if ((rc = lb2_enqueue_reset(ctx, ctx->stream))) { return rc; }
	LB2_CK(cudaEventRecord(ctx->ev[0], ctx->stream));
	LB2_CK(cudaEventRecord(ctx->ev_ready, ctx->stream));
	LB2_CK(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_ready, 0));
	LB2_CK(cudaStreamWaitEvent(ctx->wstream[0], ctx->ev_ready, 0)); LB2_CK(cudaStreamWaitEvent(ctx->wstream[1], ctx->ev_ready, 0));
	uint64_t h2d = 0;
	// (LB2_TIMING=1: a device time line of the segments -- copy + pre-pack done, window launch begin / end)
	std::vector<cudaEvent_t> tl; if (timing) { tl.resize(3 * LB2_MAX_SEG); for (auto &e : tl) { cudaEventCreate(&e); } }
#define LB2_PIECE(buf, ptr, esz, from, to) do { if ((to) > (from)) { LB2_CK(cudaMemcpyAsync((char *)ctx->buf.p + (size_t)(from) * (esz), (const char *)(ptr) + (size_t)(from) * (esz), \
		(size_t)((to) - (from)) * (esz), cudaMemcpyHostToDevice, ctx->copy_stream)); h2d += (uint64_t)((to) - (from)) * (esz); } } while (0)
	LB2_PIECE(d_ref_off, batch->ref_off, 4, (size_t)0, (size_t)W + 1); LB2_PIECE(d_ref_start, batch->ref_start, 4, (size_t)0, (size_t)W); LB2_PIECE(d_wr_off, batch->wr_off, 4, (size_t)0, (size_t)W + 1);
	// segments: the first one small (the kernels start after it), then doubling; a segment = consecutive windows + the pack
	// blocks of the pool they touch that no earlier segment has uploaded (a bitmap over the blocks)
	std::vector<lb2_seg> segs;
	const uint64_t nb = batch->n_base_bytes; const uint32_t PB = LB2_PACK_BLOCK, nblk = (R + PB - 1) / PB;
	const uint64_t first = std::max<uint64_t>(env_u32("LB2_SEG_FIRST", 4u << 20), 1u << 16), min_w = std::max(1u, env_u32("LB2_SEG_MIN_WINDOWS", 512));
	std::vector<uint8_t> up(nblk + 1, 0);
	auto blk_bytes = [&](uint32_t k) -> uint64_t { const uint32_t ra_ = k * PB, rb_ = std::min(R, (k + 1) * PB); return batch->base_off[rb_] - batch->base_off[ra_]; };
	uint64_t chunk = first; uint32_t wa = 0;
	const bool split = pinned && nb > 2 * first;
	while (wa < W) {
		const size_t s = segs.size();
		lb2_seg sg; sg.w0 = wa; sg.nr = 0;
		{
			uint32_t lo[2] = { 0xFFFFFFFFu, 0xFFFFFFFFu }, hi[2] = { 0, 0 }; uint64_t fresh = 0; uint32_t wb = wa;      // block ranges of the two stretches
			auto cover = [&](int which, uint32_t r0_, uint32_t r1_) {      // extend stretch `which` to cover reads [r0_, r1_); count the bytes of blocks not uploaded yet
				if (r1_ <= r0_) { return; }
				const uint32_t k0 = r0_ / PB, k1 = (r1_ + PB - 1) / PB;
				if (lo[which] == 0xFFFFFFFFu) { lo[which] = k0; hi[which] = k0; }
				for (uint32_t k = k0; k < lo[which]; ++k) { if (!up[k]) { fresh += blk_bytes(k); } }
				for (uint32_t k = std::max(hi[which], k0); k < k1; ++k) { if (!up[k]) { fresh += blk_bytes(k); } }
				lo[which] = std::min(lo[which], k0); hi[which] = std::max(hi[which], k1);
			};
			while (wb < W) {
				if (split && s + 1 < LB2_MAX_SEG && wb - wa >= min_w && fresh > chunk && W - wb >= min_w) { break; }
				if (planning) {      // (the plan is still being made: this window's stretches are worked out on the spot)
					const lb2_win_plan wp = lb2_plan_one(batch, wb);
					if (wp.bad) { cudaStreamSynchronize(ctx->copy_stream); cudaStreamSynchronize(ctx->pack_stream); cudaStreamSynchronize(ctx->wstream[0]); cudaStreamSynchronize(ctx->wstream[1]); return LB2_ERR_ARG; }
					cover(0, wp.rng[0], wp.rng[1]); cover(1, wp.rng[2], wp.rng[3]);
				} else { const uint32_t *rg = ctx->h_rng.data() + 4 * (size_t)wb; cover(0, rg[0], rg[1]); cover(1, rg[2], rg[3]); }
				++wb;
			}
			sg.w1 = wb;
			if (lo[0] != 0xFFFFFFFFu && lo[1] != 0xFFFFFFFFu && !(hi[0] < lo[1] || hi[1] < lo[0])) { lo[0] = std::min(lo[0], lo[1]); hi[0] = std::max(hi[0], hi[1]); lo[1] = 0xFFFFFFFFu; }      // the two stretches met
			for (int q = 0; q < 2; ++q) {      // blocks of each stretch that are not on the device yet: at most two runs per stretch are kept apart, more are bridged
				if (lo[q] == 0xFFFFFFFFu) { continue; }
				uint32_t k = lo[q]; int runs = 0;
				while (k < hi[q]) {
					while (k < hi[q] && up[k]) { ++k; }
					if (k >= hi[q]) { break; }
					uint32_t e = k; while (e < hi[q] && !up[e]) { ++e; }
					if (runs == 1) { uint32_t last = hi[q]; while (last > e && up[last - 1]) { --last; } e = last; }      // (second run of this stretch: take the rest in one piece)
					for (uint32_t z = k; z < e; ++z) { up[z] = 1; }
					sg.r0[sg.nr] = k * PB; sg.r1[sg.nr] = std::min(R, e * PB); ++sg.nr; ++runs; k = e;
				}
			}
			segs.push_back(sg); wa = wb; chunk *= 2;
		}
		LB2_PIECE(d_ref_seq, batch->ref_seq, 1, (size_t)batch->ref_off[sg.w0], (size_t)batch->ref_off[sg.w1]);
		LB2_PIECE(d_wr_idx, batch->wr_idx, 4, (size_t)batch->wr_off[sg.w0], (size_t)batch->wr_off[sg.w1]);
		for (uint32_t q = 0; q < sg.nr; ++q) {
			const uint32_t r0 = sg.r0[q], r1 = sg.r1[q]; if (r1 <= r0) { continue; }
			LB2_PIECE(d_base_off, batch->base_off, 8, (size_t)r0, (size_t)r1 + 1);
			LB2_PIECE(d_flags, batch->flags, 1, (size_t)r0, (size_t)r1); LB2_PIECE(d_name_rank, batch->name_rank, 4, (size_t)r0, (size_t)r1);
			LB2_PIECE(d_seq, batch->seq, 1, (size_t)batch->base_off[r0], (size_t)batch->base_off[r1]);
			LB2_PIECE(d_qual, batch->qual, 1, (size_t)batch->base_off[r0], (size_t)batch->base_off[r1]);
		}
		// the pre-pack pass runs on its own stream: when its blocks have to wait for room beside the window CTAs, the copies of
		// the next segment go on regardless
		LB2_CK(cudaEventRecord(ctx->ev_cp[s], ctx->copy_stream));
		LB2_CK(cudaStreamWaitEvent(ctx->pack_stream, ctx->ev_cp[s], 0));
		for (uint32_t q = 0; q < sg.nr; ++q) { if (sg.r1[q] > sg.r0[q]) { if ((rc = lb2_enqueue_pack(ctx, sg.r0[q], sg.r1[q], ctx->pack_stream))) { return rc; } } }
		LB2_CK(cudaEventRecord(ctx->ev_seg[s], ctx->pack_stream));
		if (timing) { cudaEventRecord(tl[3 * s], ctx->pack_stream); }
		cudaStream_t ws = ctx->wstream[s & 1];
		LB2_CK(cudaStreamWaitEvent(ws, ctx->ev_seg[s], 0));
		if (timing) { cudaEventRecord(tl[3 * s + 1], ws); }
		if ((rc = lb2_enqueue_windows(ctx, sg.w0, sg.w1, pinned ? (uint32_t)(s & 1) : 0u, (uint32_t)s, ws))) { return rc; }
		if (timing) { cudaEventRecord(tl[3 * s + 2], ws); }
		// the plan is taken over as soon as it is there (until then the segments work their windows' stretches out themselves,
		// which keeps the copy and compute streams fed); it is only NEEDED for the sizes of the escalation passes, below
		if (planning && (plan_done.load(std::memory_order_acquire) || wa >= W)) {
			planner.join(); planning = false;
			if (plan_rc == LB2_OK) { plan_rc = lb2_prepare_escalation(ctx); }
			if (plan_rc != LB2_OK) { cudaStreamSynchronize(ctx->copy_stream); cudaStreamSynchronize(ctx->pack_stream); cudaStreamSynchronize(ctx->wstream[0]); cudaStreamSynchronize(ctx->wstream[1]); return plan_rc; }
		}
	}
#undef LB2_PIECE
	ctx->h2d_bytes = h2d;
	for (int i = 0; i < 2; ++i) { LB2_CK(cudaEventRecord(ctx->ev_w[i], ctx->wstream[i])); LB2_CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_w[i], 0)); }
	LB2_CK(cudaEventRecord(ctx->ev[1], ctx->stream)); LB2_CK(cudaEventRecord(ctx->ev[2], ctx->stream));
	if (W) { if ((rc = lb2_enqueue_finish(ctx, ctx->stream, ctx->ev[3]))) { return rc; } } else { LB2_CK(cudaEventRecord(ctx->ev[3], ctx->stream)); }
	LB2_CK(cudaEventRecord(ctx->ev[4], ctx->stream));
	ctx->ran = true; ctx->resident = true;
	const auto t2 = std::chrono::steady_clock::now();
	LB2_CK(cudaStreamSynchronize(ctx->copy_stream)); LB2_CK(cudaStreamSynchronize(ctx->pack_stream));      // the caller's buffers are free again when the call returns
	const auto t3 = std::chrono::steady_clock::now();
	rc = lb2_download(ctx, result);
	if (timing) {
		const auto t4 = std::chrono::steady_clock::now();
		auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b_) { return std::chrono::duration<double, std::milli>(b_ - a).count(); };
		for (size_t s_ = 0; s_ < segs.size() && !tl.empty(); ++s_) {
			float a = 0, b_ = 0, c = 0; cudaEventElapsedTime(&a, ctx->ev[0], tl[3 * s_]); cudaEventElapsedTime(&b_, ctx->ev[0], tl[3 * s_ + 1]); cudaEventElapsedTime(&c, ctx->ev[0], tl[3 * s_ + 2]);
			fprintf(stderr, "  segment %zu: windows [%u, %u) reads [%u, %u)%s: copied+packed at %.2f ms, windows %.2f .. %.2f ms\n", s_, segs[s_].w0, segs[s_].w1, segs[s_].nr ? segs[s_].r0[0] : 0u, segs[s_].nr ? segs[s_].r1[0] : 0u, segs[s_].nr > 1 ? " + more" : "", a, b_, c);
		}
		for (auto &e : tl) { cudaEventDestroy(e); }
		fprintf(stderr, "lb2_process: %zu segment(s)%s, plan %.2f ms, enqueue %.2f ms, copies done +%.2f ms, kernels+download +%.2f ms, total %.2f ms (device span %.2f ms)\n",
		        segs.size(), pinned ? "" : " (pageable)", ms(t0, t1), ms(t1, t2), ms(t2, t3), ms(t3, t4), ms(t0, t4), (double)result->kernel_ms);
	}
	return rc;
}

// cycles per pipeline phase (lane 0 of every CTA), summed since the context was created: zero unless built with -DLB2_PROFILE
extern "C" int lb2_phase_cycles(lb2_ctx *ctx, unsigned long long *out24, int reset)
{
	if (!ctx || !out24) { return LB2_ERR_ARG; }
	LB2_CK(cudaMemcpy(out24, ctx->d_prof, 24 * 8, cudaMemcpyDeviceToHost));
	if (reset) { LB2_CK(cudaMemset(ctx->d_prof, 0, 24 * 8)); }
	return LB2_OK;
}
extern "C" uint64_t lb2_last_h2d_bytes(const lb2_ctx *ctx) { return ctx ? ctx->h2d_bytes : 0; }
extern "C" uint64_t lb2_last_d2h_bytes(const lb2_ctx *ctx) { return ctx ? ctx->d2h_bytes : 0; }
extern "C" uint32_t lb2_resident_ctas(const lb2_ctx *ctx) { return ctx ? ctx->C.n_slots : 0; }
extern "C" uint32_t lb2_smem_per_cta(const lb2_ctx *ctx) { return ctx ? ctx->C.smem_bytes : 0; }
extern "C" int lb2_wait(lb2_ctx *ctx) { if (!ctx) return LB2_ERR_ARG; LB2_CK(cudaStreamSynchronize(ctx->stream)); return LB2_OK; }
extern "C" float lb2_last_kernel_ms(lb2_ctx *ctx) { if (!ctx || !ctx->ran) return 0; float ms = 0; cudaEventSynchronize(ctx->ev[4]); cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[4]); return ms; }
extern "C" float lb2_last_kernel_ms_of(lb2_ctx *ctx, int which)
{
	if (!ctx || !ctx->ran || which < 0 || which > 3) { return 0; }
	float ms = 0; cudaEventSynchronize(ctx->ev[4]);
	if (cudaEventElapsedTime(&ms, ctx->ev[which], ctx->ev[which + 1]) != cudaSuccess) { cudaGetLastError(); return 0; }
	return ms;
}

extern "C" void *lb2_alloc_pinned(size_t bytes) { void *p = nullptr; if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return nullptr; } return p; }
extern "C" void lb2_free_pinned(void *p) { if (p) { cudaFreeHost(p); } }

extern "C" int lb2_rank_names(const char *const *names, uint32_t n, uint32_t *rank_out)
{
	if (!names || !rank_out) { return LB2_ERR_ARG; }
	std::vector<uint32_t> idx(n); std::iota(idx.begin(), idx.end(), 0u);
	std::sort(idx.begin(), idx.end(), [&](uint32_t a, uint32_t b) { return strcmp(names[a], names[b]) < 0; });
	uint32_t r = 0;
	for (uint32_t i = 0; i < n; ++i) { if (i && strcmp(names[idx[i]], names[idx[i - 1]]) != 0) { ++r; } rank_out[idx[i]] = r; }
	return LB2_OK;
}

// ---- record gather over NCCL (declared in include/lancet_b200.h) --------------------------------------------------------
// NCCL is bound at run time (dlopen of libnccl.so.2 on first use, the already loaded one if the process has one): a
// process that also loads PyTorch must end up with ONE libnccl, and PyTorch ships its own.
#include <dlfcn.h>
namespace {
struct lb2_nccl_api {
	ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr; ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
	ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr; ncclResult_t (*CommAbort)(ncclComm_t) = nullptr; const char *(*GetErrorString)(ncclResult_t) = nullptr;
	ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*GroupStart)() = nullptr; ncclResult_t (*GroupEnd)() = nullptr;
	bool ok = false;
	lb2_nccl_api() {
		void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
		if (!h) { h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL); }
		if (!h) { h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL); }
		if (!h) { return; }
#define LB2_SYM(f) do { *(void **)(&f) = dlsym(h, "nccl" #f); if (!f) { return; } } while (0)
		LB2_SYM(GetUniqueId); LB2_SYM(CommInitRank); LB2_SYM(CommDestroy); LB2_SYM(CommAbort); LB2_SYM(GetErrorString); LB2_SYM(AllGather); LB2_SYM(Send); LB2_SYM(Recv); LB2_SYM(GroupStart); LB2_SYM(GroupEnd);
#undef LB2_SYM
		ok = true;
	}
};
lb2_nccl_api &lb2_nccl() { static lb2_nccl_api api; return api; }
}
static void lb2_comm_release(ncclComm_t c) { if (lb2_nccl().ok) { lb2_nccl().CommAbort(c); } }      // (no rendezvous: the other ranks may be gone already)
#define ncclGetUniqueId lb2_nccl().GetUniqueId
#define ncclCommInitRank lb2_nccl().CommInitRank
#define ncclGetErrorString lb2_nccl().GetErrorString
#define ncclAllGather lb2_nccl().AllGather
#define ncclSend lb2_nccl().Send
#define ncclRecv lb2_nccl().Recv
#define ncclGroupStart lb2_nccl().GroupStart
#define ncclGroupEnd lb2_nccl().GroupEnd
#define LB2_NC(call) do { ncclResult_t r_ = (call); if (r_ != ncclSuccess) { ctx->err = std::string(#call) + ": " + ncclGetErrorString(r_); return LB2_ERR_CUDA; } } while (0)
extern "C" int lb2_comm_unique_id(char *id_out)
{
	static_assert(sizeof(ncclUniqueId) <= LB2_COMM_ID_BYTES, "NCCL id does not fit");
	if (!id_out) { return LB2_ERR_ARG; }
	if (!lb2_nccl().ok) { return LB2_ERR_CUDA; }
	ncclUniqueId id; if (ncclGetUniqueId(&id) != ncclSuccess) { return LB2_ERR_CUDA; }
	memset(id_out, 0, LB2_COMM_ID_BYTES); memcpy(id_out, &id, sizeof id);
	return LB2_OK;
}
extern "C" int lb2_comm_init(lb2_ctx *ctx, const char *id_in, int rank, int world)
{
	if (!ctx || !id_in || world < 1 || rank < 0 || rank >= world) { return LB2_ERR_ARG; }
	if (!lb2_nccl().ok) { ctx->err = "libnccl.so.2 not found"; return LB2_ERR_CUDA; }
	LB2_CK(cudaSetDevice(ctx->device));
	if (ctx->comm) { lb2_nccl().CommAbort(ctx->comm); ctx->comm = nullptr; }
	ncclUniqueId id; memcpy(&id, id_in, sizeof id);
	LB2_NC(ncclCommInitRank(&ctx->comm, world, id, rank));
	ctx->comm_rank = rank; ctx->comm_world = world;
	return LB2_OK;
}
extern "C" int lb2_comm_gather(lb2_ctx *ctx, const lb2_variant *vars, uint32_t n_vars, const char *strs, uint64_t n_str, uint64_t *stats, int n_stats, int root, lb2_result *merged)
{
	if (!ctx || !merged || n_stats < 0 || n_stats > 8 || (n_vars && !vars) || (n_str && !strs) || (n_stats && !stats)) { return LB2_ERR_ARG; }
	if (!ctx->comm) { return LB2_ERR_STATE; }
	LB2_CK(cudaSetDevice(ctx->device));
	const int W = ctx->comm_world, me = ctx->comm_rank, M = 2 + n_stats; int rc;
	if (root < 0 || root >= W) { return LB2_ERR_ARG; }
	// counts of every rank: [records, string bytes, stats...]
	if ((rc = lb2_reserve(ctx, ctx->d_comm_cnt, sizeof(uint64_t) * (size_t)M * W))) { return rc; }
	std::vector<uint64_t> cnt((size_t)M * W, 0); uint64_t *mine = cnt.data() + (size_t)M * me;
	mine[0] = n_vars; mine[1] = n_str; for (int i = 0; i < n_stats; ++i) { mine[2 + i] = stats[i]; }
	uint64_t *d_cnt = (uint64_t *)ctx->d_comm_cnt.p;
	LB2_CK(cudaMemcpyAsync(d_cnt + (size_t)M * me, mine, sizeof(uint64_t) * M, cudaMemcpyHostToDevice, ctx->stream));
	LB2_NC(ncclAllGather(d_cnt + (size_t)M * me, d_cnt, (size_t)M, ncclUint64, ctx->comm, ctx->stream));
	LB2_CK(cudaMemcpyAsync(cnt.data(), d_cnt, sizeof(uint64_t) * (size_t)M * W, cudaMemcpyDeviceToHost, ctx->stream));
	LB2_CK(cudaStreamSynchronize(ctx->stream));
	// payloads: [records | strings] of every rank to root
	auto bytes_of = [&](int r) -> size_t { return (size_t)cnt[(size_t)M * r] * sizeof(lb2_variant) + (((size_t)cnt[(size_t)M * r + 1] + 15) & ~(size_t)15); };
	const size_t my_bytes = bytes_of(me);
	if ((rc = lb2_reserve(ctx, ctx->d_comm_send, my_bytes + 16))) { return rc; }
	if (n_vars) { LB2_CK(cudaMemcpyAsync(ctx->d_comm_send.p, vars, sizeof(lb2_variant) * (size_t)n_vars, cudaMemcpyHostToDevice, ctx->stream)); }
	if (n_str) { LB2_CK(cudaMemcpyAsync((char *)ctx->d_comm_send.p + sizeof(lb2_variant) * (size_t)n_vars, strs, n_str, cudaMemcpyHostToDevice, ctx->stream)); }
	std::vector<size_t> off(W + 1, 0); for (int r = 0; r < W; ++r) { off[r + 1] = off[r] + bytes_of(r); }
	if (me == root) { if ((rc = lb2_reserve(ctx, ctx->d_comm_recv, off[W] + 16))) { return rc; } }
	LB2_NC(ncclGroupStart());
	if (me == root) {
		for (int r = 0; r < W; ++r) { if (r != root && bytes_of(r)) { LB2_NC(ncclRecv((char *)ctx->d_comm_recv.p + off[r], bytes_of(r), ncclUint8, r, ctx->comm, ctx->stream)); } }
	} else if (my_bytes) { LB2_NC(ncclSend(ctx->d_comm_send.p, my_bytes, ncclUint8, root, ctx->comm, ctx->stream)); }
	LB2_NC(ncclGroupEnd());
	memset(merged, 0, sizeof *merged);
	if (me != root) { LB2_CK(cudaStreamSynchronize(ctx->stream)); return LB2_OK; }
	if (my_bytes) { LB2_CK(cudaMemcpyAsync((char *)ctx->d_comm_recv.p + off[root], ctx->d_comm_send.p, my_bytes, cudaMemcpyDeviceToDevice, ctx->stream)); }
	std::vector<char> raw(off[W]);
	if (off[W]) { LB2_CK(cudaMemcpyAsync(raw.data(), ctx->d_comm_recv.p, off[W], cudaMemcpyDeviceToHost, ctx->stream)); }
	LB2_CK(cudaStreamSynchronize(ctx->stream));
	ctx->g_vars.clear(); ctx->g_str.clear();
	for (int r = 0; r < W; ++r) {
		const size_t nv = (size_t)cnt[(size_t)M * r], ns = (size_t)cnt[(size_t)M * r + 1]; const size_t sbase = ctx->g_str.size();
		const lb2_variant *v = (const lb2_variant *)(raw.data() + off[r]);
		for (size_t i = 0; i < nv; ++i) { lb2_variant x = v[i]; x.str_off += (uint32_t)sbase; ctx->g_vars.push_back(x); }
		ctx->g_str.insert(ctx->g_str.end(), raw.data() + off[r] + nv * sizeof(lb2_variant), raw.data() + off[r] + nv * sizeof(lb2_variant) + ns);
	}
	for (int i = 0; i < n_stats; ++i) { uint64_t t = 0; for (int r = 0; r < W; ++r) { t += cnt[(size_t)M * r + 2 + i]; } stats[i] = t; }
	merged->n_variants = (uint32_t)ctx->g_vars.size(); merged->variants = ctx->g_vars.data(); merged->strings = ctx->g_str.data(); merged->n_string_bytes = ctx->g_str.size();
	return LB2_OK;
}
