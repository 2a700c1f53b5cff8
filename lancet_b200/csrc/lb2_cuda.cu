// lb2_cuda.cu -- the sm_100a kernels and the extern "C" boundary declared in include/lancet_b200.h.
//
// One persistent CTA per resident slot; each CTA pulls window indices from a global counter and runs
// the whole micro-assembly of that window (lb2_process_window) out of its own workspace slab, with the
// window's reads staged 2-bit-packed in shared memory.  No CPU fallback exists in this file: every
// entry point fails with LB2_ERR_CUDA when there is no usable device.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>
#include <algorithm>
#include <numeric>
#include <thread>
#include <chrono>

#include "lb2_pipeline.cuh"

struct lb2_launch {
	lb2_params P; lb2_cfg C; lb2_dev_batch B; lb2_dev_out O;
	uint8_t *ws_base; size_t ws_stride; uint32_t *counter;
	const uint32_t *avail;      // streamed lb2_process: windows [0, *avail) have their reads in HBM (NULL: the batch is resident)
	uint32_t *stalled;          // set when a window fetch gave up waiting for the watermark (the host then redoes the batch resident)
	const uint32_t *win_list; const uint32_t *n_list;   // escalation pass: indices of the windows to redo (NULL = all windows)
	uint32_t *retry_list; uint32_t *retry_count;
	// compaction outputs
	uint32_t *var_off; uint32_t *str_off; uint32_t *totals; lb2_variant *cvars; char *cstr;
};

__global__ void __launch_bounds__(256, 3)
lb2_window_kernel(const lb2_launch *Lp)
{
	extern __shared__ __align__(16) uint8_t smem[];
	__shared__ uint32_t s_next;
	// the window descriptor (some 150 pointers, identical for every lane) lives in SHARED memory: as a kernel-local
	// struct handed by reference to the pipeline's functions it sat in per-thread local memory, which with three 72 KB
	// CTAs per SM has next to no L1 behind it -- every pointer fetch was an L2 round trip
	__shared__ lb2_win sW;
	// ... and so do the parameter blocks it points to (read inside lane-0 loops: a global load each time otherwise)
	__shared__ lb2_params sP; __shared__ lb2_cfg sC; __shared__ lb2_dev_batch sB; __shared__ lb2_dev_out sO;
	lb2_win &W = sW;
	if (threadIdx.x == 0) {
		sP = Lp->P; sC = Lp->C; sB = Lp->B; sO = Lp->O;
		W.P = &sP; W.C = &sC; W.B = &sB; W.O = &sO; W.escal = (Lp->win_list != nullptr);
		lb2_ws_layout(Lp->C, Lp->ws_base + (size_t)blockIdx.x * Lp->ws_stride, &W.ws); W.ws0 = W.ws;
		W.sh = (lb2_sh *)smem;
		W.ref_raw = (char *)smem + ((sizeof(lb2_sh) + 15) & ~(size_t)15);
		W.bits = (uint32_t *)(W.ref_raw + LB2_MAX_REF);
		W.lowq = W.bits + (Lp->C.max_bp / 16 + 4);
		W.treg = smem + ((lb2_smem_fixed(Lp->C.max_bp) + 15) & ~(size_t)15);
	}
	__syncthreads();
	const uint32_t nwin = Lp->win_list ? *Lp->n_list : Lp->B.n_windows;
	while (true) {
		if (threadIdx.x == 0) {
			const uint32_t nx = atomicAdd(Lp->counter, 1u);
			uint32_t nx2 = nx;
			if (Lp->avail && nx < nwin) {      // the read pool is still arriving on the copy stream: wait for this window's watermark
				// (never forever: if something serialises the copy stream behind this kernel -- a profiler, a debugger --
				// the fetch gives up after two seconds and the host redoes the batch the resident way)
				unsigned long long t0 = 0, t1 = 0; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
				while (*(volatile const uint32_t *)Lp->avail <= nx) {
					__nanosleep(200);
					asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
					if (t1 - t0 > 2000000000ull || *(volatile uint32_t *)Lp->stalled) { *(volatile uint32_t *)Lp->stalled = 1u; nx2 = nwin; break; }
				}
				__threadfence_system();
			}
			s_next = nx2;
		}
		__syncthreads();
		uint32_t w = s_next;
		__syncthreads();
		if (w >= nwin) { break; }
		if (Lp->win_list) { w = Lp->win_list[w]; }
		lb2_process_window(W, w);
	}
}

// windows that ran out of a per-CTA capacity in the first pass are redone by the escalation pass
// (same kernel, larger table / arena / BFS queue, one CTA per SM)
__global__ void lb2_collect_kernel(const lb2_launch *Lp)
{
	const uint32_t n = Lp->B.n_windows;
	for (uint32_t w = blockIdx.x * blockDim.x + threadIdx.x; w < n; w += gridDim.x * blockDim.x) {
		lb2_window_info wi = Lp->O.info[w];
		if (wi.status == LB2_WIN_OVERFLOW) {
			uint32_t d = wi.detail;
			if (d == LB2_D_HASH_FULL || d == LB2_D_NODES || d == LB2_D_ARENA || d == LB2_D_QUEUE || d == LB2_D_SMEM || d == LB2_D_BUCKETS ||
			    d == LB2_D_STACK || d == LB2_D_EDGES || d == LB2_D_SPECIAL || d == LB2_D_READS ||
			    d == LB2_D_VARIANTS || d == LB2_D_STRINGS) {      // (escalated windows emit into the large output slabs)
				Lp->retry_list[atomicAdd(Lp->retry_count, 1u)] = w;
			}
		}
	}
}

// exclusive scans of per-window variant counts and string bytes (one block)
__global__ void lb2_scan_kernel(const lb2_launch *Lp)
{
	__shared__ uint32_t pv[1024], ps[1024];
	const uint32_t n = Lp->B.n_windows, t = threadIdx.x, nt = blockDim.x;
	const uint32_t chunk = (n + nt - 1) / nt, lo = t * chunk, hi = min(n, lo + chunk);
	uint32_t sv = 0, ss = 0;
	for (uint32_t w = lo; w < hi; ++w) { sv += Lp->O.info[w].n_variants; ss += (Lp->O.info[w].n_variants ? Lp->O.str_used[w] : 0); }
	pv[t] = sv; ps[t] = ss; __syncthreads();
	if (t == 0) {
		uint32_t av = 0, as = 0;
		for (uint32_t i = 0; i < nt; ++i) { uint32_t v = pv[i], s = ps[i]; pv[i] = av; ps[i] = as; av += v; as += s; }
		Lp->totals[0] = av; Lp->totals[1] = as;
	}
	__syncthreads();
	uint32_t av = pv[t], as = ps[t];
	for (uint32_t w = lo; w < hi; ++w) {
		Lp->var_off[w] = av; Lp->str_off[w] = as;
		uint32_t nv = Lp->O.info[w].n_variants; av += nv; as += (nv ? Lp->O.str_used[w] : 0);
	}
}

// gather the per-window slabs into dense arrays (one block per window)
__global__ void lb2_gather_kernel(const lb2_launch *Lp)
{
	const uint32_t w = blockIdx.x; const uint32_t nv = Lp->O.info[w].n_variants;
	if (!nv) { return; }
	const uint32_t vo = Lp->var_off[w], so = Lp->str_off[w], sb = Lp->O.str_used[w];
	const uint32_t big = Lp->O.big_slot[w];
	const lb2_variant *vsrc = (big != 0xFFFFFFFFu) ? Lp->O.big_variants + (size_t)big * Lp->O.big_max_var : Lp->O.variants + (size_t)w * Lp->C.max_var;
	for (uint32_t i = threadIdx.x; i < nv; i += blockDim.x) {
		lb2_variant v = vsrc[i]; v.str_off += so; Lp->cvars[vo + i] = v;
	}
	const char *src = (big != 0xFFFFFFFFu) ? Lp->O.big_strings + (size_t)big * Lp->O.big_str_bytes : Lp->O.strings + (size_t)w * Lp->C.str_bytes;
	for (uint32_t i = threadIdx.x; i < sb; i += blockDim.x) { Lp->cstr[so + i] = src[i]; }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
#define LB2_MAX_MARKS 1024
struct lb2_ctx {
	int device; cudaStream_t stream; cudaEvent_t ev0, ev1;
	cudaStream_t copy_stream = nullptr; uint32_t *d_avail = nullptr, *d_stalled = nullptr;      // streamed lb2_process
	cudaEvent_t ev_avail = nullptr; uint32_t *h_marks = nullptr;          // (pinned) per chunk: windows ready
	std::vector<uint32_t> h_need;                                         // per window: leading pool reads the windows up to it use
	lb2_params P; lb2_cfg C;
	int sm_count; uint32_t threads = 256;
	std::string err;
	// device buffers of the resident batch
	struct Buf { void *p = nullptr; size_t cap = 0; };
	Buf d_ref_off, d_ref_start, d_wr_off, d_wr_idx, d_base_off, d_flags, d_name_rank, d_ref_seq, d_seq, d_qual;
	Buf d_info, d_vars, d_strs, d_str_used, d_var_off, d_str_off, d_cvars, d_cstr, d_ws, d_big_vars, d_big_strs, d_big_slot;
	uint32_t *d_big_count = nullptr; uint32_t big_cap = 256, big_max_var = 1024, big_str_bytes = 128u << 10;
	uint32_t *d_counter = nullptr, *d_totals = nullptr; lb2_launch *d_launch = nullptr; unsigned long long *d_prof = nullptr;
	lb2_launch L;
	uint32_t n_windows = 0; bool resident = false, ran = false;
	size_t ws_stride = 0; uint32_t ws_slots = 0;
	// escalation pass
	lb2_cfg C2; lb2_launch L2; lb2_launch *d_launch2 = nullptr; Buf d_ws2, d_retry; uint32_t *d_counter2 = nullptr, *d_retry_count = nullptr;
	size_t ws2_stride = 0; uint32_t ws2_slots = 0; bool escalate = true;
	uint64_t launches = 0;
	float kernel_ms = 0;
	// host result
	std::vector<lb2_window_info> h_info; std::vector<lb2_variant> h_vars; std::vector<char> h_str;
	uint64_t h2d_bytes = 0, d2h_bytes = 0;
};

#define LB2_CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_); return LB2_ERR_CUDA; } } while (0)

static int lb2_reserve(lb2_ctx *ctx, lb2_ctx::Buf &b, size_t bytes, bool zero = false)
{
	if (bytes <= b.cap && b.p) { return LB2_OK; }
	if (b.p) { cudaFree(b.p); b.p = nullptr; b.cap = 0; }
	size_t cap = bytes + bytes / 4 + 256;
	LB2_CK(cudaMalloc(&b.p, cap));
	b.cap = cap;
	if (zero) { LB2_CK(cudaMemsetAsync(b.p, 0, cap, ctx->stream)); }
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

static uint32_t env_u32(const char *name, uint32_t dflt) { const char *s = getenv(name); return s ? (uint32_t)strtoul(s, nullptr, 10) : dflt; }

extern "C" int lb2_create(lb2_ctx **out, const lb2_params *params, int device)
{
	if (!out || !params) { return LB2_ERR_ARG; }
	*out = nullptr;
	int ndev = 0;
	if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) { return LB2_ERR_CUDA; }
	lb2_ctx *ctx = new lb2_ctx();
	ctx->device = device; ctx->P = *params;
	cudaDeviceProp prop;
	if (cudaSetDevice(device) != cudaSuccess || cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete ctx; return LB2_ERR_CUDA; }
	if (prop.major < 10) { delete ctx; return LB2_ERR_CUDA; }
	ctx->sm_count = prop.multiProcessorCount;
	if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return LB2_ERR_CUDA; }
	if (cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking) != cudaSuccess || cudaMalloc(&ctx->d_avail, 4) != cudaSuccess || cudaMalloc(&ctx->d_stalled, 4) != cudaSuccess || cudaMemset(ctx->d_stalled, 0, 4) != cudaSuccess ||
	    cudaEventCreateWithFlags(&ctx->ev_avail, cudaEventDisableTiming) != cudaSuccess || cudaHostAlloc((void **)&ctx->h_marks, sizeof(uint32_t) * LB2_MAX_MARKS, cudaHostAllocDefault) != cudaSuccess) { delete ctx; return LB2_ERR_CUDA; }
	cudaEventCreate(&ctx->ev0); cudaEventCreate(&ctx->ev1);
	lb2_cfg &C = ctx->C; memset(&C, 0, sizeof C);
	C.table_slots = env_u32("LB2_TABLE_SLOTS", 4096); C.max_nodes = C.table_slots - C.table_slots / 4;
	C.max_reads = 4096; C.max_bp = 0;
	C.arena_bytes = env_u32("LB2_ARENA_BYTES", 512u << 10); C.deficit_bytes = env_u32("LB2_DEFICIT_BYTES", 1u << 20);
	C.max_inst = env_u32("LB2_MAX_INST", 1u << 18);
	C.queue_cap = env_u32("LB2_QUEUE_CAP", 1u << 16); C.max_var = env_u32("LB2_MAX_VAR", 32); C.str_bytes = env_u32("LB2_STR_BYTES", 4096);
	C.debug_flags = env_u32("LB2_DEBUG_FLAGS", 0); C.max_special = env_u32("LB2_MAX_SPECIAL", 128); C.bucket_cap = 10273; C.max_k = 127; C.graph_bytes = env_u32("LB2_GRAPH_BYTES", 48u << 10);
	if (cudaMalloc(&ctx->d_prof, 24 * 8) != cudaSuccess || cudaMemset(ctx->d_prof, 0, 24 * 8) != cudaSuccess) { delete ctx; return LB2_ERR_CUDA; }
	if (cudaMalloc(&ctx->d_counter2, 4) != cudaSuccess || cudaMalloc(&ctx->d_retry_count, 4) != cudaSuccess || cudaMalloc(&ctx->d_launch2, sizeof(lb2_launch)) != cudaSuccess) { delete ctx; return LB2_ERR_CUDA; }
	ctx->escalate = env_u32("LB2_ESCALATE", 1) != 0;
	ctx->big_cap = env_u32("LB2_BIG_SLABS", 256); ctx->big_max_var = env_u32("LB2_BIG_MAX_VAR", 1024); ctx->big_str_bytes = env_u32("LB2_BIG_STR_BYTES", 128u << 10);
	if (cudaMalloc(&ctx->d_big_count, 4) != cudaSuccess) { delete ctx; return LB2_ERR_CUDA; }
	ctx->threads = env_u32("LB2_THREADS", 256); if (ctx->threads < 32 || ctx->threads > 256 || (ctx->threads & 31)) { ctx->threads = 256; }
	if (cudaMalloc(&ctx->d_counter, 4) != cudaSuccess || cudaMalloc(&ctx->d_totals, 8) != cudaSuccess || cudaMalloc(&ctx->d_launch, sizeof(lb2_launch)) != cudaSuccess) {
		delete ctx; return LB2_ERR_CUDA;
	}
	*out = ctx;
	return LB2_OK;
}

extern "C" void lb2_destroy(lb2_ctx *ctx)
{
	if (!ctx) { return; }
	cudaSetDevice(ctx->device);
	lb2_ctx::Buf *bufs[] = { &ctx->d_ref_off, &ctx->d_ref_start, &ctx->d_wr_off, &ctx->d_wr_idx, &ctx->d_base_off, &ctx->d_flags, &ctx->d_name_rank,
		&ctx->d_ref_seq, &ctx->d_seq, &ctx->d_qual, &ctx->d_info, &ctx->d_vars, &ctx->d_strs, &ctx->d_str_used, &ctx->d_var_off, &ctx->d_str_off,
		&ctx->d_cvars, &ctx->d_cstr, &ctx->d_ws, &ctx->d_big_vars, &ctx->d_big_strs, &ctx->d_big_slot };
	for (auto b : bufs) { if (b->p) { cudaFree(b->p); } }
	if (ctx->d_ws2.p) { cudaFree(ctx->d_ws2.p); } if (ctx->d_retry.p) { cudaFree(ctx->d_retry.p); }
	cudaFree(ctx->d_big_count); cudaFree(ctx->d_counter2); cudaFree(ctx->d_retry_count); cudaFree(ctx->d_launch2);
	cudaFree(ctx->d_counter); cudaFree(ctx->d_totals); cudaFree(ctx->d_launch);
	cudaEventDestroy(ctx->ev0); cudaEventDestroy(ctx->ev1); cudaStreamDestroy(ctx->stream);
	if (ctx->copy_stream) { cudaStreamDestroy(ctx->copy_stream); } cudaFree(ctx->d_avail); cudaFree(ctx->d_stalled);
	if (ctx->ev_avail) { cudaEventDestroy(ctx->ev_avail); } if (ctx->h_marks) { cudaFreeHost(ctx->h_marks); }
	delete ctx;
}

extern "C" uint64_t lb2_kernel_launches(const lb2_ctx *ctx) { return ctx ? ctx->launches : 0; }

// streamed = false: the whole batch is copied and the call returns when it is resident (lb2_upload).
// streamed = true (lb2_process): everything except the read pool's bases/qualities is enqueued; h_need[w] = number of
// leading pool reads the windows [0, w] use, so that the pool can follow in chunks while the kernel already runs.
static int lb2_upload_impl(lb2_ctx *ctx, const lb2_batch *b, bool streamed)
{
	if (!ctx || !b) { return LB2_ERR_ARG; }
	if (cudaSetDevice(ctx->device) != cudaSuccess) { return LB2_ERR_CUDA; }
	ctx->resident = false; ctx->ran = false;
	const uint32_t W = b->n_windows, R = b->n_reads;
	// staging bound: every read rounded up to 32 bases + reference + padding (untrimmed lengths)
	uint32_t max_bp = 0, max_reads = 0;
	if (streamed) { ctx->h_need.resize(W); }
	{
		// one pass over the windows' read lists (millions of entries for a 1 Mb region): split over a few host threads
		const unsigned hw = std::thread::hardware_concurrency();
		const unsigned T = (W >= 2048 && hw > 1) ? std::min<unsigned>(std::min<unsigned>(hw, 8u), env_u32("LB2_HOST_THREADS", 8)) : 1u;
		std::vector<uint32_t> t_bp(T, 0), t_rd(T, 0); std::vector<int> t_bad(T, 0);
		auto work = [&](unsigned t) {
			const uint32_t w0 = (uint32_t)((uint64_t)W * t / T), w1 = (uint32_t)((uint64_t)W * (t + 1) / T);
			uint32_t mbp = 0, mrd = 0;
			for (uint32_t w = w0; w < w1; ++w) {
				uint64_t bp = 0; uint32_t top = 0;
				for (uint32_t x = b->wr_off[w]; x < b->wr_off[w + 1]; ++x) {
					const uint32_t r = b->wr_idx[x]; if (r >= R) { t_bad[t] = 1; return; }
					bp += ((b->base_off[r + 1] - b->base_off[r]) + 15) & ~15ull;
					if (r >= top) { top = r + 1; }
				}
				if (streamed) { ctx->h_need[w] = top; }
				bp += ((b->ref_off[w + 1] - b->ref_off[w]) + 31) & ~31u; bp += 128;
				if (bp > mbp) { mbp = (uint32_t)std::min<uint64_t>(bp, 1u << 30); }
				mrd = std::max(mrd, b->wr_off[w + 1] - b->wr_off[w]);
			}
			t_bp[t] = mbp; t_rd[t] = mrd;
		};
		if (T == 1) { work(0); }
		else { std::vector<std::thread> th; for (unsigned t = 0; t < T; ++t) { th.emplace_back(work, t); } for (auto &x : th) { x.join(); } }
		for (unsigned t = 0; t < T; ++t) { if (t_bad[t]) { return LB2_ERR_ARG; } max_bp = std::max(max_bp, t_bp[t]); max_reads = std::max(max_reads, t_rd[t]); }
		if (streamed) { uint32_t need = 0; for (uint32_t w = 0; w < W; ++w) { need = std::max(need, ctx->h_need[w]); ctx->h_need[w] = need; } }      // leading pool reads the windows [0, w] use
	}
	max_bp = (max_bp + 1023) & ~1023u; if (max_bp < 32768) { max_bp = 32768; }
	const uint32_t smem_cap = 220u << 10;
	if (max_bp > (1u << 20) - 1024) { max_bp = (1u << 20) - 1024; }   // representative base index has 20 bits in the table key
	lb2_cfg &C = ctx->C;
	C.table_slots = std::min<uint32_t>(env_u32("LB2_TABLE_SLOTS", 4096), 16384u);      // (occurrence words hold 14-bit slot numbers)
	while (C.table_slots > 1024 && lb2_smem_bytes(max_bp, C.table_slots, C.graph_bytes) > smem_cap) { C.table_slots >>= 1; }
	while (lb2_smem_bytes(max_bp, C.table_slots, C.graph_bytes) > smem_cap) { max_bp -= 1024; }   // windows beyond this report LB2_WIN_OVERFLOW
	C.max_nodes = C.table_slots - C.table_slots / 4;
	C.max_bp = max_bp; C.smem_bytes = (uint32_t)lb2_smem_bytes(max_bp, C.table_slots, C.graph_bytes); C.max_reads = std::max(max_reads + 2, 64u);
	LB2_CK(cudaFuncSetAttribute(lb2_window_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C.smem_bytes));
	int occ = 0;
	LB2_CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, lb2_window_kernel, (int)ctx->threads, C.smem_bytes));
	if (occ < 1) { ctx->err = "kernel does not fit on an SM"; return LB2_ERR_CUDA; }
	uint32_t max_occ = env_u32("LB2_MAX_CTAS_PER_SM", 16);
	C.n_slots = (uint32_t)ctx->sm_count * std::min<uint32_t>((uint32_t)occ, max_occ);
	if (C.n_slots > std::max(W, 1u)) { C.n_slots = std::max(W, 1u); }
	size_t stride = lb2_ws_layout(C, nullptr, nullptr);
	if (stride != ctx->ws_stride || C.n_slots > ctx->ws_slots) {
		if (ctx->d_ws.p) { cudaFree(ctx->d_ws.p); ctx->d_ws.p = nullptr; ctx->d_ws.cap = 0; }
		size_t bytes = stride * C.n_slots;
		LB2_CK(cudaMalloc(&ctx->d_ws.p, bytes)); ctx->d_ws.cap = bytes;
		LB2_CK(cudaMemsetAsync(ctx->d_ws.p, 0, bytes, ctx->stream));   // the hash table must start all-zero
		ctx->ws_stride = stride; ctx->ws_slots = C.n_slots;
	}
	uint64_t h2d = 0;
#define LB2_UP(buf, ptr, bytes) do { int rc_ = lb2_reserve(ctx, ctx->buf, (bytes)); if (rc_) return rc_; \
		LB2_CK(cudaMemcpyAsync(ctx->buf.p, (ptr), (bytes), cudaMemcpyHostToDevice, ctx->stream)); h2d += (bytes); } while (0)
	// streamed: only the three per-window offset tables go ahead of the kernels, everything else follows in chunks (lb2_process)
#define LB2_UPS(buf, ptr, bytes) do { if (!streamed) { LB2_UP(buf, ptr, bytes); } else { int rc2_ = lb2_reserve(ctx, ctx->buf, (bytes)); if (rc2_) return rc2_; h2d += (bytes); } } while (0)
	LB2_UP(d_ref_off, b->ref_off, sizeof(uint32_t) * (size_t)(W + 1));
	LB2_UP(d_ref_start, b->ref_start, sizeof(int32_t) * (size_t)W);
	LB2_UP(d_wr_off, b->wr_off, sizeof(uint32_t) * (size_t)(W + 1));
	LB2_UPS(d_wr_idx, b->wr_idx, sizeof(uint32_t) * (size_t)b->n_wr);
	LB2_UPS(d_base_off, b->base_off, sizeof(uint64_t) * (size_t)(R + 1));
	LB2_UPS(d_flags, b->flags, (size_t)R);
	LB2_UPS(d_name_rank, b->name_rank, sizeof(uint32_t) * (size_t)R);
	LB2_UPS(d_ref_seq, b->ref_seq, (size_t)b->n_ref_bytes);
	LB2_UPS(d_seq, b->seq, (size_t)b->n_base_bytes);
	LB2_UPS(d_qual, b->qual, (size_t)b->n_base_bytes);
#undef LB2_UPS
#undef LB2_UP
	ctx->h2d_bytes = h2d;
	int rc;
	if ((rc = lb2_reserve(ctx, ctx->d_info, sizeof(lb2_window_info) * (size_t)W))) return rc;
	if ((rc = lb2_reserve(ctx, ctx->d_vars, sizeof(lb2_variant) * (size_t)W * C.max_var))) return rc;
	if ((rc = lb2_reserve(ctx, ctx->d_strs, (size_t)W * C.str_bytes))) return rc;
	if ((rc = lb2_reserve(ctx, ctx->d_str_used, sizeof(uint32_t) * (size_t)W))) return rc;
	if ((rc = lb2_reserve(ctx, ctx->d_var_off, sizeof(uint32_t) * (size_t)W))) return rc;
	if ((rc = lb2_reserve(ctx, ctx->d_str_off, sizeof(uint32_t) * (size_t)W))) return rc;
	const uint32_t nbig = ctx->escalate ? std::min<uint32_t>(ctx->big_cap, std::max(W, 1u)) : 0u;
	if ((rc = lb2_reserve(ctx, ctx->d_cvars, sizeof(lb2_variant) * ((size_t)W * C.max_var + (size_t)nbig * ctx->big_max_var)))) return rc;
	if ((rc = lb2_reserve(ctx, ctx->d_cstr, (size_t)W * C.str_bytes + (size_t)nbig * ctx->big_str_bytes))) return rc;
	if ((rc = lb2_reserve(ctx, ctx->d_big_vars, sizeof(lb2_variant) * (size_t)nbig * ctx->big_max_var + 64))) return rc;
	if ((rc = lb2_reserve(ctx, ctx->d_big_strs, (size_t)nbig * ctx->big_str_bytes + 64))) return rc;
	if ((rc = lb2_reserve(ctx, ctx->d_big_slot, sizeof(uint32_t) * (size_t)(W + 1)))) return rc;
	lb2_launch &L = ctx->L;
	L.P = ctx->P; L.C = C;
	L.B.n_windows = W; L.B.ref_off = (const uint32_t *)ctx->d_ref_off.p; L.B.ref_start = (const int32_t *)ctx->d_ref_start.p;
	L.B.wr_off = (const uint32_t *)ctx->d_wr_off.p; L.B.wr_idx = (const uint32_t *)ctx->d_wr_idx.p;
	L.B.base_off = (const uint64_t *)ctx->d_base_off.p; L.B.flags = (const uint8_t *)ctx->d_flags.p;
	L.B.name_rank = (const uint32_t *)ctx->d_name_rank.p; L.B.ref_seq = (const char *)ctx->d_ref_seq.p;
	L.B.seq = (const char *)ctx->d_seq.p; L.B.qual = (const char *)ctx->d_qual.p;
	L.O.info = (lb2_window_info *)ctx->d_info.p; L.O.variants = (lb2_variant *)ctx->d_vars.p; L.O.strings = (char *)ctx->d_strs.p;
	L.O.str_used = (uint32_t *)ctx->d_str_used.p; L.O.prof = ctx->d_prof;
	L.O.big_variants = (lb2_variant *)ctx->d_big_vars.p; L.O.big_strings = (char *)ctx->d_big_strs.p; L.O.big_slot = (uint32_t *)ctx->d_big_slot.p;
	L.O.big_count = ctx->d_big_count; L.O.big_cap = nbig; L.O.big_max_var = ctx->big_max_var; L.O.big_str_bytes = ctx->big_str_bytes;
	L.ws_base = (uint8_t *)ctx->d_ws.p; L.ws_stride = ctx->ws_stride; L.counter = ctx->d_counter;
	L.win_list = nullptr; L.n_list = nullptr; L.avail = streamed ? ctx->d_avail : nullptr; L.stalled = ctx->d_stalled;
	if ((rc = lb2_reserve(ctx, ctx->d_retry, sizeof(uint32_t) * (size_t)(W + 1)))) return rc;
	L.retry_list = (uint32_t *)ctx->d_retry.p; L.retry_count = ctx->d_retry_count;
	L.var_off = (uint32_t *)ctx->d_var_off.p; L.str_off = (uint32_t *)ctx->d_str_off.p; L.totals = ctx->d_totals;
	L.cvars = (lb2_variant *)ctx->d_cvars.p; L.cstr = (char *)ctx->d_cstr.p;
	LB2_CK(cudaMemcpyAsync(ctx->d_launch, &L, sizeof L, cudaMemcpyHostToDevice, ctx->stream));
	if (ctx->escalate) {
		// escalation pass: the largest table that still fits beside the staged reads, big arena / BFS queue, one CTA per SM
		lb2_cfg &C2 = ctx->C2; C2 = C;
		C2.table_slots = std::min<uint32_t>(env_u32("LB2_TABLE_SLOTS2", 16384), 16384u); C2.graph_bytes = env_u32("LB2_GRAPH_BYTES2", 184u << 10);
		while (C2.graph_bytes > C.graph_bytes && lb2_smem_bytes(max_bp, C2.table_slots, C2.graph_bytes) > smem_cap) { C2.graph_bytes -= 4096; }
		while (C2.table_slots > C.table_slots && lb2_smem_bytes(max_bp, C2.table_slots, C2.graph_bytes) > smem_cap) { C2.table_slots >>= 1; }
		C2.max_nodes = C2.table_slots - C2.table_slots / 4;
		C2.smem_bytes = (uint32_t)lb2_smem_bytes(max_bp, C2.table_slots, C2.graph_bytes);
		C2.arena_bytes = env_u32("LB2_ARENA_BYTES2", 8u << 20); C2.deficit_bytes = env_u32("LB2_DEFICIT_BYTES2", 16u << 20);
		C2.queue_cap = env_u32("LB2_QUEUE_CAP2", 1u << 22); C2.max_inst = env_u32("LB2_MAX_INST2", 1u << 20); C2.max_special = env_u32("LB2_MAX_SPECIAL2", 2048);
		C2.n_slots = (uint32_t)std::min<uint32_t>((uint32_t)ctx->sm_count, env_u32("LB2_SLOTS2", 64));
		size_t stride2 = lb2_ws_layout(C2, nullptr, nullptr);
		if (stride2 != ctx->ws2_stride || C2.n_slots > ctx->ws2_slots) {
			if (ctx->d_ws2.p) { cudaFree(ctx->d_ws2.p); ctx->d_ws2.p = nullptr; }
			LB2_CK(cudaMalloc(&ctx->d_ws2.p, stride2 * C2.n_slots)); ctx->d_ws2.cap = stride2 * C2.n_slots;
			ctx->ws2_stride = stride2; ctx->ws2_slots = C2.n_slots;
		}
		lb2_launch &L2 = ctx->L2; L2 = L; L2.C = C2; L2.avail = nullptr;      // (the first pass ends after the last pool chunk has arrived)
		L2.ws_base = (uint8_t *)ctx->d_ws2.p; L2.ws_stride = stride2; L2.counter = ctx->d_counter2;
		L2.win_list = (const uint32_t *)ctx->d_retry.p; L2.n_list = ctx->d_retry_count;
		LB2_CK(cudaMemcpyAsync(ctx->d_launch2, &L2, sizeof L2, cudaMemcpyHostToDevice, ctx->stream));
		LB2_CK(cudaFuncSetAttribute(lb2_window_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max(C.smem_bytes, C2.smem_bytes)));
	}
	if (!streamed) { LB2_CK(cudaStreamSynchronize(ctx->stream)); }
	ctx->n_windows = W; ctx->resident = true;
	return LB2_OK;
}

extern "C" int lb2_upload(lb2_ctx *ctx, const lb2_batch *b) { return lb2_upload_impl(ctx, b, false); }

extern "C" int lb2_run(lb2_ctx *ctx)
{
	if (!ctx) { return LB2_ERR_ARG; }
	if (!ctx->resident) { return LB2_ERR_STATE; }
	if (cudaSetDevice(ctx->device) != cudaSuccess) { return LB2_ERR_CUDA; }
	const uint32_t W = ctx->n_windows;
	LB2_CK(cudaMemsetAsync(ctx->d_counter, 0, 4, ctx->stream));
	LB2_CK(cudaMemsetAsync(ctx->d_big_count, 0, 4, ctx->stream));
	LB2_CK(cudaMemsetAsync(ctx->d_big_slot.p, 0xFF, sizeof(uint32_t) * (size_t)(W + 1), ctx->stream));
	LB2_CK(cudaEventRecord(ctx->ev0, ctx->stream));
	if (W) {
		lb2_window_kernel<<<ctx->C.n_slots, ctx->threads, ctx->C.smem_bytes, ctx->stream>>>(ctx->d_launch);
		if (ctx->escalate) {
			LB2_CK(cudaMemsetAsync(ctx->d_counter2, 0, 4, ctx->stream)); LB2_CK(cudaMemsetAsync(ctx->d_retry_count, 0, 4, ctx->stream));
			lb2_collect_kernel<<<64, 256, 0, ctx->stream>>>(ctx->d_launch);
			lb2_window_kernel<<<ctx->C2.n_slots, ctx->threads, ctx->C2.smem_bytes, ctx->stream>>>(ctx->d_launch2);
			ctx->launches += 2;
		}
		lb2_scan_kernel<<<1, 1024, 0, ctx->stream>>>(ctx->d_launch);
		lb2_gather_kernel<<<W, 64, 0, ctx->stream>>>(ctx->d_launch);
		ctx->launches += 3;
	}
	LB2_CK(cudaEventRecord(ctx->ev1, ctx->stream));
	LB2_CK(cudaGetLastError());
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
		LB2_CK(cudaMemcpyAsync(totals, ctx->d_totals, 8, cudaMemcpyDeviceToHost, ctx->stream));
		LB2_CK(cudaMemcpyAsync(ctx->h_info.data(), ctx->d_info.p, sizeof(lb2_window_info) * (size_t)W, cudaMemcpyDeviceToHost, ctx->stream));
	}
	LB2_CK(cudaStreamSynchronize(ctx->stream));
	ctx->h_vars.resize(totals[0]); ctx->h_str.resize(totals[1]);
	if (totals[0]) { LB2_CK(cudaMemcpyAsync(ctx->h_vars.data(), ctx->d_cvars.p, sizeof(lb2_variant) * (size_t)totals[0], cudaMemcpyDeviceToHost, ctx->stream)); }
	if (totals[1]) { LB2_CK(cudaMemcpyAsync(ctx->h_str.data(), ctx->d_cstr.p, totals[1], cudaMemcpyDeviceToHost, ctx->stream)); }
	LB2_CK(cudaStreamSynchronize(ctx->stream));
	float ms = 0; cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1); ctx->kernel_ms = ms;
	ctx->d2h_bytes = 8 + sizeof(lb2_window_info) * (size_t)W + sizeof(lb2_variant) * (size_t)totals[0] + totals[1];
	res->n_windows = W; res->n_variants = totals[0]; res->windows = ctx->h_info.data(); res->variants = ctx->h_vars.data();
	res->strings = ctx->h_str.data(); res->n_string_bytes = totals[1]; res->kernel_ms = ms;
	return LB2_OK;
}

// Host buffers in, host buffers out.  The window tables go first, the kernels are enqueued behind them, and the read
// pool (bases + qualities, ~85 % of the bytes) follows on a second stream in chunks; after every chunk a watermark in
// device memory tells the running kernel how many windows have all their reads in HBM.  Chunk boundaries are 128-byte
// aligned and a window only counts as ready once 128 bytes past its last read have arrived, so no cache line (and no
// 16-byte over-read of the staging loads) is ever touched before it is complete.
extern "C" int lb2_process(lb2_ctx *ctx, const lb2_batch *batch, lb2_result *result)
{
	if (!ctx || !batch || !result) { return LB2_ERR_ARG; }
	if (cudaSetDevice(ctx->device) != cudaSuccess) { return LB2_ERR_CUDA; }
	{	// copies from pageable memory do not overlap a running kernel (a kernel waiting for them would wait forever):
		// the pool is only streamed from page-locked buffers, otherwise the batch is made resident first
		bool pinned = batch->n_base_bytes > 0 && batch->n_windows > 0 && env_u32("LB2_STREAM", 1) != 0;
		// tools that make kernel launches synchronous would park the kernel in front of the copies it waits for
		if ((getenv("CUDA_INJECTION64_PATH") || getenv("NV_COMPUTE_PROFILER_PERFWORKS_DIR") || getenv("NSYS_PROFILING_SESSION_ID") || env_u32("CUDA_LAUNCH_BLOCKING", 0)) &&
		    !env_u32("LB2_STREAM_FORCE", 0)) { pinned = false; }      // (LB2_STREAM_FORCE: test hook for the time-out path below)
		const void *arrs[] = { batch->seq, batch->qual, batch->base_off, batch->flags, batch->name_rank, batch->wr_idx, batch->ref_seq };
		const uint64_t sizes[] = { batch->n_base_bytes, batch->n_base_bytes, 1, batch->n_reads, batch->n_reads, batch->n_wr, batch->n_ref_bytes };
		for (int i = 0; i < 7 && pinned; ++i) {
			if (!sizes[i]) { continue; }
			cudaPointerAttributes at; if (cudaPointerGetAttributes(&at, arrs[i]) != cudaSuccess || at.type != cudaMemoryTypeHost) { pinned = false; }
		}
		cudaGetLastError();
		if (!pinned) {
			int rc0 = lb2_upload_impl(ctx, batch, false); if (rc0) { return rc0; }
			rc0 = lb2_run(ctx); if (rc0) { return rc0; }
			return lb2_download(ctx, result);
		}
	}
	const bool timing = env_u32("LB2_TIMING", 0) != 0; const auto t0 = std::chrono::steady_clock::now();
	// the watermark starts at 0 before the kernel may look at it (same stream as the updates that follow)
	LB2_CK(cudaMemsetAsync(ctx->d_avail, 0, 4, ctx->copy_stream));
	LB2_CK(cudaMemsetAsync(ctx->d_stalled, 0, 4, ctx->copy_stream));
	LB2_CK(cudaEventRecord(ctx->ev_avail, ctx->copy_stream));
	LB2_CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_avail, 0));
	int rc = lb2_upload_impl(ctx, batch, true); if (rc) { return rc; }
	const uint32_t W = ctx->n_windows; const uint64_t nb = batch->n_base_bytes;
	const auto t1 = std::chrono::steady_clock::now();
	rc = lb2_run(ctx); if (rc) { return rc; }
	// chunks of consecutive windows: everything the windows [wa, wb) read -- their reference bases, their read lists, and
	// the leading part of the pool they use (per-read tables, bases, qualities) -- then the watermark wb.  Every array's
	// upload boundary is a multiple of 128 bytes, at least 128 bytes past the last byte the ready windows touch.
	const uint32_t R = batch->n_reads; const uint64_t n_wr = batch->n_wr, n_ref = batch->n_ref_bytes;
	const uint64_t chunk_max = std::max<uint64_t>((uint64_t)env_u32("LB2_STREAM_CHUNK", 8u << 20), 1u << 16);
	uint64_t chunk = std::min<uint64_t>(chunk_max, 1u << 20);      // short chunks first: the kernel starts on the first one
	uint64_t up_by = 0, up_rd = 0, up_bo = 0, up_wr = 0, up_ref = 0; uint32_t wa = 0; size_t c = 0;
	auto up128 = [](uint64_t x, uint64_t unit, uint64_t cap) { const uint64_t per = 128 / unit; x = (x + per - 1) / per * per; return x > cap ? cap : x; };
#define LB2_PIECE(buf, ptr, esz, from, to) do { if ((to) > (from)) { LB2_CK(cudaMemcpyAsync((char *)ctx->buf.p + (size_t)(from) * (esz), (const char *)(ptr) + (size_t)(from) * (esz), \
		(size_t)((to) - (from)) * (esz), cudaMemcpyHostToDevice, ctx->copy_stream)); } } while (0)
	while (wa < W) {
		uint32_t wb = wa + 1;
		if (c + 2 >= LB2_MAX_MARKS) { wb = W; }
		else { const uint64_t lim = up_by + chunk; while (wb < W && batch->base_off[ctx->h_need[wb]] <= lim) { ++wb; } }
		const uint64_t rd_t = ctx->h_need[wb - 1];
		const uint64_t to_by = up128(batch->base_off[rd_t] + 128, 1, nb), to_rd = up128(rd_t + 1, 1, R), to_bo = up128(rd_t + 2, 8, (uint64_t)R + 1);
		const uint64_t to_wr = up128((uint64_t)batch->wr_off[wb] + 32, 4, n_wr), to_ref = up128((uint64_t)batch->ref_off[wb] + 128, 1, n_ref);
		LB2_PIECE(d_ref_seq, batch->ref_seq, 1, up_ref, to_ref); if (to_ref > up_ref) { up_ref = to_ref; }
		LB2_PIECE(d_wr_idx, batch->wr_idx, 4, up_wr, to_wr); if (to_wr > up_wr) { up_wr = to_wr; }
		LB2_PIECE(d_base_off, batch->base_off, 8, up_bo, to_bo); if (to_bo > up_bo) { up_bo = to_bo; }
		LB2_PIECE(d_flags, batch->flags, 1, up_rd, to_rd);
		LB2_PIECE(d_name_rank, batch->name_rank, 4, up_rd, to_rd); if (to_rd > up_rd) { up_rd = to_rd; }
		LB2_PIECE(d_seq, batch->seq, 1, up_by, to_by);
		LB2_PIECE(d_qual, batch->qual, 1, up_by, to_by); if (to_by > up_by) { up_by = to_by; }
		ctx->h_marks[c] = wb;
		LB2_CK(cudaMemcpyAsync(ctx->d_avail, &ctx->h_marks[c], 4, cudaMemcpyHostToDevice, ctx->copy_stream));
		++c; wa = wb; chunk = std::min<uint64_t>(chunk_max, chunk * 2);
	}
#undef LB2_PIECE
	const auto t2 = std::chrono::steady_clock::now();
	LB2_CK(cudaStreamSynchronize(ctx->copy_stream));      // the caller's buffers are free again when the call returns
	{	// safety net: a window fetch that gave up on the watermark => the whole batch again, resident (it is by now)
		uint32_t stalled = 0;
		LB2_CK(cudaStreamSynchronize(ctx->stream));
		LB2_CK(cudaMemcpy(&stalled, ctx->d_stalled, 4, cudaMemcpyDeviceToHost));
		if (stalled) {
			ctx->L.avail = nullptr;
			LB2_CK(cudaMemcpyAsync(ctx->d_launch, &ctx->L, sizeof ctx->L, cudaMemcpyHostToDevice, ctx->stream));
			rc = lb2_run(ctx); if (rc) { return rc; }
		}
	}
	if (timing) {
		const auto t3 = std::chrono::steady_clock::now(); rc = lb2_download(ctx, result); const auto t4 = std::chrono::steady_clock::now();
		auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b_) { return std::chrono::duration<double, std::milli>(b_ - a).count(); };
		fprintf(stderr, "lb2_process: tables+config %.2f ms, enqueue %.2f ms, pool copies done +%.2f ms, kernels+download +%.2f ms, total %.2f ms (kernels alone %.2f ms)\n", ms(t0, t1), ms(t1, t2), ms(t2, t3), ms(t3, t4), ms(t0, t4), (double)result->kernel_ms);
		return rc;
	}
	return lb2_download(ctx, result);
}

// debugging aid: cycles per pipeline phase (lane 0 of every CTA), summed since the context was created
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
extern "C" float lb2_last_kernel_ms(lb2_ctx *ctx) { if (!ctx) return 0; float ms = 0; cudaEventSynchronize(ctx->ev1); cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1); return ms; }

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
