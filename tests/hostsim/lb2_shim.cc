// DEBUG-ONLY implementation of the subset of the C ABI the CLI uses, on top of the one-thread simulation
// of the device sources (see hostsim.cc for what that is and is not).  It lets `lancet_cli.cc` be linked and
// run end to end in the GPU-less development container; the shipped binary links lancet_b200/_lb2.so instead.
#define LB2_HOSTSIM 1
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <numeric>
#include <algorithm>
#include "../../lancet_b200/csrc/lb2_pipeline.cuh"
#include "sim_pack.h"

struct lb2_ctx { lb2_params P; std::vector<lb2_window_info> info; std::vector<lb2_variant> vars; std::vector<char> strs; };

extern "C" void lb2_default_params(lb2_params *p)
{
	p->min_k = 11; p->max_k = 101; p->min_qual_trim = 43; p->min_qual_call = 50; p->cov_threshold = 5; p->low_cov_threshold = 1; p->max_tip_len = 11; p->dfs_limit = 1000000;
	p->max_indel_len = 500; p->max_mismatch = 2; p->max_unit_len = 4; p->min_report_units = 3; p->min_report_len = 7; p->dist_from_str = 1; p->min_cov_ratio = 0.01;
}
extern "C" const char *lb2_strerror(const lb2_ctx *, int code) { return code == LB2_OK ? "ok" : "hostsim shim error"; }
extern "C" int lb2_create(lb2_ctx **out, const lb2_params *p, int) { *out = new lb2_ctx; (*out)->P = *p; return LB2_OK; }
extern "C" void lb2_destroy(lb2_ctx *c) { delete c; }
extern "C" int lb2_rank_names(const char *const *names, uint32_t n, uint32_t *rank_out)
{
	std::vector<uint32_t> idx(n); std::iota(idx.begin(), idx.end(), 0u);
	std::sort(idx.begin(), idx.end(), [&](uint32_t a, uint32_t b) { return strcmp(names[a], names[b]) < 0; });
	uint32_t r = 0;
	for (uint32_t i = 0; i < n; ++i) { if (i && strcmp(names[idx[i]], names[idx[i - 1]]) != 0) { ++r; } rank_out[idx[i]] = r; }
	return LB2_OK;
}
extern "C" int lb2_process(lb2_ctx *ctx, const lb2_batch *b, lb2_result *res)
{
	const uint32_t W = b->n_windows;
	lb2_cfg C; memset(&C, 0, sizeof C);
	C.table_slots = 16384; C.max_nodes = 12000; C.max_reads = 8192; C.max_bp = (1 << 20) - 1024; C.arena_bytes = 1 << 21; C.deficit_bytes = 1 << 23;
	C.queue_cap = 1 << 22; C.graph_bytes = 1 << 20; C.max_inst = 1 << 20; C.max_var = 64; C.str_bytes = 8192; C.bucket_cap = 10273; C.max_k = 127; C.n_slots = 1; C.max_special = 2048;
	lb2_dev_batch B; B.n_windows = W; B.ref_off = b->ref_off; B.ref_start = b->ref_start; B.wr_off = b->wr_off; B.wr_idx = b->wr_idx;
	B.base_off = b->base_off; B.flags = b->flags; B.name_rank = b->name_rank; B.ref_seq = b->ref_seq; std::vector<char> pseq(b->seq, b->seq + b->n_base_bytes), pqual(b->qual, b->qual + b->n_base_bytes); pseq.resize(pseq.size() + 64, 0); pqual.resize(pqual.size() + 64, 0);
	B.seq = pseq.data(); B.qual = pqual.data();
	SimPack sp; sim_pack(B, b->n_reads, ctx->P, sp);
	std::vector<lb2_window_info> info(W); std::vector<lb2_variant> vars((size_t)W * C.max_var); std::vector<char> strs((size_t)W * C.str_bytes); std::vector<uint32_t> sused(W);
	lb2_dev_out O; memset(&O, 0, sizeof O); O.info = info.data(); O.variants = vars.data(); O.strings = strs.data(); O.str_used = sused.data();
	size_t wsb = lb2_ws_layout(C, NULL, NULL);
	std::vector<uint8_t> slab(wsb, 0); std::vector<uint8_t> smem(lb2_smem_bytes(C.max_bp, C.table_slots, C.graph_bytes) + 64, 0);
	lb2_win Wn; Wn.P = &ctx->P; Wn.C = &C; Wn.B = &B; Wn.O = &O; Wn.escal = false;
	lb2_ws_layout(C, slab.data(), &Wn.ws); Wn.ws0 = Wn.ws;
	Wn.sh = (lb2_sh *)smem.data();
	Wn.ref_raw = (char *)smem.data() + ((sizeof(lb2_sh) + 15) & ~(size_t)15);
	Wn.bits = (uint32_t *)(Wn.ref_raw + LB2_MAX_REF);
	Wn.lowq = Wn.bits + (C.max_bp / 16 + 4);
	Wn.treg = smem.data() + ((lb2_smem_fixed(C.max_bp) + 15) & ~(size_t)15);
	ctx->info.assign(W, lb2_window_info()); ctx->vars.clear(); ctx->strs.clear();
	for (uint32_t w = 0; w < W; ++w) {
		lb2_process_window(Wn, w);
		ctx->info[w] = info[w];
		for (uint32_t v = 0; v < info[w].n_variants; ++v) {
			lb2_variant x = vars[(size_t)w * C.max_var + v]; const char *sp = strs.data() + (size_t)w * C.str_bytes + x.str_off;
			uint32_t n = (uint32_t)x.ref_len + x.alt_len + x.motif_len;
			x.str_off = (uint32_t)ctx->strs.size(); x.window = w; ctx->strs.insert(ctx->strs.end(), sp, sp + n); ctx->vars.push_back(x);
		}
	}
	res->n_windows = W; res->n_variants = (uint32_t)ctx->vars.size(); res->windows = ctx->info.data(); res->variants = ctx->vars.data();
	res->strings = ctx->strs.data(); res->n_string_bytes = ctx->strs.size(); res->kernel_ms = 0;
	return LB2_OK;
}

// page-locked memory and the multi-rank gather have no meaning in the one-thread simulation
extern "C" void *lb2_alloc_pinned(size_t bytes) { return malloc(bytes ? bytes : 1); }
extern "C" void lb2_free_pinned(void *p) { free(p); }
extern "C" int lb2_comm_unique_id(char *) { return LB2_ERR_CUDA; }
extern "C" int lb2_comm_init(lb2_ctx *, const char *, int, int) { return LB2_ERR_CUDA; }
extern "C" int lb2_comm_gather(lb2_ctx *, const lb2_variant *, uint32_t, const char *, uint64_t, uint64_t *, int, int, lb2_result *) { return LB2_ERR_CUDA; }
