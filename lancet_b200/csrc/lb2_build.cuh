// lb2_build.cuh -- read staging (trim + 2-bit pack into shared memory) and the parallel
// coloured de Bruijn graph build (Mer -> Node open-addressing hash insert).
//
// Restates, as an order-independent parallel computation, what the reference does sequentially in
//   Graph_t::trim            src/Graph.cc:355-384   (5'/3' quality / non-ACGT trim, junk flag)
//   Graph_t::buildgraph      src/Graph.cc:530-589   (reads in order T.., N.., reference LAST)
//   Graph_t::loadSequence    src/Graph.cc:119-349   (canonical k-mers, find-or-insert, colours,
//                                                    strand x sample counts, min-QV counts, edges)
//   Node_t::addEdge          src/Node.cc:140-175    (edge vector in first-seen order)
// Sequential facts the reference derives from processing order are recovered from per-key
// minima: node insertion order = order of first occurrence (read, offset); edge order = first
// time the (target,dir) pair was added.
#ifndef LB2_BUILD_CUH
#define LB2_BUILD_CUH

#include "lb2_state.cuh"

struct lb2_win {
	const lb2_params *P; const lb2_cfg *C; const lb2_dev_batch *B; const lb2_dev_out *O;
	lb2_ws ws; lb2_sh *sh;
	uint32_t *bits;      // smem: 2-bit packed trimmed reads, then the window reference
	uint32_t *lowq;      // smem: 1 bit per staged base: quality < MIN_QUAL_CALL
	char     *ref_raw;   // smem: window reference, ASCII
};

LB2_DEVNI void lb2_sort64(uint64_t *a, uint32_t n2);

// attribute the cycles since the previous mark to phase ph (lane 0 only)
LB2_DEV void lb2_mark(lb2_win &W, int ph) {
	if (lb2_tid() == 0) { unsigned long long t = lb2_clock(); W.sh->prof[ph] += t - W.sh->t_last; W.sh->t_last = t; }
}

LB2_DEV void lb2_fail(lb2_win &W, uint32_t status, uint32_t detail) {
	if (lb2_cas32(&W.sh->status, LB2_WIN_OK, status) == LB2_WIN_OK) { W.sh->detail = detail; }
}

// ---------------------------------------------------------------------------------------------
// stage the window: trim every read, pack bases/low-quality mask into shared memory
// ---------------------------------------------------------------------------------------------
LB2_DEVNI void lb2_stage_window(lb2_win &W, uint32_t w)
{
	lb2_sh *sh = W.sh; const lb2_dev_batch *B = W.B; lb2_ws &ws = W.ws;
	const unsigned tid = lb2_tid(), nt = lb2_nthr();
	if (tid == 0) {
		sh->w = w; sh->status = LB2_WIN_OK; sh->detail = 0;
		sh->L = B->ref_off[w + 1] - B->ref_off[w];
		sh->R = B->wr_off[w + 1] - B->wr_off[w];
		sh->ref_start = B->ref_start[w];
		sh->has_lowq = 0; sh->has_pairs = 0; sh->mapped = 0; sh->flag_a = 0; sh->totalreadbp = 0;
		sh->n_var = 0; sh->str_used = 0; sh->n_k_tried = 0; sh->final_k = 0; sh->last_nodes = 0;
		if (sh->L > LB2_MAX_REF) { sh->status = LB2_WIN_OVERFLOW; sh->detail = LB2_D_REFLEN; }
		else if (sh->R + 1 > W.C->max_reads) { sh->status = LB2_WIN_OVERFLOW; sh->detail = LB2_D_READS; }
	}
	lb2_sync();
	if (sh->status != LB2_WIN_OK) { return; }
	const uint32_t L = sh->L, R = sh->R;
	const char *ref = B->ref_seq + B->ref_off[w];
	for (uint32_t i = tid; i < L; i += nt) {
		char c = ref[i]; W.ref_raw[i] = c;
		if (lb2_code(c) < 0) { lb2_or32(&sh->flag_a, 1u); }
	}
	const uint32_t *widx = B->wr_idx + B->wr_off[w];
	const int qtrim = W.P->min_qual_trim;
	for (uint32_t r = tid; r < R; r += nt) {
		uint32_t idx = widx[r];
		uint64_t o0 = B->base_off[idx]; int len = (int)(B->base_off[idx + 1] - o0);
		const char *s = B->seq + o0; const char *q = B->qual + o0;
		uint8_t fl = B->flags[idx];
		// Graph_t::trim (src/Graph.cc:355-384)
		int t5 = 0, t3 = 0; bool junk = false;
		while (t5 < len && (lb2_code(s[t5]) < 0 || q[t5] < qtrim)) { ++t5; }
		if (t5 < len) {
			while (t3 < len && (lb2_code(s[len - 1 - t3]) < 0 || q[len - 1 - t3] < qtrim)) { ++t3; }
			for (int i = t5; i < len - t3; ++i) { if (lb2_code(s[i]) < 0) { junk = true; break; } }
		} else { junk = true; }
		int n = junk ? 0 : (len - t5 - t3);
		if (n > 4095) { lb2_fail(W, LB2_WIN_UNSUPPORTED, LB2_D_READS); n = 0; }
		ws.rd_len[r] = (uint32_t)n; ws.rd_t5[r] = (uint32_t)t5;
		uint32_t cls = ((fl & LB2_READ_NORMAL) ? 2u : 0u) | ((fl & LB2_READ_REVERSE) ? 1u : 0u);
		uint32_t mate = (fl >> LB2_READ_MATE_SHIFT) & 3u;
		ws.rd_info[r] = cls | (mate << 2);
		ws.rd_rank[r] = B->name_rank[idx];
		if (!(fl & LB2_READ_UNMAPPED)) { lb2_or32(&sh->mapped, 1u); }
		if (n) { lb2_add32(&sh->totalreadbp, (uint32_t)n); }
	}
	lb2_sync();
	if (tid == 0) {
		uint32_t cum = 0;
		for (uint32_t r = 0; r < R; ++r) { ws.rd_start[r] = cum; cum += (ws.rd_len[r] + 31u) & ~31u; }
		ws.rd_start[R] = cum; sh->ref_g = cum; cum += (L + 31u) & ~31u;
		sh->total_bp = cum;
		if (cum + 64 > W.C->max_bp) { lb2_fail(W, LB2_WIN_OVERFLOW, LB2_D_SMEM); }
		if (sh->flag_a) { lb2_fail(W, LB2_WIN_UNSUPPORTED, LB2_D_NREF); }
		if (!sh->mapped) { lb2_fail(W, LB2_WIN_NO_READS, 0); }
		sh->seq_off = 0; sh->seq_len = L; sh->trim5 = 0; sh->trim3 = 0;
	}
	lb2_sync();
	if (sh->status != LB2_WIN_OK) { return; }
	{	// does any query name occur with both mate orders?  (otherwise hasOverlappingMate can never fire)
		uint32_t r2 = 1; while (r2 < R) { r2 <<= 1; }
		for (uint32_t r = tid; r < r2; r += nt) {
			uint64_t k = ~0ull;
			if (r < R) { uint32_t mate = (ws.rd_info[r] >> 2) & 3u; if (mate == 1 || mate == 2) { k = ((uint64_t)ws.rd_rank[r] << 2) | mate; } }
			ws.sortk[r] = k;
		}
		lb2_sync();
		lb2_sort64(ws.sortk, r2);
		for (uint32_t r = tid + 1; r < r2; r += nt) {
			uint64_t a = ws.sortk[r - 1], b = ws.sortk[r];
			if (b != ~0ull && (a >> 2) == (b >> 2) && (a & 3) != (b & 3)) { sh->has_pairs = 1; }
		}
		lb2_sync();
	}
	const int qcall = W.P->min_qual_call;
	for (uint32_t r = tid; r < R; r += nt) {
		uint32_t n = ws.rd_len[r]; if (!n) { continue; }
		uint32_t idx = widx[r];
		uint64_t o0 = B->base_off[idx] + ws.rd_t5[r];
		const char *s = B->seq + o0; const char *q = B->qual + o0;
		uint32_t g = ws.rd_start[r];
		uint32_t anylow = 0;
		for (uint32_t b0 = 0; b0 < n; b0 += 32) {
			uint64_t bw = 0; uint32_t lw = 0;
			uint32_t m = (n - b0 < 32) ? (n - b0) : 32;
			for (uint32_t i = 0; i < m; ++i) {
				bw |= (uint64_t)lb2_code(s[b0 + i]) << (2 * i);
				if (q[b0 + i] < qcall) { lw |= 1u << i; }
			}
			W.bits[(g + b0) >> 4] = (uint32_t)bw; W.bits[((g + b0) >> 4) + 1] = (uint32_t)(bw >> 32);
			W.lowq[(g + b0) >> 5] = lw; anylow |= lw;
		}
		if (anylow) { lb2_or32(&sh->has_lowq, 1u); }
	}
	{
		uint32_t g = sh->ref_g;
		for (uint32_t b0 = tid * 32; b0 < ((L + 31u) & ~31u) + 64; b0 += nt * 32) {
			uint64_t bw = 0;
			for (uint32_t i = 0; i < 32 && b0 + i < L; ++i) { bw |= (uint64_t)(lb2_code(W.ref_raw[b0 + i]) & 3) << (2 * i); }
			W.bits[(g + b0) >> 4] = (uint32_t)bw; W.bits[((g + b0) >> 4) + 1] = (uint32_t)(bw >> 32);
			W.lowq[(g + b0) >> 5] = 0;
		}
	}
	lb2_sync();
}

// ---------------------------------------------------------------------------------------------
// hash table: slot word = (fingerprint32 | 0x80000000) << 32 | (g << 1 | ori) where g is the
// staged base index of one representative occurrence; matches are verified against the bases.
// ---------------------------------------------------------------------------------------------
LB2_DEV uint32_t lb2_find_or_insert(lb2_win &W, const lb2_kmer &canon, const lb2_kmer &nonc, uint32_t rep, int K, int nw, bool insert)
{
	lb2_ws &ws = W.ws; const uint32_t mask = W.C->hash_cap - 1;
	uint64_t h = lb2_table_hash(canon, nw);
	uint32_t i = (uint32_t)h & mask;
	uint64_t fp = (h >> 32) | 0x80000000ull;
	for (uint32_t probes = 0; probes <= mask; ++probes) {
		uint64_t cur = lb2_ld64(&ws.slots[i]);
		if (cur == 0) {
			if (!insert) { return LB2_NIL; }
			uint64_t mine = (fp << 32) | rep;
			uint64_t prev = lb2_cas64(&ws.slots[i], 0ull, mine);
			if (prev == 0) {
				uint32_t u = lb2_add32(&W.sh->n_used, 1u);
				if (u < W.C->max_nodes) { ws.used[u] = i; } else { lb2_or32(&W.sh->err, 1u << LB2_D_NODES); }
				return i;
			}
			cur = prev;
		}
		if ((cur >> 32) == fp) {
			uint32_t r = (uint32_t)cur;
			lb2_kmer o; lb2_extract(W.bits, r >> 1, K, o);
			if (lb2_equal(o, (r & 1) ? nonc : canon, nw)) { return i; }
		}
		i = (i + 1) & mask;
	}
	lb2_or32(&W.sh->err, 1u << LB2_D_HASH_FULL);
	return LB2_NIL;
}

LB2_DEV void lb2_add_edge(lb2_win &W, uint32_t slot, uint32_t to, uint32_t dir, uint32_t seq)
{
	lb2_ws &ws = W.ws;
	uint32_t key = ((to << 2) | dir) + 1u;
	uint32_t inv = 0xFFFFFFFFu - seq;
	uint32_t *ek = ws.ekey + (size_t)slot * LB2_ECAP; uint32_t *es = ws.eseq + (size_t)slot * LB2_ECAP;
	for (int e = 0; e < LB2_ECAP; ++e) {
		uint32_t cur = lb2_ld32(&ek[e]);
		if (cur == 0) { cur = lb2_cas32(&ek[e], 0u, key); if (cur == 0) { cur = key; } }
		if (cur == key) { lb2_max32(&es[e], inv); return; }
	}
	lb2_or32(&W.sh->err, 1u << LB2_D_EDGES);
}

// one work item: a whole read, or a 64-pair chunk of the reference "read"
LB2_DEV void lb2_walk(lb2_win &W, uint32_t g0, uint32_t n, uint32_t o_begin, uint32_t o_end, uint32_t kbase,
                      bool isref, uint32_t cls, int K, int nw)
{
	lb2_ws &ws = W.ws; lb2_sh *sh = W.sh;
	lb2_kmer f, rc;
	for (int j = 0; j < LB2_MAXW; ++j) { f.w[j] = 0; rc.w[j] = 0; }
	uint32_t wordbuf = 0; uint32_t g = g0 + o_begin;
	// warm-up: first K bases
	for (int i = 0; i < K; ++i, ++g) {
		if ((g & 15) == 0 || i == 0) { wordbuf = W.bits[g >> 4]; }
		int c = (wordbuf >> ((g & 15) << 1)) & 3;
		lb2_roll_fwd(f, K, c); lb2_roll_rc(rc, K, c);
	}
	const bool tumor = !isref && cls < 2, normal = !isref && cls >= 2;
	const bool track_q = tumor && sh->has_lowq;
	int lowcnt = 0;    // low-quality bases in [o, o+K-1]; the pair window adds base o+K
	if (track_q) { for (int i = 0; i < K; ++i) { lowcnt += lb2_getbit(W.lowq, g0 + o_begin + i); } }
	bool fless = lb2_less(f, rc, nw);
	uint32_t ori_u = fless ? 0u : 1u;
	uint32_t su = lb2_find_or_insert(W, fless ? f : rc, fless ? rc : f, ((g0 + o_begin) << 1) | ori_u, K, nw, true);
	if (su == LB2_NIL) { return; }
	lb2_max32(&ws.occ[su], 0xFFFFFFFFu - (kbase + o_begin));
	const bool rec = !isref && sh->has_pairs;
	if (isref) { ws.refnode[o_begin] = su; }
	else if (o_begin == 0) {
		lb2_add32(&ws.cnt[su * 4 + cls], 1u);
		if (normal) { lb2_or32(&ws.sflags[su], 1u); }
		if (rec) { ws.inst[kbase] = su; }
	}
	for (uint32_t o = o_begin; o < o_end; ++o, ++g) {
		if ((g & 15) == 0) { wordbuf = W.bits[g >> 4]; }
		int c = (wordbuf >> ((g & 15) << 1)) & 3;
		lb2_roll_fwd(f, K, c); lb2_roll_rc(rc, K, c);
		fless = lb2_less(f, rc, nw);
		uint32_t ori_v = fless ? 0u : 1u;
		uint32_t sv = lb2_find_or_insert(W, fless ? f : rc, fless ? rc : f, ((g0 + o + 1) << 1) | ori_v, K, nw, true);
		if (sv == LB2_NIL) { return; }
		lb2_max32(&ws.occ[sv], 0xFFFFFFFFu - (kbase + o + 1));
		if (isref) { ws.refnode[o + 1] = sv; }
		else {
			lb2_add32(&ws.cnt[sv * 4 + cls], 1u);
			if (normal) { lb2_or32(&ws.sflags[sv], 1u); }
			if (rec) { ws.inst[kbase + o + 1] = sv; }
			if (tumor) {
				bool clean = true;
				if (track_q) {
					int wl = lowcnt + lb2_getbit(W.lowq, g);          // window [o, o+K]
					clean = (wl == 0);
					lowcnt = wl - lb2_getbit(W.lowq, g0 + o);         // slide to [o+1, o+K]
				}
				if (clean) { lb2_or32(&ws.sflags[su], 2u); lb2_or32(&ws.sflags[sv], 2u); }
			}
		}
		// edge directions (src/Graph.cc:320-326)
		uint32_t fdir, rdir;
		if (!ori_u && !ori_v) { fdir = LB2_FF; rdir = LB2_RR; }
		else if (!ori_u && ori_v) { fdir = LB2_FR; rdir = LB2_FR; }
		else if (ori_u && !ori_v) { fdir = LB2_RF; rdir = LB2_RF; }
		else { fdir = LB2_RR; rdir = LB2_FF; }
		uint32_t es = 2u * (kbase + o);
		lb2_add_edge(W, su, sv, fdir, es);
		lb2_add_edge(W, sv, su, rdir, es + 1u);
		su = sv; ori_u = ori_v;
	}
}

// bitonic sort of ws.sortk[0..n2) ascending, n2 a power of two
LB2_DEVNI void lb2_sort64(uint64_t *a, uint32_t n2)
{
	const unsigned tid = lb2_tid(), nt = lb2_nthr();
	for (uint32_t k = 2; k <= n2; k <<= 1) {
		for (uint32_t j = k >> 1; j > 0; j >>= 1) {
			for (uint32_t i = tid; i < n2; i += nt) {
				uint32_t l = i ^ j;
				if (l > i) {
					uint64_t x = a[i], y = a[l];
					bool up = ((i & k) == 0);
					if ((x > y) == up) { a[i] = y; a[l] = x; }
				}
			}
			lb2_sync();
		}
	}
}

LB2_DEV void lb2_revcomp(const lb2_kmer &f, int K, lb2_kmer &rc) {
	for (int j = 0; j < LB2_MAXW; ++j) { rc.w[j] = 0; }
	for (int i = 0; i < K; ++i) {
		uint64_t c = 3 - ((f.w[i >> 5] >> ((i & 31) << 1)) & 3);
		int t = K - 1 - i;
		rc.w[t >> 5] |= c << ((t & 31) << 1);
	}
}

// ---------------------------------------------------------------------------------------------
// build the graph for k-mer size K; on return the dense node arrays are filled (insertion order)
// ---------------------------------------------------------------------------------------------
LB2_DEVNI void lb2_build_graph(lb2_win &W, int K)
{
	lb2_sh *sh = W.sh; lb2_ws &ws = W.ws; const lb2_cfg *C = W.C;
	const unsigned tid = lb2_tid(), nt = lb2_nthr();
	const int nw = lb2_nw(K);
	const uint32_t R = sh->R, L = sh->L;
	if (tid == 0) {
		uint32_t cum = 0;
		for (uint32_t r = 0; r < R; ++r) { ws.rd_kbase[r] = cum; uint32_t n = ws.rd_len[r]; cum += (n > (uint32_t)K) ? (n - K + 1) : 0; }
		ws.rd_kbase[R] = cum;
		sh->n_used = 0; sh->err = 0; sh->n_spec = 0;
		sh->K = K; sh->nw = nw;
		if (cum + L >= 0x7FFFFFF0u) { sh->err |= 1u << LB2_D_READS; }
		if (sh->has_pairs && cum > W.C->max_inst) { sh->err |= 1u << LB2_D_READS; }
	}
	for (uint32_t i = tid; i < LB2_MAX_REF; i += nt) { ws.refnode[i] = LB2_NIL; }
	lb2_sync();
	if (sh->err) { return; }
	const uint32_t nref_pairs = (L > (uint32_t)K) ? (L - K) : 0;
	const uint32_t nchunks = (nref_pairs + 63) / 64;
	for (uint32_t it = tid; it < R + nchunks; it += nt) {
		if (it < R) {
			uint32_t n = ws.rd_len[it];
			if (n > (uint32_t)K) { lb2_walk(W, ws.rd_start[it], n, 0, n - K, ws.rd_kbase[it], false, ws.rd_info[it] & 3u, K, nw); }
		} else {
			uint32_t c = it - R, ob = c * 64, oe = ob + 64; if (oe > nref_pairs) { oe = nref_pairs; }
			lb2_walk(W, sh->ref_g, L, ob, oe, ws.rd_kbase[R], true, 0, K, nw);
		}
	}
	lb2_sync();
	lb2_mark(W, LB2_PH_WALK);
	if (sh->err) { return; }
	// ---- compaction: order used slots by first occurrence = insertion order of the reference map
	const uint32_t n = sh->n_used;
	uint32_t n2 = 1; while (n2 < n) { n2 <<= 1; }
	for (uint32_t j = tid; j < n2; j += nt) {
		if (j < n) { uint32_t s = ws.used[j]; ws.sortk[j] = ((uint64_t)(0xFFFFFFFFu - ws.occ[s]) << 32) | s; }
		else { ws.sortk[j] = ~0ull; }
	}
	lb2_sync();
	lb2_sort64(ws.sortk, n2);
	for (uint32_t j = tid; j < n; j += nt) { ws.slot2id[(uint32_t)ws.sortk[j]] = j; }
	lb2_sync();
	for (uint32_t j = tid; j < n; j += nt) {
		uint32_t s = (uint32_t)ws.sortk[j];
		uint32_t rep = (uint32_t)ws.slots[s];
		ws.d_rep[j] = rep;
		lb2_kmer km; lb2_extract(W.bits, rep >> 1, K, km);
		if (rep & 1) { lb2_kmer t; lb2_revcomp(km, K, t); km = t; }
		ws.d_hash[j] = lb2_stdhash_kmer(km, K);
		uint32_t tot = 0;
		for (int c = 0; c < 4; ++c) { uint32_t v = ws.cnt[s * 4 + c]; ws.d_cov[j * 4 + c] = (float)v; ws.d_cnt[j * 4 + c] = v; tot += v; }
		uint32_t sf = ws.sflags[s];
		ws.d_stn[j] = 1; ws.d_stT[j] = (sf == 2u) ? 1u : 0u;   // cov_status of the k-mer: 'T' iff tumour-qualified and never normal
		ws.d_mincov[j] = (int32_t)tot; ws.d_mincovqv[j] = (int32_t)tot;
		ws.d_flags[j] = 0; ws.d_comp[j] = 0; ws.d_color[j] = 0;
		ws.d_str[j] = LB2_NIL; ws.d_len[j] = (uint32_t)K; ws.d_cd[j] = LB2_NIL;
		// edges: sort by first-seen, map targets to dense ids
		uint32_t ek[LB2_ECAP], es[LB2_ECAP]; int ne = 0;
		for (int e = 0; e < LB2_ECAP; ++e) {
			uint32_t k = ws.ekey[(size_t)s * LB2_ECAP + e];
			if (k) { ek[ne] = k - 1; es[ne] = 0xFFFFFFFFu - ws.eseq[(size_t)s * LB2_ECAP + e]; ++ne; }
		}
		for (int a = 1; a < ne; ++a) {
			uint32_t kk = ek[a], ss = es[a]; int b = a - 1;
			while (b >= 0 && es[b] > ss) { ek[b + 1] = ek[b]; es[b + 1] = es[b]; --b; }
			ek[b + 1] = kk; es[b + 1] = ss;
		}
		for (int e = 0; e < ne; ++e) {
			lb2_edge ed; ed.to = ws.slot2id[ek[e] >> 2]; ed.dir = (uint8_t)(ek[e] & 3); ed.flag = 0; ed.pad = 0;
			ws.d_edge[(size_t)j * LB2_ECAP + e] = ed;
		}
		ws.d_ne[j] = (uint8_t)ne;
	}
	for (uint32_t p = tid; p < LB2_MAX_REF; p += nt) { uint32_t s = ws.refnode[p]; if (s != LB2_NIL) { ws.refnode[p] = ws.slot2id[s]; } }
	lb2_sync();
	lb2_sync();
	lb2_mark(W, LB2_PH_COMPACT);
	// ---- overlapping-mate suppression (Node_t::hasOverlappingMate / addMateName, src/Node.cc:638-671;
	//      call sites src/Graph.cc:232-233, 269, 299): a k-mer occurrence is not counted when std::binary_search
	//      finds the read's name in the node's list of names of the OTHER mate order -- a list that is in push
	//      order (unsorted) at that time.  Queries of one read only look at lists fed by reads of
	//      the other mate order, so per node it suffices to replay its occurrences in read order.
	if (sh->has_pairs) {
		const uint32_t total = ws.rd_kbase[R];
		uint32_t t2 = 1; while (t2 < total) { t2 <<= 1; }
		for (uint32_t x = tid; x < t2; x += nt) { ws.sortk[x] = (x < total) ? (((uint64_t)ws.slot2id[ws.inst[x]] << 32) | x) : ~0ull; }
		lb2_sync();
		lb2_sort64(ws.sortk, t2);
		uint32_t *nstart = ws.stack;
		for (uint32_t x = tid; x < total; x += nt) {
			uint32_t nd = (uint32_t)(ws.sortk[x] >> 32);
			if (x == 0 || (uint32_t)(ws.sortk[x - 1] >> 32) != nd) { nstart[nd] = x; }
		}
		lb2_sync();
		for (uint32_t j = tid; j < n; j += nt) {
			uint32_t tot = 0; for (int c = 0; c < 4; ++c) { tot += ws.d_cnt[j * 4 + c]; }
			if (!tot) { continue; }
			uint32_t b0 = nstart[j];
			uint32_t *L1 = ws.mates + 2 * (size_t)b0; uint32_t *L2top = ws.mates + 2 * (size_t)b0 + 2 * (size_t)tot - 1;   // list 2 grows downwards
			uint32_t n1 = 0, n2 = 0;
			for (uint32_t x = b0; x < b0 + tot; ++x) {
				uint32_t s_ = (uint32_t)ws.sortk[x];
				uint32_t lo = 0, hi = R;                       // read of occurrence s_: last r with kbase[r] <= s_
				while (hi - lo > 1) { uint32_t mid = (lo + hi) >> 1; if (ws.rd_kbase[mid] <= s_) { lo = mid; } else { hi = mid; } }
				uint32_t r = lo, p = s_ - ws.rd_kbase[r];
				uint32_t info = ws.rd_info[r], mate = (info >> 2) & 3u, cls = info & 3u, name = ws.rd_rank[r];
				if (mate == 1 || mate == 2) {
					// std::binary_search(first,last,val) = lower_bound + !(val < *it)   (libstdc++ stl_algo.h)
					uint32_t len = (mate == 1) ? n2 : n1, first = 0;
					while (len > 0) {
						uint32_t half = len >> 1, mid = first + half;
						uint32_t v = (mate == 1) ? *(L2top - mid) : L1[mid];
						if (v < name) { first = mid + 1; len = len - half - 1; } else { len = half; }
					}
					uint32_t cnt_other = (mate == 1) ? n2 : n1;
					bool ovl = false;
					if (first < cnt_other) { uint32_t v = (mate == 1) ? *(L2top - first) : L1[first]; ovl = !(name < v); }
					if (ovl) { ws.inst[s_] |= 0x80000000u; ws.d_cnt[j * 4 + cls] -= 1; }
					uint32_t last = ws.rd_len[r] - (uint32_t)K;
					uint32_t pushes = (p == 0 || p == last) ? 1u : 2u;
					for (uint32_t q = 0; q < pushes; ++q) { if (mate == 1) { L1[n1++] = name; } else { *(L2top - n2) = name; ++n2; } }
				}
			}
			uint32_t t = 0; for (int c = 0; c < 4; ++c) { uint32_t v = ws.d_cnt[j * 4 + c]; ws.d_cov[j * 4 + c] = (float)v; t += v; }
			ws.d_mincov[j] = (int32_t)t; ws.d_mincovqv[j] = (int32_t)t;
		}
		lb2_sync();
	}
	lb2_mark(W, LB2_PH_MATES);
	// ---- low-quality deficits: minqv_{fwd,rev}[i] = count - deficit[i]   (Node_t::updateCovDistr, src/Node.cc:470-497)
	if (tid == 0) { sh->n_nodes = n; sh->last_nodes = n; }
	if (sh->has_lowq) {
		uint32_t *d32 = (uint32_t *)ws.deficit;
		if ((size_t)n * K * 8 > (size_t)C->deficit_bytes) { if (tid == 0) { sh->err |= 1u << LB2_D_ARENA; } }
		else {
			for (uint32_t i = tid; i < n * (uint32_t)K * 2; i += nt) { d32[i] = 0; }
			lb2_sync();
			for (uint32_t r = tid; r < R; r += nt) {
				uint32_t len = ws.rd_len[r]; if (len <= (uint32_t)K) { continue; }
				uint32_t g0 = ws.rd_start[r]; uint32_t cls = ws.rd_info[r] & 3u;
				for (uint32_t q = 0; q < len; ++q) {
					if (!lb2_getbit(W.lowq, g0 + q)) { continue; }
					uint32_t p0 = (q + 1 > (uint32_t)K) ? (q + 1 - K) : 0, p1 = (q < len - K) ? q : (len - K);
					for (uint32_t p = p0; p <= p1; ++p) {
						if (sh->has_pairs && (ws.inst[ws.rd_kbase[r] + p] & 0x80000000u)) { continue; }
						lb2_kmer f, rc; lb2_extract(W.bits, g0 + p, K, f); lb2_revcomp(f, K, rc);
						bool fl = lb2_less(f, rc, nw);
						uint32_t s = lb2_find_or_insert(W, fl ? f : rc, fl ? rc : f, 0, K, nw, false);
						if (s == LB2_NIL) { continue; }
						uint32_t id = ws.slot2id[s];
						uint32_t i = fl ? (q - p) : ((uint32_t)K - 1 - (q - p));   // qv string is reversed for ori R (src/Graph.cc:148-158)
						lb2_add32(&d32[((size_t)id * K + i) * 2 + (cls >> 1)], (cls & 1) ? 0x10000u : 1u);
					}
				}
			}
			lb2_sync();
			for (uint32_t j = tid; j < n; j += nt) {
				uint32_t mx = 0;
				for (int i = 0; i < K; ++i) {
					uint32_t a = d32[((size_t)j * K + i) * 2], b = d32[((size_t)j * K + i) * 2 + 1];
					uint32_t t = (a & 0xFFFF) + (a >> 16) + (b & 0xFFFF) + (b >> 16);
					if (t > mx) { mx = t; }
				}
				ws.d_mincovqv[j] -= (int32_t)mx;
			}
		}
	}
	lb2_sync();
	lb2_mark(W, LB2_PH_LOWQ);
	// ---- hand the slots back clean (the table is all-zero between builds)
	for (uint32_t j = tid; j < n; j += nt) {
		uint32_t s = ws.used[j];
		ws.slots[s] = 0; ws.occ[s] = 0; ws.sflags[s] = 0;
		for (int c = 0; c < 4; ++c) { ws.cnt[s * 4 + c] = 0; }
		for (int e = 0; e < LB2_ECAP; ++e) { ws.ekey[(size_t)s * LB2_ECAP + e] = 0; ws.eseq[(size_t)s * LB2_ECAP + e] = 0; }
	}
	lb2_sync();
}

#endif
