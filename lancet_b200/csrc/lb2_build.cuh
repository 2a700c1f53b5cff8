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
	lb2_ws ws, ws0; lb2_sh *sh;
	uint32_t *bits;      // smem: 2-bit packed trimmed reads, then the window reference
	uint32_t *lowq;      // smem: 1 bit per staged base: quality < MIN_QUAL_CALL
	char     *ref_raw;   // smem: window reference, ASCII
	uint8_t  *treg;      // smem region T: Mer->Node table during the build, graph-stage arrays afterwards
	uint32_t *t_key; uint16_t *t_id;   // smem: table keys; slot -> dense node id (bit 15: needs first-seen edge order)
	lb2_variant *ovar; char *ostr; uint32_t ovar_cap, ostr_cap; bool escal;   // this window's output slab (regular or large)
};

// attribute the cycles since the previous mark to phase ph (lane 0 only).  Instrumentation: compiled in only with
// -DLB2_PROFILE (tools/phase_profile.py builds that variant as lancet_b200/_lb2_prof.so); the product library has none of it
#ifdef LB2_PROFILE
LB2_DEV void lb2_mark(lb2_win &W, int ph) {
	if (lb2_tid() == 0) { unsigned long long t = lb2_clock(); W.sh->prof[ph] += t - W.sh->t_last; W.sh->t_last = t; }
}
#else
LB2_DEV void lb2_mark(lb2_win &, int) {}
#endif
#ifdef LB2_PROFILE_SEQ      // (debug builds only: the lane-0 sections of a component split over counters that are idle on unpaired, clean input)
#define LB2_SEQMARK(ph) lb2_mark(W, ph)
#else
#define LB2_SEQMARK(ph) do { } while (0)
#endif

// CTA-wide exclusive prefix sum: set(i, sum_{j<i} get(j)); returns the total.  Each lane owns a contiguous chunk.
template <class Get, class Set>
LB2_DEV uint32_t lb2_excl_scan(lb2_win &W, uint32_t n, Get get, Set set)
{
	const unsigned tid = lb2_tid(), nt = lb2_nthr(); uint32_t *sc = W.sh->scan;
	const uint32_t chunk = (n + nt - 1) / nt, lo = (tid * chunk < n) ? tid * chunk : n, hi = (lo + chunk < n) ? lo + chunk : n;
	uint32_t sum = 0;
	for (uint32_t i = lo; i < hi; ++i) { sum += get(i); }
	uint32_t total = 0;
	uint32_t acc = lb2_block_excl(sc, sum, &total);
	for (uint32_t i = lo; i < hi; ++i) { uint32_t v = get(i); set(i, acc); acc += v; }
	lb2_sync();
	return total;
}

LB2_DEV void lb2_fail(lb2_win &W, uint32_t status, uint32_t detail) {
	if (lb2_cas32(&W.sh->status, LB2_WIN_OK, status) == LB2_WIN_OK) { W.sh->detail = detail; }
}

// ---------------------------------------------------------------------------------------------
// Staging.  The window's reads come out of the packed pool (lb2_pack.cuh) as whole 16-base words: a read lands in shared
// memory UNtrimmed, its first kept base at staged index g = 16 * word + trm5.  Pool reads that follow each other in the
// window's list (the usual case: a window's tumour reads are one run of the coordinate-sorted pool, its normal reads
// another) keep their pool layout, so a run moves with one bulk-async copy (cp.async.bulk, completion on an mbarrier).
// Bulk copies want 16-byte alignment on both sides; the quality mask has 2 bytes per word, hence a run is placed so that
// its shared-memory word index equals its pool word index modulo 8, with up to 7 pad words before and after it (the
// copy moves whole 8-word blocks, the pad words receive whatever surrounds the run in the pool and are never looked at).
// Both shared-memory areas are lent to the graph stage after every build, so they are staged again (copies only) before
// the build of a later k.
// ---------------------------------------------------------------------------------------------
LB2_DEVNI void lb2_stage_pack(lb2_win &W, bool do_bits, bool do_lowq)
{
	lb2_sh *sh = W.sh; const lb2_dev_batch *B = W.B; lb2_ws &ws = W.ws;
	const unsigned tid = lb2_tid(), nt = lb2_nthr(), lane = lb2_lane();
	const uint32_t R = sh->R, L = sh->L;
	lb2_sync();                  // every ordinary access to the two areas is over ...
	lb2_async_fence();           // ... and ordered before this lane's bulk copies
	const uint32_t *const pbits = B->pk_bits; const uint16_t *const plowq = B->pk_lowq; uint16_t *const lowq16 = (uint16_t *)W.lowq;
	// pieces: maximal stretches of reads that are consecutive in the pool AND in shared memory, cut at groups of LB2_WARP reads;
	// the first lane of a piece issues its copies
	for (uint32_t base = (tid / LB2_WARP) * LB2_WARP; base < R; base += (nt / LB2_WARP) * LB2_WARP) {
		const uint32_t i = base + lane; const bool valid = i < R;
		uint32_t gw = 0, nw = 0, sw = 0;
		if (valid) { const uint64_t v = ws.rd_src[i]; gw = (uint32_t)v; nw = (uint32_t)(v >> 32); sw = (ws.rd_start[i] - ws.rd_t5[i]) >> 4; }
		const uint32_t pg = lb2_shfl_up1(gw + nw), ps = lb2_shfl_up1(sw + nw);
		const bool head = valid && (lane == 0 || pg != gw || ps != sw);
		const uint32_t heads = lb2_ballot(head), above = (lane + 1 < LB2_WARP) ? (heads >> (lane + 1)) : 0u;
		uint32_t tail = above ? lane + (uint32_t)lb2_ctz32(above) : LB2_WARP - 1u;      // last lane of the piece that starts here
		if (base + tail >= R) { tail = R - 1u - base; }
		const uint32_t gend = lb2_shfl(gw + nw, tail);
		if (head) {
			const uint32_t g0 = gw & ~7u, g1 = (gend + 7u) & ~7u, s0 = sw - (gw & 7u);
			if (g1 > g0) {
				if (do_bits) { lb2_bulk_g2s(W.bits + s0, pbits + g0, (g1 - g0) * 4u, &sh->mbar); }
				if (do_lowq) { lb2_bulk_g2s(lowq16 + s0, plowq + g0, (g1 - g0) * 2u, &sh->mbar); }
			}
		}
	}
	lb2_mbar_arrive(&sh->mbar);
	if (do_bits) {      // the window reference, packed from its ASCII copy
		const uint32_t g = sh->ref_g;
		for (uint32_t b0 = tid * 16; b0 < ((L + 15u) & ~15u) + 64; b0 += nt * 16) {
			uint32_t bw = 0;
			for (uint32_t i = 0; i < 16 && b0 + i < L; ++i) { bw |= (uint32_t)(lb2_code(W.ref_raw[b0 + i]) & 3) << (2 * i); }
			W.bits[(g + b0) >> 4] = bw;
		}
	}
	if (do_lowq) { for (uint32_t i = (sh->ref_g >> 5) + tid; i < (sh->total_bp >> 5) + 4; i += nt) { W.lowq[i] = 0; } }      // (the reference "read" has quality 'K' everywhere)
	lb2_mbar_wait(&sh->mbar, sh->mbar_phase);
	lb2_sync();
	if (tid == 0) { sh->mbar_phase ^= 1u; if (do_bits) { sh->bits_live = 1; } if (do_lowq) { sh->lowq_live = 1; } }
	lb2_sync();
}
LB2_DEV void lb2_stage_bits(lb2_win &W) { lb2_stage_pack(W, true, false); }
LB2_DEV void lb2_stage_lowq(lb2_win &W) { lb2_stage_pack(W, false, true); }

// ---------------------------------------------------------------------------------------------
// stage the window: per-read records from the packed pool, shared-memory layout, bulk copies
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
		sh->has_lowq = 0; sh->has_pairs = 0; sh->mapped = 0; sh->flag_a = 0; sh->flag_b = 0; sh->totalreadbp = 0; sh->ref_hasN = 0;
		sh->n_var = 0; sh->str_used = 0; sh->n_k_tried = 0; sh->final_k = 0; sh->last_nodes = 0;
		sh->err = 0;      // (a window whose every k is skipped never reaches lb2_build_graph, which resets it per k)
		if (sh->L > LB2_MAX_REF) { sh->status = LB2_WIN_OVERFLOW; sh->detail = LB2_D_REFLEN; }
		else if (sh->R + 1 > W.C->max_reads) { sh->status = LB2_WIN_OVERFLOW; sh->detail = LB2_D_READS; }
	}
	lb2_sync();
	if (sh->status != LB2_WIN_OK) { return; }
	const uint32_t L = sh->L, R = sh->R;
	const char *ref = B->ref_seq + B->ref_off[w];
	for (uint32_t i = tid; i < LB2_MAX_REF / 32 + 2; i += nt) { sh->refn[i] = 0; }
	lb2_sync();
	for (uint32_t i = tid; i < L; i += nt) {
		char c = ref[i]; W.ref_raw[i] = c;
		if (lb2_code(c) < 0) { if (c == 'N') { lb2_or32(&sh->refn[i >> 5], 1u << (i & 31)); sh->ref_hasN = 1; } else { lb2_or32(&sh->flag_a, 1u); } }
	}
	const uint32_t *widx = B->wr_idx + B->wr_off[w];
	const lb2_pkread *const PK = B->pk; const uint32_t *const NR_ = B->name_rank;
	{	// per-read records; words each read adds to the layout (its own, plus the pad words where a run of pool-consecutive reads starts / ends)
		uint32_t bp = 0, fl = 0;
		for (uint32_t i = tid; i < R; i += nt) {
			const uint32_t idx = widx[i]; const lb2_pkread rec = PK[idx];
			const bool first = (i == 0) || (widx[i - 1] + 1u != idx), last = (i + 1 == R) || (widx[i + 1] != idx + 1u);
			const uint32_t lead = first ? (rec.woff & 7u) : 0u, trail = last ? ((0u - (rec.woff + rec.nw)) & 7u) : 0u;
			ws.rd_len[i] = rec.n; ws.rd_t5[i] = rec.t5; ws.rd_info[i] = rec.info & 15u; ws.rd_rank[i] = NR_[idx];
			ws.rd_src[i] = (uint64_t)rec.woff | ((uint64_t)rec.nw << 32);
			ws.rd_kbase[i] = ((uint32_t)rec.nw + lead + trail) | (lead << 28);      // (scratch until the build of the first k)
			bp += rec.n;
			fl |= (rec.info & LB2_PK_UNMAPPED) ? 0u : 1u; fl |= (rec.info & LB2_PK_TOOLONG) ? 2u : 0u; fl |= rec.lowq ? 4u : 0u;
		}
		if (bp) { lb2_add32(&sh->totalreadbp, bp); }
		if (fl) { lb2_or32(&sh->flag_b, fl); }
	}
	lb2_sync();
	const uint32_t words = lb2_excl_scan(W, R, [&](uint32_t r) -> uint32_t { return ws.rd_kbase[r] & 0x0FFFFFFFu; },
	                                     [&](uint32_t r, uint32_t v) { ws.rd_start[r] = 16u * (v + (ws.rd_kbase[r] >> 28)) + ws.rd_t5[r]; });
	if (tid == 0) {
		uint32_t cum = 16u * words;      // (a multiple of 8 words: every run is padded to whole 8-word blocks)
		ws.rd_start[R] = cum; sh->ref_g = cum; cum += (L + 31u) & ~31u;
		sh->total_bp = cum;
		const uint32_t fl = sh->flag_b; sh->mapped = fl & 1u; sh->has_lowq = (fl & 4u) ? 1u : 0u; sh->flag_b = 0;
		if (cum + 64 > W.C->max_bp) { lb2_fail(W, LB2_WIN_OVERFLOW, LB2_D_SMEM); }
		if (fl & 2u) { lb2_fail(W, LB2_WIN_UNSUPPORTED, LB2_D_READS); }
		if (sh->flag_a) { lb2_fail(W, LB2_WIN_UNSUPPORTED, LB2_D_NREF); }      // a reference character that is neither ACGT nor N (callers map IUPAC codes to N)
		if (!(fl & 1u)) { lb2_fail(W, LB2_WIN_NO_READS, 0); }
		sh->seq_off = 0; sh->seq_len = L; sh->trim5 = 0; sh->trim3 = 0;
	}
	lb2_sync();
	if (sh->status != LB2_WIN_OK) { return; }
	{	// does any query name occur with both mate orders?  (otherwise hasOverlappingMate can never fire)
		// open-addressing set of name ranks in the (idle) table region, two mate-order bits per entry
		// ... and which read is whose mate: rd_mate[r] = the window's other read of the same name (LB2_NIL: none).  A name that
		// occurs more than twice, or twice with the same mate order, makes the window "complex" (has_pairs = 2: the replay
		// then takes no shortcut)
		uint32_t *hs = (uint32_t *)W.treg; const uint32_t TS = W.C->table_slots, hmask = TS - 1; uint16_t *hi = (uint16_t *)(hs + TS);
		for (uint32_t i = tid; i < TS; i += nt) { hs[i] = 0xFFFFFFFFu; }
		for (uint32_t r = tid; r < R; r += nt) { ws.rd_mate[r] = LB2_NIL; }
		lb2_sync();
		if (2 * R > TS || R > 0xFFF0u) { if (tid == 0) { lb2_max32(&sh->has_pairs, 2u); } }      // (would not fit: run the full replay unconditionally, it is exact either way)
		else {
			for (uint32_t r = tid; r < R; r += nt) {      // first read of every name claims a slot
				const uint32_t mate = (ws.rd_info[r] >> 2) & 3u; if (mate != 1 && mate != 2) { continue; }
				const uint32_t rank = ws.rd_rank[r];
				if (rank >= 0x3FFFFFFFu) { lb2_max32(&sh->has_pairs, 2u); continue; }
				uint32_t h = (rank * 2654435761u) & hmask;
				for (uint32_t probes = 0; probes <= hmask; ++probes) {
					uint32_t cur = lb2_ld32(&hs[h]);
					if (cur == 0xFFFFFFFFu) { cur = lb2_cas32(&hs[h], 0xFFFFFFFFu, (rank << 2) | mate); if (cur == 0xFFFFFFFFu) { hi[h] = (uint16_t)r; break; } }
					if ((cur >> 2) == rank) { break; }
					h = (h + 1) & hmask;
				}
			}
			lb2_sync();
			for (uint32_t r = tid; r < R; r += nt) {      // every other read of the name meets the first one
				const uint32_t mate = (ws.rd_info[r] >> 2) & 3u; if (mate != 1 && mate != 2) { continue; }
				const uint32_t rank = ws.rd_rank[r]; if (rank >= 0x3FFFFFFFu) { continue; }
				uint32_t h = (rank * 2654435761u) & hmask;
				for (uint32_t probes = 0; probes <= hmask; ++probes) {
					const uint32_t cur = lb2_ld32(&hs[h]);
					if (cur == 0xFFFFFFFFu) { break; }
					if ((cur >> 2) == rank) {
						const uint32_t r0 = hi[h];
						if (r0 != r) {
							if ((cur & 3u) == mate) { lb2_max32(&sh->has_pairs, 2u); }
							else {
								if (lb2x_cas32(&ws.rd_mate[r0], LB2_NIL, r) != LB2_NIL) { lb2_max32(&sh->has_pairs, 2u); }      // a third read of that name
								ws.rd_mate[r] = r0; lb2_max32(&sh->has_pairs, 1u);
							}
						}
						break;
					}
					h = (h + 1) & hmask;
				}
			}
		}
		lb2_sync();
	}
	lb2_stage_pack(W, true, sh->has_lowq != 0);
}

// ---------------------------------------------------------------------------------------------
// Mer -> Node table in SHARED memory (region T), TS = cfg.table_slots slots, per slot:
//   t_key  u32  0x80000000 | fingerprint10 << 21 | (g << 1 | ori): g = staged base index of the FIRST occurrence seen
//               so far (kept minimal with a shared-memory min: reads are staged in the reference's processing
//               order, so the smallest g is the occurrence that inserts the node into the reference's map);
//               matches are verified against the packed bases
//   t_id   u16  during the walk: bits 0-7 edge types seen (start orientation*4 + appended base), bit 8 normal,
//               bit 9 tumour-qualified (tested with a plain load, set with a shared-memory or only when missing);
//               after the compaction: dense node id, bit 15 = "branching" (needs first-seen edge order)
// and, in GLOBAL memory:
//   g_cnt  2xu32 per slot  tumour / normal occurrence counts, fwd in the low half, rev in the high half
//               (fire-and-forget reductions: nothing on the lane's critical path reads them back)
//   g_em   u32 per DENSE node  the slot's mask, saved before t_id is reused
//   inst   u16 per occurrence  slot (14 bits) | orientation << 15 (| 0x4000 suppressed by the mate replay), indexed
//               ibase + offset * istride: transposed (offset-major) so that the lanes of a warp -- one read each, same
//               offset -- store to consecutive words; read-major when the transposed array would not fit
// ---------------------------------------------------------------------------------------------
#define LB2_EM_NORMAL 0x100u
#define LB2_EM_TUMOR  0x200u
#define LB2_ID_BRANCH 0x8000u

// one-word k-mers (K <= 32; KT = uint32_t when K <= 16): table and packed bases through precomputed shared addresses
template <class KT> LB2_DEV uint32_t lb2_foi_small(lb2_win &W, lb2_sp tk, lb2_sp bits, uint32_t mask, KT canon, KT nonc, uint32_t rep, KT kmask, bool insert)
{
	const uint32_t h = lb2_hash1((uint32_t)canon, sizeof(KT) > 4 ? (uint32_t)((uint64_t)canon >> 32) : 0u);
	uint32_t i = h & mask;
	const uint32_t fp = 0x80000000u | ((h >> 22) << 21);
	// One exit (the outer loop's end): the lanes of a warp leave the probe sequence at different times and meet again
	// right behind it.  The inner loop only skips slots of other fingerprints -- three instructions per slot, which is
	// what the lanes that are already done wait for; bases are fetched and compared once a fingerprint matches.
	uint32_t res = LB2_NIL, probes = 0; bool full = true;
#pragma unroll 1
	while (true) {
		lb2_sp at = lb2_sp_at(tk, i); uint32_t cur = lb2s_ldv(at);
#pragma unroll 1
		while (cur != 0 && (cur & 0xFFE00000u) != fp && probes <= mask) { i = (i + 1) & mask; ++probes; at = lb2_sp_at(tk, i); cur = lb2s_ldv(at); }
		if (probes > mask) { break; }
		if (cur == 0) {
			if (!insert) { full = false; break; }
			cur = lb2s_cas(at, 0u, fp | rep);
			if (cur == 0) {      // this lane created the slot
				uint32_t u = lb2_add32(&W.sh->n_used, 1u);
				if (u < W.C->max_nodes && u < (mask + 1) - ((mask + 1) >> 2)) { W.ws.used[u] = i; } else { lb2_or32(&W.sh->err, 1u << LB2_D_HASH_FULL); }
				res = i; full = false; break;
			}
		}
		if ((cur & 0xFFE00000u) == fp) {
			const uint32_t r = cur & 0x1FFFFFu;
			const KT o = lb2_extract_small<KT>(bits, r >> 1) & kmask;
			if (o == ((r & 1u) ? nonc : canon)) {
				if (insert && rep < r) { lb2s_min(at, fp | rep); }      // same k-mer => same fingerprint: the word orders by rep
				res = i; full = false; break;
			}
		}
		i = (i + 1) & mask; ++probes;
	}
	if (full) { lb2_or32(&W.sh->err, 1u << LB2_D_HASH_FULL); }
	return res;
}

template <int NWT = LB2_MAXW> LB2_DEV uint32_t lb2_find_or_insert(lb2_win &W, const lb2_kmer &canon, const lb2_kmer &nonc, uint32_t rep, int K, int nw, bool insert)
{
	if (NWT == 1) {      // (the same hash and fingerprint as the one-word walk)
		const uint64_t kmask = (~0ull) >> (64 - 2 * K);
		return lb2_foi_small<uint64_t>(W, lb2_sp_of(W.t_key), lb2_sp_of(W.bits), W.C->table_slots - 1, canon.w[0], nonc.w[0], rep, kmask, insert);
	}
	uint32_t *tk = W.t_key; const uint32_t mask = W.C->table_slots - 1;
	uint64_t h = lb2_table_hash<NWT>(canon, nw);
	uint32_t i = (uint32_t)h & mask;
	const uint32_t fp = 0x80000000u | (((uint32_t)(h >> 40) & 0x3FFu) << 21);
	for (uint32_t probes = 0; probes <= mask; ++probes) {
		uint32_t cur = lb2_ld32(&tk[i]);
		if (cur == 0) {
			if (!insert) { return LB2_NIL; }
			uint32_t prev = lb2_cas32(&tk[i], 0u, fp | rep);
			if (prev == 0) {
				uint32_t u = lb2_add32(&W.sh->n_used, 1u);
				if (u < W.C->max_nodes && u < (mask + 1) - ((mask + 1) >> 2)) { W.ws.used[u] = i; } else { lb2_or32(&W.sh->err, 1u << LB2_D_HASH_FULL); }
				return i;
			}
			cur = prev;
		}
		if ((cur & 0xFFE00000u) == fp) {
			uint32_t r = cur & 0x1FFFFFu;
			lb2_kmer o; lb2_extract<NWT>(W.bits, r >> 1, K, o);
			if (lb2_equal<NWT>(o, (r & 1) ? nonc : canon, nw)) {
				if (insert && rep < r) { lb2_min32(&tk[i], fp | rep); }      // same k-mer => same fingerprint: the word orders by rep
				return i;
			}
		}
		i = (i + 1) & mask;
	}
	lb2_or32(&W.sh->err, 1u << LB2_D_HASH_FULL);
	return LB2_NIL;
}

// set mask bits of a slot (16-bit lanes of the words that later hold t_id)
LB2_DEV void lb2_em_or(lb2_win &W, uint32_t slot, uint32_t bits) {
	uint32_t *wp = (uint32_t *)W.t_id + (slot >> 1); const uint32_t sh = (slot & 1u) << 4;
	if ((((lb2_ld32(wp)) >> sh) & bits) != bits) { lb2_or32(wp, bits << sh); }
}

// one work item: a whole read, or a 64-pair chunk of the reference "read"
template <int NWT> LB2_DEV void lb2_walk(lb2_win &W, uint32_t g0, uint32_t n, uint32_t o_begin, uint32_t o_end, uint32_t ibase, uint32_t istride,
                      bool isref, uint32_t cls, int K, int nw)
{
	lb2_ws &ws = W.ws; lb2_sh *sh = W.sh;
	lb2_kmer f, rc;
	for (int j = 0; j < LB2_MAXW; ++j) { f.w[j] = 0; rc.w[j] = 0; }
	uint32_t wordbuf = 0; uint32_t g = g0 + o_begin;
	for (int i = 0; i < K; ++i, ++g) {          // warm-up: first K bases
		if ((g & 15) == 0 || i == 0) { wordbuf = lb2_lds(&W.bits[g >> 4]); }
		int c = (wordbuf >> ((g & 15) << 1)) & 3;
		lb2_roll_fwd<NWT>(f, K, c); lb2_roll_rc<NWT>(rc, K, c);
	}
	const bool tumor = !isref && cls < 2, normal = !isref && cls >= 2;
	const bool track_q = tumor && sh->has_lowq;
	const uint32_t cadd = (cls & 1) ? 0x10000u : 1u, csel = cls >> 1;
	int lowcnt = 0;    // low-quality bases in [o, o+K-1]; the pair window adds base o+K
	if (track_q) { for (int i = 0; i < K; ++i) { lowcnt += lb2_getbit(W.lowq, g0 + o_begin + i); } }
	bool fless = lb2_less<NWT>(f, rc, nw);
	uint32_t ori_u = fless ? 0u : 1u;
	uint32_t su = lb2_find_or_insert<NWT>(W, lb2_pick<NWT>(fless, f, rc), lb2_pick<NWT>(fless, rc, f), ((g0 + o_begin) << 1) | ori_u, K, nw, true);
	if (su == LB2_NIL) { return; }
	ws.inst[ibase + o_begin * istride] = (uint16_t)(su | (ori_u << 15));
	if (isref) { ws.refnode[o_begin] = su; }
	else if (o_begin == 0) {
		lb2g_red_add(&ws.g_cnt[su * 2 + csel], cadd);
		if (normal) { lb2_em_or(W, su, LB2_EM_NORMAL); }
	}
	for (uint32_t o = o_begin; o < o_end; ++o, ++g) {
		if ((g & 15) == 0) { wordbuf = lb2_lds(&W.bits[g >> 4]); }
		int c = (wordbuf >> ((g & 15) << 1)) & 3;
		int a = (int)(f.w[0] & 3);                                 // base that leaves the window (first base of u)
		lb2_roll_fwd<NWT>(f, K, c); lb2_roll_rc<NWT>(rc, K, c);
		fless = lb2_less<NWT>(f, rc, nw);
		uint32_t ori_v = fless ? 0u : 1u;
		uint32_t sv = lb2_find_or_insert<NWT>(W, lb2_pick<NWT>(fless, f, rc), lb2_pick<NWT>(fless, rc, f), ((g0 + o + 1) << 1) | ori_v, K, nw, true);
		if (sv == LB2_NIL) { return; }
		ws.inst[ibase + (o + 1) * istride] = (uint16_t)(sv | (ori_v << 15));
		uint32_t emu = 1u << (ori_u * 4 + (uint32_t)c);            // u leaves in orientation ori_u appending c
		uint32_t emv = 1u << ((1u - ori_v) * 4 + (uint32_t)(3 - a)); // v leaves in the flipped orientation appending comp(a)
		if (isref) { ws.refnode[o + 1] = sv; }
		else {
			lb2g_red_add(&ws.g_cnt[sv * 2 + csel], cadd);
			if (normal) { emv |= LB2_EM_NORMAL; }
			if (tumor) {
				bool clean = true;
				if (track_q) {
					int wl = lowcnt + lb2_getbit(W.lowq, g);          // window [o, o+K]
					clean = (wl == 0);
					lowcnt = wl - lb2_getbit(W.lowq, g0 + o);         // slide to [o+1, o+K]
				}
				if (clean) { emu |= LB2_EM_TUMOR; emv |= LB2_EM_TUMOR; }
			}
		}
		lb2_em_or(W, su, emu); lb2_em_or(W, sv, emv);
		su = sv; ori_u = ori_v;
	}
}

LB2_DEV uint64_t lb2_revcomp1(uint64_t f, int K);
// the same work item for one-word k-mers.  With base i at bits [2i, 2i+1], the reverse complement's integer is
// (4^K - 1) minus the base-0-first ("big-endian") integer of the k-mer, so the reference's string comparison
// mer < rc is the integer comparison rc > f of the two little-endian words the walk rolls anyway; one mask update per k-mer (the bits a k-mer receives as the v of one
// pair and as the u of the next are merged).  Warp-converged: EVERY lane of the warp calls it (active = false: no item),
// the pair loop runs to the warp's longest piece with the lanes re-joined at the top of every round.
template <class KT, bool isref> LB2_DEV void lb2_walk_small(lb2_win &W, bool active, uint32_t g0, uint32_t o_begin, uint32_t o_end, uint32_t ibase, uint32_t istride,
                      uint32_t cls, int K)
{
	lb2_ws &ws = W.ws; lb2_sh *sh = W.sh;
	const lb2_sp tk = lb2_sp_of(W.t_key), tidw = lb2_sp_of(W.t_id), bits = lb2_sp_of(W.bits);
	const uint32_t mask = W.C->table_slots - 1;
	// (array bases in registers: the descriptor lives in shared memory and would be re-read behind every store)
	uint16_t *inst_p = ws.inst + ibase + (size_t)o_begin * istride; uint32_t *const GC = ws.g_cnt + (cls >> 1); uint32_t *const RN = ws.refnode;
	const int topsh = 2 * (K - 1); const KT kmask = (KT)(~(KT)0) >> (sizeof(KT) * 8 - 2 * K);
	KT f = 0, rc = 0;
	uint32_t wordbuf = 0; uint32_t g = g0 + o_begin;
	if (active) {      // the piece's first k-mer straight from the packed bases
		f = lb2_extract_small<KT>(bits, g) & kmask;
		rc = (KT)lb2_revcomp1((uint64_t)f, K);
		g += (uint32_t)K;
		wordbuf = lb2s_ld(lb2_sp_at(bits, g >> 4));
	}
	const bool tumor = !isref && cls < 2, normal = !isref && cls >= 2;
	const bool track_q = active && tumor && sh->has_lowq;
	const uint32_t cadd = (cls & 1) ? 0x10000u : 1u;
	int lowcnt = 0;    // low-quality bases in [o, o+K-1]; the pair window adds base o+K
	if (track_q) { for (int i = 0; i < K; ++i) { lowcnt += lb2_getbit(W.lowq, g0 + o_begin + i); } }
	bool fless = rc > f;
	uint32_t ori_u = fless ? 0u : 1u, su = LB2_NIL, pend = 0;      // pend: mask bits owed to su
	if (active) {
		su = lb2_foi_small<KT>(W, tk, bits, mask, fless ? f : rc, fless ? rc : f, ((g0 + o_begin) << 1) | ori_u, kmask, true);
		if (su != LB2_NIL) {
			*inst_p = (uint16_t)(su | (ori_u << 15));
			if (isref) { RN[o_begin] = su; }
			else if (o_begin == 0) {
				lb2g_red_add(&GC[su * 2], cadd);
				if (normal) { pend = LB2_EM_NORMAL; }
			}
		}
	}
	const uint32_t nsteps = (active && su != LB2_NIL) ? (o_end - o_begin) : 0u, maxsteps = lb2_warp_max(nsteps);
	uint32_t o = o_begin;
#pragma unroll 1
	for (uint32_t st = 0; st < maxsteps; ++st) {
		lb2_warp_sync();
		if (st >= nsteps || su == LB2_NIL) { continue; }
		if ((g & 15) == 0) { wordbuf = lb2s_ld(lb2_sp_at(bits, g >> 4)); }
		const uint32_t c = (wordbuf >> ((g & 15) << 1)) & 3u;
		const uint32_t a = (uint32_t)f & 3u;                           // base that leaves the window (first base of u)
		f = (f >> 2) | ((KT)c << topsh); rc = ((rc << 2) | (KT)(3u - c)) & kmask;
		fless = rc > f;
		const uint32_t ori_v = fless ? 0u : 1u;
		const uint32_t sv = lb2_foi_small<KT>(W, tk, bits, mask, fless ? f : rc, fless ? rc : f, ((g0 + o + 1) << 1) | ori_v, kmask, true);
		uint32_t emu = 1u << (ori_u * 4 + c);                         // u leaves in orientation ori_u appending c
		uint32_t emv = 1u << ((1u - ori_v) * 4 + (3u - a));           // v leaves in the flipped orientation appending comp(a)
		if (sv != LB2_NIL) {
			inst_p += istride; *inst_p = (uint16_t)(sv | (ori_v << 15));
			if (isref) { RN[o + 1] = sv; }
			else {
				lb2g_red_add(&GC[sv * 2], cadd);
				if (normal) { emv |= LB2_EM_NORMAL; }
				if (tumor) {
					bool clean = true;
					if (track_q) {
						int wl = lowcnt + lb2_getbit(W.lowq, g);          // window [o, o+K]
						clean = (wl == 0);
						lowcnt = wl - lb2_getbit(W.lowq, g0 + o);         // slide to [o+1, o+K]
					}
					if (clean) { emu |= LB2_EM_TUMOR; emv |= LB2_EM_TUMOR; }
				}
			}
			pend |= emu;
			{ const lb2_sp wp = lb2_sp_at(tidw, su >> 1); const uint32_t s4 = (su & 1u) << 4; if (((lb2s_ldv(wp) >> s4) & pend) != pend) { lb2s_or(wp, pend << s4); } }
			pend = emv;
		} else { pend = 0; }      // (table full: the window is redone by the escalation pass)
		su = sv; ori_u = ori_v; ++o; ++g;
	}
	if (pend && su != LB2_NIL) { const lb2_sp wp = lb2_sp_at(tidw, su >> 1); const uint32_t s4 = (su & 1u) << 4; if (((lb2s_ldv(wp) >> s4) & pend) != pend) { lb2s_or(wp, pend << s4); } }
}

// reverse complement: complement, reverse the 2-bit groups of the 256-bit integer, shift down to 2K bits
LB2_DEV uint64_t lb2_rev2(uint64_t x) {
#ifndef LB2_HOSTSIM
	{ const uint64_t y = __brevll(x); return ((y >> 1) & 0x5555555555555555ull) | ((y & 0x5555555555555555ull) << 1); }      // bit reversal, then the two bits of every base back in order
#endif
	x = ((x >> 2) & 0x3333333333333333ull) | ((x & 0x3333333333333333ull) << 2);
	x = ((x >> 4) & 0x0F0F0F0F0F0F0F0Full) | ((x & 0x0F0F0F0F0F0F0F0Full) << 4);
	x = ((x >> 8) & 0x00FF00FF00FF00FFull) | ((x & 0x00FF00FF00FF00FFull) << 8);
	x = ((x >> 16) & 0x0000FFFF0000FFFFull) | ((x & 0x0000FFFF0000FFFFull) << 16);
	return (x >> 32) | (x << 32);
}
LB2_DEV uint64_t lb2_revcomp1(uint64_t f, int K) { return lb2_rev2(~f) >> (64 - 2 * K); }      // one-word k-mer (K <= 32)
template <int NWT = LB2_MAXW> LB2_DEV void lb2_revcomp(const lb2_kmer &f, int K, lb2_kmer &rc) {
	const int nw = lb2_nw(K);
	uint64_t t[NWT];
	// reversed words of the complement, as if the k-mer filled nw*32 bases; word order reversed within the nw words
	// (all indexing is static after unrolling: the words stay in registers)
#pragma unroll
	for (int j = 0; j < NWT; ++j) {
		uint64_t v = 0;
#pragma unroll
		for (int s = 0; s < NWT; ++s) { if (s == nw - 1 - j) { v = lb2_rev2(~f.w[s]); } }
		t[j] = (j < nw) ? v : 0;
	}
	// the (nw*32 - K) pad bases now sit at the low end: shift right by 2*pad bits
	const int sh = (nw * 32 - K) * 2;
#pragma unroll
	for (int j = 0; j < NWT; ++j) {
		uint64_t hi = (j + 1 < NWT) ? t[(j + 1 < NWT) ? j + 1 : j] : 0;
		rc.w[j] = sh ? ((t[j] >> sh) | (hi << (64 - sh))) : t[j];
	}
	lb2_mask_top<NWT>(rc, K);
#pragma unroll
	for (int j = 0; j < LB2_MAXW; ++j) { if (j >= nw || j >= NWT) { rc.w[j] = 0; } }
}
// canonical k-mer of a dense node (from its representative occurrence)
template <int NWT = LB2_MAXW> LB2_DEV void lb2_rep_kmer(lb2_win &W, uint32_t rep, int K, lb2_kmer &km) {
	lb2_extract<NWT>(W.bits, rep >> 1, K, km);
#pragma unroll
	for (int j = NWT; j < LB2_MAXW; ++j) { km.w[j] = 0; }
	if (rep & 1) { lb2_kmer t; lb2_revcomp<NWT>(km, K, t); km = lb2_pick<NWT>(true, t, t); }
}

// edges of one surviving node: every edge type (start orientation, appended base) names one neighbour.  A group of
// LB2_GS lanes per node, one edge type each (all lanes of a warp call this together; act = false: no node)
template <int NWT> LB2_DEV void lb2_node_edges(lb2_win &W, bool act, uint32_t j, int K, int nw)
{
	lb2_ws &ws = W.ws; lb2_sh *sh = W.sh; const uint32_t gl = lb2_glane();
	uint32_t livemask = 0; lb2_bedge edl[8 / LB2_GS];
	if (act && NWT == 1) {      // one-word k-mers as plain 64-bit integers (mer < rc  <=>  rc > mer as integers, see lb2_walk_small)
		const uint32_t em = ws.g_em[j] & 0xFFu, rep = ws.b_rep[j];
		const uint64_t kmask = (~0ull) >> (64 - 2 * K); const int topsh = 2 * (K - 1);
		const lb2_sp tk = lb2_sp_of(W.t_key), bits = lb2_sp_of(W.bits); const uint32_t mask = W.C->table_slots - 1;
		uint64_t c0 = lb2_extract_small<uint64_t>(bits, rep >> 1) & kmask;
		if (rep & 1u) { c0 = lb2_revcomp1(c0, K); }
		const uint64_t c1 = lb2_revcomp1(c0, K);
		for (uint32_t t = gl; t < 8; t += LB2_GS) {
			if (!(em & (1u << t))) { continue; }
			const uint32_t o = t >> 2, b = t & 3u;
			const uint64_t V = ((o ? c1 : c0) >> 2) | ((uint64_t)b << topsh), Vr = lb2_revcomp1(V, K);
			const bool fl = Vr > V;
			const uint32_t ts = lb2_foi_small<uint64_t>(W, tk, bits, mask, fl ? V : Vr, fl ? Vr : V, 0, kmask, false);
			if (ts == LB2_NIL) { lb2_or32(&sh->err, 1u << LB2_D_EDGES); continue; }
			const uint32_t to = W.t_id[ts] & 0x7FFFu;
			if (ws.b_flags[to] & LB2_NF_DEAD) { continue; }
			lb2_bedge ed; ed.to = to; ed.dir = o * 2 + (fl ? 0u : 1u); ed.flag = 0; ed.type = t;
			edl[t / LB2_GS] = ed; livemask |= 1u << t;
		}
	} else if (act) {
		const uint32_t em = ws.g_em[j] & 0xFFu;
		lb2_kmer C0; lb2_rep_kmer<NWT>(W, ws.b_rep[j], K, C0);
		lb2_kmer C1; lb2_revcomp<NWT>(C0, K, C1);
		for (uint32_t t = gl; t < 8; t += LB2_GS) {
			if (!(em & (1u << t))) { continue; }
			const int o = (int)(t >> 2), b = (int)(t & 3);
			lb2_kmer V = lb2_pick<NWT>(o != 0, C1, C0); lb2_roll_fwd<NWT>(V, K, b);
			lb2_kmer Vr; lb2_revcomp<NWT>(V, K, Vr);
			bool fl = lb2_less<NWT>(V, Vr, nw);
			uint32_t ts = lb2_find_or_insert<NWT>(W, lb2_pick<NWT>(fl, V, Vr), lb2_pick<NWT>(fl, Vr, V), 0, K, nw, false);
			if (ts == LB2_NIL) { lb2_or32(&sh->err, 1u << LB2_D_EDGES); continue; }
			uint32_t to = W.t_id[ts] & 0x7FFFu;
			if (ws.b_flags[to] & LB2_NF_DEAD) { continue; }
			lb2_bedge ed; ed.to = to; ed.dir = (uint32_t)(o * 2 + (fl ? 0 : 1)); ed.flag = 0; ed.type = t;
			edl[t / LB2_GS] = ed; livemask |= 1u << t;
		}
	}
	livemask = lb2_gor(livemask);      // the node's live edge types; edges are stored in type order
	if (!act) { return; }
	for (uint32_t t = gl; t < 8; t += LB2_GS) {
		if (livemask & (1u << t)) { ws.b_edge[(size_t)j * LB2_BECAP + (uint32_t)lb2_popc32(livemask & ((1u << t) - 1u))] = edl[t / LB2_GS]; }
	}
	const int nF = lb2_popc32(livemask & 0x0Fu), nR = lb2_popc32(livemask & 0xF0u);
	if (gl == 0) { ws.b_ne[j] = (uint8_t)(nF + nR); }
	if (nF > 1 || nR > 1) {   // first-seen order matters only among edges leaving in the same orientation
		if (gl == 0) { ws.b_flags[j] |= 0x20; sh->flag_a = 1; }      // slot's branch bit is set after the barrier (t_id is being read by other lanes)
		for (uint32_t t = gl; t < 8; t += LB2_GS) { ws.bseq[(size_t)j * 8 + t] = 0xFFFFFFFFu; }
	}
}

// ---------------------------------------------------------------------------------------------
// build the graph for k-mer size K.  On return: dense nodes in insertion order (all nodes), the
// first low-coverage sweep already applied as a per-node predicate (LB2_NF_DEAD), edges of the
// survivors in the reference's first-seen order.
// ---------------------------------------------------------------------------------------------
LB2_DEVNI void lb2_build_graph(lb2_win &W, int K)
{
	lb2_sh *sh = W.sh; lb2_ws &ws = W.ws; const lb2_cfg *C = W.C;
	const unsigned tid = lb2_tid(), nt = lb2_nthr();
	const int nw = lb2_nw(K);
	const uint32_t R = sh->R, L = sh->L, TS = C->table_slots;
	if (!sh->bits_live) { lb2_stage_bits(W); }
	if (!sh->lowq_live && sh->has_lowq) { lb2_stage_lowq(W); }
	if (tid == 0) { sh->maxnk = 0; W.t_key = (uint32_t *)W.treg; W.t_id = (uint16_t *)(W.t_key + TS); }
	lb2_sync();
	{ uint32_t mx = 0; for (uint32_t r = tid; r < R; r += nt) { uint32_t n = ws.rd_len[r]; if (n > (uint32_t)K && n - K + 1 > mx) { mx = n - K + 1; } } if (mx) { lb2_max32(&sh->maxnk, mx); } }
	uint32_t kcum = lb2_excl_scan(W, R, [&](uint32_t r) -> uint32_t { uint32_t n = ws.rd_len[r]; return (n > (uint32_t)K) ? (n - K + 1) : 0; },
	                              [&](uint32_t r, uint32_t v) { ws.rd_kbase[r] = v; });
	if (tid == 0) {
		uint32_t cum = kcum;
		ws.rd_kbase[R] = cum;
		sh->n_used = 0; sh->err = 0; sh->n_spec = 0; sh->flag_a = 0; sh->n_nk = 0; sh->n_refitems = 0;
		sh->K = K; sh->nw = nw;
		// up to four pieces per read, none shorter than 16 pairs (every piece re-derives its first k-mer)
		{ const uint32_t mp = sh->maxnk > 1 ? sh->maxnk - 1 : 1u; uint32_t np_ = mp / 16u; if (np_ < 1) { np_ = 1; } if (np_ > 4) { np_ = 4; }
		  sh->walk_np = np_; sh->walk_pl = (mp + np_ - 1) / np_; if (sh->walk_pl < 16) { sh->walk_pl = 16; } sh->walk_next = 0; }
		// occurrence array: offset-major (coalesced stores) when it fits, read-major otherwise
		const uint32_t Rs = (R + 31u) & ~31u;
		bool transposed = (uint64_t)sh->maxnk * Rs + L + 2 <= (uint64_t)C->max_inst;
#ifdef LB2_HOSTSIM
		if (getenv("LB2_SIM_COMPACT")) { transposed = false; }
#endif
		if (transposed) { sh->inst_stride = Rs; sh->inst_ref = sh->maxnk * Rs; }
		else { sh->inst_stride = 0; sh->inst_ref = cum; if (cum + L + 2 > C->max_inst) { sh->err |= 1u << LB2_D_READS; } }
	}
	if (tid == 0 && sh->ref_hasN) {
		// the reference walk stops at N: its work items are pieces (<= walk_pl pairs, ob | oe << 16) of the N-free stretches
		// that hold at least one k-mer; k-mers with an N get no table slot (they become dead map entries further down)
		uint32_t ni = 0, a = 0; const uint32_t PL_ = sh->walk_pl;
		for (uint32_t i = 0; i <= L; ++i) {
			if (i < L && !((sh->refn[i >> 5] >> (i & 31)) & 1u)) { continue; }
			if (i - a >= (uint32_t)K) { uint32_t ob = a; const uint32_t lastk = i - (uint32_t)K; while (true) { uint32_t oe = ob + PL_; if (oe > lastk) { oe = lastk; } ws.jobs[ni++] = ob | (oe << 16); if (oe == lastk) { break; } ob = oe; } }
			a = i + 1;
		}
		sh->n_refitems = ni;
	}
	for (uint32_t i = tid; i < LB2_MAX_REF; i += nt) { ws.refnode[i] = LB2_NIL; }
	for (uint32_t i = tid; i < TS; i += nt) { W.t_key[i] = 0; }
	for (uint32_t i = tid; i < TS / 2; i += nt) { ((uint32_t *)W.t_id)[i] = 0; }
	for (uint32_t i = tid; i < TS * 2; i += nt) { ws.g_cnt[i] = 0; }
	lb2_sync();
	if (sh->err) { return; }
	const uint32_t istr = sh->inst_stride ? sh->inst_stride : 1u;
	const uint32_t nref_pairs = (L > (uint32_t)K) ? (L - K) : 0;
	// work items: (piece of a read) and (piece of the reference), handed out to whole warps from a shared counter so that
	// the warps finish together.  Item = piece * R + read: the lanes of a warp walk the same piece of consecutive reads
	// (equal lengths, consecutive words of the offset-major occurrence array).
	const uint32_t PL = sh->walk_pl, NP = sh->walk_np;
	const bool refN = sh->ref_hasN != 0;
	const uint32_t nchunks = refN ? sh->n_refitems : (nref_pairs + PL - 1) / PL, nitems = NP * R + nchunks;
	while (true) {
		const uint32_t it = lb2_batch_next(&sh->walk_next);
		if (it >= ((nitems + 31u) & ~31u)) { break; }
		if (lb2_ballot(lb2_ld32(&sh->err) != 0)) { break; }      // (table full: the window is redone with a larger one; filling this one to the brim only makes the probes longer)
		// this lane's item (none: active = false; the one-word walks are called by all lanes of the warp together)
		bool active = false, isref = false; uint32_t g0 = 0, n = 0, ob = 0, oe = 0, ib = 0, st = 1, cls = 0;
		if (it < NP * R) {
			const uint32_t piece = it / R, r = it - piece * R;
			n = ws.rd_len[r];
			if (n > (uint32_t)K) {
				const uint32_t np_ = n - K; ob = piece * PL; oe = ob + PL; if (oe > np_) { oe = np_; }
				if (ob < np_) { active = true; g0 = ws.rd_start[r]; ib = sh->inst_stride ? r : ws.rd_kbase[r]; st = istr; cls = ws.rd_info[r] & 3u; }
			}
		} else if (it < nitems) {
			if (refN) { const uint32_t v = ws.jobs[it - NP * R]; ob = v & 0xFFFFu; oe = v >> 16; }
			else { ob = (it - NP * R) * PL; oe = ob + PL; if (oe > nref_pairs) { oe = nref_pairs; } }
			active = true; isref = true; g0 = sh->ref_g; n = L; ib = sh->inst_ref; st = 1u;
		}
		const bool hasreads = (it & ~31u) < NP * R, hasref = (it | 31u) >= NP * R;      // (the same for all lanes of the warp)
		if (K <= 16) {
			if (hasreads) { lb2_walk_small<uint32_t, false>(W, active && !isref, g0, ob, oe, ib, st, cls, K); }
			if (hasref) { lb2_walk_small<uint32_t, true>(W, active && isref, g0, ob, oe, ib, st, cls, K); }
		} else if (nw == 1) {
			if (hasreads) { lb2_walk_small<uint64_t, false>(W, active && !isref, g0, ob, oe, ib, st, cls, K); }
			if (hasref) { lb2_walk_small<uint64_t, true>(W, active && isref, g0, ob, oe, ib, st, cls, K); }
		} else if (active) {
			if (nw == 2) { lb2_walk<2>(W, g0, n, ob, oe, ib, st, isref, cls, K, nw); }
			else { lb2_walk<LB2_MAXW>(W, g0, n, ob, oe, ib, st, isref, cls, K, nw); }
		}
	}
	lb2_sync();
	lb2_mark(W, LB2_PH_WALK);
	if (sh->err) { return; }
	// ---- compaction: dense ids in order of first occurrence = insertion order of the reference map.  The first
	//      occurrences are distinct staged base indices, so the rank of a node is the number of set bits below its
	//      index in a bitmap over the staged bases (one popcount scan instead of a sort).
	// ---- reference k-mers that contain an N.  The reference is loaded as an untrimmed read (src/Graph.cc:534-540; isNseq
	//      never fires, src/util.cc:259-273), so each distinct one is an entry of the reference's map: never covered by a
	//      read, removed by the first low-coverage sweep, but present while the map grows -- it shifts the iteration order
	//      of everything else.  Here: canonical form on the ASCII window (rrc('N') = 'N', 'G' < 'N' < 'T'), std::hash of
	//      it, first occurrence of every distinct one; they join the dense nodes below as dead nodes without a table slot.
	uint64_t *const nkh = (uint64_t *)ws.mates; uint8_t *const nkf = (uint8_t *)(nkh + LB2_MAX_REF);      // (the mate lists are idle until the replay)
	if (refN) {
		const char *const RR = W.ref_raw;
		auto canon_at = [&](uint32_t o, uint32_t ori, uint32_t i) -> char { return ori ? lb2_comp(RR[o + (uint32_t)K - 1u - i]) : RR[o + i]; };
		for (uint32_t o = tid; o + (uint32_t)K <= L; o += nt) {
			bool hasn = false; for (uint32_t i = o; i < o + (uint32_t)K && !hasn; ++i) { hasn = ((sh->refn[i >> 5] >> (i & 31)) & 1u) != 0; }
			uint8_t f = 0;
			if (hasn) {
				uint32_t ori = 1;      // CanonicalMer_t::set: F iff mer < rc2(mer) as strings (a palindrome is R)
				for (int i = 0; i < K; ++i) { const char a = RR[o + i], b = lb2_comp(RR[o + K - 1 - i]); if (a != b) { ori = ((unsigned char)a < (unsigned char)b) ? 0u : 1u; break; } }
				lb2_stdhash hs; lb2_sh_init(hs, (uint32_t)K);
				for (int i = 0; i < K; ++i) { lb2_sh_byte(hs, (unsigned char)canon_at(o, ori, (uint32_t)i)); }
				nkh[o] = lb2_sh_final(hs); f = (uint8_t)(1u | (ori << 1));
			}
			nkf[o] = f;
		}
		lb2_sync();
		uint32_t mine = 0;
		for (uint32_t o = tid; o + (uint32_t)K <= L; o += nt) {
			const uint8_t f = nkf[o]; if (!(f & 1u)) { continue; }
			bool first = true;
			for (uint32_t p_ = 0; p_ < o && first; ++p_) {
				const uint8_t fp_ = nkf[p_]; if (!(fp_ & 1u) || nkh[p_] != nkh[o]) { continue; }
				bool same = true; for (int i = 0; i < K && same; ++i) { same = canon_at(o, (f >> 1) & 1u, (uint32_t)i) == canon_at(p_, (fp_ >> 1) & 1u, (uint32_t)i); }
				if (same) { first = false; }
			}
			if (first) { nkf[o] = (uint8_t)(f | 4u); ++mine; }      // (bit 2 is only read after the barrier below)
		}
		if (mine) { lb2_add32(&sh->n_nk, mine); }
		lb2_sync();
		if (tid == 0 && (sh->n_used + sh->n_nk > C->max_nodes || sh->n_used + sh->n_nk >= 0x7FF0u)) { sh->err |= 1u << LB2_D_HASH_FULL; }
		lb2_sync();
		if (sh->err) { return; }
	}
	const uint32_t n_slots_used = sh->n_used, n = n_slots_used + sh->n_nk;
	{
		uint32_t *bm = (uint32_t *)ws.sortk; const uint32_t nwords = (sh->total_bp >> 5) + 2; uint32_t *pref = bm + nwords;
		for (uint32_t i = tid; i < nwords; i += nt) { bm[i] = 0; }
		lb2_sync();
		for (uint32_t j = tid; j < n_slots_used; j += nt) { const uint32_t g = (W.t_key[ws.used[j]] & 0x1FFFFFu) >> 1; lb2g_red_or(&bm[g >> 5], 1u << (g & 31)); }
		if (refN) { for (uint32_t o = tid; o + (uint32_t)K <= L; o += nt) { if (nkf[o] & 4u) { const uint32_t g = sh->ref_g + o; lb2g_red_or(&bm[g >> 5], 1u << (g & 31)); } } }
		lb2_sync();
		lb2_excl_scan(W, nwords, [&](uint32_t i) -> uint32_t { return (uint32_t)lb2_popc32(bm[i]); }, [&](uint32_t i, uint32_t v) { pref[i] = v; });
		for (uint32_t j = tid; j < n_slots_used; j += nt) {
			const uint32_t s = ws.used[j]; const uint32_t g = (W.t_key[s] & 0x1FFFFFu) >> 1;
			ws.b_row[pref[g >> 5] + (uint32_t)lb2_popc32(bm[g >> 5] & ((1u << (g & 31)) - 1u))] = s;      // b_row is free until lb2_order_and_pack
		}
		if (refN) {      // an N k-mer has no slot: its entry is 0x80000000 | reference offset
			for (uint32_t o = tid; o + (uint32_t)K <= L; o += nt) {
				if (nkf[o] & 4u) { const uint32_t g = sh->ref_g + o; ws.b_row[pref[g >> 5] + (uint32_t)lb2_popc32(bm[g >> 5] & ((1u << (g & 31)) - 1u))] = 0x80000000u | o; }
			}
		}
		lb2_sync();
		for (uint32_t j = tid; j < n; j += nt) { const uint32_t s = ws.b_row[j]; ws.used[j] = s; ws.g_em[j] = (s & 0x80000000u) ? 0u : ((((const uint32_t *)W.t_id)[s >> 1] >> ((s & 1u) << 4)) & 0xFFFFu); }   // used := dense id -> slot
		lb2_sync();
		for (uint32_t j = tid; j < n; j += nt) { const uint32_t s = ws.used[j]; if (!(s & 0x80000000u)) { W.t_id[s] = (uint16_t)j; } }      // slot -> dense id (the mask bits were saved above)
		lb2_sync();
	}
	for (uint32_t j = tid; j < n; j += nt) {
		uint32_t s = ws.used[j];
		if (s & 0x80000000u) {      // N k-mer: no coverage, no edges (dies in the first low-coverage sweep below)
			ws.b_rep[j] = 0; ws.b_hash[j] = nkh[s & 0xFFFFu];
			for (int c = 0; c < 4; ++c) { ws.b_cnt[j * 4 + c] = 0; }
			ws.b_stT[j] = 0; ws.b_mincovqv[j] = 0; ws.b_flags[j] = 0; ws.b_ne[j] = 0;
			continue;
		}
		uint32_t rep = W.t_key[s] & 0x1FFFFFu;
		ws.b_rep[j] = rep;
		lb2_kmer km;
		if (nw == 1) { lb2_rep_kmer<1>(W, rep, K, km); ws.b_hash[j] = lb2_stdhash_kmer<1>(km, K); }
		else if (nw == 2) { lb2_rep_kmer<2>(W, rep, K, km); ws.b_hash[j] = lb2_stdhash_kmer<2>(km, K); }
		else { lb2_rep_kmer<LB2_MAXW>(W, rep, K, km); ws.b_hash[j] = lb2_stdhash_kmer<LB2_MAXW>(km, K); }
		uint32_t ct = ws.g_cnt[s * 2], cn = ws.g_cnt[s * 2 + 1];
		uint32_t v[4] = { ct & 0xFFFFu, ct >> 16, cn & 0xFFFFu, cn >> 16 }, tot = 0;
		for (int c = 0; c < 4; ++c) { ws.b_cnt[j * 4 + c] = v[c]; tot += v[c]; }
		uint32_t em = ws.g_em[j];
		ws.b_stT[j] = ((em & (LB2_EM_NORMAL | LB2_EM_TUMOR)) == LB2_EM_TUMOR) ? 1u : 0u;   // cov_status 'T': tumour-qualified, never normal
		ws.b_mincovqv[j] = (int32_t)tot;
		ws.b_flags[j] = 0; ws.b_ne[j] = 0;
	}
	for (uint32_t p = tid; p < LB2_MAX_REF; p += nt) { uint32_t s = ws.refnode[p]; if (s != LB2_NIL) { ws.refnode[p] = W.t_id[s]; } }
	if (tid == 0) { sh->n_nodes = n; sh->last_nodes = n; }
	lb2_sync();
	lb2_mark(W, LB2_PH_COMPACT);
	// ---- overlapping-mate suppression (Node_t::hasOverlappingMate / addMateName, src/Node.cc:638-671;
	//      call sites src/Graph.cc:232-233, 269, 299): a k-mer occurrence is not counted when std::binary_search
	//      finds the read's name in the node's list of names of the OTHER mate order -- a list that is in push
	//      order (unsorted) at that time.  Queries of one read only look at lists fed by reads of
	//      the other mate order, so per node it suffices to replay its occurrences in read order.
	// Which nodes need the replay at all?  A read's name can only be found in the other mate order's list if its mate
	// (rd_mate, the window's other read of that name) has had an occurrence in the same node before -- binary_search on
	// the unsorted list may miss a name that is there, it never finds one that is not.  So: the (node, read) pairs of every
	// read whose mate comes later in the window go into an exact open-addressing set (global scratch), the occurrences of
	// the later mates look their (node, mate) up, hits mark the node in a bitmap.  Mates that do not overlap mark nothing,
	// and nothing below runs; otherwise only the marked nodes' occurrences are sorted and replayed, and inside the replay
	// only the reads whose mate is in the set search the list (for everybody else the answer is known to be "not found").
	uint32_t *const nbm = (uint32_t *)(W.treg + (size_t)TS * 6);      // (region T behind the table: idle until the graph stage)
	const uint32_t nbm_words = (n + 31u) >> 5;
	const bool nbm_fits = lb2_treg_bytes(TS, C->graph_bytes, C->max_bp) >= (size_t)TS * 6 + (size_t)nbm_words * 4 + 16;
	uint32_t *const pset = (uint32_t *)ws.deficit;      // (the deficit counters come after the replay)
	uint32_t smask = 0, hshift = 32; bool use_set = false;
	if (tid == 0) { sh->flag_b = (sh->has_pairs == 2) ? 1u : 0u; }
	if (sh->has_pairs && nbm_fits) { for (uint32_t i = tid; i < nbm_words; i += nt) { nbm[i] = 0; } }
	lb2_sync();
	if (sh->has_pairs == 1) {
		const uint32_t total = ws.rd_kbase[R];
		uint32_t cap = 1024; while (cap < total + 16u) { cap <<= 1; }
		if ((size_t)cap * 4 > (size_t)C->deficit_bytes || R > 0x3FFFu || !nbm_fits) { if (tid == 0) { sh->flag_b = 1; } }      // no room: replay every node (exact either way)
		else {
			use_set = true; smask = cap - 1u; while ((1u << (32 - hshift)) < cap) { --hshift; }      // slot = top bits of key * odd constant
			for (uint32_t i = tid; i < cap; i += nt) { pset[i] = 0; }
			lb2_sync();
			// one lane per read, its occurrences in offset order (offset-major storage: the lanes of a warp, on consecutive reads,
			// load consecutive words), four loads in flight per lane; pass 0 inserts the reads whose mate comes later, pass 1
			// looks the mates of the others up
			const uint16_t *const INST = ws.inst; const uint16_t *const TID = W.t_id;
			for (int pass = 0; pass < 2; ++pass) {
				for (uint32_t r = tid; r < R; r += nt) {
					const uint32_t q = ws.rd_mate[r]; if (q == LB2_NIL) { continue; }
					if ((pass == 0) != (q > r)) { continue; }
					const uint32_t kb = ws.rd_kbase[r], nk = ws.rd_kbase[r + 1] - kb, ib = sh->inst_stride ? r : kb, who = ((pass == 0) ? r : q) & 0x3FFFu;
					for (uint32_t o0 = 0; o0 < nk; o0 += 8) {      // eight occurrences at a time: their loads, then their first probes, are all in flight together
						uint32_t iw[8], key[8], h[8], got[8];
#pragma unroll
						for (uint32_t u = 0; u < 8; ++u) { iw[u] = (o0 + u < nk) ? INST[ib + (o0 + u) * istr] : 0xFFFFFFFFu; }
#pragma unroll
						for (uint32_t u = 0; u < 8; ++u) {
							if (iw[u] == 0xFFFFFFFFu) { key[u] = 0; h[u] = 0; got[u] = 0; continue; }
							const uint32_t j = TID[iw[u] & 0x3FFFu] & 0x7FFFu; key[u] = ((j + 1u) << 14) | who; h[u] = ((key[u] * 2654435761u) >> hshift) & smask;
							got[u] = (pass == 0) ? lb2x_cas32(&pset[h[u]], 0u, key[u]) : pset[h[u]];
						}
#pragma unroll
						for (uint32_t u = 0; u < 8; ++u) {
							if (!key[u]) { continue; }
							if (pass == 0) { while (!(got[u] == 0u || got[u] == key[u])) { h[u] = (h[u] + 1u) & smask; got[u] = lb2x_cas32(&pset[h[u]], 0u, key[u]); } }
							else {
								while (got[u] != 0u && got[u] != key[u]) { h[u] = (h[u] + 1u) & smask; got[u] = pset[h[u]]; }
								if (got[u] == key[u]) { const uint32_t j = (key[u] >> 14) - 1u; lb2_or32(&nbm[j >> 5], 1u << (j & 31u)); sh->flag_b = 1; }
							}
						}
					}
				}
				lb2_sync();
			}
		}
		lb2_sync();
	}
	if (sh->has_pairs && sh->flag_b) {
		const bool all_nodes = !use_set;      // (complex names, or no room for the set / the bitmap)
		auto marked = [&](uint32_t j) -> bool { return all_nodes || ((lb2_lds(&nbm[j >> 5]) >> (j & 31u)) & 1u); };
		// The occurrences of every marked node in read order: a counting sort by node.  Every warp owns a contiguous stretch
		// of reads (about the same number of occurrences each); counts per (warp, node) are turned into first positions
		// (node's start + the counts of the warps before); the warp then walks its occurrences by occurrence number, 32 at
		// a time, handing out positions in order (lanes holding the same node rank themselves by lane number:
		// __match_any).  occ[pos] = read << 12 | offset.
		const uint32_t total = ws.rd_kbase[R];
		uint32_t *nstart = ws.b_row;                  // free until lb2_order_and_pack
		uint32_t *occ = (uint32_t *)ws.sortk, *wcnt = ws.bseq;      // (bseq: 8 words per node, idle until the edges are built)
		const uint32_t NW = nt / LB2_WARP, wid = tid / LB2_WARP, lane = lb2_lane();
		lb2_excl_scan(W, n, [&](uint32_t j) -> uint32_t { return marked(j) ? ws.b_cnt[j * 4] + ws.b_cnt[j * 4 + 1] + ws.b_cnt[j * 4 + 2] + ws.b_cnt[j * 4 + 3] : 0u; },
		              [&](uint32_t j, uint32_t v) { nstart[j] = v; });
		if (NW > 8) { if (tid == 0) { sh->err |= 1u << LB2_D_READS; } }
		for (uint32_t i = tid; i < NW * n && NW <= 8; i += nt) { wcnt[i] = 0; }
		// this warp's reads: [ra, rb) with rd_kbase[ra] the first read start >= total * wid / NW
		uint32_t ra, rb;
		{
			auto first_read = [&](uint32_t x) -> uint32_t { uint32_t lo = 0, hi = R; while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (ws.rd_kbase[mid] < x) { lo = mid + 1; } else { hi = mid; } } return lo; };
			ra = first_read((uint32_t)((uint64_t)total * wid / NW)); rb = (wid + 1 == NW) ? R : first_read((uint32_t)((uint64_t)total * (wid + 1) / NW));
		}
		lb2_sync();
		if (sh->err) { return; }
		{	// pass 1: counts per (owning warp, node), one lane per read, four loads in flight
			uint32_t *bounds = sh->scan;      // first read of every warp's stretch (scan[] is idle between the block scans)
			if (lane == 0) { bounds[wid] = ra; if (wid + 1 == NW) { bounds[NW] = R; } }
			lb2_sync();
			const uint16_t *const INST = ws.inst; const uint16_t *const TID = W.t_id;
			for (uint32_t r = tid; r < R; r += nt) {
				uint32_t w = 0; while (w + 1 < NW && bounds[w + 1] <= r) { ++w; }
				const uint32_t kb = ws.rd_kbase[r], nk = ws.rd_kbase[r + 1] - kb, ib = sh->inst_stride ? r : kb;
				for (uint32_t o0 = 0; o0 < nk; o0 += 4) {
					uint32_t iw[4];
#pragma unroll
					for (uint32_t u = 0; u < 4; ++u) { iw[u] = (o0 + u < nk) ? INST[ib + (o0 + u) * istr] : 0xFFFFFFFFu; }
#pragma unroll
					for (uint32_t u = 0; u < 4; ++u) { if (iw[u] != 0xFFFFFFFFu) { const uint32_t j = TID[iw[u] & 0x3FFFu] & 0x7FFFu; if (marked(j)) { lb2g_red_add(&wcnt[w * n + j], 1u); } } }
				}
			}
		}
		lb2_sync();
		for (uint32_t j = tid; j < n; j += nt) { uint32_t run = nstart[j]; for (uint32_t w = 0; w < NW; ++w) { const uint32_t c = wcnt[w * n + j]; wcnt[w * n + j] = run; run += c; } }
		lb2_sync();
		{	// pass 3: the warp's occurrences by occurrence number, 32 at a time, the next tile's word on its way while this one is placed
			const uint16_t *const INST = ws.inst; const uint16_t *const TID = W.t_id;
			const uint32_t xa = (ra < R) ? ws.rd_kbase[ra] : total, xb = (rb < R) ? ws.rd_kbase[rb] : total;
			uint32_t rr = ra;      // this lane's read cursor (occurrence numbers grow by LB2_WARP per tile)
			auto fetch = [&](uint32_t x, uint32_t &r_, uint32_t &o_) -> uint32_t {
				while (rr + 1 < R && x >= ws.rd_kbase[rr + 1]) { ++rr; }
				r_ = rr; o_ = x - ws.rd_kbase[rr];
				return INST[(sh->inst_stride ? rr : ws.rd_kbase[rr]) + o_ * istr];
			};
			uint32_t x = xa + lane, r_n = 0, o_n = 0, iw_n = 0; bool act_n = x < xb;
			if (act_n) { iw_n = fetch(x, r_n, o_n); }
			for (uint32_t t0 = xa; t0 < xb; t0 += LB2_WARP) {
				bool act = act_n; const uint32_t iw = iw_n, r = r_n, o = o_n;
				x += LB2_WARP; act_n = x < xb; if (act_n) { iw_n = fetch(x, r_n, o_n); }
				uint32_t j = act ? (uint32_t)(TID[iw & 0x3FFFu] & 0x7FFFu) : 0xFFFFFFFFu;
				if (act && !marked(j)) { act = false; j = 0xFFFFFFFFu; }
				if (!lb2_ballot(act)) { continue; }      // (the same for every lane of the warp)
				const uint32_t same = lb2_match_any(j), leader = (uint32_t)lb2_ctz32(same), below = same & ((1u << lane) - 1u);
				uint32_t pos = 0;
				if (act && lane == leader) { pos = wcnt[wid * n + j]; wcnt[wid * n + j] = pos + (uint32_t)lb2_popc32(same); }
				pos = lb2_shfl(pos, leader);
				if (act) { occ[pos + (uint32_t)lb2_popc32(below)] = (r << 12) | o; }      // (read, offset of the occurrence: offsets < 4096)
				lb2_warp_sync();
			}
		}
		lb2_sync();
		for (uint32_t j = tid; j < n; j += nt) {
			if (!marked(j)) { continue; }      // no read of this node has its mate in it: nothing is suppressed here
			uint32_t tot = 0; for (int c = 0; c < 4; ++c) { tot += ws.b_cnt[j * 4 + c]; }
			if (!tot) { continue; }
			uint32_t b0 = nstart[j];
			uint32_t *L1 = ws.mates + 2 * (size_t)b0; uint32_t *L2top = ws.mates + 2 * (size_t)b0 + 2 * (size_t)tot - 1;   // list 2 grows downwards
			uint32_t n1 = 0, n2_ = 0;
			for (uint32_t x = b0; x < b0 + tot; ++x) {
				const uint32_t s_ = occ[x], r = s_ >> 12, p = s_ & 0xFFFu;
				uint32_t info = ws.rd_info[r], mate = (info >> 2) & 3u, cls = info & 3u, name = ws.rd_rank[r];
				if (mate == 1 || mate == 2) {
					bool search = true;
					if (use_set) {      // only a read whose mate has been in this node can be found
						search = false; const uint32_t q = ws.rd_mate[r];
						if (q != LB2_NIL && q < r) {
							const uint32_t key = ((j + 1u) << 14) | (q & 0x3FFFu); uint32_t h = ((key * 2654435761u) >> hshift) & smask;
							while (true) { const uint32_t v = pset[h]; if (v == key) { search = true; break; } if (v == 0u) { break; } h = (h + 1u) & smask; }
						}
					}
					bool ovl = false;
					if (search) {
						// std::binary_search(first,last,val) = lower_bound + !(val < *it)   (libstdc++ stl_algo.h)
						uint32_t len = (mate == 1) ? n2_ : n1, first = 0;
						while (len > 0) {
							uint32_t half = len >> 1, mid = first + half;
							uint32_t v = (mate == 1) ? *(L2top - mid) : L1[mid];
							if (v < name) { first = mid + 1; len = len - half - 1; } else { len = half; }
						}
						uint32_t cnt_other = (mate == 1) ? n2_ : n1;
						if (first < cnt_other) { uint32_t v = (mate == 1) ? *(L2top - first) : L1[first]; ovl = !(name < v); }
					}
					if (ovl) { ws.inst[(sh->inst_stride ? r : ws.rd_kbase[r]) + p * istr] |= 0x4000u; ws.b_cnt[j * 4 + cls] -= 1; }
					uint32_t last = ws.rd_len[r] - (uint32_t)K;
					uint32_t pushes = (p == 0 || p == last) ? 1u : 2u;
					for (uint32_t q = 0; q < pushes; ++q) { if (mate == 1) { L1[n1++] = name; } else { *(L2top - n2_) = name; ++n2_; } }
				}
			}
			uint32_t t = 0; for (int c = 0; c < 4; ++c) { t += ws.b_cnt[j * 4 + c]; }
			ws.b_mincovqv[j] = (int32_t)t;
		}
		lb2_sync();
	}
	lb2_mark(W, LB2_PH_MATES);
	// ---- low-quality deficits: minqv_{fwd,rev}[i] = count - deficit[i]   (Node_t::updateCovDistr, src/Node.cc:470-497)
	if (sh->has_lowq) {
		uint32_t *d32 = (uint32_t *)ws.deficit;
		if ((size_t)n * K * 8 > (size_t)C->deficit_bytes) { if (tid == 0) { sh->err |= 1u << LB2_D_ARENA; } }
		else {
			for (uint32_t i = tid; i < n * (uint32_t)K * 2; i += nt) { d32[i] = 0; }
			// the low-quality bases of the kept stretches as a list (read << 12 | position): few per read, so one lane per read
			// only collects them; the K k-mers covering each base are then handled by all lanes at once
			uint32_t *lq = (uint32_t *)ws.sortk; const uint32_t lq_cap = C->max_inst;      // (sortk: >= max_inst 8-byte words)
			if (tid == 0) { sh->n_jobs = 0; }
			lb2_sync();
			for (uint32_t r = tid; r < R; r += nt) {
				const uint32_t len = ws.rd_len[r]; if (len <= (uint32_t)K) { continue; }
				const uint32_t g0 = ws.rd_start[r];
				for (uint32_t q0 = 0; q0 < len; ) {      // the mask words of the read, 32 - (g & 31) bases at a time
					const uint32_t g = g0 + q0, sh_ = g & 31u; uint32_t m = lb2_lds(&W.lowq[g >> 5]) >> sh_; const uint32_t nb_ = (32u - sh_ < len - q0) ? 32u - sh_ : len - q0;
					if (nb_ < 32u) { m &= (1u << nb_) - 1u; }
					while (m) { const uint32_t bpos = (uint32_t)lb2_ctz32(m); m &= m - 1u; const uint32_t e = lb2_add32(&sh->n_jobs, 1u); if (e < lq_cap) { lq[e] = (r << 12) | (q0 + bpos); } }
					q0 += nb_;
				}
			}
			lb2_sync();
			if (sh->n_jobs > lq_cap) { if (tid == 0) { sh->err |= 1u << LB2_D_ARENA; } }
			else {
				const uint32_t ne_ = sh->n_jobs;
				for (uint32_t x = tid; x < ne_ * (uint32_t)K; x += nt) {
					const uint32_t e = lq[x / (uint32_t)K], kk = x % (uint32_t)K, r = e >> 12, q = e & 0xFFFu;
					const uint32_t len = ws.rd_len[r];
					if (kk > q || q - kk > len - (uint32_t)K) { continue; }      // k-mer p = q - kk must exist: 0 <= p <= len - K
					const uint32_t p_ = q - kk, kb = sh->inst_stride ? r : ws.rd_kbase[r], cls = ws.rd_info[r] & 3u;
					const uint32_t iw = ws.inst[kb + p_ * istr];
					if (iw & 0x4000u) { continue; }                   // suppressed (overlapping mate)
					const uint32_t id = W.t_id[iw & 0x3FFFu] & 0x7FFFu;
					const uint32_t i = (iw >> 15) ? ((uint32_t)K - 1u - kk) : kk;   // qv string is reversed for ori R (src/Graph.cc:148-158)
					lb2g_red_add(&d32[((size_t)id * K + i) * 2 + (cls >> 1)], (cls & 1) ? 0x10000u : 1u);
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
				ws.b_mincovqv[j] -= (int32_t)mx;
			}
		}
	}
	lb2_sync();
	lb2_mark(W, LB2_PH_LOWQ);
	if (sh->err) { return; }
	// ---- first low-coverage sweep, Graph_t::removeLowCov(false,0) (src/Graph.cc:2790-2827): the rule only looks
	//      at the node itself, so it is a per-node predicate; dead nodes simply never get edges below
	{
		double avgcov = ((double)(int)sh->totalreadbp) / ((double)sh->L);
		double thr = W.P->min_cov_ratio * avgcov;
		for (uint32_t j = tid; j < n; j += nt) {
			int mq = ws.b_mincovqv[j];
			float tt = (float)ws.b_cnt[j * 4 + 0] + (float)ws.b_cnt[j * 4 + 1], tn = (float)ws.b_cnt[j * 4 + 2] + (float)ws.b_cnt[j * 4 + 3];
			if (mq <= W.P->low_cov_threshold || (double)mq <= thr || (tt == 1 && tn == 1)) { ws.b_flags[j] = LB2_NF_DEAD; }
		}
	}
	lb2_sync();
	// ---- edges of the survivors: every edge type (start orientation, appended base) names one neighbour
	{
		uint32_t *slist = (uint32_t *)ws.sortk;      // the survivors (the sort keys are idle by now)
		const uint32_t nlive = lb2_excl_scan(W, n, [&](uint32_t j) -> uint32_t { return (ws.b_flags[j] & LB2_NF_DEAD) ? 0u : 1u; },
		                                     [&](uint32_t j, uint32_t v) { if (!(ws.b_flags[j] & LB2_NF_DEAD)) { slist[v] = j; } });
		const uint32_t ng = lb2_ngroups(), grp = lb2_group();
		for (uint32_t j0 = 0; j0 < nlive; j0 += ng) {
			const bool act = j0 + grp < nlive; const uint32_t j = act ? slist[j0 + grp] : 0u;
			if (nw == 1) { lb2_node_edges<1>(W, act, j, K, nw); } else if (nw == 2) { lb2_node_edges<2>(W, act, j, K, nw); } else { lb2_node_edges<LB2_MAXW>(W, act, j, K, nw); }
		}
	}
	lb2_sync();
	if (sh->flag_a && !sh->err) {
		for (uint32_t j = tid; j < n; j += nt) { if (ws.b_flags[j] & 0x20) { W.t_id[ws.used[j]] |= LB2_ID_BRANCH; } }
		if (tid == 0) { sh->walk_next = 0; }
		lb2_sync();
		// second pass over the occurrence array, same work items as the walk; the occurrence words are fetched eight at a
		// time (the loads do not depend on each other, the lane would otherwise sit out one memory latency per k-mer)
		while (true) {
			const uint32_t it = lb2_batch_next(&sh->walk_next);
			if (it >= ((nitems + 31u) & ~31u)) { break; }
			if (it >= nitems) { continue; }
			uint32_t g0, ob, oe, kb, ks, st;      // kb: occurrence number (first-seen stamps), ks/st: where the occurrences are stored
			if (it < NP * R) {
				const uint32_t piece = it / R, r = it - piece * R, n_ = ws.rd_len[r];
				if (n_ <= (uint32_t)K) { continue; }
				ob = piece * PL; oe = ob + PL; if (oe > n_ - K) { oe = n_ - K; }
				if (ob >= n_ - K) { continue; }
				g0 = ws.rd_start[r]; kb = ws.rd_kbase[r]; ks = sh->inst_stride ? r : kb; st = istr;
			} else {
				if (refN) { const uint32_t v = ws.jobs[it - NP * R]; ob = v & 0xFFFFu; oe = v >> 16; }
				else { ob = (it - NP * R) * PL; oe = ob + PL; if (oe > nref_pairs) { oe = nref_pairs; } }
				g0 = sh->ref_g; kb = ws.rd_kbase[R]; ks = sh->inst_ref; st = 1u;
			}
			uint32_t iu = ws.inst[ks + ob * st];
			for (uint32_t o8 = ob; o8 < oe; o8 += 8) {
				uint32_t v[8];
#pragma unroll
				for (uint32_t q = 0; q < 8; ++q) { v[q] = (o8 + q < oe) ? ws.inst[ks + (o8 + q + 1) * st] : 0u; }
#pragma unroll
				for (uint32_t q = 0; q < 8; ++q) {
					if (o8 + q >= oe) { continue; }
					const uint32_t o = o8 + q, iv = v[q];
					const uint32_t su = iu & 0x3FFFu, sv = iv & 0x3FFFu;
					const uint32_t idu = W.t_id[su], idv = W.t_id[sv];
					if (idu & LB2_ID_BRANCH) {
						uint32_t t = ((iu >> 15) & 1u) * 4 + (uint32_t)lb2_getbase(W.bits, g0 + o + K);
						lb2g_min32(&ws.bseq[(size_t)(idu & 0x7FFFu) * 8 + t], 2 * (kb + o));
					}
					if (idv & LB2_ID_BRANCH) {
						uint32_t t = (1u - ((iv >> 15) & 1u)) * 4 + (uint32_t)(3 - lb2_getbase(W.bits, g0 + o));
						lb2g_min32(&ws.bseq[(size_t)(idv & 0x7FFFu) * 8 + t], 2 * (kb + o) + 1);
					}
					iu = iv;
				}
			}
		}
		lb2_sync();
		for (uint32_t j = tid; j < n; j += nt) {
			if (!(ws.b_flags[j] & 0x20)) { continue; }
			lb2_bedge *e = ws.b_edge + (size_t)j * LB2_BECAP; int ne = ws.b_ne[j];
			for (int a = 1; a < ne; ++a) {
				lb2_bedge x = e[a]; uint32_t sx = ws.bseq[(size_t)j * 8 + x.type]; int b = a - 1;
				while (b >= 0 && ws.bseq[(size_t)j * 8 + e[b].type] > sx) { e[b + 1] = e[b]; --b; }
				e[b + 1] = x;
			}
		}
	}
	lb2_sync();
	lb2_mark(W, LB2_PH_CLEAR);
}

#endif
