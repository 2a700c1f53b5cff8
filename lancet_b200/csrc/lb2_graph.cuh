// lb2_graph.cuh -- the order-sensitive graph stages, executed in the reference's own sweep order
// by one lane of the window's CTA (the graphs are ~600 -> ~10 nodes; throughput comes from the
// hundreds of windows resident on the GPU).
//
//   libstdc++ unordered_map iteration order      SURVEY.md Appendix D  (hashtable.h / hashtable_policy.h)
//   Graph_t::removeLowCov / removeNode / cleanDead   src/Graph.cc:2790-2827, 2768-2784, 2737-2762
//   Graph_t::markConnectedComponents             src/Graph.cc:2252-2336
//   Graph_t::markRefEnds                         src/Graph.cc:2028-2228
//   Graph_t::hasCycle / hasCycleRec              src/Graph.cc:593-681
//   Graph_t::compress / compressNode             src/Graph.cc:2712-2732, 2486-2706
//   Graph_t::removeTips / removeShortLinks       src/Graph.cc:2885-2926, 2833-2880
//   findTandems                                  src/util.cc:574-758
#ifndef LB2_GRAPH_CUH
#define LB2_GRAPH_CUH

#include "lb2_build.cuh"

// ---- edge direction algebra (src/Edge.hh:62-110, src/Edge.cc:25-30) ----------------------------
LB2_DEV int  lb2_dir_start(int d) { return (d == LB2_FF || d == LB2_FR) ? 0 : 1; }   // 0=F 1=R
LB2_DEV int  lb2_dir_dest(int d)  { return (d == LB2_FF || d == LB2_RF) ? 0 : 1; }
LB2_DEV bool lb2_is_dir(int d, int ori) { return lb2_dir_start(d) == ori; }
LB2_DEV int  lb2_flipme(int d)   { return d == LB2_FF ? LB2_RF : d == LB2_FR ? LB2_RR : d == LB2_RF ? LB2_FF : LB2_FR; }
LB2_DEV int  lb2_fliplink(int d) { return d == LB2_FF ? LB2_RR : d == LB2_FR ? LB2_FR : d == LB2_RF ? LB2_RF : LB2_FF; }

// ---- edge storage: 4 half-edges inline per row, nodes that need more move to a 12-slot block of a small pool
#define LB2_EINL 4
#define LB2_EOV_BLOCKS 96
LB2_DEV lb2_edge *lb2_edges(lb2_ws &ws, uint32_t id) {
	uint32_t ov = ws.d_eov[id];
	return ov ? (ws.e_pool + (size_t)(ov - 1) * LB2_ECAP) : (ws.d_edge + (size_t)id * LB2_EINL);
}

// ---- node accessors ------------------------------------------------------------------------------
LB2_DEV bool lb2_special(lb2_win &W, uint32_t id) { return (W.ws.d_flags[id] & LB2_NF_SPECIAL) != 0; }
LB2_DEV uint32_t lb2_strlen(lb2_win &W, uint32_t id) { return lb2_special(W, id) ? 0u : W.ws.d_len[id]; }   // Node_t::strlen src/Node.cc:340-345

LB2_DEV char lb2_node_char(lb2_win &W, uint32_t id, uint32_t i) {
	lb2_ws &ws = W.ws;
	if (ws.d_str[id] == LB2_NIL) {
		uint32_t rep = ws.d_rep[id], g = rep >> 1; int K = W.sh->K;
		if (rep & 1) { return lb2_base(3 - lb2_getbase(W.bits, g + K - 1 - i)); }
		return lb2_base(lb2_getbase(W.bits, g + i));
	}
	return (char)ws.arena[ws.d_str[id] + i];
}
LB2_DEV lb2_cov lb2_node_cov(lb2_win &W, uint32_t id, uint32_t i, int sample /*0=T 1=N*/) {
	lb2_ws &ws = W.ws; lb2_cov c;
	if (ws.d_cd[id] == LB2_NIL) {
		uint32_t f = ws.d_cnt[id * 4 + sample * 2], r = ws.d_cnt[id * 4 + sample * 2 + 1];
		uint32_t df = 0, dr = 0;
		if (W.sh->has_lowq) {
			uint32_t v = ((const uint32_t *)ws.deficit)[((size_t)ws.d_orig[id] * W.sh->K + i) * 2 + sample];
			df = v & 0xFFFF; dr = v >> 16;
		}
		c.fwd = (uint16_t)f; c.rev = (uint16_t)r; c.mqf = (uint16_t)(f - df); c.mqr = (uint16_t)(r - dr);
		return c;
	}
	const lb2_cov *a = (const lb2_cov *)(ws.arena + ws.d_cd[id]);
	return a[(size_t)sample * ws.d_len[id] + i];
}
// a node's cold fields loaded once (they live in global memory); element access is then cheap
struct lb2_nview { uint32_t str, cd, rep, len, orig; uint32_t cnt[4]; };
LB2_DEV void lb2_view(lb2_win &W, uint32_t id, lb2_nview &v) {
	lb2_ws &ws = W.ws;
	v.str = ws.d_str[id]; v.cd = ws.d_cd[id]; v.rep = ws.d_rep[id]; v.len = ws.d_len[id]; v.orig = ws.d_orig[id];
	for (int c = 0; c < 4; ++c) { v.cnt[c] = ws.d_cnt[id * 4 + c]; }
}
LB2_DEV char lb2_vchar(lb2_win &W, const lb2_nview &v, uint32_t i) {
	if (v.str == LB2_NIL) {
		uint32_t g = v.rep >> 1; int K = W.sh->K;
		if (v.rep & 1) { return lb2_base(3 - lb2_getbase(W.bits, g + K - 1 - i)); }
		return lb2_base(lb2_getbase(W.bits, g + i));
	}
	return (char)W.ws.arena[v.str + i];
}
LB2_DEV lb2_cov lb2_vcov(lb2_win &W, const lb2_nview &v, uint32_t i, int sample) {
	lb2_cov c;
	if (v.cd == LB2_NIL) {
		uint32_t f = v.cnt[sample * 2], r = v.cnt[sample * 2 + 1], df = 0, dr = 0;
		if (W.sh->has_lowq) {
			uint32_t x = ((const uint32_t *)W.ws.deficit)[((size_t)v.orig * W.sh->K + i) * 2 + sample];
			df = x & 0xFFFF; dr = x >> 16;
		}
		c.fwd = (uint16_t)f; c.rev = (uint16_t)r; c.mqf = (uint16_t)(f - df); c.mqr = (uint16_t)(r - dr);
		return c;
	}
	const lb2_cov *a = (const lb2_cov *)(W.ws.arena + v.cd);
	return a[(size_t)sample * v.len + i];
}
LB2_DEV float lb2_totcov(lb2_win &W, uint32_t id) {   // Node_t::getTotCov (src/Node.hh:151), same association order
	float *c = W.ws.d_cov + id * 4; return c[0] + c[1] + c[2] + c[3];
}
LB2_DEV uint32_t lb2_arena_alloc(lb2_win &W, uint32_t bytes) {
	uint32_t o = (W.sh->arena_used + 7u) & ~7u;
	if (o + bytes > W.C->arena_bytes) { W.sh->err |= 1u << LB2_D_ARENA; return 0; }
	W.sh->arena_used = o + bytes; return o;
}

// ---- libstdc++ _Hashtable order emulation -----------------------------------------------------------
LB2_DEV uint32_t lb2_bget(lb2_win &W, uint32_t b) { uint32_t v = W.ws.buckets[b]; return v >= 0xFFFEu ? (v | 0xFFFF0000u) : v; }   // 0xFFFF -> LB2_NIL, 0xFFFE -> LB2_SENT
LB2_DEV void lb2_bset(lb2_win &W, uint32_t b, uint32_t v) { W.ws.buckets[b] = (uint16_t)v; }
LB2_DEV uint32_t lb2_oe_next(lb2_win &W, uint32_t x) { return x == LB2_SENT ? W.sh->lhead : W.ws.d_lnext[x]; }
LB2_DEV void lb2_oe_setnext(lb2_win &W, uint32_t x, uint32_t v) { if (x == LB2_SENT) { W.sh->lhead = v; } else { W.ws.d_lnext[x] = v; } }
LB2_DEV uint32_t lb2_oe_next_bkt(uint32_t x) {
	const uint32_t chain[] = { 13, 29, 59, 127, 257, 541, 1109, 2357, 5087, 10273, 20753, 42043, 85229, 172933 };
	for (int i = 0; i < 14; ++i) { if (chain[i] >= x) { return chain[i]; } }
	return 0;
}
LB2_DEV void lb2_oe_reset(lb2_win &W) {
	lb2_sh *sh = W.sh; sh->bkt_count = 1; sh->elem_count = 0; sh->next_resize = 0; sh->lhead = LB2_NIL;
	lb2_bset(W, 0, LB2_NIL);
}
LB2_DEV void lb2_oe_rehash(lb2_win &W, uint32_t nb) {   // _M_rehash_aux (unique keys); rare here (only a source/sink insert can trigger it)
	lb2_ws &ws = W.ws; lb2_sh *sh = W.sh;
	if (nb > W.sh->bkt_cap) { sh->err |= 1u << LB2_D_BUCKETS; return; }
	for (uint32_t b = 0; b < nb; ++b) { lb2_bset(W, b, LB2_NIL); }
	uint32_t p = sh->lhead; sh->lhead = LB2_NIL; uint32_t bbegin = 0;
	while (p != LB2_NIL) {
		uint32_t nx = ws.d_lnext[p];
		uint32_t b = (uint32_t)(ws.d_hash[p] % nb); ws.d_bk[p] = (uint16_t)b;
		if (lb2_bget(W, b) == LB2_NIL) {
			ws.d_lnext[p] = sh->lhead; sh->lhead = p; lb2_bset(W, b, LB2_SENT);
			if (ws.d_lnext[p] != LB2_NIL) { lb2_bset(W, bbegin, p); }
			bbegin = b;
		} else {
			uint32_t before = lb2_bget(W, b);
			ws.d_lnext[p] = lb2_oe_next(W, before); lb2_oe_setnext(W, before, p);
		}
		p = nx;
	}
	sh->bkt_count = nb;
}
LB2_DEV void lb2_oe_insert(lb2_win &W, uint32_t id) {   // _M_insert_unique_node + _Prime_rehash_policy::_M_need_rehash
	lb2_ws &ws = W.ws; lb2_sh *sh = W.sh;
	uint32_t n = sh->elem_count;
	if (n + 1 > sh->next_resize) {
		uint32_t min_bkts = n + 1; if (!sh->next_resize && min_bkts < 11) { min_bkts = 11; }
		if (min_bkts >= sh->bkt_count) {
			uint32_t want = min_bkts + 1; if (want < sh->bkt_count * 2) { want = sh->bkt_count * 2; }
			uint32_t nb = lb2_oe_next_bkt(want);
			if (nb == 0) { sh->err |= 1u << LB2_D_BUCKETS; return; }
			sh->next_resize = nb;
			lb2_oe_rehash(W, nb);
			if (sh->err) { return; }
		} else { sh->next_resize = sh->bkt_count; }
	}
	uint32_t b = (uint32_t)(ws.d_hash[id] % sh->bkt_count); ws.d_bk[id] = (uint16_t)b;
	if (lb2_bget(W, b) != LB2_NIL) {
		uint32_t before = lb2_bget(W, b);
		ws.d_lnext[id] = lb2_oe_next(W, before); lb2_oe_setnext(W, before, id);
	} else {
		ws.d_lnext[id] = sh->lhead; sh->lhead = id;
		if (ws.d_lnext[id] != LB2_NIL) { lb2_bset(W, ws.d_bk[ws.d_lnext[id]], id); }
		lb2_bset(W, b, LB2_SENT);
	}
	sh->elem_count = n + 1;
}
LB2_DEV void lb2_oe_erase(lb2_win &W, uint32_t id) {    // _M_erase(bkt, prev, n)
	lb2_ws &ws = W.ws; lb2_sh *sh = W.sh;
	uint32_t b = ws.d_bk[id];
	uint32_t prev = lb2_bget(W, b);
	while (lb2_oe_next(W, prev) != id) { prev = lb2_oe_next(W, prev); }
	uint32_t nx = ws.d_lnext[id];
	if (prev == lb2_bget(W, b)) {
		uint32_t nb = (nx != LB2_NIL) ? ws.d_bk[nx] : 0;
		if (nx == LB2_NIL || nb != b) {
			if (nx != LB2_NIL) { lb2_bset(W, nb, lb2_bget(W, b)); }
			lb2_bset(W, b, LB2_NIL);   // (before_begin.next is updated by the unlink below)
		}
	} else if (nx != LB2_NIL) {
		uint32_t nb = ws.d_bk[nx];
		if (nb != b) { lb2_bset(W, nb, prev); }
	}
	lb2_oe_setnext(W, prev, nx);
	sh->elem_count -= 1;
	ws.d_flags[id] |= LB2_NF_GONE;
}

// ---- edge helpers (Node_t::removeEdge / updateEdge, src/Node.cc:177-233) ------------------------------
LB2_DEV void lb2_remove_edge(lb2_win &W, uint32_t id, uint32_t to, int dir) {
	lb2_ws &ws = W.ws; lb2_edge *e = lb2_edges(ws, id); int ne = ws.d_ne[id];
	for (int i = 0; i < ne; ++i) {
		if (e[i].to == to && e[i].dir == dir) {
			for (int j = i; j + 1 < ne; ++j) { e[j] = e[j + 1]; }
			ws.d_ne[id] = (uint8_t)(ne - 1); return;
		}
	}
	W.sh->err |= 1u << LB2_D_EDGES;   // the reference would assert here
}
LB2_DEV void lb2_update_edge(lb2_win &W, uint32_t id, uint32_t oldto, int olddir, uint32_t newto, int newdir) {
	lb2_ws &ws = W.ws; lb2_edge *e = lb2_edges(ws, id); int ne = ws.d_ne[id];
	for (int i = 0; i < ne; ++i) {
		if (e[i].to == oldto && e[i].dir == olddir) { e[i].to = (uint16_t)newto; e[i].dir = (uint16_t)newdir; return; }
	}
	W.sh->err |= 1u << LB2_D_EDGES;
}
LB2_DEV void lb2_push_edge(lb2_win &W, uint32_t id, uint32_t to, int dir, int flag) {
	lb2_ws &ws = W.ws; int ne = ws.d_ne[id];
	if (ne >= LB2_ECAP) { W.sh->err |= 1u << LB2_D_EDGES; return; }
	if (ne == LB2_EINL && !ws.d_eov[id]) {          // outgrew the inline slots: move to a pool block
		uint32_t blk = W.sh->n_eov;
		if (blk >= LB2_EOV_BLOCKS) { W.sh->err |= 1u << LB2_D_EDGES; return; }
		W.sh->n_eov = blk + 1;
		lb2_edge *src = ws.d_edge + (size_t)id * LB2_EINL, *dst = ws.e_pool + (size_t)blk * LB2_ECAP;
		for (int i = 0; i < LB2_EINL; ++i) { dst[i] = src[i]; }
		ws.d_eov[id] = (uint8_t)(blk + 1);
	}
	lb2_edge ed; ed.to = (uint16_t)to; ed.dir = (uint16_t)dir; ed.flag = (uint16_t)flag; ed.pad = 0;
	lb2_edges(ws, id)[ne] = ed; ws.d_ne[id] = (uint8_t)(ne + 1);
}
LB2_DEV void lb2_add_edge_node(lb2_win &W, uint32_t id, uint32_t to, int dir) {   // Node_t::addEdge without read ids
	lb2_ws &ws = W.ws; lb2_edge *e = lb2_edges(ws, id); int ne = ws.d_ne[id];
	for (int i = 0; i < ne; ++i) { if (e[i].to == to && e[i].dir == dir) { return; } }
	lb2_push_edge(W, id, to, dir, 0);
}
LB2_DEV void lb2_remove_node(lb2_win &W, uint32_t id) {   // Graph_t::removeNode
	lb2_ws &ws = W.ws; ws.d_flags[id] |= LB2_NF_DEAD;
	lb2_edge *e = lb2_edges(ws, id); int ne = ws.d_ne[id];
	for (int i = 0; i < ne; ++i) { if (e[i].to != id) { lb2_remove_edge(W, e[i].to, id, lb2_fliplink(e[i].dir)); } }
}
// ---- the sweeps of one component (lane 0).  The reference walks the whole map and skips the nodes of other components;
//      with many small components (large k, short reads) that is components x nodes dependent loads.  Once a component
//      has been compacted for the first time its nodes are threaded on a list of their own, in map order (an erase keeps
//      the relative order, and nothing is inserted while a component is being processed); nodes the map has dropped since
//      are unlinked on the way.  f(p) returns false to stop the walk.
LB2_DEV bool lb2_special(lb2_win &W, uint32_t id);
template <class F> LB2_DEV void lb2_each_node(lb2_win &W, int compid, F f) {
	lb2_ws &ws = W.ws; lb2_sh *sh = W.sh;
	if (!sh->cm_valid) {
		uint32_t p = sh->lhead;
		while (p != LB2_NIL) { const uint32_t nx = ws.d_lnext[p]; if (ws.d_comp[p] == compid && !lb2_special(W, p)) { if (!f(p)) { return; } } p = nx; }
		return;
	}
	uint32_t prev = LB2_NIL, p = sh->chead;
	while (p != LB2_NIL) {
		const uint16_t v = ws.d_cnext[p]; const uint32_t nx = (v == 0xFFFFu) ? LB2_NIL : (uint32_t)v;
		bool go = true;
		if (!(ws.d_flags[p] & LB2_NF_GONE)) { go = f(p); }
		if (ws.d_flags[p] & LB2_NF_GONE) { if (prev == LB2_NIL) { sh->chead = nx; } else { ws.d_cnext[prev] = v; } } else { prev = p; }
		if (!go) { return; }
		p = nx;
	}
}
LB2_DEV void lb2_build_members(lb2_win &W, int compid) {
	lb2_ws &ws = W.ws; lb2_sh *sh = W.sh; uint32_t last = LB2_NIL; sh->chead = LB2_NIL;
	for (uint32_t p = sh->lhead; p != LB2_NIL; p = ws.d_lnext[p]) {
		if (ws.d_comp[p] != compid || lb2_special(W, p)) { continue; }
		if (last == LB2_NIL) { sh->chead = p; } else { ws.d_cnext[last] = (uint16_t)p; }
		last = p;
	}
	if (last != LB2_NIL) { ws.d_cnext[last] = 0xFFFFu; }
	sh->cm_valid = 1;
}
LB2_DEV void lb2_clean_dead(lb2_win &W, int compid) {     // Graph_t::cleanDead (every dead node belongs to the component being swept)
	lb2_ws &ws = W.ws;
	lb2_each_node(W, compid, [&](uint32_t p) -> bool { if (ws.d_flags[p] & LB2_NF_DEAD) { lb2_oe_erase(W, p); } return true; });
}
LB2_DEV bool lb2_is_tandem(lb2_win &W, uint32_t id) {     // Node_t::isTandem
	lb2_edge *e = lb2_edges(W.ws, id); int ne = W.ws.d_ne[id];
	for (int i = 0; i < ne; ++i) { if (e[i].to == id) { return true; } }
	return false;
}
LB2_DEV int lb2_get_buddy(lb2_win &W, uint32_t id, int ori) {   // Node_t::getBuddy
	if (lb2_special(W, id)) { return -1; }
	lb2_edge *e = lb2_edges(W.ws, id); int ne = W.ws.d_ne[id]; int r = -1;
	for (int i = 0; i < ne; ++i) { if (lb2_is_dir(e[i].dir, ori)) { if (r != -1) { return -1; } r = i; } }
	if (r != -1 && e[r].to == id) { return -1; }
	return r;
}

// ---- findTandems (src/util.cc:574-758) for ONE query position, scalar, no scratch memory.  Same event formulation as
//      the CTA-wide lb2_path_tandems (lb2_paths.cuh): for a unit length m the reference compares the unit at i with the
//      unit at the last position of the same phase where that comparison failed -- and every unit in between equals
//      both, so the unit one period back (i - m) serves as well.  A position whose comparison fails (or that reaches the
//      end of the string) closes a run of period m; the run's first unit is found by stepping back over positions whose
//      comparison succeeded.  Runs that are long enough, are not preceded by a further copy of the unit's last base and
//      have no shorter period are the reference's candidate repeats, met in (i, m) order; those within `delta` of `pos`
//      set LEN (the last one wins) and append their unit to the motif.
template <class GetC>
LB2_DEV bool lb2_find_tandems(GetC getc, uint32_t slen, const lb2_params *P, int pos, int &len, char *motif, uint32_t &mlen, uint32_t mcap, bool &movf,
                              uint32_t i_first = 0, uint32_t i_step = 1)      // (i_first, i_step: this caller's share of the positions, for callers that only need "any")
{
	const uint32_t maxu = (uint32_t)P->max_unit_len < 16u ? (uint32_t)P->max_unit_len : 16u;
	const int delta = P->dist_from_str;
	// bases of the unit at i that agree with the unit one period back (the first unit of a phase is compared with itself)
	auto agree = [&](uint32_t i, uint32_t m) -> uint32_t {
		const uint32_t back = (i >= m) ? i - m : i; uint32_t j = 0;
		while (j < m && i + j < slen && getc(i + j) == getc(back + j)) { ++j; }
		return j;
	};
	auto closes = [&](uint32_t i, uint32_t m, uint32_t j) -> bool { return j != m || i + j + 1 == slen; };
	bool found = false;
	for (uint32_t i = i_first; i < slen; i += i_step) {
		for (uint32_t m = 1; m <= maxu; ++m) {
			const uint32_t j = agree(i, m);
			if (!closes(i, m, j)) { continue; }
			int q = (int)i - (int)m;
			while (q >= 0 && !closes((uint32_t)q, m, agree((uint32_t)q, m))) { q -= (int)m; }
			const uint32_t first = (q >= 0) ? (uint32_t)q : i % m, span = i - first;
			if (span / m < (uint32_t)P->min_report_units || span < (uint32_t)P->min_report_len) { continue; }
			// seq[first-1] at first == 0 reads the byte before the buffer: 0 in practice
			const char before = first ? getc(first - 1) : (char)0;
			if (before == getc(first + m - 1)) { continue; }
			// a shorter period d < m that tiles the whole run (whole units of length d only) disqualifies the unit
			bool shorter = false;
			for (uint32_t d = 1; d < m && !shorter; ++d) {
				const uint32_t copies = (span + j) / d; bool tiles = true;
				for (uint32_t c = 1; tiles && c < copies; ++c) { for (uint32_t x = 0; x < d; ++x) { if (getc(first + x) != getc(first + c * d + x)) { tiles = false; break; } } }
				shorter = tiles;
			}
			if (shorter) { continue; }
			const int lo = (int)first, hi = (int)(i + j);
			if (pos < lo - delta || pos > hi + delta) { continue; }
			found = true; len = hi - lo;
			for (uint32_t z = 0; z < m; ++z) { if (mlen < mcap) { motif[mlen++] = getc(first + z); } else { movf = true; } }
		}
	}
	return found;
}

// ---------------------------------------------------------------------------------------------------
// sequential stages (lane 0)
// ---------------------------------------------------------------------------------------------------
// ---------------------------------------------------------------------------------------------------
// Iteration order of the reference's unordered_map over ALL nodes (insertion order = dense id), the
// first low-coverage sweep, and the hand-over to the graph stage:
//  1. emulate the map in shared memory with 16-bit links.  The bucket count only changes when element
//     number B+1 arrives (13, 29, 59, ... SURVEY App. D), so the work is done level by level: all lanes
//     compute hash % B for the level, lane 0 relinks (rehash) and inserts the level's elements;
//  2. survivors of removeLowCov(false,0) get compact "row" ids; lane 0 walks the emulated list, drops the
//     dead nodes (unordered_map::erase keeps the relative order) and rebuilds the bucket heads;
//  3. the hot per-node arrays of the graph stage are laid out in shared memory (the region that held
//     the low-quality mask and the Mer->Node table) and filled from the build-space arrays.
// ---------------------------------------------------------------------------------------------------
LB2_DEV uint32_t lb2_level_bkt(uint32_t n_before) {   // bucket count in force while inserting element index n_before (0-based)
	const uint32_t chain[] = { 13, 29, 59, 127, 257, 541, 1109, 2357, 5087, 10273, 20753, 42043, 85229, 172933 };
	for (int i = 0; i < 14; ++i) { if (n_before < chain[i]) { return chain[i]; } }
	return 0;
}

// Iteration order of the reference's map after the build, computed by all lanes (no list is walked):
// an insert into an empty bucket goes to the list head, an insert into a non-empty bucket goes to the front of that
// bucket's run, and _M_rehash_aux re-places every element by the same two rules in old list order.  Hence after any
// sequence of placements the list is: buckets by DEscending creation time, inside a bucket elements by DEscending
// placement time.  Per level (bucket count B) the placement time of an element is its list position
// before the rehash (old elements) or its dense id (elements inserted at this level), and
//     new position = (number of elements in buckets created later) + (number of later elements in its own bucket).
// tm/tm2/bk: u16[n] (tm = time, tm2 = new position / rank, bk = bucket); nx/S/cm: u16[n]; head: u32[B].
// Returns the array holding the final list position of every dense node; *bk_out = bucket at the final bucket count,
// *free_out = the other u16[n] array (free for the caller).
LB2_DEVNI uint16_t *lb2_emulate_order(lb2_win &W, uint32_t n, uint16_t *tm, uint16_t *tm2, uint16_t *bk, uint16_t *nx, uint16_t *S, uint16_t *cm, uint32_t *head, uint16_t **free_out)
{
	lb2_ws &ws = W.ws; const unsigned tid = lb2_tid(), nt = lb2_nthr();
	uint32_t done = 0;
	while (done < n) {
		const uint32_t B = lb2_level_bkt(done); const uint32_t end = B < n ? B : n;     // elements [done, end) arrive at this bucket count
		for (uint32_t j = tid; j < end; j += nt) { bk[j] = (uint16_t)(ws.b_hash[j] % B); if (j >= done) { tm[j] = (uint16_t)j; } }
		for (uint32_t b = tid; b < B; b += nt) { head[b] = LB2_NIL; }
		lb2_sync();
		for (uint32_t j = tid; j < end; j += nt) { nx[j] = (uint16_t)lb2x_exch32(&head[bk[j]], j); }
		lb2_sync();
		for (uint32_t j = tid; j < end; j += nt) {
			const uint32_t tj = tm[j]; uint32_t cmin = tj, cnt = 0, rank = 0;
			for (uint32_t y = head[bk[j]]; y != LB2_NIL; ) {
				uint32_t ty = tm[y]; ++cnt; if (ty < cmin) { cmin = ty; } if (ty > tj) { ++rank; }
				uint32_t nxt = nx[y]; y = (nxt == 0xFFFFu) ? LB2_NIL : nxt;
			}
			S[tj] = (uint16_t)((tj == cmin) ? cnt : 0u); cm[j] = (uint16_t)cmin; tm2[j] = (uint16_t)rank;
		}
		lb2_sync();
		// S[t] := number of elements in buckets created after time t (exclusive suffix sum)
		lb2_excl_scan(W, end, [&](uint32_t i) -> uint32_t { return S[end - 1 - i]; }, [&](uint32_t i, uint32_t v) { S[end - 1 - i] = (uint16_t)v; });
		for (uint32_t j = tid; j < end; j += nt) { tm2[j] = (uint16_t)(S[cm[j]] + tm2[j]); }
		lb2_sync();
		uint16_t *t = tm; tm = tm2; tm2 = t;
		done = end;
	}
	*free_out = tm2;
	return tm;
}

LB2_DEVNI void lb2_order_and_pack(lb2_win &W)
{
	lb2_ws &ws = W.ws; lb2_sh *sh = W.sh; const unsigned tid = lb2_tid(), nt = lb2_nthr();
	const uint32_t n = sh->n_nodes; const int K = sh->K;
	const uint32_t Bfinal = n ? lb2_level_bkt(n - 1) : 13;
	// graph region: [low-quality mask | table region], both dead now
	uint8_t *G = (uint8_t *)W.lowq; const size_t Gbytes = (size_t)(W.treg - (uint8_t *)W.lowq) + lb2_treg_bytes(W.C->table_slots, W.C->graph_bytes, W.C->max_bp);
	// rows for the survivors (dense-id order); their k-mer spellings leave the packed reads for the arena, because the
	// packed reads' shared memory is about to hold the all-node emulation arrays
	uint32_t nrows = lb2_excl_scan(W, n, [&](uint32_t j) -> uint32_t { return (ws.b_flags[j] & LB2_NF_DEAD) ? 0u : 1u; },
	                               [&](uint32_t j, uint32_t v) { ws.b_row[j] = (ws.b_flags[j] & LB2_NF_DEAD) ? LB2_NIL : v; });
	if (tid == 0) {
		sh->n_rows = nrows; sh->n_spec = 0; sh->arena_used = 8 + nrows * (uint32_t)K;
		if (sh->arena_used + 64 > W.C->arena_bytes) { sh->err |= 1u << LB2_D_ARENA; }
	}
	lb2_sync();
	if (sh->err) { return; }
	for (uint32_t j = tid; j < n; j += nt) {
		uint32_t r = ws.b_row[j]; if (r == LB2_NIL) { continue; }
		uint32_t rep = ws.b_rep[j], g0 = rep >> 1; char *dst = (char *)ws.arena + 8 + (size_t)r * K;
		if (rep & 1) { for (int i = 0; i < K; ++i) { dst[i] = lb2_base(3 - lb2_getbase(W.bits, g0 + K - 1 - i)); } }
		else { for (int i = 0; i < K; ++i) { dst[i] = lb2_base(lb2_getbase(W.bits, g0 + i)); } }
	}
	lb2_sync();
	// row-space layout (filled after the emulation: its temporaries borrow the same shared memory)
	// rows = survivors + room for the source/sink nodes of the anchored components: as many as the configuration allows
	// (cfg.max_special) and the shared-memory graph region holds next to the survivors, never fewer than 32
	const uint32_t NR = sh->n_rows; const size_t bits_bytes = ((size_t)W.C->max_bp / 16 + 4) * 4;
	uint32_t spec_cap = W.C->max_special < 32u ? 32u : W.C->max_special, NT = 0, bcap = 0;
	size_t rows_bytes = 0;
	while (true) {
		NT = NR + spec_cap;
		bcap = Bfinal; if (NT > Bfinal) { bcap = lb2_level_bkt(NT); }      // a source/sink insert may still trigger a rehash
		size_t off = 0;
#define LB2_GT(field, type, count) do { off = (off + 7) & ~(size_t)7; if (tid == 0) { ws.field = (type *)(G + off); } off += sizeof(type) * (size_t)(count); } while (0)
		LB2_GT(d_lnext, uint32_t, NT); LB2_GT(d_bk, uint16_t, NT); LB2_GT(d_cnext, uint16_t, NT); LB2_GT(buckets, uint16_t, bcap);
		LB2_GT(d_cov, float, NT * 4); LB2_GT(stack, uint32_t, NT + 8); LB2_GT(cpos, uint32_t, NT + 8);
		LB2_GT(d_edge, lb2_edge, NT * LB2_EINL); LB2_GT(e_pool, lb2_edge, LB2_EOV_BLOCKS * LB2_ECAP);
		LB2_GT(d_len, uint16_t, NT); LB2_GT(d_stn, uint16_t, NT); LB2_GT(d_stT, uint16_t, NT); LB2_GT(d_comp, int16_t, NT);
		LB2_GT(d_ne, uint8_t, NT); LB2_GT(d_flags, uint8_t, NT); LB2_GT(d_color, uint8_t, NT); LB2_GT(d_eov, uint8_t, NT);
#undef LB2_GT
		if (tid == 0) { ws.chain = ws.stack; } rows_bytes = (off + 15) & ~(size_t)15;
		if (spec_cap <= 32u || (NT <= LB2_MAX_ROWS && rows_bytes <= Gbytes && (size_t)NT * 2 + 16 <= bits_bytes)) {
			if (tid == 0) { sh->spec_cap = spec_cap; }
			// the node list of the current path (lane-0 code walks it over and over) moves in as well when there is room
			const size_t pbytes = sizeof(uint32_t) * (2 * (size_t)LB2_MAX_PNODES + 1) + 2 * (size_t)LB2_MAX_PNODES + 16;
			if (rows_bytes <= Gbytes && rows_bytes + pbytes <= Gbytes && tid == 0) {
				uint8_t *q = G + rows_bytes;
				ws.pnodes = (uint32_t *)q; q += sizeof(uint32_t) * LB2_MAX_PNODES; ws.pstart = (uint32_t *)q; q += sizeof(uint32_t) * (LB2_MAX_PNODES + 1);
				ws.pdirs = q; q += LB2_MAX_PNODES; ws.peidx = q;
			}
			break;
		}
		spec_cap >>= 1;
	}
	// scratch of the graph stage in the (now dead) packed-read words: list index of every row, then the parallel
	// compaction's words.  The emulation arrays: three u16[n] here, the rest over the graph region; when either does not
	// fit, all of them live in the workspace slab instead.
	if (tid == 0) {
		ws.d_pos = (uint16_t *)W.bits; ws.px = (uint32_t *)((uint8_t *)W.bits + (((size_t)NT * 2 + 15) & ~(size_t)15));
		ws.px_words = (bits_bytes > (((size_t)NT * 2 + 15) & ~(size_t)15)) ? (uint32_t)((bits_bytes - (((size_t)NT * 2 + 15) & ~(size_t)15)) / 4) : 0u;
	}
	lb2_sync();
	const size_t n6 = ((size_t)n * 6 + 15) & ~(size_t)15;
	if (Bfinal == 0 || n >= 0x7FF0u || NT > LB2_MAX_ROWS || rows_bytes > Gbytes || (size_t)NT * 2 + 16 > bits_bytes || n > W.C->max_nodes) { if (tid == 0) { sh->err |= 1u << LB2_D_SMEM; } lb2_sync(); return; }
	uint16_t *tm, *tm2, *bk, *nx, *S, *cm; uint32_t *head;
	if (n6 <= bits_bytes && n6 + (size_t)Bfinal * 4 <= Gbytes) {
		tm = (uint16_t *)W.bits; tm2 = tm + n; bk = tm2 + n;
		nx = (uint16_t *)G; S = nx + n; cm = S + n; head = (uint32_t *)(G + n6);
	} else {
		tm = (uint16_t *)ws.emu; tm2 = tm + n; bk = tm2 + n; nx = bk + n; S = nx + n; cm = S + n; head = (uint32_t *)((uint8_t *)ws.emu + (((size_t)n * 12 + 15) & ~(size_t)15));
	}
	if (tid == 0) { sh->lowq_live = 0; sh->bits_live = 0; }
	uint16_t *ord = nullptr;
	uint16_t *pos = lb2_emulate_order(W, n, tm, tm2, bk, nx, S, cm, head, &ord);
	for (uint32_t j = tid; j < n; j += nt) { ord[pos[j]] = (uint16_t)j; }     // dense node at every list position
	lb2_sync();
	lb2_mark(W, LB2_PH_ORDER);
	if (tid == 0) { sh->bkt_cap = bcap; sh->bkt_count = Bfinal; sh->next_resize = Bfinal; sh->elem_count = NR; sh->n_eov = 0; }
	for (uint32_t r = tid; r < NT; r += nt) { ws.d_eov[r] = 0; ws.d_ne[r] = 0; }
	lb2_sync();
	for (uint32_t j = tid; j < n; j += nt) {
		uint32_t r = ws.b_row[j]; if (r == LB2_NIL) { continue; }
		ws.d_bk[r] = bk[j];
		uint32_t tot = 0;
		for (int c = 0; c < 4; ++c) { uint32_t v = ws.b_cnt[j * 4 + c]; ws.d_cov[r * 4 + c] = (float)v; ws.d_cnt[r * 4 + c] = v; tot += v; }
		ws.d_len[r] = (uint32_t)K; ws.d_stn[r] = 1; ws.d_stT[r] = ws.b_stT[j]; ws.d_comp[r] = 0;
		ws.d_flags[r] = 0; ws.d_color[r] = 0;
		ws.d_rep[r] = ws.b_rep[j]; ws.d_hash[r] = ws.b_hash[j]; ws.d_orig[r] = j;
		ws.d_mincov[r] = (int32_t)tot; ws.d_mincovqv[r] = ws.b_mincovqv[j]; ws.d_str[r] = 8 + r * (uint32_t)K; ws.d_cd[r] = LB2_NIL;
		const int ne = ws.b_ne[j];
		lb2_edge *dst = ws.d_edge + (size_t)r * LB2_EINL;
		if (ne > LB2_EINL) {          // more than the inline slots: a block of the pool
			uint32_t blk = lb2_add32(&sh->n_eov, 1u);
			if (blk >= LB2_EOV_BLOCKS) { lb2_or32(&sh->err, 1u << LB2_D_EDGES); continue; }
			ws.d_eov[r] = (uint8_t)(blk + 1); dst = ws.e_pool + (size_t)blk * LB2_ECAP;
		}
		for (int e = 0; e < ne; ++e) {
			lb2_bedge be = ws.b_edge[(size_t)j * LB2_BECAP + e];
			lb2_edge ed; ed.to = (uint16_t)ws.b_row[be.to]; ed.dir = (uint16_t)be.dir; ed.flag = 0; ed.pad = 0;
			dst[e] = ed;
		}
		ws.d_ne[r] = (uint8_t)ne;
	}
	for (uint32_t p = tid; p < LB2_MAX_REF; p += nt) { uint32_t j = ws.refnode[p]; if (j != LB2_NIL) { ws.refnode[p] = ws.b_row[j]; } }
	for (uint32_t b = tid; b < bcap; b += nt) { lb2_bset(W, b, LB2_NIL); }
	// the survivors in list order (erase keeps the relative order): rows at list index 0..NR-1
	uint32_t *rowlist = ws.stack;
	lb2_excl_scan(W, n, [&](uint32_t p) -> uint32_t { return (ws.b_row[ord[p]] != LB2_NIL) ? 1u : 0u; },
	              [&](uint32_t p, uint32_t v) { uint32_t r = ws.b_row[ord[p]]; if (r != LB2_NIL) { rowlist[v] = r; } });
	if (sh->n_eov > LB2_EOV_BLOCKS && tid == 0) { sh->n_eov = LB2_EOV_BLOCKS; }
	lb2_sync();
	for (uint32_t li = tid; li < NR; li += nt) {
		const uint32_t r = rowlist[li], b = ws.d_bk[r];
		ws.d_lnext[r] = (li + 1 < NR) ? rowlist[li + 1] : LB2_NIL;
		if (li == 0) { lb2_bset(W, b, LB2_SENT); }
		else { uint32_t pr = rowlist[li - 1]; if (ws.d_bk[pr] != b) { lb2_bset(W, b, pr); } }      // bucket -> node before its first node
	}
	if (tid == 0) { sh->lhead = NR ? rowlist[0] : LB2_NIL; }
	lb2_sync();
	for (uint32_t li = tid; li < NR; li += nt) { ws.d_pos[rowlist[li]] = (uint16_t)li; }     // (d_pos overlays the emulation arrays: they are dead now)
	lb2_sync();
	lb2_mark(W, LB2_PH_LOWCOV_CC);
}

// removeLowCov(docompression=false path and the sweep part) src/Graph.cc:2790-2827
LB2_DEVNI void lb2_remove_lowcov(lb2_win &W, int compid) {
	lb2_ws &ws = W.ws; lb2_sh *sh = W.sh;
	double avgcov = ((double)(int)sh->totalreadbp) / ((double)sh->L);
	double thr = W.P->min_cov_ratio * avgcov;
	uint32_t removed = 0;
	lb2_each_node(W, compid, [&](uint32_t p) -> bool {
		int mq = ws.d_mincovqv[p];
		float tt = ws.d_cov[p * 4 + 0] + ws.d_cov[p * 4 + 1], tn = ws.d_cov[p * 4 + 2] + ws.d_cov[p * 4 + 3];
		if (mq <= W.P->low_cov_threshold || (double)mq <= thr || (tt == 1 && tn == 1)) { lb2_remove_node(W, p); ++removed; }
		return true;
	});
	sh->flag_b = removed; sh->n_changed = removed;      // (nothing removed: the graph is still fully compacted and the compaction that follows is a no-op)
	if (removed) { lb2_clean_dead(W, compid); }
}

// Graph_t::markConnectedComponents (src/Graph.cc:2252-2336) by all lanes: the reference numbers a component when its
// first node comes up in map iteration order and labels everything reachable from it; edges are reciprocal, so the
// labels are the connected components numbered by their smallest list index.  Min-label hooking + pointer jumping
// over list indices (P lives in ws.cpos, the numbering in ws.stack).  Returns the number of components.
LB2_DEVNI int lb2_mark_components(lb2_win &W) {
	lb2_ws &ws = W.ws; lb2_sh *sh = W.sh; const unsigned tid = lb2_tid(), nt = lb2_nthr();
	const uint32_t NR = sh->n_rows; uint32_t *P = ws.cpos; uint32_t *num = ws.stack;
	for (uint32_t i = tid; i < NR; i += nt) { P[i] = i; }
	lb2_sync();
	for (uint32_t round = 0; round < 2 * LB2_MAX_ROWS; ++round) {
		for (uint32_t r = tid; r < NR; r += nt) {
			const uint32_t pi = P[ws.d_pos[r]];
			const lb2_edge *e = lb2_edges(ws, r); const int ne = ws.d_ne[r];
			for (int x = 0; x < ne; ++x) {
				const uint32_t pj = P[ws.d_pos[e[x].to]];
				if (pi < pj) { lb2_min32(&P[pj], pi); } else if (pj < pi) { lb2_min32(&P[pi], pj); }
			}
		}
		if (tid == 0) { sh->flag_b = 0; }
		lb2_sync();
		for (uint32_t i = tid; i < NR; i += nt) { uint32_t p = lb2_ld32(&P[i]); while (true) { uint32_t q = lb2_ld32(&P[p]); if (q == p) { break; } p = q; } P[i] = p; }
		lb2_sync();
		for (uint32_t r = tid; r < NR; r += nt) {
			const uint32_t pi = P[ws.d_pos[r]]; bool diff = false;
			const lb2_edge *e = lb2_edges(ws, r); const int ne = ws.d_ne[r];
			for (int x = 0; x < ne; ++x) { if (P[ws.d_pos[e[x].to]] != pi) { diff = true; } }
			if (diff) { sh->flag_b = 1; }
		}
		lb2_sync();
		if (!sh->flag_b) { break; }
		lb2_sync();
	}
	uint32_t ncomp = lb2_excl_scan(W, NR, [&](uint32_t i) -> uint32_t { return P[i] == i ? 1u : 0u; }, [&](uint32_t i, uint32_t v) { num[i] = v; });
	for (uint32_t r = tid; r < NR; r += nt) { ws.d_comp[r] = (int16_t)(num[P[ws.d_pos[r]]] + 1u); }
	lb2_sync();
	return (int)ncomp;
}

// special node creation: key string "source<c>" / "sink<c>" hashed like any other map key
LB2_DEV uint32_t lb2_new_special(lb2_win &W, bool source, int compid) {
	lb2_ws &ws = W.ws; lb2_sh *sh = W.sh;
	if (sh->n_spec >= sh->spec_cap) { sh->err |= 1u << LB2_D_SPECIAL; return LB2_NIL; }
	uint32_t id = sh->n_rows + sh->n_spec++;
	char buf[24]; int n = 0;
	const char *pre = source ? "source" : "sink";
	while (*pre) { buf[n++] = *pre++; }
	char dig[12]; int nd = 0; int c = compid; do { dig[nd++] = (char)('0' + c % 10); c /= 10; } while (c);
	while (nd) { buf[n++] = dig[--nd]; }
	ws.d_hash[id] = lb2_stdhash_bytes(buf, (uint32_t)n);
	ws.d_flags[id] = source ? LB2_NF_SOURCE : LB2_NF_SINK;
	ws.d_comp[id] = (int16_t)compid; ws.d_ne[id] = 0; ws.d_eov[id] = 0; ws.d_len[id] = 0; ws.d_str[id] = LB2_NIL; ws.d_cd[id] = LB2_NIL; ws.d_orig[id] = 0; ws.d_rep[id] = 0;
	for (int k = 0; k < 4; ++k) { ws.d_cov[id * 4 + k] = 0; ws.d_cnt[id * 4 + k] = 0; }
	ws.d_stn[id] = 0; ws.d_stT[id] = 0; ws.d_color[id] = 0; ws.d_mincov[id] = 0; ws.d_mincovqv[id] = 0;
	return id;
}

// markRefEnds src/Graph.cc:2028-2228.  The reference looks every window k-mer up in the map; here the
// dense node of the reference k-mer at each offset was recorded at build time (ws.refnode).
// all lanes: first / last reference offset whose node qualifies as an anchor, and the "same node matched twice" test
LB2_DEVNI void lb2_find_anchors(lb2_win &W, int compid) {
	lb2_ws &ws = W.ws; lb2_sh *sh = W.sh; const int K = sh->K; const int L = (int)sh->L;
	const unsigned tid = lb2_tid(), nt = lb2_nthr();
	const float thr = (float)W.P->cov_threshold;
	if (tid == 0) { sh->anc_src = LB2_NIL; sh->anc_snk = 0; sh->anc_amb = 0; }
	lb2_sync();
	for (int off = (int)tid; off + K <= L; off += (int)nt) {
		uint32_t nd = ws.refnode[off];
		if (nd == LB2_NIL || (ws.d_flags[nd] & LB2_NF_GONE)) { continue; }
		if (lb2_totcov(W, nd) >= thr && ws.d_comp[nd] == compid) { lb2_min32(&sh->anc_src, (uint32_t)off); lb2_max32(&sh->anc_snk, (uint32_t)off + 1u); }
	}
	lb2_sync();
	if (sh->anc_src != LB2_NIL) {
		const int so = (int)sh->anc_src, ko = (int)sh->anc_snk - 1;
		const uint32_t sn = ws.refnode[so], kn = ws.refnode[ko];
		uint32_t amb = 0;
		for (int off = (int)tid; off + K <= L; off += (int)nt) {
			uint32_t nd = ws.refnode[off];
			if (off > so && nd == sn) { amb |= 1u; }
			if (off < ko && nd == kn) { amb |= 2u; }
		}
		if (amb) { lb2_or32(&sh->anc_amb, amb); }
	}
	lb2_sync();
}

LB2_DEVNI void lb2_mark_ref_ends(lb2_win &W, int compid) {   // lane 0, after lb2_find_anchors
	lb2_ws &ws = W.ws; lb2_sh *sh = W.sh; const int K = sh->K; const int L = (int)sh->L;
	sh->trim5 = 0xFFFF; sh->trim3 = 0xFFFF; sh->source = LB2_NIL; sh->sink = LB2_NIL;
	if (sh->anc_src == LB2_NIL || sh->anc_amb) { return; }     // no match / ambiguous match (source first, then sink)
	const int src_off = (int)sh->anc_src, snk_off = (int)sh->anc_snk - 1;
	const uint32_t src = ws.refnode[src_off], snk = ws.refnode[snk_off];
	int ref_dist = snk_off - src_off + K;
	// std::string::substr(pos, n) clamps n; a negative n (sink left of source) becomes npos => to the end
	sh->seq_off = (uint32_t)src_off;
	sh->seq_len = (ref_dist < 0) ? (uint32_t)(L - src_off) : (uint32_t)((src_off + ref_dist > L) ? (L - src_off) : ref_dist);
	sh->trim5 = (uint32_t)src_off & 0xFFFF; sh->trim3 = (uint32_t)(L - snk_off - K) & 0xFFFF;
	// orientation of the anchor k-mers in the reference: F iff the reference spelling is the canonical one
	auto ref_ori = [&](int off) -> int {      // CanonicalMer_t::set on the reference spelling (src/Mer.hh:57-71)
		const char *m = W.ref_raw + off;
		for (int i = 0; i < K; ++i) { char a = m[i], b = lb2_comp(m[K - 1 - i]); if (a != b) { return a < b ? 0 : 1; } }
		return 1;
	};
	int sori = ref_ori(src_off), kori = ref_ori(snk_off);
	uint32_t ns = lb2_new_special(W, true, compid); if (ns == LB2_NIL) { return; }
	int sourcedir = sori ? LB2_FR : LB2_FF;
	{
		lb2_edge *e = lb2_edges(ws, src);
		for (int i = (int)ws.d_ne[src] - 1; i >= 0; --i) {
			if (lb2_dir_start(e[i].dir) == (sori ^ 1)) {
				uint32_t other = e[i].to;
				if (other != src) {
					lb2_remove_edge(W, other, src, lb2_fliplink(e[i].dir));
					int ne = ws.d_ne[src]; for (int j = i; j + 1 < ne; ++j) { e[j] = e[j + 1]; } ws.d_ne[src] = (uint8_t)(ne - 1);
				}
			}
		}
	}
	lb2_add_edge_node(W, ns, src, sourcedir);
	lb2_add_edge_node(W, src, ns, lb2_fliplink(sourcedir));
	lb2_oe_insert(W, ns);
	uint32_t nk = lb2_new_special(W, false, compid); if (nk == LB2_NIL) { return; }
	int sinkdir = kori ? LB2_FF : LB2_RR;
	{
		lb2_edge *e = lb2_edges(ws, snk);
		for (int i = (int)ws.d_ne[snk] - 1; i >= 0; --i) {
			if (lb2_dir_start(e[i].dir) == kori) {
				uint32_t other = e[i].to;
				if (other != snk) {
					lb2_remove_edge(W, other, snk, lb2_fliplink(e[i].dir));
					int ne = ws.d_ne[snk]; for (int j = i; j + 1 < ne; ++j) { e[j] = e[j + 1]; } ws.d_ne[snk] = (uint8_t)(ne - 1);
				}
			}
		}
	}
	lb2_add_edge_node(W, nk, snk, sinkdir);
	lb2_add_edge_node(W, snk, nk, lb2_fliplink(sinkdir));
	lb2_oe_insert(W, nk);
	sh->source = ns; sh->sink = nk;
}

// hasCycle / hasCycleRec with an explicit stack; frame = node << 16 | incoming orientation << 15 | next edge index
LB2_DEV bool lb2_cycle_from(lb2_win &W, uint32_t start, int ori) {
	lb2_ws &ws = W.ws; uint32_t *const st = ws.stack; uint32_t sp = 0; bool ans = false;
	const uint32_t cap = W.sh->n_rows + W.sh->spec_cap;
	// (array bases in registers: the descriptor lives in shared memory and would be re-read behind every store)
	uint8_t *const COL = ws.d_color; const uint8_t *const NE = ws.d_ne, *const FL = ws.d_flags, *const EOV = ws.d_eov; const lb2_edge *const ED = ws.d_edge, *const EP = ws.e_pool;
	COL[start] = 2; st[sp++] = (start << 16) | ((uint32_t)ori << 15);
	while (sp) {
		uint32_t v = st[sp - 1]; uint32_t node = v >> 16; int o = (int)((v >> 15) & 1); int i = (int)(v & 0x7FFF);
		if (ans || i >= (int)NE[node]) { COL[node] = 3; --sp; continue; }
		st[sp - 1] = v + 1;
		const uint32_t ov = EOV[node]; const lb2_edge ed = (ov ? (EP + (size_t)(ov - 1) * LB2_ECAP) : (ED + (size_t)node * LB2_EINL))[i];
		if (!lb2_is_dir(ed.dir, o)) { continue; }
		uint32_t other = ed.to;
		if (FL[other] & LB2_NF_SPECIAL) { continue; }
		if (COL[other] == 2) { ans = true; continue; }
		if (COL[other] == 1) {
			if (sp + 1 > cap) { W.sh->err |= 1u << LB2_D_STACK; return true; }
			COL[other] = 2; st[sp++] = (other << 16) | ((uint32_t)lb2_dir_dest(ed.dir) << 15);
		}
	}
	return ans;
}
LB2_DEVNI bool lb2_has_cycle(lb2_win &W, int compid) {
	lb2_ws &ws = W.ws; lb2_sh *sh = W.sh;
	if (sh->source == LB2_NIL || sh->sink == LB2_NIL) { return false; }
	// (the reference whitens every node of the map; the search from the source never leaves the source's component)
	lb2_each_node(W, compid, [&](uint32_t p) -> bool { ws.d_color[p] = 1; return true; });
	bool a1 = lb2_cycle_from(W, sh->source, 0);
	bool a2 = lb2_cycle_from(W, sh->source, 1);
	return a1 || a2;
}

// ---- compressNode (src/Graph.cc:2486-2706) for one direction.  The reference rewrites the seed's edge vector after
//      every absorbed buddy; only the last rewrite is observable (each intermediate edge is the unique edge erased by
//      the next step, and the intermediate updateEdge targets are absorbed next), so the walk keeps the "current edge"
//      in registers and performs the literal edge surgery once, for the last buddy.  Float averages, status counts and
//      dead flags are applied per buddy in the reference's order.
// chain entry: node id | flip<<31 (flip: buddy is reverse-complemented in the seed's frame)
LB2_DEV uint32_t lb2_compress_dir(lb2_win &W, uint32_t node, int dir, uint32_t *chain, uint32_t nchain, uint32_t &curlen) {
	lb2_ws &ws = W.ws; const int K = W.sh->K;
	uint8_t *const NE = ws.d_ne; uint8_t *const FL = ws.d_flags;
	float *const COV = ws.d_cov; uint16_t *const LEN = ws.d_len; uint16_t *const STN = ws.d_stn; uint16_t *const STT = ws.d_stT;
	int uid = lb2_get_buddy(W, node, dir);
	if (uid == -1) { return nchain; }
	if (lb2_is_tandem(W, node)) { return nchain; }
	uint32_t cur_to = lb2_edges(ws, node)[uid].to; int cur_dir = lb2_edges(ws, node)[uid].dir;
	uint32_t last = LB2_NIL; int last_buid = -1, last_edir = 0;
	float c0 = COV[node * 4 + 0], c1 = COV[node * 4 + 1], c2 = COV[node * 4 + 2], c3 = COV[node * 4 + 3];
	uint32_t stn = STN[node], stt = STT[node];
	while (true) {
		const int edir = cur_dir; const uint32_t buddy = cur_to;
		const int bdir = (edir == LB2_FF || edir == LB2_RF) ? 1 : 0;
		if (FL[buddy] & LB2_NF_SPECIAL) { break; }                    // getBuddy of a special node is -1
		const lb2_edge *be = lb2_edges(ws, buddy); const int bne = NE[buddy];
		int buid = -1, nb = 0; bool tandem = false;
		for (int i = 0; i < bne; ++i) {
			lb2_edge e = be[i];
			if (e.to == buddy) { tandem = true; }
			if (lb2_is_dir(e.dir, bdir)) { ++nb; buid = i; }
		}
		if (tandem || nb != 1) { break; }                              // buddy->isTandem(), buddy->getBuddy(bdir) == -1
		// absorb the buddy
		const int dest_r = lb2_dir_dest(edir);
		const uint32_t flip = (dir == 0) ? (uint32_t)dest_r : (uint32_t)(dest_r ^ 1);
		chain[nchain++] = buddy | (flip << 31);
		const int amerlen = (int)curlen - K + 1, bmerlen = (int)LEN[buddy] - K + 1;
		// same expression order as src/Graph.cc:2631-2636
		c0 = lb2_wavg(c0, amerlen, COV[buddy * 4 + 0], bmerlen); c1 = lb2_wavg(c1, amerlen, COV[buddy * 4 + 1], bmerlen);
		c2 = lb2_wavg(c2, amerlen, COV[buddy * 4 + 2], bmerlen); c3 = lb2_wavg(c3, amerlen, COV[buddy * 4 + 3], bmerlen);
		curlen += (uint32_t)bmerlen; stn += STN[buddy]; stt += STT[buddy];
		FL[buddy] |= LB2_NF_DEAD;
		last = buddy; last_buid = buid; last_edir = edir;
		if (bne - 1 != 1) { break; }                                   // the seed would have 0 or >= 2 edges in `dir`
		const lb2_edge f = be[buid == 0 ? 1 : 0];
		int nd = f.dir; if (edir == LB2_FR || edir == LB2_RF) { nd = lb2_flipme(nd); }
		const uint32_t nto = (f.to == buddy) ? node : f.to;
		if (nto == node) { break; }                                    // self edge: getBuddy -> -1
		cur_to = nto; cur_dir = nd;
	}
	if (last == LB2_NIL) { return nchain; }
	COV[node * 4 + 0] = c0; COV[node * 4 + 1] = c1; COV[node * 4 + 2] = c2; COV[node * 4 + 3] = c3; STN[node] = stn; STT[node] = stt;
	// the literal edge surgery of the last step: erase the seed's edge in `dir`, move over the last buddy's other edges
	{ lb2_edge *ne_ = lb2_edges(ws, node); int ne = NE[node]; for (int j = uid; j + 1 < ne; ++j) { ne_[j] = ne_[j + 1]; } NE[node] = (uint8_t)(ne - 1); }
	const lb2_edge *be = lb2_edges(ws, last); const int bne = NE[last];
	for (int i = 0; i < bne; ++i) {
		if (i == last_buid) { continue; }
		int nd = be[i].dir; if (last_edir == LB2_FR || last_edir == LB2_RF) { nd = lb2_flipme(nd); }
		uint32_t other = be[i].to;
		if (other == last) { lb2_push_edge(W, node, node, nd, be[i].flag); }   // "circle to buddy"
		else {
			lb2_push_edge(W, node, other, nd, be[i].flag);
			lb2_update_edge(W, other, last, lb2_fliplink(be[i].dir), node, lb2_fliplink(nd));
		}
	}
	return nchain;
}

// job record of one compacted unitig (lane 0 writes it during the sweep, all lanes lay the bases out)
struct lb2_job { uint32_t node, cbeg, nF, nAll, len0, curlen, so, co, leftlen, pad; };

LB2_DEVNI void lb2_compress_sweep(lb2_win &W, int compid) {
	lb2_ws &ws = W.ws; lb2_sh *sh = W.sh; const int K = sh->K;
	uint32_t cused = 0, njobs = 0; const uint32_t ccap = sh->n_rows + sh->spec_cap;
	lb2_job *jobs = (lb2_job *)ws.jobs;
	lb2_each_node(W, compid, [&](uint32_t p) -> bool {
		if (ws.d_flags[p] & LB2_NF_DEAD) { return true; }
		uint32_t *chain = ws.chain + cused;
		uint32_t len0 = ws.d_len[p], curlen = len0;
		uint32_t nF = lb2_compress_dir(W, p, 0, chain, 0, curlen);
		uint32_t nAll = lb2_compress_dir(W, p, 1, chain, nF, curlen);
		if (sh->err) { return false; }
		if (nAll == 0) { return true; }
		if (cused + nAll > ccap || njobs >= LB2_MAX_ROWS) { sh->err |= 1u << LB2_D_STACK; return false; }
		lb2_job jb; jb.node = p; jb.cbeg = cused; jb.nF = nF; jb.nAll = nAll; jb.len0 = len0; jb.curlen = curlen; jb.pad = 0;
		jb.so = lb2_arena_alloc(W, curlen); jb.co = lb2_arena_alloc(W, curlen * 2 * (uint32_t)sizeof(lb2_cov));
		if (sh->err) { return false; }
		// destination offsets: [R-chain, last absorbed first] seed [F-chain]; chain entries get their start position
		uint32_t leftlen = 0;
		for (uint32_t c = nF; c < nAll; ++c) { leftlen += ws.d_len[chain[c] & 0x7FFFFFFFu] - K + 1; }
		jb.leftlen = leftlen;
		// (cpos: k-mers absorbed before the entry, in chain order -- F entries then R entries; lb2_materialize turns that
		// into positions: F entry at leftlen + len0 + cpos, R entry ending at leftlen - (cpos - cpos[first R entry]))
		uint32_t pos = 0;
		for (uint32_t c = 0; c < nAll; ++c) { ws.cpos[cused + c] = pos; pos += ws.d_len[chain[c] & 0x7FFFFFFFu] - K + 1; }
		ws.d_mincov[p] = 10000000; ws.d_mincovqv[p] = 10000000;
		jobs[njobs++] = jb; cused += nAll;
		return true;
	});
	sh->n_jobs = njobs;
}

// lay out bases and per-base coverage of every compacted unitig (all lanes); Node_t::computeMinCov on the fly
LB2_DEVNI void lb2_materialize(lb2_win &W) {
	lb2_ws &ws = W.ws; lb2_sh *sh = W.sh; const int K = sh->K; const unsigned tid = lb2_tid(), nt = lb2_nthr();
	const lb2_job *jobs = (const lb2_job *)ws.jobs;
	const uint32_t njobs = sh->n_jobs;
	if (!njobs) { return; }
	// one member's contribution: `count` bases from position `first` of node `id` (reverse-complemented when flip) to
	// position dst of its job's unitig; lanes `l0, l0 + lstep, ...` of the member take part
	auto put = [&](const lb2_job &jb, uint32_t m, uint32_t l0, uint32_t lstep, int &mn, int &mnq) {
		char *S = (char *)ws.arena + jb.so; lb2_cov *CT = (lb2_cov *)(ws.arena + jb.co); lb2_cov *CN = CT + jb.curlen;
		uint32_t id, first, count, dst; bool flip;
		if (m == jb.nAll) { id = jb.node; flip = false; first = 0; count = jb.len0; dst = jb.leftlen; }
		else {
			uint32_t ce = ws.chain[jb.cbeg + m]; id = ce & 0x7FFFFFFFu; flip = (ce >> 31) != 0;
			uint32_t bl = ws.d_len[id]; count = bl - K + 1; const uint32_t pre = ws.cpos[jb.cbeg + m] - ws.cpos[jb.cbeg];
			dst = (m < jb.nF) ? jb.leftlen + jb.len0 + pre : jb.leftlen - (pre - (ws.cpos[jb.cbeg + jb.nF] - ws.cpos[jb.cbeg]) + count);
			first = (m < jb.nF) ? (uint32_t)K - 1 : 0;       // F: oriented[K-1..], R: frame[0..bl-K]
		}
		lb2_nview v; lb2_view(W, id, v);
		for (uint32_t i = l0; i < count; i += lstep) {
			uint32_t o = first + i, src = flip ? (v.len - 1 - o) : o;
			char ch = lb2_vchar(W, v, src); S[dst + i] = flip ? lb2_comp(ch) : ch;
			lb2_cov ct = lb2_vcov(W, v, src, 0), cn = lb2_vcov(W, v, src, 1);
			CT[dst + i] = ct; CN[dst + i] = cn;
			int t = ct.fwd + ct.rev + cn.fwd + cn.rev, qq = ct.mqf + ct.mqr + cn.mqf + cn.mqr;
			if (t < mn) { mn = t; } if (qq < mnq) { mnq = qq; }
		}
	};
	const uint32_t totc = jobs[njobs - 1].cbeg + jobs[njobs - 1].nAll;      // chain slots of all jobs (jobs are in slot order)
	if (totc + njobs >= 64) {
		// many members (first compaction: hundreds of single k-mer nodes in a dozen unitigs): one lane per member over ALL
		// jobs at once; the member's job is found by bisection on the jobs' first slots
		for (uint32_t it = tid; it < totc + njobs; it += nt) {
			uint32_t q, m;
			if (it < totc) {
				uint32_t lo = 0, hi = njobs;
				while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (jobs[mid].cbeg <= it) { lo = mid; } else { hi = mid; } }
				q = lo; m = it - jobs[q].cbeg;
			} else { q = it - totc; m = jobs[q].nAll; }
			const lb2_job jb = jobs[q];
			if (m > jb.nAll) { continue; }      // (a seed without absorbed nodes owns no slots: cannot be hit, kept as a guard)
			int mn = 10000000, mnq = 10000000;
			put(jb, m, 0u, 1u, mn, mnq);
			if (mn != 10000000) { lb2g_min32((uint32_t *)&ws.d_mincov[jb.node], (uint32_t)mn); lb2g_min32((uint32_t *)&ws.d_mincovqv[jb.node], (uint32_t)mnq); }
		}
	} else {
		// a few long members (later compactions merge unitigs): job by job, member by member, lanes across the bases
		for (uint32_t q = 0; q < njobs; ++q) {
			const lb2_job jb = jobs[q];
			int mn = 10000000, mnq = 10000000;
			for (uint32_t m = 0; m <= jb.nAll; ++m) { put(jb, m, tid, nt, mn, mnq); }
			if (mn != 10000000) { lb2g_min32((uint32_t *)&ws.d_mincov[jb.node], (uint32_t)mn); lb2g_min32((uint32_t *)&ws.d_mincovqv[jb.node], (uint32_t)mnq); }
		}
	}
}

// Graph_t::cleanDead by all lanes: every live node finds its next live node (unordered_map::erase keeps the relative
// order of the survivors), then the bucket array is rebuilt from its invariant (bucket -> node before the bucket's
// first node).  Dead nodes are only read, live nodes only written by their owner lane.
LB2_DEVNI void lb2_clean_dead_par(lb2_win &W)
{
	lb2_ws &ws = W.ws; lb2_sh *sh = W.sh; const unsigned tid = lb2_tid(), nt = lb2_nthr();
	const uint32_t NT = sh->n_rows + sh->n_spec;
	if (tid == 0) { sh->n_dead = 0; }
	lb2_sync();
	// runs of dead nodes are shortcut first (pointer jumping on the dead nodes' own links, in place: a dead node's link is
	// never followed again once the node is erased; a stale read is still a valid, shorter shortcut)
	while (true) {
		if (tid == 0) { sh->flag_d = 0; }
		lb2_sync();
		bool ch = false;
		for (uint32_t r = tid; r < NT; r += nt) {
			if ((ws.d_flags[r] & (LB2_NF_DEAD | LB2_NF_GONE)) != LB2_NF_DEAD) { continue; }
			const uint32_t x = ws.d_lnext[r];
			if (x != LB2_NIL && (ws.d_flags[x] & LB2_NF_DEAD)) { ws.d_lnext[r] = ws.d_lnext[x]; ch = true; }
		}
		if (ch) { sh->flag_d = 1; }
		lb2_sync();
		if (!sh->flag_d) { break; }
		lb2_sync();
	}
	uint32_t nd = 0;
	for (uint32_t r = tid; r < NT; r += nt) {
		const uint8_t f = ws.d_flags[r];
		if (f & LB2_NF_GONE) { continue; }
		if (f & LB2_NF_DEAD) { ++nd; continue; }
		uint32_t x = ws.d_lnext[r];
		while (x != LB2_NIL && (ws.d_flags[x] & LB2_NF_DEAD)) { x = ws.d_lnext[x]; }
		ws.d_lnext[r] = x;
	}
	if (nd) { lb2_add32(&sh->n_dead, nd); }
	if (tid == 0) { uint32_t x = sh->lhead; while (x != LB2_NIL && (ws.d_flags[x] & LB2_NF_DEAD)) { x = ws.d_lnext[x]; } sh->lhead = x; }
	lb2_sync();
	if (!sh->n_dead) { return; }
	for (uint32_t b = tid; b < sh->bkt_count; b += nt) { lb2_bset(W, b, LB2_NIL); }
	lb2_sync();
	for (uint32_t r = tid; r < NT; r += nt) {
		const uint8_t f = ws.d_flags[r];
		if (f & LB2_NF_GONE) { continue; }
		if (f & LB2_NF_DEAD) { ws.d_flags[r] = f | LB2_NF_GONE; continue; }
		const uint32_t nb = ws.d_lnext[r];
		if (nb != LB2_NIL && ws.d_bk[nb] != ws.d_bk[r]) { lb2_bset(W, ws.d_bk[nb], r); }
	}
	if (tid == 0) { if (sh->lhead != LB2_NIL) { lb2_bset(W, ws.d_bk[sh->lhead], LB2_SENT); } sh->elem_count -= sh->n_dead; }
	lb2_sync();
}

// Graph_t::compress (all lanes)
LB2_DEVNI void lb2_compress(lb2_win &W, int compid) {
	lb2_ws &ws = W.ws; lb2_sh *sh = W.sh;
	lb2_mark(W, LB2_PH_COMP_SEQ);
	if (lb2_tid() == 0) { lb2_compress_sweep(W, compid); }
	lb2_mark(W, LB2_PH_CSWEEP);
	lb2_sync();
	if (!sh->err) { lb2_materialize(W); }
	lb2_sync();
	lb2_mark(W, LB2_PH_CMAT);
	if (lb2_tid() == 0 && !sh->err) {
		const lb2_job *jobs = (const lb2_job *)ws.jobs;
		for (uint32_t q = 0; q < sh->n_jobs; ++q) { const lb2_job &jb = jobs[q]; ws.d_str[jb.node] = jb.so; ws.d_cd[jb.node] = jb.co; ws.d_len[jb.node] = jb.curlen; }
	}
	lb2_sync();
	if (!sh->err) { lb2_clean_dead_par(W); }
	lb2_mark(W, LB2_PH_CCLEAN);
	lb2_sync();
}

// ---------------------------------------------------------------------------------------------------
// The FIRST compaction of a component (hundreds of single-k-mer nodes -> a handful of unitigs) by all lanes.
// compressNode's stop conditions (src/Graph.cc:2486-2706; Node_t::getBuddy/isTandem src/Node.cc:235-266) only look
// at the two nodes of a step, and absorbing a node never changes them for any other pair, so "x absorbs y through
// its side o" is a static, symmetric relation (a link): x, y not special, not tandem, x has exactly one edge leaving
// in orientation o (to y != x) and y exactly one edge leaving towards x.  The maximal chains of links are the
// unitigs; a chain is swallowed whole by its first node in map iteration order (compressNode(F) then (R) run to the
// chain's two ends).  So: successor of every oriented node (x,o), pointer jumping to the chain end (distance + end
// state), seed = smallest list index of the chain, and every member's slot in the seed's absorb order
// (F-side in order, then R-side).  The float coverage averages are still folded in that order, one lane per chain;
// the edge surgery is evaluated from the untouched edge arrays into a side buffer (seed keeps its non-link edges, then
// the F-end's, then the R-end's outward edges; every target is renamed to its chain's seed, orientation bits flipped
// for reverse-complemented members) and written back after a barrier.  A closed ring of links (no chain end) makes
// the function return false before anything is modified; the caller then runs the sequential sweep.
// ---------------------------------------------------------------------------------------------------
LB2_DEV uint32_t lb2_pc_succ(lb2_ws &ws, uint32_t r, int o) {       // oriented successor state or LB2_NIL
	if (!ws.d_color[r]) { return LB2_NIL; }
	const lb2_edge *e = lb2_edges(ws, r); const int ne = ws.d_ne[r]; int cnt = 0; lb2_edge pick; pick.to = 0; pick.dir = 0;
	for (int i = 0; i < ne; ++i) { if ((int)(e[i].dir >> 1) == o) { ++cnt; pick = e[i]; } }
	if (cnt != 1) { return LB2_NIL; }
	const uint32_t v = pick.to; if (!ws.d_color[v]) { return LB2_NIL; }
	const int ov = (int)(pick.dir & 1), back = 1 - ov;
	const lb2_edge *f = lb2_edges(ws, v); const int nf = ws.d_ne[v]; int cb = 0;
	for (int i = 0; i < nf; ++i) { if ((int)(f[i].dir >> 1) == back) { ++cb; } }
	if (cb != 1) { return LB2_NIL; }
	return 2u * v + (uint32_t)ov;
}
#define LB2_PI_MEMBER 0x20000u
#define LB2_PI_SEED   0x40000u
LB2_DEV lb2_edge lb2_pc_map(const uint32_t *INF, lb2_edge e) {       // rename the target to its chain's seed
	const uint32_t iy = INF[e.to];
	if (iy & LB2_PI_MEMBER) { e.to = (uint16_t)(iy & 0xFFFFu); e.dir = (uint16_t)(e.dir ^ ((iy >> 16) & 1u)); }
	return e;
}

#ifdef LB2_HOSTSIM
static unsigned long lb2_dbg_par[4];     // debug build only: [0] taken, [1] no scratch, [2] ring
#define LB2_DBG(i) (++lb2_dbg_par[i])
#else
#define LB2_DBG(i) ((void)0)
#endif
LB2_DEVNI bool lb2_compress_par(lb2_win &W, int compid)
{
	lb2_ws &ws = W.ws; lb2_sh *sh = W.sh; const int K = sh->K; const unsigned tid = lb2_tid(), nt = lb2_nthr();
	const uint32_t NT = sh->n_rows + sh->n_spec;
#ifdef LB2_HOSTSIM
	if (getenv("LB2_SIM_NOPAR")) { LB2_DBG(1); return false; }
#endif
	if ((size_t)NT * 4 + 16 > ws.px_words || (W.C->debug_flags & 1u)) { LB2_DBG(1); return false; }
	// (the reciprocals are written once and read once per chain slot: they live in the slab's scratch, not in shared memory)
	uint32_t *J = ws.px, *SEED = J + 2 * NT, *INF = SEED + NT; float *RCP = (float *)ws.emu;
	lb2_edge *etmp = ws.etmp; uint8_t *etn = (uint8_t *)(ws.etmp + (size_t)LB2_MAX_ROWS * LB2_ECAP);
	lb2_job *jobs = (lb2_job *)ws.jobs;
	// eligibility (d_color is idle between the cycle checks)
	for (uint32_t r = tid; r < NT; r += nt) {
		uint8_t ok = (ws.d_comp[r] == compid && !(ws.d_flags[r] & (LB2_NF_DEAD | LB2_NF_GONE | LB2_NF_SPECIAL))) ? 1 : 0;
		if (ok) { const lb2_edge *e = lb2_edges(ws, r); const int ne = ws.d_ne[r]; for (int i = 0; i < ne; ++i) { if (e[i].to == r) { ok = 0; } } }
		ws.d_color[r] = ok; SEED[r] = LB2_NIL; INF[r] = 0;
	}
	if (tid == 0) { sh->flag_b = 0; }
	lb2_sync();
	for (uint32_t s = tid; s < 2 * NT; s += nt) {
		const uint32_t nx = lb2_pc_succ(ws, s >> 1, (int)(s & 1));
		J[s] = (nx == LB2_NIL) ? (s << 16) : ((nx << 16) | 1u);
	}
	lb2_sync();
	uint32_t rounds = 2; while ((1u << (rounds - 2)) < 2 * NT) { ++rounds; }
	for (uint32_t it = 0; it < rounds; ++it) {
		for (uint32_t s = tid; s < 2 * NT; s += nt) {
			const uint32_t w = J[s], j = w >> 16;
			if (j == s) { continue; }
			const uint32_t w2 = lb2_lds(&J[j]), j2 = w2 >> 16;
			if (j2 == j) { continue; }                      // j is a chain end
			uint32_t d = (w & 0xFFFFu) + (w2 & 0xFFFFu); if (d > 0xFFFFu) { d = 0xFFFFu; }
			J[s] = (j2 << 16) | d;
		}
		lb2_sync();
	}
	for (uint32_t s = tid; s < 2 * NT; s += nt) { const uint32_t j = J[s] >> 16; if ((J[j] >> 16) != j) { sh->flag_b = 1; } }   // never reached an end: ring
	lb2_sync();
	if (sh->flag_b) { lb2_sync(); LB2_DBG(2); return false; }
	LB2_DBG(0);
	// seed of every chain: smallest list index; key of a chain = smaller node id of its two end states
	for (uint32_t r = tid; r < NT; r += nt) {
		if (!ws.d_color[r]) { continue; }
		const uint32_t a = J[2 * r] >> 17, b = J[2 * r + 1] >> 17;
		lb2_min32(&SEED[a < b ? a : b], ((uint32_t)ws.d_pos[r] << 16) | r);
	}
	lb2_sync();
	for (uint32_t r = tid; r < NT; r += nt) {
		if (!ws.d_color[r]) { continue; }
		const uint32_t a = J[2 * r] >> 17, b = J[2 * r + 1] >> 17;
		const uint32_t sd = SEED[a < b ? a : b] & 0xFFFFu;
		if (sd == r) { INF[r] = LB2_PI_SEED | r; }
		else { const uint32_t ostar = ((J[2 * r] >> 16) == (J[2 * sd] >> 16)) ? 0u : 1u; INF[r] = LB2_PI_MEMBER | (ostar << 16) | sd; }
	}
	lb2_sync();
	// jobs (seeds that absorb something) and their chain segments
	const uint32_t tot = lb2_excl_scan(W, NT,
		[&](uint32_t r) -> uint32_t { if (!(INF[r] & LB2_PI_SEED)) { return 0u; } uint32_t na = (J[2 * r] & 0xFFFFu) + (J[2 * r + 1] & 0xFFFFu); return na ? ((1u << 16) | na) : 0u; },
		[&](uint32_t r, uint32_t v) { if (INF[r] & LB2_PI_SEED) { SEED[r] = v; } });
	const uint32_t njobs = tot >> 16;
	for (uint32_t r = tid; r < NT; r += nt) {
		const uint32_t ir = INF[r];
		if (ir & LB2_PI_SEED) {
			const uint32_t nF = J[2 * r] & 0xFFFFu, nR = J[2 * r + 1] & 0xFFFFu;
			if (nF + nR) { lb2_job jb; jb.node = r; jb.cbeg = SEED[r] & 0xFFFFu; jb.nF = nF; jb.nAll = nF + nR; jb.len0 = ws.d_len[r]; jb.curlen = 0; jb.so = 0; jb.co = 0; jb.leftlen = 0; jb.pad = 0; jobs[SEED[r] >> 16] = jb; }
		} else if (ir & LB2_PI_MEMBER) {
			const uint32_t sd = ir & 0xFFFFu, ostar = (ir >> 16) & 1u, cbeg = SEED[sd] & 0xFFFFu;
			const uint32_t dFs = J[2 * sd] & 0xFFFFu, dRs = J[2 * sd + 1] & 0xFFFFu;
			const uint32_t dx = J[2 * r + ostar] & 0xFFFFu;
			uint32_t slot;
			if (dx < dFs) { slot = cbeg + (dFs - dx - 1); }
			else { const uint32_t dy = J[2 * r + (1u - ostar)] & 0xFFFFu; slot = cbeg + dFs + (dRs - dy - 1); }
			ws.chain[slot] = r | (ostar << 31);
		}
	}
	lb2_sync();
	// fold the float coverages in the reference's absorb order (src/Graph.cc:2631-2636): a serial recurrence per chain and
	// channel (every step rounds), so one lane per (chain, channel); the integer bookkeeping rides along in every lane
	// cpos[slot] = k-mers absorbed before the slot (prefix over ALL chain slots; a job's own prefix is the difference to
	// its first slot), one more entry for the total; RCP[slot] = correctly rounded reciprocal of the slot's divisor
	const uint32_t totc = tot & 0xFFFFu;
	{
		const uint32_t K1 = (uint32_t)K - 1u;
		const uint32_t tk_ = lb2_excl_scan(W, totc, [&](uint32_t x) -> uint32_t { return (uint32_t)ws.d_len[ws.chain[x] & 0x7FFFFFFFu] - K1; }, [&](uint32_t x, uint32_t v) { ws.cpos[x] = v; });
		if (tid == 0) { ws.cpos[totc] = tk_; }
		lb2_sync();
		for (uint32_t x = tid; x < totc; x += nt) {
			const uint32_t b = ws.chain[x] & 0x7FFFFFFFu, sd = INF[b] & 0xFFFFu, cbeg = SEED[sd] & 0xFFFFu;
			RCP[x] = lb2_rcp_int((uint32_t)ws.d_len[sd] - K1 + (ws.cpos[x + 1] - ws.cpos[cbeg]));
			lb2g_red_add(&jobs[SEED[sd] >> 16].pad, (uint32_t)ws.d_stn[b] | ((uint32_t)ws.d_stT[b] << 16));      // status counts of the absorbed nodes (sums < 2^16)
		}
		lb2_sync();
	}
	lb2_mark(W, LB2_PH_CP_LINK);
	{
		// fold the float coverages in the reference's absorb order (src/Graph.cc:2631-2636): a serial recurrence per chain
		// and channel (every step rounds), so one lane per (chain, channel).  Only product, sum and the division's three
		// fused corrections depend on the running value; everything else of a step is loaded / converted one step ahead.
		// (array bases in registers: the descriptor lives in shared memory and would be re-read after every store)
		const uint16_t *const LEN = ws.d_len; const float *const COV = ws.d_cov;
		const uint32_t *const CPOS = ws.cpos; const int K1 = K - 1;
		for (uint32_t q = tid / LB2_FQ; q < njobs; q += nt / LB2_FQ) {
			const lb2_job jb = jobs[q]; const uint32_t node = jb.node, nAll = jb.nAll, nF = jb.nF; const uint32_t *const chain = ws.chain + jb.cbeg;
			const uint32_t *const P = CPOS + jb.cbeg; const float *const RC = RCP + jb.cbeg;
			for (uint32_t ch = tid % LB2_FQ; ch < 4; ch += LB2_FQ) {
				float cv = COV[node * 4 + ch];
				int am = (int)jb.len0 - K1;
				uint32_t b1 = nAll ? (chain[0] & 0x7FFFFFFFu) : 0u;
				uint32_t len_c = LEN[b1]; float cov_c = COV[b1 * 4 + ch], r_c = RC[0];
				b1 = (nAll > 1) ? (chain[1] & 0x7FFFFFFFu) : 0u;
				for (uint32_t c = 0; c < nAll; ++c) {
					const uint32_t len_n = LEN[b1]; const float cov_n = COV[b1 * 4 + ch], r_n = RC[(c + 1 < nAll) ? c + 1 : c];
					const uint32_t b2 = (c + 2 < nAll) ? (chain[c + 2] & 0x7FFFFFFFu) : 0u;
					const int bm = (int)len_c - K1;
					cv = lb2_wavg_rcp(cv, am, cov_c, bm, r_c);
					am += bm;
					len_c = len_n; cov_c = cov_n; r_c = r_n; b1 = b2;
				}
				ws.d_cov[node * 4 + ch] = cv;
				if (ch == 0) {
					ws.d_stn[node] = (uint16_t)(ws.d_stn[node] + (jb.pad & 0xFFFFu)); ws.d_stT[node] = (uint16_t)(ws.d_stT[node] + (jb.pad >> 16));
					ws.d_mincov[node] = 10000000; ws.d_mincovqv[node] = 10000000;
					jobs[q].curlen = jb.len0 + (P[nAll] - P[0]); jobs[q].leftlen = P[nAll] - P[nF];
				}
			}
		}
	}
	lb2_sync();
	lb2_mark(W, LB2_PH_CP_FOLD);
	if (tid == 0) {
		for (uint32_t q = 0; q < njobs && !sh->err; ++q) { jobs[q].so = lb2_arena_alloc(W, jobs[q].curlen); jobs[q].co = lb2_arena_alloc(W, jobs[q].curlen * 2 * (uint32_t)sizeof(lb2_cov)); }
		sh->n_jobs = njobs;
	}
	// edge surgery, evaluated from the untouched edge arrays into the side buffer
	for (uint32_t z = tid; z < NT; z += nt) {
		if (ws.d_comp[z] != compid || (ws.d_flags[z] & (LB2_NF_DEAD | LB2_NF_GONE)) || (INF[z] & LB2_PI_MEMBER)) { continue; }
		lb2_edge *out = etmp + (size_t)z * LB2_ECAP; int no = 0; bool ovf = false;
		const bool seed = (INF[z] & LB2_PI_SEED) != 0;
		const uint32_t nF = seed ? (J[2 * z] & 0xFFFFu) : 0u, nR = seed ? (J[2 * z + 1] & 0xFFFFu) : 0u;
		const lb2_edge *e = lb2_edges(ws, z); const int ne = ws.d_ne[z];
		for (int i = 0; i < ne; ++i) {
			const int st = (int)(e[i].dir >> 1);
			if ((st == 0 && nF) || (st == 1 && nR)) { continue; }        // the link edge of that side
			if (no >= LB2_ECAP) { ovf = true; break; }
			out[no++] = lb2_pc_map(INF, e[i]);
		}
		for (int side = 0; side < 2 && !ovf; ++side) {
			const uint32_t cnt = side ? nR : nF; if (!cnt) { continue; }
			const uint32_t cbeg = SEED[z] & 0xFFFFu;
			const uint32_t ce = ws.chain[cbeg + (side ? nF + nR : nF) - 1], last = ce & 0x7FFFFFFFu, fl = ce >> 31;
			const int ot = side ? (int)(1u - fl) : (int)fl;                // orientation in which the end node is traversed
			const lb2_edge *f = lb2_edges(ws, last); const int nf = ws.d_ne[last];
			for (int i = 0; i < nf; ++i) {
				if ((int)(f[i].dir >> 1) != ot) { continue; }               // (the single edge of the other side leads back into the chain)
				if (no >= LB2_ECAP) { ovf = true; break; }
				lb2_edge x = f[i]; x.dir = (uint16_t)(x.dir ^ (fl << 1));
				out[no++] = lb2_pc_map(INF, x);
			}
		}
		if (ovf) { lb2_or32(&sh->err, 1u << LB2_D_EDGES); no = 0; }
		etn[z] = (uint8_t)no;
	}
	lb2_sync();
	if (sh->err) { return true; }
	for (uint32_t z = tid; z < NT; z += nt) {
		if (ws.d_comp[z] != compid || (ws.d_flags[z] & (LB2_NF_DEAD | LB2_NF_GONE))) { continue; }
		if (INF[z] & LB2_PI_MEMBER) { ws.d_flags[z] |= LB2_NF_DEAD; continue; }
		const int no = etn[z];
		if (no > LB2_EINL && !ws.d_eov[z]) {
			uint32_t blk = lb2_add32(&sh->n_eov, 1u);
			if (blk >= LB2_EOV_BLOCKS) { lb2_or32(&sh->err, 1u << LB2_D_EDGES); continue; }
			ws.d_eov[z] = (uint8_t)(blk + 1);
		}
		lb2_edge *dst = lb2_edges(ws, z); const lb2_edge *src = etmp + (size_t)z * LB2_ECAP;
		for (int i = 0; i < no; ++i) { dst[i] = src[i]; }
		ws.d_ne[z] = (uint8_t)no;
	}
	lb2_sync();
	if (sh->err) { if (tid == 0 && sh->n_eov > LB2_EOV_BLOCKS) { sh->n_eov = LB2_EOV_BLOCKS; } lb2_sync(); return true; }
	lb2_mark(W, LB2_PH_CSWEEP);
	lb2_materialize(W);
	lb2_sync();
	lb2_mark(W, LB2_PH_CMAT);
	for (uint32_t q = tid; q < njobs; q += nt) { const lb2_job &jb = jobs[q]; ws.d_str[jb.node] = jb.so; ws.d_cd[jb.node] = jb.co; ws.d_len[jb.node] = (uint16_t)jb.curlen; }
	lb2_sync();
	lb2_clean_dead_par(W);
	lb2_mark(W, LB2_PH_CCLEAN);
	return true;
}

LB2_DEVNI void lb2_remove_tips(lb2_win &W, int compid) {      // all lanes
	lb2_ws &ws = W.ws; lb2_sh *sh = W.sh; const int K = sh->K;
	while (true) {
		if (lb2_tid() == 0) {
			int tips = 0;
			lb2_each_node(W, compid, [&](uint32_t p) -> bool {
				int deg = ws.d_ne[p]; int len = (int)lb2_strlen(W, p) - K + 1;
				if (deg <= 1 && len < W.P->max_tip_len) { lb2_remove_node(W, p); ++tips; }
				return true;
			});
			sh->flag_b = (uint32_t)tips; sh->n_changed += (uint32_t)tips;
		}
		lb2_sync();
		if (!sh->flag_b || sh->err) { break; }
		LB2_SEQMARK(LB2_PH_LOWQ);
		lb2_compress(W, compid);
	}
}

// removeShortLinks (src/Graph.cc:3015-3062).  Whether a node's string holds a tandem repeat near its k-1'th base does not
// change while the sweep runs (nor do its length and minimum coverage; only degrees drop as neighbours go), so the
// findTandems calls -- the expensive part: every short low-coverage node of a large-k graph is a candidate -- are made
// up front by all lanes, one warp per candidate, and lane 0 then applies the sweep in map order with the degree test at
// its original place.
#define LB2_LINK_BUF 256      /* bytes per warp, >= 127 + 127 / 2; longer strings are read in place */
LB2_DEVNI void lb2_remove_short_links(lb2_win &W, int compid) {   // all lanes
	lb2_ws &ws = W.ws; lb2_sh *sh = W.sh; const int K = sh->K; const unsigned tid = lb2_tid(), nt = lb2_nthr();
	uint32_t *cand = ws.stack;      // (idle between the cycle checks and the compactions: rows + 8 words)
	if (tid == 0) {
		uint32_t n = 0;
		double avgcov = ((double)(int)sh->totalreadbp) / ((double)sh->L);
		const int max_link = (int)floor((double)K / 2.0);
		const double lim = floor(sqrt(avgcov));
		lb2_each_node(W, compid, [&](uint32_t p) -> bool {
			int deg = ws.d_ne[p]; int len = (int)ws.d_len[p] - K + 1;
			if (deg >= 2 && len < max_link && (double)ws.d_mincov[p] <= lim) { cand[n++] = p; }
			return true;
		});
		sh->n_jobs = n;
	}
	lb2_sync();
	// one warp per candidate, the string in the warp's piece of the idle compaction scratch, the positions dealt to the
	// lanes (a reported repeat always has a positive length: "LEN == 0" is "no position reports one")
	const uint32_t nc = sh->n_jobs, nwarp = (nt / LB2_WARP) ? nt / LB2_WARP : 1u, wid = tid / LB2_WARP, lane = lb2_lane();
	char *const wbuf = ((size_t)ws.px_words * 4 >= (size_t)nwarp * LB2_LINK_BUF) ? (char *)ws.px + (size_t)wid * LB2_LINK_BUF : nullptr;
	for (uint32_t i = wid; i < nc; i += nwarp) {
		const uint32_t p = cand[i];
		int LEN = 0; char motif[4]; uint32_t ml = 0; bool ov = false, found;
		lb2_nview v; lb2_view(W, p, v);
		if (wbuf && v.len <= LB2_LINK_BUF) {
			for (uint32_t x = lane; x < v.len; x += LB2_WARP) { wbuf[x] = lb2_vchar(W, v, x); }
			lb2_warp_sync();
			found = lb2_find_tandems([&](uint32_t x) -> char { return wbuf[x]; }, v.len, W.P, K - 1, LEN, motif, ml, 0, ov, lane, LB2_WARP);
		} else {
			found = lb2_find_tandems([&](uint32_t x) -> char { return lb2_vchar(W, v, x); }, v.len, W.P, K - 1, LEN, motif, ml, 0, ov, lane, LB2_WARP);
		}
		const uint32_t any = lb2_ballot(found);
		if (any && lane == 0) { cand[i] = p | 0x80000000u; }
		lb2_warp_sync();
	}
	lb2_sync();
	LB2_SEQMARK(LB2_PH_SCAN);
	if (tid == 0) {
		int links = 0;
		for (uint32_t i = 0; i < nc; ++i) {
			const uint32_t p = cand[i];
			if (p & 0x80000000u) { continue; }
			if (ws.d_ne[p] >= 2) { lb2_remove_node(W, p); ++links; }
		}
		sh->flag_b = (uint32_t)links; sh->n_changed += (uint32_t)links;
	}
	lb2_sync();
	if (sh->flag_b && !sh->err) { lb2_compress(W, compid); }
}

#endif
