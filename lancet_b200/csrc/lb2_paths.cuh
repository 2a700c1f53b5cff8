// lb2_paths.cuh -- source->sink path enumeration, contig/reference alignment, variant extraction.
//
//   Graph_t::bfs / eka                      src/Graph.cc:1299-1425, 1430-1501
//   Graph_t::findRepeatsInGraphPaths        src/Graph.cc:686-730
//   Path_t::str / covDistr / pathcontig     src/Path.cc:69-108, 110-180, 291-314
//   Graph_t::processPath                    src/Graph.cc:788-1220
//   global_align_aff                        src/align.cc:235-364   (full-matrix Gotoh, CTA-wide wavefront)
//   Transcript_t / computeStats             src/Transcript.hh:79-226
//   isRepeat / isAlmostRepeat / kMismatch   src/util.cc:295-360
#ifndef LB2_PATHS_CUH
#define LB2_PATHS_CUH

#include "lb2_graph.cuh"

// ---------------------------------------------------------------------------------------------------
// repeat predicates over a character string, CTA-wide, for ALL k at once (one lane per diagonal).
//   isRepeat(seq,K)             (src/util.cc:295-315)  <=>  K   <= sh->scan_emax
//   isAlmostRepeat(seq,K,max)   (src/util.cc:317-360)  <=>  K+1 <= sh->scan_wmax
// where, over the diagonals d = i - s0 >= 1 with mismatch bits m_d[p] = (seq[p] != seq[p+d]):
//   emax = longest run of zeros of m_d restricted to p <= len-2-d   (both offsets must be < len-K)
//   wmax = longest window of m_d (p <= len-1-d) holding <= max ones  (kMismatch compares K+1 columns)
// ---------------------------------------------------------------------------------------------------
// 16 mismatch flags (bit 2i = base p+i differs from base p+i+d) from the 2-bit packed sequence starting at base g0
LB2_DEV uint32_t lb2_mm16(const uint32_t *bits, uint32_t g0, uint32_t p, uint32_t d) {
	uint32_t ia = g0 + p, ib = ia + d;
	uint32_t a0 = lb2_lds(&bits[ia >> 4]), a1 = lb2_lds(&bits[(ia >> 4) + 1]), sa = (ia & 15) << 1;
	uint32_t b0 = lb2_lds(&bits[ib >> 4]), b1 = lb2_lds(&bits[(ib >> 4) + 1]), sb = (ib & 15) << 1;
	uint32_t A = sa ? ((a0 >> sa) | (a1 << (32 - sa))) : a0;
	uint32_t B = sb ? ((b0 >> sb) | (b1 << (32 - sb))) : b0;
	uint32_t x = A ^ B;
	return (x | (x >> 1)) & 0x55555555u;
}

// the sequence is ACGT only and 2-bit packed at base index g0 of `bits` (two readable words past the end).
// Exact rule per diagonal (lane-serial): walking the MISMATCH positions q (ffs over the 16 flags of one XOR) with the
// previous mismatches m1 > m2 > ..., the zero run ending at q-1 has length q-m1-1 and the longest window ending at q-1
// with <= max mismatches has length q - m_{max+1} - 1.
// Word filter: the results are only ever compared with k >= min_k (callers: lb2_pipeline.cuh), so runs and windows
// shorter than 11 are irrelevant when min_k >= 11.  A run/window of 11+ positions contains two adjacent aligned
// 4-position blocks holding <= max mismatches between them ("sparse pair"), and when it is evaluated at a mismatch q
// inside word i that pair starts in word i-1 or i.  So a word needs the exact rule only if it or its predecessor holds a
// sparse pair or one straddles their boundary (random sequence: ~2-3 % of the words).
// Pass 1 (one lane per diagonal, dealt boustrophedon because the lengths fall with d) only classifies the words -- a
// dozen instructions each -- and queues the START of every maximal stretch of exact-rule words.  Pass 2 (one lane per
// queued stretch) applies the exact rule along its stretch.  The word before a stretch was skipped, hence holds
// >= 2(max+1) mismatches: the mismatch history the stretch starts from is the top set bits of that word.
struct lb2_wflags { uint32_t fw; bool sp; uint32_t c0, c3; };
// 16 bits of a one-bit-per-position mask starting at position p, spread to the even bits of a word
LB2_DEV uint32_t lb2_nbits16(const uint32_t *nm, uint32_t p) {
	const uint32_t lo = nm[p >> 5], hi = nm[(p >> 5) + 1], sh = p & 31u;
	uint32_t x = (sh ? ((lo >> sh) | (hi << (32u - sh))) : lo) & 0xFFFFu;
	x = (x | (x << 8)) & 0x00FF00FFu; x = (x | (x << 4)) & 0x0F0F0F0Fu; x = (x | (x << 2)) & 0x33333333u; x = (x | (x << 1)) & 0x55555555u;
	return x;
}
// nm (may be NULL): positions holding an 'N' (the window reference only).  'N' equals 'N' and differs from every base,
// whatever 2-bit code the packed copy holds in its place.
LB2_DEV lb2_wflags lb2_scan_word(const uint32_t *bits, uint32_t g0, int p0, int d, int np, uint32_t bias, const uint32_t *nm) {
	lb2_wflags r; r.fw = lb2_mm16(bits, g0, (uint32_t)p0, (uint32_t)d);
	if (nm) { const uint32_t na = lb2_nbits16(nm, (uint32_t)p0), nb = lb2_nbits16(nm, (uint32_t)(p0 + d)); r.fw = (r.fw & ~(na & nb)) | (na ^ nb); }
	if (np - p0 < 16) { r.fw &= (1u << (2 * (np - p0))) - 1u; }
	const uint32_t t = (r.fw & 0x11111111u) + ((r.fw >> 2) & 0x11111111u);
	const uint32_t c = (t + (t >> 4)) & 0x0F0F0F0Fu;      // mismatches per 4-position block
	const uint32_t s2 = c + (c >> 8);                     // blocks (0,1) (1,2) (2,3) in bytes 0..2
	r.sp = ((~(s2 + bias)) & 0x808080u) != 0;             // some pair inside the word is sparse
	r.c0 = c & 0xFFu; r.c3 = c >> 24;
	return r;
}
LB2_DEVNI void lb2_diag_scan(lb2_win &W, const uint32_t *bits, uint32_t g0, int len, int maxmm, const uint32_t *nm = nullptr)
{
	lb2_sh *sh = W.sh; const unsigned tid = lb2_tid(), nt = lb2_nthr();
	uint32_t *tasks = (uint32_t *)W.ws0.sortk;      // (the build's sort keys: idle whenever a scan runs)
	if (tid == 0) { sh->scan_emax = 0; sh->scan_wmax = 0; sh->walk_next = 0; }
	lb2_sync();
	if (maxmm > 3) { maxmm = 3; if (tid == 0) { sh->err |= 1u << LB2_D_KMAX; } }
	const bool filter = W.P->min_k >= 11 && maxmm >= 1;
	const uint32_t bias = (0x80u - (uint32_t)(maxmm + 1)) * 0x010101u;
	const int ndiag = len - 1;
	// ---- pass 1: classify, queue the stretches
	for (int r = 0; r * (int)nt < ndiag; ++r) {
		const int j = r * (int)nt + ((r & 1) ? (int)(nt - 1 - tid) : (int)tid);
		if (j >= ndiag) { continue; }
		const int d = j + 1, np = len - d;                 // positions p in [0, np)
		if (!filter) { tasks[lb2_add32(&sh->walk_next, 1u)] = (uint32_t)d << 12; continue; }
		uint32_t prev_c3 = 4; bool prev_sp = false, prev_slow = false;
		for (int p0 = 0; p0 < np; p0 += 16) {
			const lb2_wflags wf = lb2_scan_word(bits, g0, p0, d, np, bias, nm);
			const bool slow = wf.sp || prev_sp || (prev_c3 + wf.c0 <= (uint32_t)maxmm);
			if (slow && !prev_slow) { tasks[lb2_add32(&sh->walk_next, 1u)] = ((uint32_t)d << 12) | (uint32_t)(p0 >> 4); }
			prev_sp = wf.sp; prev_c3 = wf.c3; prev_slow = slow;
		}
	}
	lb2_sync();
	// ---- pass 2: the exact rule along every queued stretch
	const uint32_t ntask = sh->walk_next; int emax = 0, wmax = 0;
	for (uint32_t t = tid; t < ntask; t += nt) {
		const uint32_t tk = tasks[t]; const int d = (int)(tk >> 12), np = len - d; int p0 = (int)(tk & 0xFFFu) << 4;
		int m1 = -1, m2 = -1, m3 = -1, m4 = -1;             // previous mismatch positions (virtual mismatch at -1)
		uint32_t prev_c3 = 4; bool prev_sp = false;
		if (p0 > 0) {      // history and block count of the skipped word before the stretch
			const lb2_wflags pw = lb2_scan_word(bits, g0, p0 - 16, d, np, bias, nm);
			uint32_t x = pw.fw; const int pb = p0 - 16; prev_c3 = pw.c3;
			if (x) { int b = 31 - lb2_clz32(x); m1 = pb + (b >> 1); x &= ~(1u << b); }
			if (x) { int b = 31 - lb2_clz32(x); m2 = pb + (b >> 1); x &= ~(1u << b); }
			if (x) { int b = 31 - lb2_clz32(x); m3 = pb + (b >> 1); x &= ~(1u << b); }
			if (x) { int b = 31 - lb2_clz32(x); m4 = pb + (b >> 1); }
		}
		bool at_end = false;
		for (; p0 < np; p0 += 16) {
			const lb2_wflags wf = lb2_scan_word(bits, g0, p0, d, np, bias, nm);
			if (filter && !(wf.sp || prev_sp || (prev_c3 + wf.c0 <= (uint32_t)maxmm))) { break; }      // the stretch ends here
			prev_sp = wf.sp; prev_c3 = wf.c3;
			uint32_t fw = wf.fw;
			while (fw) {
				int b = lb2_ctz32(fw); fw &= fw - 1;
				int q = p0 + (b >> 1);
				int run = q - m1 - 1;                       // (q == np-1 is outside the exact-repeat range, but then run ends at np-2 anyway)
				if (run > emax) { emax = run; }
				int far = (maxmm == 0) ? m1 : (maxmm == 1) ? m2 : (maxmm == 2) ? m3 : m4;
				int win = q - far - 1; if (win > wmax) { wmax = win; }
				m4 = m3; m3 = m2; m2 = m1; m1 = q;
			}
			at_end = p0 + 16 >= np;
		}
		if (at_end) {      // end of the diagonal: exact runs may use positions <= np-2, near-repeat windows positions <= np-1
			{ int run = (np - 1) - m1 - 1; if (run > emax) { emax = run; } }
			{ int far = (maxmm == 0) ? m1 : (maxmm == 1) ? m2 : (maxmm == 2) ? m3 : m4; int win = np - far - 1; if (win > wmax) { wmax = win; } }
		}
	}
	if (emax > 0) { lb2_max32(&sh->scan_emax, (uint32_t)emax); }
	if (wmax > 0) { lb2_max32(&sh->scan_wmax, (uint32_t)wmax); }
	lb2_sync();
}

// pack the loaded path (ws.pathseq, ACGT) 2-bit into `dst` (all lanes); dst must hold plen/16 + 3 words
LB2_DEVNI void lb2_pack_path(lb2_win &W, uint32_t *dst)
{
	const unsigned tid = lb2_tid(), nt = lb2_nthr(); const uint32_t plen = W.sh->plen; const char *s = W.ws.pathseq;
	for (uint32_t w = tid; w < (plen >> 4) + 3; w += nt) {
		uint32_t v = 0;
		for (uint32_t i = 0; i < 16; ++i) { uint32_t q = w * 16 + i; if (q < plen) { v |= (uint32_t)(lb2_code(s[q]) & 3) << (2 * i); } }
		dst[w] = v;
	}
	lb2_sync();
}

// ---------------------------------------------------------------------------------------------------
// bfs (src/Graph.cc:1299-1425): best = first complete path (dequeue order) with the most not-yet-flagged edges.
// By all lanes: the FIFO queue is expanded a CTA-width of entries at a time.  Lane i takes the i-th unvisited entry,
// counts the children it will push, an exclusive scan over the lanes places every entry's children behind those of the
// entries before it -- the order the reference's one-at-a-time loop pushes them in -- and the visit limit (DFS_LIMIT:
// the reference stops at its (limit+1)-th dequeue) becomes "entries with index >= limit are never expanded".  Every lane
// keeps the first best candidate among its own (increasing) indices; the winner is the highest score, lowest index.
// Returns the queue index of the best complete path or LB2_NIL, the same value in every lane.
// ---------------------------------------------------------------------------------------------------
LB2_DEVNI uint32_t lb2_bfs(lb2_win &W)
{
	lb2_ws &ws = W.ws; lb2_sh *sh = W.sh; const int K = sh->K; const unsigned tid = lb2_tid(), nt = lb2_nthr();
	const uint32_t cap = W.C->queue_cap;
	// the head of the queue lives in the idle shared-memory scratch (most searches never leave it), the rest in the slab
	lb2_qent *const Qs = (lb2_qent *)ws.px, *const Qg = ws.queue; const uint32_t qs = (uint32_t)(((size_t)ws.px_words * 4) / sizeof(lb2_qent));
#define Q_AT(i) ((i) < qs ? Qs[(i)] : Qg[(i)])
	const int reflen = (int)sh->seq_len;
	const uint8_t *const NE = ws.d_ne, *const FL = ws.d_flags, *const EOV = ws.d_eov; const lb2_edge *const ED = ws.d_edge, *const EP = ws.e_pool;
	const uint16_t *const LEN = ws.d_len; const uint32_t sink = sh->sink; const int maxlen = reflen + W.P->max_indel_len;
	const uint32_t limit = W.P->dfs_limit ? (uint32_t)W.P->dfs_limit : 0xFFFFFFFFu;
	if (tid == 0) {
		sh->q_smem = qs; sh->bfs_score = 0; sh->bfs_best = LB2_NIL;
		lb2_qent root; root.parent = LB2_NIL; root.node = sh->source; root.len = K; root.score = 0; root.eidx = 0; root.dirflag = 2;   // dir F, flag 1
		Q_AT(0) = root;
	}
	lb2_sync();
	uint32_t qh = 0, qt = 1, my_best = LB2_NIL; int my_score = -1;      // (qh, qt: the same in every lane that takes part)
	// one round: the next `width` unvisited entries, one per lane; returns the children pushed (LB2_NIL: queue overflow)
	auto expand = [&](uint32_t lane_, uint32_t width, bool whole_cta) -> uint32_t {
		const uint32_t end = qt < limit ? qt : limit, idx = qh + lane_;
		lb2_qent e; e.parent = 0; e.node = 0; e.len = 0; e.score = 0; e.eidx = 0; e.dirflag = 0;
		const lb2_edge *ed = nullptr; int ne = 0, pdir = 0, pflag = 0; uint32_t nchild = 0;
		if (idx < end && lane_ < width) {
			e = Q_AT(idx); const uint32_t cur = e.node; pdir = e.dirflag & 1; pflag = (e.dirflag >> 1) & 1;
			if (cur == sink && pflag == 0) { if ((int)e.score > my_score) { my_best = idx; my_score = (int)e.score; } }
			else if (e.len > maxlen) { }
			else {
				const uint32_t ov = EOV[cur]; ed = ov ? (EP + (size_t)(ov - 1) * LB2_ECAP) : (ED + (size_t)cur * LB2_EINL); ne = NE[cur];
				for (int i = 0; i < ne; ++i) { if (lb2_is_dir(ed[i].dir, pdir)) { ++nchild; } }
			}
		}
		uint32_t total = 0; const uint32_t off = whole_cta ? lb2_block_excl(sh->scan, nchild, &total) : lb2_warp_excl(nchild, &total);
		if (total > cap - qt) { return LB2_NIL; }
		if (nchild) {
			uint32_t o = qt + off;
			for (int i = 0; i < ne; ++i) {
				const lb2_edge ei = ed[i];
				if (!lb2_is_dir(ei.dir, pdir)) { continue; }
				const uint32_t other = ei.to;
				lb2_qent c; c.parent = idx; c.node = other; c.eidx = (uint8_t)i;
				c.len = e.len + (int)((FL[other] & LB2_NF_SPECIAL) ? 0u : (uint32_t)LEN[other]) - K + 1;
				const int nflag = pflag * (int)ei.flag;
				c.score = (uint16_t)(e.score + (ei.flag == 0 ? 1 : 0));
				c.dirflag = (uint8_t)(lb2_dir_dest(ei.dir) | (nflag << 1));
				Q_AT(o) = c; ++o;
			}
		}
		qh = (end - qh > width) ? qh + width : end; qt += total;
		return total;
	};
	// small searches (nearly all of them) never leave the first warp: a warp-width of entries per round, warp barriers only;
	// a frontier that keeps growing is handed to the whole CTA
	bool overflow = false;
	if (tid < LB2_WARP) {
		while (qh < qt && qh < limit && qt - qh <= 8u * LB2_WARP) {
			if (expand(tid, LB2_WARP, false) == LB2_NIL) { overflow = true; break; }
			lb2_warp_sync();
		}
		if (tid == 0) { sh->bfs_qh = qh; sh->bfs_qt = overflow ? LB2_NIL : qt; }
	}
	lb2_sync();
	qh = sh->bfs_qh; qt = sh->bfs_qt;
	if (qt == LB2_NIL) { if (tid == 0) { sh->err |= 1u << LB2_D_QUEUE; } lb2_sync(); return LB2_NIL; }
	while (qh < qt && qh < limit) {
		if (expand(tid, nt, true) == LB2_NIL) { if (tid == 0) { sh->err |= 1u << LB2_D_QUEUE; } lb2_sync(); return LB2_NIL; }
		lb2_sync();
	}
#undef Q_AT
	const uint32_t key = (my_best == LB2_NIL) ? 0u : (uint32_t)my_score + 1u;
	if (key) { lb2_max32(&sh->bfs_score, key); }
	lb2_sync();
	if (key && key == sh->bfs_score) { lb2_min32(&sh->bfs_best, my_best); }
	lb2_sync();
	const uint32_t best = sh->bfs_best;
	lb2_sync();
	return best;
}

// materialise the chosen path: node list, string, per-base tumour/normal coverage
LB2_DEVNI void lb2_load_path(lb2_win &W, uint32_t best)
{
	lb2_ws &ws = W.ws; lb2_sh *sh = W.sh; const int K = sh->K;
	lb2_qent *const Qs = (lb2_qent *)ws.px, *const Qg = ws.queue; const uint32_t qs = sh->q_smem;
#define Q_AT(i) ((i) < qs ? Qs[(i)] : Qg[(i)])
	uint32_t n = 0;
	for (uint32_t x = best; x != LB2_NIL; x = Q_AT(x).parent) { ++n; }
	if (n > LB2_MAX_PNODES) { sh->err |= 1u << LB2_D_PATH; return; }
	uint32_t k = n;
	for (uint32_t x = best; x != LB2_NIL; x = Q_AT(x).parent) { --k; const lb2_qent e = Q_AT(x); ws.pnodes[k] = e.node; ws.peidx[k] = e.eidx; }
#undef Q_AT
	sh->pn = n;
	for (uint32_t i = 1; i < n; ++i) { ws.pdirs[i - 1] = lb2_edges(ws, ws.pnodes[i - 1])[ws.peidx[i]].dir; }
	// Path_t::str / covDistr: where every node's contribution starts (lane 0); the copy itself is lb2_copy_path
	int dir = lb2_dir_start(ws.pdirs[0]);
	uint32_t plen = 0;
	for (uint32_t i = 0; i < n; ++i) {
		uint32_t nd = ws.pnodes[i];
		ws.pstart[i] = plen | ((uint32_t)dir << 31);
		if (!lb2_special(W, nd)) {
			uint32_t bl = ws.d_len[nd]; uint32_t from = plen ? (uint32_t)K - 1 : 0;
			if (plen + (bl - from) > LB2_MAX_PATH) { sh->err |= 1u << LB2_D_PATH; return; }
			plen += bl - from;
		}
		if (i + 1 < n) { dir = lb2_dir_dest(ws.pdirs[i]); }
	}
	ws.pstart[n] = plen;
	sh->plen = plen;
}

// all lanes: bases and per-base tumour/normal coverage of the loaded path
LB2_DEVNI void lb2_copy_path(lb2_win &W)
{
	lb2_ws &ws = W.ws; lb2_sh *sh = W.sh; const int K = sh->K; const unsigned tid = lb2_tid(), nt = lb2_nthr();
	const uint32_t n = sh->pn;
	for (uint32_t i = 0; i < n; ++i) {
		uint32_t nd = ws.pnodes[i];
		if (lb2_special(W, nd)) { continue; }
		uint32_t ps = ws.pstart[i] & 0x7FFFFFFFu; bool dir = (ws.pstart[i] >> 31) != 0;
		uint32_t from = ps ? (uint32_t)K - 1 : 0;
		lb2_nview v; lb2_view(W, nd, v);
		for (uint32_t j = from + tid; j < v.len; j += nt) {
			uint32_t src = dir ? (v.len - 1 - j) : j, dst = ps + (j - from);
			char ch = lb2_vchar(W, v, src);
			ws.pathseq[dst] = dir ? lb2_comp(ch) : ch;
			ws.pcovT[dst] = lb2_vcov(W, v, src, 0); ws.pcovN[dst] = lb2_vcov(W, v, src, 1);
		}
	}
	lb2_sync();
}

LB2_DEV uint32_t lb2_pathcontig(lb2_win &W, int pos) {   // Path_t::pathcontig
	lb2_ws &ws = W.ws; const int K = W.sh->K; int curpos = 0;
	for (uint32_t i = 0; i < W.sh->pn; ++i) {
		uint32_t nd = ws.pnodes[i];
		if (!lb2_special(W, nd)) {
			int span = (int)ws.d_len[nd];
			if (curpos + span >= pos) { return nd; }
			curpos += span - K + 1;
		}
	}
	return LB2_NIL;
}
LB2_DEV bool lb2_status_T(lb2_win &W, uint32_t nd) {     // Node_t::isStatusCnt('T') src/Node.cc:423-440
	double prc = (double)(int)W.ws.d_stT[nd] / (double)W.ws.d_stn[nd];
	return prc > 0.8;
}
LB2_DEV void lb2_flag_path(lb2_win &W, int flag) {
	lb2_ws &ws = W.ws;
	for (uint32_t i = 1; i < W.sh->pn; ++i) { lb2_edges(ws, ws.pnodes[i - 1])[ws.peidx[i]].flag = (uint16_t)flag; }
}

// ---------------------------------------------------------------------------------------------------
// global_align_aff(S = trimmed reference, T = path): anti-diagonal wavefront over the CTA.
// tb byte per cell: M.tb (0 '\\', 1 '<', 2 '^', 3 '*') | X.tb<<2 (0 '<', 1 '-', 2 other) | Y.tb<<4 (0 '^', 1 '|', 2 other)
// ---------------------------------------------------------------------------------------------------
// (the traceback bytes are stored anti-diagonal-major, tb[(i+j) * (n+1) + i]: the lanes of a wavefront write consecutive bytes)
// One cell record {M, X, Y} per row index and anti-diagonal, three rotating anti-diagonals: a cell reads the records
// [i-1] and [i] of the previous anti-diagonal and [i-1] of the one before with one load each.  The matrix borders are
// written by two designated lanes, the loop body is interior cells only.
struct lb2_cell16 { int16_t m, x, y, pad; };
struct lb2_cell32 { int32_t m, x, y, pad; };
template <class CT> LB2_DEV void lb2_align_fill_t(lb2_win &W, CT *dp, const char *T)
{
	lb2_ws &ws = W.ws; lb2_sh *sh = W.sh; const unsigned tid = lb2_tid(), nt = lb2_nthr();
	const int n = (int)sh->seq_len, m = (int)sh->plen;
	const char *S = W.ref_raw + sh->seq_off;
	const int stride = n + 2;
	const size_t row = (size_t)n + 1;
	for (int d = 0; d <= n + m; ++d) {
		CT *C0 = dp + (d % 3) * stride; const CT *C1 = dp + ((d + 2) % 3) * stride, *C2 = dp + ((d + 1) % 3) * stride;
		uint8_t *tbd = ws.tb + (size_t)d * row;
		if (tid == 0 && d <= m) {          // cell (0, d)
			CT c; c.pad = 0;
			if (d == 0) { c.m = 0; c.x = -8; c.y = -8; tbd[0] = 3 | (2 << 2) | (2 << 4); }
			else { c.m = (decltype(c.m))(-8 - d); c.x = c.m; c.y = 0; tbd[0] = 2 | (2 << 2) | (2 << 4); }
			C0[0] = c;
		}
		if (tid == (nt > 1 ? 1u : 0u) && d >= 1 && d <= n) {      // cell (d, 0)
			CT c; c.pad = 0; c.m = (decltype(c.m))(-8 - d); c.y = c.m; c.x = 0; C0[d] = c; tbd[d] = 1 | (2 << 2) | (2 << 4);
		}
		int lo = d - m; if (lo < 1) { lo = 1; } const int hi = (d - 1 < n) ? d - 1 : n;      // interior cells: 1 <= i <= n, 1 <= j = d - i <= m
		for (int i = lo + (int)tid; i <= hi; i += (int)nt) {
			const int j = d - i;
			const CT up = C1[i - 1], lf = C1[i], dg = C2[i - 1];      // (i-1, j), (i, j-1), (i-1, j-1)
			int xe = (int)up.x - 1, xo = (int)up.m - 8; int x, xt; if (xe > xo) { x = xe; xt = 1; } else { x = xo; xt = 0; }
			int ye = (int)lf.y - 1, yo = (int)lf.m - 8; int y, yt; if (ye > yo) { y = ye; yt = 1; } else { y = yo; yt = 0; }
			int z = (int)dg.m + ((S[i - 1] == T[j - 1]) ? 2 : -4); int mt = 0;
			if (x > z) { z = x; mt = 1; }
			if (y > z) { z = y; mt = 2; }
			CT c; c.m = (decltype(c.m))z; c.x = (decltype(c.m))x; c.y = (decltype(c.m))y; c.pad = 0; C0[i] = c;
			tbd[i] = (uint8_t)(mt | (xt << 2) | (yt << 4));
		}
		lb2_sync();
	}
}
LB2_DEVNI void lb2_align_fill(lb2_win &W)
{
	lb2_ws &ws = W.ws; lb2_sh *sh = W.sh; const unsigned tid = lb2_tid(), nt = lb2_nthr();
	const uint32_t n = sh->seq_len, m = sh->plen;
	// three anti-diagonals of 16-bit cell records (|score| <= 8 + 4 * (n + m) < 2^15) and the path in the idle shared-memory scratch
	const uint32_t dpw = 3u * (n + 2u) * 2u, tw = (m + 4u) / 4u;
	if (dpw <= ws.px_words) {
		const char *T = ws.pathseq;
		if (dpw + tw <= ws.px_words) {
			char *Ts = (char *)(ws.px + dpw);
			for (uint32_t i = tid; i < m; i += nt) { Ts[i] = ws.pathseq[i]; }
			T = Ts;
			lb2_sync();
		}
		lb2_align_fill_t<lb2_cell16>(W, (lb2_cell16 *)ws.px, T);
	} else { lb2_align_fill_t<lb2_cell32>(W, (lb2_cell32 *)ws.dp, ws.pathseq); }
}

// traceback by the first warp.  Every step reads one traceback byte whose address depends on the previous step -- a
// chain of dependent global loads -- but alignments are mostly runs of plain diagonal moves: lane l looks at the cell l
// steps down the diagonal, a ballot gives the length of the leading run of diagonal moves, and the lanes emit those
// columns together.  Anything else (gaps, the forcex / forcey continuation states) takes one step by the scalar rule.
LB2_DEVNI void lb2_align_trace(lb2_win &W)
{
	lb2_ws &ws = W.ws; lb2_sh *sh = W.sh; const unsigned lane = lb2_tid();
	if (lane >= LB2_WARP) { return; }
	const int n = (int)sh->seq_len, m = (int)sh->plen; const size_t row = (size_t)n + 1;
	const char *S = W.ref_raw + sh->seq_off; const char *T = ws.pathseq;
	int i = n, j = m; bool forcex = false, forcey = false; uint32_t L = 0; bool bad = false;      // (identical in every lane)
	while (i > 0 || j > 0) {
		if (!forcex && !forcey) {
			const int ii = i - (int)lane, jj = j - (int)lane; bool diag = false;
			if (ii > 0 && jj > 0) { diag = (ws.tb[(size_t)(ii + jj) * row + ii] & 3) == 0; }
			const uint32_t bal = lb2_ballot(diag);
			const int nrun = (bal == 0xFFFFFFFFu) ? 32 : lb2_ctz32(~bal);
			if (nrun > 0) {
				if ((int)lane < nrun) { ws.aln_ref[L + lane] = S[ii - 1]; ws.aln_path[L + lane] = T[jj - 1]; }
				L += (uint32_t)nrun; i -= nrun; j -= nrun;
				continue;
			}
		}
		uint8_t tb = ws.tb[(size_t)(i + j) * row + i]; int t = tb & 3, x = (tb >> 2) & 3, y = (tb >> 4) & 3;
		char a, b;
		if (t == 3) { break; }
		else if (forcex) { if (i <= 0) { bad = true; break; } a = S[i - 1]; b = '-'; if (x == 0) { forcex = false; } --i; }
		else if (t == 1) { if (i <= 0) { bad = true; break; } a = S[i - 1]; b = '-'; if (x == 1) { forcex = true; } --i; }
		else if (forcey) { if (j <= 0) { bad = true; break; } a = '-'; b = T[j - 1]; if (y == 0) { forcey = false; } --j; }
		else if (t == 2) { if (j <= 0) { bad = true; break; } a = '-'; b = T[j - 1]; if (y == 1) { forcey = true; } --j; }
		else { a = S[i - 1]; b = T[j - 1]; --i; --j; }
		if (lane == 0) { ws.aln_ref[L] = a; ws.aln_path[L] = b; }
		++L;
	}
	if (bad) { if (lane == 0) { sh->err |= 1u << LB2_D_ALIGN; } return; }
	lb2_warp_sync();
	for (uint32_t k = lane; k < L / 2; k += LB2_WARP) {
		char c = ws.aln_ref[k]; ws.aln_ref[k] = ws.aln_ref[L - 1 - k]; ws.aln_ref[L - 1 - k] = c;
		c = ws.aln_path[k]; ws.aln_path[k] = ws.aln_path[L - 1 - k]; ws.aln_path[L - 1 - k] = c;
	}
	if (lane == 0) { sh->aln_len = L; }
}

// ---------------------------------------------------------------------------------------------------
// transcripts
// ---------------------------------------------------------------------------------------------------
LB2_DEV void lb2_tr_add(lb2_trans &t, int list, lb2_cov c) {
	uint16_t v[4] = { c.fwd, c.rev, c.mqf, c.mqr };
	if (t.n[list] == 0) { for (int k = 0; k < 4; ++k) { t.mn[list][k] = v[k]; t.mn0[list][k] = v[k]; } t.sum[list][0] = 0; t.sum[list][1] = 0; }
	t.n[list] += 1;
	t.sum[list][0] = (uint16_t)(t.sum[list][0] + c.fwd); t.sum[list][1] = (uint16_t)(t.sum[list][1] + c.rev);
	for (int k = 0; k < 4; ++k) {
		if (v[k] < t.mn[list][k]) { t.mn[list][k] = v[k]; }
		if (v[k] < t.mn0[list][k] && v[k] != 0) { t.mn0[list][k] = v[k]; }
	}
}
LB2_DEV lb2_cov lb2_refcov_at(lb2_win &W, uint32_t pos, int sample) {   // Ref_t::getCovStructAt: only fwd/rev are ever set
	lb2_cov c; c.fwd = 0; c.rev = 0; c.mqf = 0; c.mqr = 0;
	if (pos < W.sh->L) { c.fwd = W.ws.refcov[((size_t)sample * LB2_MAX_REF + pos) * 2]; c.rev = W.ws.refcov[((size_t)sample * LB2_MAX_REF + pos) * 2 + 1]; }
	return c;
}
LB2_DEV bool lb2_isACGT(char c) { return c == 'A' || c == 'C' || c == 'G' || c == 'T'; }

// ---------------------------------------------------------------------------------------------------
// findTandems (src/util.cc:574-758) of the loaded path for EVERY query position at once, by all lanes.  The reference
// walks i = 0..len-1 and, per unit length m, compares the block at i with the block at offsets[m][i % m] -- the last
// position of that phase where the comparison failed ("flagged").  Between two flagged positions all blocks are equal,
// so the block at i can just as well be compared with the block at i-m: every (i,m) is independent.  A flagged (i,m)
// finds its run start by stepping back over unflagged positions, applies the reference's length / left-neighbour /
// minimal-unit tests and becomes an event; the per-variant answer is a filter over the events in (i,m) order.
// ---------------------------------------------------------------------------------------------------
LB2_DEVNI void lb2_path_tandems(lb2_win &W)
{
	lb2_ws &ws = W.ws; lb2_sh *sh = W.sh; const lb2_params *P = W.P; const unsigned tid = lb2_tid(), nt = lb2_nthr();
	const uint32_t MAXU = (uint32_t)P->max_unit_len, slen = sh->plen;
	uint8_t *JV = (uint8_t *)ws.px; char *sq = (char *)JV + (size_t)MAXU * slen;
	const bool fits = MAXU >= 1 && MAXU <= 16 && slen > 0 && slen < 0xFFF0u && (size_t)MAXU * slen + slen + 32 <= (size_t)ws.px_words * 4;
	if (tid == 0) { sh->n_tev = fits ? 0u : LB2_NIL; sh->tev_ovf = 0; }
	if (!fits) { lb2_sync(); return; }
	for (uint32_t i = tid; i < slen + 16; i += nt) { sq[i] = (i < slen) ? ws.pathseq[i] : (char)0; }
	lb2_sync();
	for (uint32_t idx = tid; idx < MAXU * slen; idx += nt) {
		const uint32_t m = idx / slen + 1, i = idx % slen, ref = (i >= m) ? i - m : i;
		uint32_t j = 0;
		while (j < m && i + j < slen && sq[i + j] == sq[ref + j]) { ++j; }
		JV[idx] = (uint8_t)(j | ((j != m || i + j + 1 == slen) ? 0x80u : 0u));
	}
	lb2_sync();
	for (uint32_t idx = tid; idx < MAXU * slen; idx += nt) {
		const uint32_t v = JV[idx]; if (!(v & 0x80u)) { continue; }
		const uint32_t m = idx / slen + 1, i = idx % slen, j = v & 0x7Fu;
		const uint8_t *row = JV + (size_t)(m - 1) * slen;
		int q = (int)i - (int)m; while (q >= 0 && !(row[q] & 0x80u)) { q -= (int)m; }
		const uint32_t offset = (q >= 0) ? (uint32_t)q : i % m, span = i - offset;
		if (span / m < (uint32_t)P->min_report_units || span < (uint32_t)P->min_report_len) { continue; }
		const char left = (offset >= 1) ? sq[offset - 1] : (char)0;
		if (left == sq[offset + m - 1]) { continue; }
		uint32_t ml = 1;
		while (ml < m) {
			const uint32_t units = (span + j) / ml; bool allmatch = true;
			for (uint32_t index = 1; allmatch && index < units; ++index) {
				for (uint32_t x = 0; x < ml; ++x) { if (sq[offset + x] != sq[offset + index * ml + x]) { allmatch = false; break; } }
			}
			if (!allmatch) { ++ml; } else { break; }
		}
		if (ml != m) { continue; }
		const uint32_t slot = lb2_add32(&sh->n_tev, 1u);
		if (slot < LB2_MAX_TEV) { lb2_tev e; e.i = (uint16_t)i; e.off = (uint16_t)offset; e.m = (uint8_t)m; e.j = (uint8_t)j; e.pad = 0; sh->tev[slot] = e; }
		else { sh->tev_ovf = 1; }
	}
	lb2_sync();
	if (tid == 0) {
		if (sh->tev_ovf) { sh->n_tev = LB2_NIL; }
		else {      // the reference meets the events in (i, m) order
			for (uint32_t a = 1; a < sh->n_tev; ++a) {
				lb2_tev x = sh->tev[a]; int b = (int)a - 1;
				while (b >= 0 && (sh->tev[b].i > x.i || (sh->tev[b].i == x.i && sh->tev[b].m > x.m))) { sh->tev[b + 1] = sh->tev[b]; --b; }
				sh->tev[b + 1] = x;
			}
		}
	}
	lb2_sync();
}

// column scan + stats + emission.  aligned strings are in ws.aln_ref / ws.aln_path.  The reference walks every column;
// only the columns that are not '=' do anything, so all lanes classify the columns, two prefix sums give every column's
// reference / path position and the list of non-'=' columns, and lane 0 walks that list (lb2_scan_columns).
LB2_DEV uint32_t lb2_col_class(char r, char p) { return r == '-' ? 1u : p == '-' ? 2u : (r != p) ? 3u : 0u; }   // 0 '=', 1 '^', 2 'v', 3 'x'
LB2_DEVNI void lb2_scan_columns(lb2_win &W, const uint32_t *pre, const uint32_t *lst, uint32_t nne)
{
	lb2_ws &ws = W.ws; lb2_sh *sh = W.sh; const int K = sh->K;
	const char *ra = ws.aln_ref, *pa = ws.aln_path;
	const uint32_t plen = sh->plen;
	lb2_trans *tr = ws.trans; uint32_t ts = 0;
	char *poolR = ws.tstr, *poolQ = ws.tstr + (LB2_MAX_PATH + LB2_MAX_REF + 8); uint32_t usedR = 0, usedQ = 0;
	const uint32_t trim5 = sh->trim5;
	uint32_t prev_i = LB2_NIL;
	for (uint32_t e = 0; e < nne; ++e) {
		const uint32_t i = lst[e], cl = lb2_col_class(ra[i], pa[i]), pv = pre[i];
		const char code = (cl == 1) ? '^' : (cl == 2) ? 'v' : 'x';
		const uint32_t pos_in_ref = pv & 0xFFFFu, pathpos = (pv >> 16) + ((cl != 2) ? 1u : 0u);
		const bool prev_noneq = (i == 0) || (prev_i != LB2_NIL && prev_i + 1 == i);      // prev_code != '='
		prev_i = i;
		uint32_t spanner = lb2_pathcontig(W, (int)pathpos);
		if (spanner == LB2_NIL) { break; }
		bool within_tumor = lb2_status_T(W, spanner);
		if (pathpos == 0 || pathpos > plen || i == 0) { sh->err |= 1u << LB2_D_ALIGN; return; }   // reference: UB / assert
		int P = (int)pathpos - 1;
		lb2_cov COVn = ws.pcovN[P], COVt = ws.pcovT[P];
		lb2_cov REFn = lb2_refcov_at(W, pos_in_ref + trim5, 1), REFt = lb2_refcov_at(W, pos_in_ref + trim5, 0);
		uint32_t rrpos = pos_in_ref + (uint32_t)sh->ref_start + trim5;
		int pr = (int)i - 1, pq = (int)i - 1;
		while (pr >= 0 && !lb2_isACGT(ra[pr])) { --pr; }
		while (pq >= 0 && !lb2_isACGT(pa[pq])) { --pq; }
		if (pr < 0 || pq < 0) { sh->err |= 1u << LB2_D_ALIGN; return; }
		if (ts > 0 && prev_noneq) {
			lb2_trans &t = tr[ts - 1];
			if (within_tumor) { t.isSomatic = 1; }
			if (usedR >= LB2_MAX_PATH + LB2_MAX_REF || usedQ >= LB2_MAX_PATH + LB2_MAX_REF) { sh->err |= 1u << LB2_D_TRANS; return; }
			poolR[usedR++] = ra[i]; t.ref_len++; poolQ[usedQ++] = pa[i]; t.qry_len++;
			t.end_pos = (uint32_t)P; t.ref_end_pos = pos_in_ref;
			if (code == '^' && t.code == code && t.pos == rrpos) { lb2_tr_add(t, 0, COVn); lb2_tr_add(t, 1, COVt); }
			else if (code == 'v' && t.code == code && (t.pos + t.ref_len) == rrpos) { lb2_tr_add(t, 2, REFn); lb2_tr_add(t, 3, REFt); }
			else if (code == 'x' || t.code != code) {
				t.code = 'c';
				lb2_tr_add(t, 0, COVn); lb2_tr_add(t, 1, COVt); lb2_tr_add(t, 2, REFn); lb2_tr_add(t, 3, REFt);
			}
		} else {
			if (ts >= LB2_MAX_TRANS) { sh->err |= 1u << LB2_D_TRANS; return; }
			lb2_trans &t = tr[ts++];
			t.pos = rrpos; t.ref_pos = pos_in_ref; t.start_pos = (uint32_t)P + 1; t.code = (uint8_t)code;
			t.end_pos = (uint32_t)P; t.ref_end_pos = pos_in_ref; t.isSomatic = within_tumor ? 1 : 0;
			t.prev_bp_ref = (uint8_t)ra[pr]; t.prev_bp_alt = (uint8_t)pa[pq];
			t.ref_off = usedR; t.qry_off = usedQ; t.ref_len = 1; t.qry_len = 1;
			poolR[usedR++] = ra[i]; poolQ[usedQ++] = pa[i];
			for (int l = 0; l < 4; ++l) { t.n[l] = 0; }
			lb2_tr_add(t, 0, COVn); lb2_tr_add(t, 1, COVt); lb2_tr_add(t, 2, REFn); lb2_tr_add(t, 3, REFt);
		}
	}
	// trailing columns, stats, emission (src/Graph.cc:1036-1190)
	for (uint32_t ti = 0; ti < ts; ++ti) {
		lb2_trans &t = tr[ti];
		if (t.code != 'x') {
			for (int j = 0; j <= K; ++j) {
				uint32_t idx1 = t.end_pos + (uint32_t)j;
				if (idx1 < plen) {
					uint32_t sp = lb2_pathcontig(W, (int)idx1);
					if (sp == LB2_NIL) { break; }
					if (lb2_status_T(W, sp)) { t.isSomatic = 1; }
					lb2_tr_add(t, 0, ws.pcovN[idx1]); lb2_tr_add(t, 1, ws.pcovT[idx1]);
				}
				uint32_t idx2 = t.ref_end_pos + trim5 + (uint32_t)j;
				lb2_tr_add(t, 2, lb2_refcov_at(W, idx2, 1)); lb2_tr_add(t, 3, lb2_refcov_at(W, idx2, 0));
			}
		}
		const bool snv = (t.code == 'x');
		uint16_t RCNF = t.mn[2][0], RCNR = t.mn[2][1], RCTF = t.mn[3][0], RCTR = t.mn[3][1];
		uint16_t ACNF = snv ? t.mn[0][2] : t.mn[0][0], ACNR = snv ? t.mn[0][3] : t.mn[0][1];
		if (!snv) { ACNF = t.mn0[0][0]; ACNR = t.mn0[0][1]; }
		uint16_t ACTF = snv ? t.mn[1][2] : t.mn[1][0], ACTR = snv ? t.mn[1][3] : t.mn[1][1];
		if (t.isSomatic) {
			// mean = (float)sum/(float)n stored into an unsigned short (src/Transcript.hh:201-203)
			RCNF = (uint16_t)((float)t.sum[2][0] / (float)t.n[2]); RCNR = (uint16_t)((float)t.sum[2][1] / (float)t.n[2]);
			RCTF = (uint16_t)((float)t.sum[3][0] / (float)t.n[3]); RCTR = (uint16_t)((float)t.sum[3][1] / (float)t.n[3]);
			ACNF = 0; ACNR = 0;
		}
		if (ACNF > 0 || ACNR > 0 || ACTF > 0 || ACTR > 0) {
			uint32_t w = sh->w;
			if (sh->n_var >= W.ovar_cap) { sh->err |= 1u << LB2_D_VARIANTS; return; }
			uint32_t need = t.ref_len + t.qry_len;
			char *spool = W.ostr;
			if (sh->str_used + need + 64 > W.ostr_cap) { sh->err |= 1u << LB2_D_STRINGS; return; }
			lb2_variant v;
			v.window = w; v.pos = (int32_t)t.pos - 1; v.str_off = sh->str_used;
			v.ref_len = (uint16_t)t.ref_len; v.alt_len = (uint16_t)t.qry_len;
			char *dst = spool + sh->str_used;
			for (uint32_t k = 0; k < t.ref_len; ++k) { dst[k] = poolR[t.ref_off + k]; }
			for (uint32_t k = 0; k < t.qry_len; ++k) { dst[t.ref_len + k] = poolQ[t.qry_off + k]; }
			int LEN = 0; uint32_t ml = 0; bool movf = false;
			const char *ps = ws.pathseq;
			bool ans = false;
			if (sh->n_tev != LB2_NIL) {
				const int pos = (int)t.start_pos, delta = W.P->dist_from_str;
				for (uint32_t e = 0; e < sh->n_tev; ++e) {
					const lb2_tev ev = sh->tev[e]; const int start = (int)ev.off, end = (int)ev.i + (int)ev.j;
					if (pos >= start - delta && pos <= end + delta) {
						ans = true; LEN = end - start;
						for (uint32_t z = 0; z < ev.m; ++z) { if (ml < 64) { dst[need + ml++] = ps[ev.off + z]; } else { movf = true; } }
					}
				}
			} else { ans = lb2_find_tandems([&](uint32_t q) -> char { return ps[q]; }, plen, W.P, (int)t.start_pos, LEN, dst + need, ml, 64, movf); }
			if (movf) { sh->err |= 1u << LB2_D_MOTIF; return; }
			v.motif_len = (uint16_t)(ans ? ml : 0); v.str_len = (uint16_t)(ans ? LEN : 0);
			sh->str_used += need + v.motif_len;
			v.rcn_fwd = RCNF; v.rcn_rev = RCNR; v.rct_fwd = RCTF; v.rct_rev = RCTR;
			v.acn_fwd = ACNF; v.acn_rev = ACNR; v.act_fwd = ACTF; v.act_rev = ACTR;
			v.code = t.code; v.prev_bp_ref = t.prev_bp_ref; v.prev_bp_alt = t.prev_bp_alt; v.kmer = (uint8_t)K;
			W.ovar[sh->n_var] = v;
			sh->n_var += 1;
		}
	}
}

LB2_DEVNI void lb2_scan_alignment(lb2_win &W)      // all lanes
{
	lb2_ws &ws = W.ws; lb2_sh *sh = W.sh; const char *ra = ws.aln_ref, *pa = ws.aln_path; const uint32_t alen = sh->aln_len;
	uint32_t *pre = (uint32_t *)ws.sortk, *lst = pre + ((alen + 3u) & ~3u);      // (the build's sort keys are idle in the graph stage)
	// reference position before the column (low half) | path position before the column (high half)
	lb2_excl_scan(W, alen, [&](uint32_t i) -> uint32_t { const uint32_t c = lb2_col_class(ra[i], pa[i]); return ((c != 1u) ? 1u : 0u) | ((c != 2u) ? 0x10000u : 0u); },
	              [&](uint32_t i, uint32_t v) { pre[i] = v; });
	const uint32_t nne = lb2_excl_scan(W, alen, [&](uint32_t i) -> uint32_t { return lb2_col_class(ra[i], pa[i]) ? 1u : 0u; },
	                                   [&](uint32_t i, uint32_t v) { if (lb2_col_class(ra[i], pa[i])) { lst[v] = i; } });
	if (lb2_tid() == 0) { lb2_scan_columns(W, pre, lst, nne); }
}

// does the loaded path spell exactly the trimmed reference?  (all lanes)
LB2_DEVNI bool lb2_path_is_ref(lb2_win &W)
{
	lb2_ws &ws = W.ws; lb2_sh *sh = W.sh; const unsigned tid = lb2_tid(), nt = lb2_nthr();
	if (sh->seq_len != sh->plen) { return false; }
	if (tid == 0) { sh->flag_b = 0; }
	lb2_sync();
	const char *S = W.ref_raw + sh->seq_off; bool diff = false;
	for (uint32_t i = tid; i < sh->plen; i += nt) { if (S[i] != ws.pathseq[i]) { diff = true; } }
	if (diff) { sh->flag_b = 1; }
	lb2_sync();
	const bool same = sh->flag_b == 0;
	lb2_sync();      // (flag_b is reused right away)
	return same;
}

// processPath for the path loaded in ws.pathseq (all lanes: the alignment is CTA-wide)
LB2_DEVNI void lb2_process_path(lb2_win &W)
{
	lb2_ws &ws = W.ws; lb2_sh *sh = W.sh; const unsigned tid = lb2_tid(), nt = lb2_nthr();
	// HammingDistance cut-off (src/Graph.cc:818-826): equal lengths and <= 5 mismatches => the identity alignment
	const bool samelen = sh->seq_len == sh->plen;
	if (tid == 0) { sh->flag_b = 0; }
	lb2_sync();
	if (samelen) {
		const char *S = W.ref_raw + sh->seq_off; uint32_t hd = 0;
		for (uint32_t i = tid; i < sh->plen; i += nt) { const char a = S[i], b = ws.pathseq[i]; ws.aln_ref[i] = a; ws.aln_path[i] = b; if (a != b) { ++hd; } }
		if (hd) { lb2_add32(&sh->flag_b, hd); }
	}
	lb2_sync();
	if (tid == 0) { sh->need_align = (!samelen || sh->flag_b > 5) ? 1u : 0u; if (!sh->need_align) { sh->aln_len = sh->plen; } }
	lb2_sync();
	if (samelen && sh->flag_b == 0) {      // the path IS the reference: every column is '=', no transcript, nothing to emit
		lb2_mark(W, LB2_PH_ALIGN);
		return;
	}
	if (sh->need_align) {
		lb2_align_fill(W);
		lb2_align_trace(W);
		lb2_sync();
	}
	lb2_mark(W, LB2_PH_ALIGN);
	lb2_path_tandems(W);
	if (!sh->err) { lb2_scan_alignment(W); }
	lb2_sync();
	lb2_mark(W, LB2_PH_SCAN);
}

#endif
