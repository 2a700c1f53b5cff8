// lb2_pipeline.cuh -- one window through the whole micro-assembly: the k sweep of
// Microassembler::processGraph (reference src/Microassembler.cc:73-249) on one CTA.
#ifndef LB2_PIPELINE_CUH
#define LB2_PIPELINE_CUH

#include "lb2_paths.cuh"

// byte layout of one CTA's workspace slab; the same function runs on host (sizing) and device (pointers)

LB2_HD size_t lb2_ws_layout(const lb2_cfg &c, uint8_t *base, lb2_ws *ws)
{
	size_t off = 0;
#define LB2_TAKE(field, type, count) do { off = (off + 15) & ~(size_t)15; if (ws) { ws->field = (type *)(base + off); } off += sizeof(type) * (size_t)(count); } while (0)
	const size_t MN = (size_t)c.max_nodes + 16, MR = (size_t)c.max_reads + 2;
	size_t n2 = 1; while (n2 < c.max_nodes || n2 < c.max_inst || n2 < MR) { n2 <<= 1; }
	LB2_TAKE(used, uint32_t, c.max_nodes); LB2_TAKE(bseq, uint32_t, MN * 8); LB2_TAKE(g_cnt, uint32_t, (size_t)c.table_slots * 2); LB2_TAKE(g_em, uint32_t, c.table_slots); LB2_TAKE(sortk, uint64_t, n2); LB2_TAKE(inst, uint16_t, c.max_inst + 16); LB2_TAKE(mates, uint32_t, 2 * (size_t)c.max_inst + 16);
	LB2_TAKE(rd_start, uint32_t, MR); LB2_TAKE(rd_len, uint32_t, MR); LB2_TAKE(rd_t5, uint32_t, MR);
	LB2_TAKE(rd_info, uint32_t, MR); LB2_TAKE(rd_rank, uint32_t, MR); LB2_TAKE(rd_kbase, uint32_t, MR); LB2_TAKE(rd_mate, uint32_t, MR); LB2_TAKE(rd_src, uint64_t, MR);
	LB2_TAKE(b_rep, uint32_t, MN); LB2_TAKE(b_hash, uint64_t, MN); LB2_TAKE(b_cnt, uint32_t, MN * 4); LB2_TAKE(b_mincovqv, int32_t, MN);
	LB2_TAKE(b_flags, uint8_t, MN); LB2_TAKE(b_stT, uint8_t, MN); LB2_TAKE(b_ne, uint8_t, MN); LB2_TAKE(b_edge, lb2_bedge, MN * LB2_BECAP); LB2_TAKE(b_row, uint32_t, MN);
	LB2_TAKE(d_rep, uint32_t, LB2_MAX_ROWS); LB2_TAKE(d_hash, uint64_t, LB2_MAX_ROWS); LB2_TAKE(d_cnt, uint32_t, LB2_MAX_ROWS * 4); LB2_TAKE(d_orig, uint32_t, LB2_MAX_ROWS);
	LB2_TAKE(d_mincov, int32_t, LB2_MAX_ROWS); LB2_TAKE(d_mincovqv, int32_t, LB2_MAX_ROWS); LB2_TAKE(d_str, uint32_t, LB2_MAX_ROWS); LB2_TAKE(d_cd, uint32_t, LB2_MAX_ROWS);
	LB2_TAKE(deficit, uint16_t, c.deficit_bytes / 2);
	LB2_TAKE(refnode, uint32_t, LB2_MAX_REF); LB2_TAKE(refcov, uint16_t, 2 * LB2_MAX_REF * 2);
	LB2_TAKE(arena, uint8_t, c.arena_bytes); LB2_TAKE(etmp, lb2_edge, (size_t)LB2_MAX_ROWS * LB2_ECAP + LB2_MAX_ROWS / 2 + 8); LB2_TAKE(emu, uint32_t, (size_t)c.max_nodes * 3 + (size_t)c.max_nodes * 9 / 4 + 64); LB2_TAKE(queue, lb2_qent, c.queue_cap); LB2_TAKE(jobs, uint32_t, LB2_MAX_ROWS * 10); LB2_TAKE(pstart, uint32_t, LB2_MAX_PNODES + 1);
	LB2_TAKE(pathseq, char, LB2_MAX_PATH + 16); LB2_TAKE(pcovN, lb2_cov, LB2_MAX_PATH + 16); LB2_TAKE(pcovT, lb2_cov, LB2_MAX_PATH + 16);
	LB2_TAKE(pnodes, uint32_t, LB2_MAX_PNODES); LB2_TAKE(pdirs, uint8_t, LB2_MAX_PNODES); LB2_TAKE(peidx, uint8_t, LB2_MAX_PNODES);
	LB2_TAKE(aln_ref, char, LB2_MAX_PATH + LB2_MAX_REF + 16); LB2_TAKE(aln_path, char, LB2_MAX_PATH + LB2_MAX_REF + 16);
	LB2_TAKE(dp, int32_t, 12 * (LB2_MAX_REF + 2)); LB2_TAKE(tb, uint8_t, (size_t)(LB2_MAX_REF + 1) * (LB2_MAX_PATH + LB2_MAX_REF + 2));      /* anti-diagonal-major */
	LB2_TAKE(trans, lb2_trans, LB2_MAX_TRANS); LB2_TAKE(tstr, char, 2 * (LB2_MAX_PATH + LB2_MAX_REF + 8));
#undef LB2_TAKE
	return (off + 255) & ~(size_t)255;
}

LB2_DEV uint32_t lb2_first_err(uint32_t e) { for (uint32_t b = 0; b < 32; ++b) { if (e & (1u << b)) { return b; } } return 0; }

// reference k-mer coverage tracks (Ref_t::indexMers/updateCoverage/computeCoverage, src/Ref.cc:40-64,128-149,173-250)
LB2_DEVNI void lb2_ref_coverage(lb2_win &W)
{
	lb2_ws &ws = W.ws; lb2_sh *sh = W.sh; const int K = sh->K; const unsigned tid = lb2_tid(), nt = lb2_nthr();
	const uint32_t L = sh->L;
	for (uint32_t i = tid; i < 2 * LB2_MAX_REF * 2; i += nt) { ws.refcov[i] = 0; }
	// mers indexed from the (possibly already trimmed) ref->seq: offsets i with i + K < seq.length()
	for (uint32_t i = tid; i + K < sh->seq_len; i += nt) {
		uint32_t nd = ws.refnode[sh->seq_off + i];
		if (nd != LB2_NIL) { ws.b_flags[nd] |= 0x10; }      // same bit from every writer; the low-coverage bit was set before the barrier
	}
	lb2_sync();
	for (uint32_t i = tid; i + K < L; i += nt) {
		uint32_t nd = ws.refnode[i];
		uint16_t v[4] = { 0, 0, 0, 0 };
		if (nd != LB2_NIL && (ws.b_flags[nd] & 0x10)) { for (int c = 0; c < 4; ++c) { v[c] = (uint16_t)ws.b_cnt[nd * 4 + c]; } }
		for (int s = 0; s < 2; ++s) {
			uint16_t *rc = ws.refcov + (size_t)s * LB2_MAX_REF * 2;
			if (i == 0) { for (int j = 0; j < K; ++j) { rc[j * 2] = v[s * 2]; rc[j * 2 + 1] = v[s * 2 + 1]; } }
			else { rc[(i + K - 1) * 2] = v[s * 2]; rc[(i + K - 1) * 2 + 1] = v[s * 2 + 1]; }
		}
	}
	lb2_sync();
}

LB2_DEVNI void lb2_process_window(lb2_win &W, uint32_t w)
{
	lb2_sh *sh = W.sh; const lb2_params *P = W.P; const unsigned tid = lb2_tid();
	if (tid == 0) {
#ifdef LB2_PROFILE
		for (int i = 0; i < 24; ++i) { sh->prof[i] = 0; } sh->t_last = lb2_clock();
#endif
		uint32_t slot = LB2_NIL;
		if (W.escal && W.O->big_count) {      // (a window redone twice keeps the large slab it was given the first time)
			slot = W.O->big_slot[w];
			if (slot == LB2_NIL) { slot = lb2g_add32(W.O->big_count, 1u); if (slot >= W.O->big_cap) { slot = LB2_NIL; } W.O->big_slot[w] = slot; }
		}
		sh->big = slot;
		// (the descriptor is shared by the CTA: one lane writes it, everybody reads it after the barrier)
		if (slot != LB2_NIL) { W.ovar = W.O->big_variants + (size_t)slot * W.O->big_max_var; W.ostr = W.O->big_strings + (size_t)slot * W.O->big_str_bytes; W.ovar_cap = W.O->big_max_var; W.ostr_cap = W.O->big_str_bytes; }
		else { W.ovar = W.O->variants + (size_t)w * W.C->max_var; W.ostr = W.O->strings + (size_t)w * W.C->str_bytes; W.ovar_cap = W.C->max_var; W.ostr_cap = W.C->str_bytes; }
	}
	lb2_sync();
	lb2_stage_window(W, w);
	lb2_mark(W, LB2_PH_STAGE);
	if (sh->status == LB2_WIN_OK) {
		if (P->max_unit_len > 16 || P->max_k > (int32_t)W.C->max_k || P->max_mismatch > 3) { if (tid == 0) { sh->status = LB2_WIN_UNSUPPORTED; sh->detail = LB2_D_KMAX; } lb2_sync(); }
	}
	if (sh->status == LB2_WIN_OK) {
		// one pass over the window reference answers isRepeat / isAlmostRepeat for every k
		lb2_diag_scan(W, W.bits, sh->ref_g, (int)sh->L, P->max_mismatch, sh->ref_hasN ? sh->refn : nullptr);
		if (tid == 0) {
			sh->ref_emax = sh->scan_emax; sh->ref_wmax = sh->scan_wmax;
			// window pre-skip: isRepeat(rawseq, maxK)  (src/Microassembler.cc:800)
			if ((uint32_t)P->max_k <= sh->ref_emax) { sh->status = LB2_WIN_SKIP_REPEAT; }
		}
		lb2_sync();
		lb2_mark(W, LB2_PH_REFSCAN);
	}
	if (sh->status == LB2_WIN_OK) {
		for (int k = P->min_k; k <= P->max_k; k += 2) {
			// isRepeat(rawseq,k) || isAlmostRepeat(rawseq,k,MAX_MISMATCH)  (src/Microassembler.cc:118-131)
			if ((uint32_t)k <= sh->ref_emax || (uint32_t)k + 1 <= sh->ref_wmax) { continue; }
			lb2_sync();
			if (tid == 0) { W.ws = W.ws0; }       // row-space pointers go back to their global homes
			lb2_sync();
			lb2_build_graph(W, k);
			if (tid == 0) { sh->n_k_tried += 1; sh->final_k = (uint32_t)k; }
			lb2_sync();
			lb2_mark(W, LB2_PH_CLEAR);
			if (sh->err) { break; }
			lb2_ref_coverage(W);
			lb2_mark(W, LB2_PH_REFCOV);
			lb2_order_and_pack(W);
			if (sh->err) { break; }
			{
				int nc = lb2_mark_components(W);
				if (tid == 0) { sh->flag_c = 0; sh->numcomp = nc; }
				lb2_mark(W, LB2_PH_LOWCOV_CC);
			}
			lb2_sync();
			if (sh->err) { break; }
			bool retry = false;
			const int numcomp = sh->numcomp;
			for (int c = 1; c <= numcomp; ++c) {
				lb2_find_anchors(W, c);
				if (tid == 0) { lb2_mark_ref_ends(W, c); sh->flag_c = 0; sh->cm_valid = 0; lb2_mark(W, LB2_PH_ANCHOR); }
				lb2_sync();
				// A component without anchors is invisible from here on: both cycle checks, the path-repeat scan and the path
				// enumeration return at once without a source (src/Graph.cc:602, :689, :2422), and the sweeps in between only
				// touch this component's own nodes, which nothing looks at again -- so it is not compacted or swept at all
				if (!sh->err && (sh->source == LB2_NIL || sh->sink == LB2_NIL)) { continue; }
				if (!sh->err) {
					// the cycle check before the first compaction (src/Microassembler.cc:179) gives the same answer on the
					// compacted graph (a chain of links is always traversed whole), where it costs a handful of nodes
					const bool par = lb2_compress_par(W, c);
					if (!par) {
						if (tid == 0) { sh->flag_c = lb2_has_cycle(W, c) ? 1u : 0u; }
						lb2_sync();
						if (!sh->flag_c) { lb2_compress(W, c); }
					} else {
						// (from here on the sweeps walk the component's own node list, not the whole map)
						if (tid == 0 && !sh->err) { lb2_build_members(W, c); sh->flag_c = lb2_has_cycle(W, c) ? 1u : 0u; }
						lb2_sync();
					}
				}
				LB2_SEQMARK(LB2_PH_PRESCAN);
				if (!sh->flag_c && !sh->err) {
					if (tid == 0 && !sh->err) { if (!sh->cm_valid) { lb2_build_members(W, c); } lb2_remove_lowcov(W, c); }      // removeLowCov(true,c): sweep, cleanDead, then compress
					lb2_sync();
					if (!sh->err && sh->flag_b) { lb2_compress(W, c); }          // (compaction is idempotent: skipped when the sweep removed nothing)
					LB2_SEQMARK(LB2_PH_MATES);
					if (!sh->err) { lb2_remove_tips(W, c); }
					LB2_SEQMARK(LB2_PH_LOWQ);
					if (!sh->err) { lb2_remove_short_links(W, c); }
					// (the graph the first check called acyclic is only checked again if one of the three sweeps changed it)
					if (tid == 0 && !sh->err && sh->n_changed) { sh->flag_c = lb2_has_cycle(W, c) ? 1u : 0u; }
				}
				lb2_mark(W, LB2_PH_COMP_SEQ);
				lb2_sync();
				if (sh->err) { break; }
				if (sh->flag_c) { retry = true; break; }
				if (sh->source == LB2_NIL || sh->sink == LB2_NIL) { continue; }
				// ---- findRepeatsInGraphPaths: enumerate the covering paths, near-repeat test on each
				bool rpt = false; uint32_t nflag = 0;
				while (true) {
					if (nflag > 8 * LB2_MAX_ROWS) { if (tid == 0) { sh->err |= 1u << LB2_D_STACK; } lb2_sync(); break; }   // every round flags >= 1 new edge
					{
						const uint32_t best = lb2_bfs(W);
						if (tid == 0) {
							sh->path_found = (best != LB2_NIL && !sh->err) ? 1u : 0u;
							if (sh->path_found) { lb2_load_path(W, best); }
							lb2_mark(W, LB2_PH_BFS_SEQ);
						}
					}
					lb2_sync();
					if (sh->err || !sh->path_found) { lb2_mark(W, LB2_PH_BFS); break; }
					lb2_copy_path(W);
					lb2_mark(W, LB2_PH_BFS);
					// A path that spells the (trimmed) reference is a substring of the window reference: every window of every
					// diagonal inside it exists in the window's own scan, whose wmax this k has already passed (k + 1 >
					// ref_wmax, or the k loop would have skipped this k) -- no need to scan it again
					if (!lb2_path_is_ref(W)) {
						lb2_pack_path(W, W.ws.cpos);               // cpos is idle outside lb2_compress
						lb2_diag_scan(W, W.ws.cpos, 0, (int)sh->plen, P->max_mismatch);
						lb2_mark(W, LB2_PH_PATHSCAN);
						if ((uint32_t)k + 1 <= sh->scan_wmax) { rpt = true; break; }
					}
					if (tid == 0) { lb2_flag_path(W, 1); }
					++nflag;
					lb2_sync();
				}
				if (sh->err) { break; }
				if (tid == 0) {   // clear the edge flags again (all flags of this component were 0 before)
					auto unflag = [&](uint32_t p) -> bool { for (int e = 0; e < (int)W.ws.d_ne[p]; ++e) { lb2_edges(W.ws, p)[e].flag = 0; } return true; };
					lb2_each_node(W, c, unflag);
					unflag(sh->source); unflag(sh->sink);      // (the component's only source/sink nodes: the node walk leaves them out)
				}
				lb2_sync();
				if (rpt) { retry = true; break; }
				// ---- eka: repeat { best path; processPath; flag its edges }.  The enumeration is the one the loop above just did
				// (same graph, flags cleared): if that found exactly one path, the path is still loaded and copied and there is
				// nothing to search for, neither now nor after it has been flagged
				const bool single = (nflag == 1);
				for (uint32_t round = 0; ; ++round) {
					if (round > 8 * LB2_MAX_ROWS) { if (tid == 0) { sh->err |= 1u << LB2_D_STACK; } lb2_sync(); break; }
					if (single) { if (round == 1) { break; } }
					else {
						{
							const uint32_t best = lb2_bfs(W);
							if (tid == 0) {
								sh->path_found = (best != LB2_NIL && !sh->err) ? 1u : 0u;
								if (sh->path_found) { lb2_load_path(W, best); }
								lb2_mark(W, LB2_PH_BFS_SEQ);
							}
						}
						lb2_sync();
						if (sh->err || !sh->path_found) { lb2_mark(W, LB2_PH_BFS); break; }
						lb2_copy_path(W);
						lb2_mark(W, LB2_PH_BFS);
					}
					lb2_process_path(W);
					if (sh->err) { break; }
					if (tid == 0) { lb2_flag_path(W, 1); }
					lb2_sync();
				}
				if (sh->err) { break; }
			}
			if (sh->err) { break; }
			if (!retry) { break; }
		}
		if (sh->err) { if (tid == 0) { sh->status = LB2_WIN_OVERFLOW; sh->detail = lb2_first_err(sh->err); } lb2_sync(); }
	}
	if (tid == 0) {
		lb2_window_info wi; wi.status = (uint8_t)sh->status; wi.final_k = (uint8_t)sh->final_k; wi.n_k_tried = (uint16_t)sh->n_k_tried;
		wi.n_variants = (sh->status == LB2_WIN_OK) ? sh->n_var : 0; wi.n_nodes = sh->last_nodes; wi.detail = sh->detail;
		lb2_mark(W, LB2_PH_OTHER);
#ifdef LB2_PROFILE_SEQ      // (debug: this window's cycles / 256 in the detail word)
		{ unsigned long long t = 0; for (int i = 0; i < LB2_PH_N; ++i) { t += sh->prof[i]; } wi.detail = (uint32_t)(t >> 8); }
#endif
		W.O->info[w] = wi; W.O->str_used[w] = sh->str_used;
#if defined(LB2_PROFILE) && !defined(LB2_HOSTSIM)
		if (W.O->prof) { for (int i = 0; i < LB2_PH_N; ++i) { if (sh->prof[i]) { atomicAdd(&W.O->prof[i], sh->prof[i]); } } }
#endif
	}
	lb2_sync();
}

#endif
