// lb2_state.cuh -- per-CTA workspace layout and shared state of the per-window pipeline.
#ifndef LB2_STATE_CUH
#define LB2_STATE_CUH

#include "lb2_common.h"
#include "lb2_pack.cuh"
#include "lb2_kmer.cuh"

#define LB2_NIL 0xFFFFFFFFu
#define LB2_SENT 0xFFFFFFFEu      // "before_begin" of the emulated libstdc++ forward list

// node flag bits
#define LB2_NF_DEAD    0x01
#define LB2_NF_SOURCE  0x02
#define LB2_NF_SINK    0x04
#define LB2_NF_GONE    0x08       // erased from the map (cleanDead)
#define LB2_NF_SPECIAL (LB2_NF_SOURCE | LB2_NF_SINK)

// build-space half-edge (dense ids over ALL nodes of the map) and row-space half-edge (ids over the
// survivors of the first low-coverage sweep + source/sink nodes, < 4096)
struct lb2_bedge { uint32_t to : 24, dir : 2, flag : 1, type : 5; };
struct lb2_edge  { uint16_t to : 12, dir : 2, flag : 1, pad : 1; };
#define LB2_BECAP 8               // a k-mer node has at most 8 neighbours (orientation x appended base)
#define LB2_MAX_ROWS 4096

struct lb2_qent {                 // one partial path of bfs() (reference src/Graph.cc:1299-1425)
	uint32_t parent;              // queue index of the prefix
	uint32_t node;                // current node
	int32_t  len;
	uint16_t score;
	uint8_t  eidx;                // edge index in parent's node that led here
	uint8_t  dirflag;             // bit0: dir (0=F,1=R), bit1: flag
};

struct lb2_trans {                // Transcript_t (reference src/Transcript.hh:33-120), running stats only
	uint32_t pos, ref_pos, start_pos, end_pos, ref_end_pos;
	uint32_t ref_off, ref_len, qry_off, qry_len;   // strings in the scratch pool
	uint8_t  code, isSomatic, prev_bp_ref, prev_bp_alt;
	// running computeStats state per list: n, sum(fwd,rev) [u16 wrap], min(fwd,rev,mqf,mqr), min_non0(fwd,rev,mqf,mqr)
	uint32_t n[4];                // 0 altN 1 altT 2 refN 3 refT
	uint16_t sum[4][2];
	uint16_t mn[4][4];
	uint16_t mn0[4][4];
};

// pointers into this CTA's global-memory slab (identical layout for every CTA) and, for the hot
// graph-stage arrays, into the CTA's shared memory (assigned per (window,k) by lb2_order_and_pack)
struct lb2_ws {
	// --- build stage ---
	uint32_t *used; uint64_t *sortk; uint16_t *inst; uint32_t *mates; uint32_t *bseq;
	uint32_t *g_cnt; uint32_t *g_em;   // per-slot accumulators of the build (fed by fire-and-forget reductions)
	uint32_t *b_rep; uint64_t *b_hash; uint32_t *b_cnt; int32_t *b_mincovqv; uint8_t *b_flags; uint8_t *b_stT; uint8_t *b_ne;
	lb2_bedge *b_edge; uint32_t *b_row;
	// --- reads ---
	uint32_t *rd_start; uint32_t *rd_len; uint32_t *rd_t5; uint32_t *rd_info; uint32_t *rd_rank; uint32_t *rd_kbase; uint32_t *rd_mate; uint64_t *rd_src;      // rd_src: first word in the packed pool | words << 32
	// --- graph stage, row space.  hot (shared memory): ---
	uint32_t *d_lnext; uint16_t *d_bk; uint16_t *d_cnext; uint16_t *buckets; uint8_t *d_ne; uint8_t *d_flags; uint8_t *d_color; uint8_t *d_eov; int16_t *d_comp;
	uint16_t *d_pos; uint32_t *px; uint32_t px_words;   // (packed-read words, dead in the graph stage) list index of every row; scratch of the parallel compaction
	lb2_edge *d_edge; lb2_edge *e_pool; float *d_cov; uint16_t *d_len; uint16_t *d_stn; uint16_t *d_stT; uint32_t *stack; uint32_t *chain; uint32_t *cpos;
	// cold (global):
	uint32_t *d_rep; uint64_t *d_hash; uint32_t *d_cnt; uint32_t *d_orig; int32_t *d_mincov; int32_t *d_mincovqv; uint32_t *d_str; uint32_t *d_cd;
	uint16_t *deficit;            // [dense node][K][4] low-quality deficits (only when the window has low-qual bases)
	uint32_t *refnode;            // [LB2_MAX_REF] node of the reference k-mer at each offset (dense id, then row id)
	uint16_t *refcov;             // [2 samples][LB2_MAX_REF][2] fwd,rev
	uint8_t  *arena;
	lb2_edge *etmp;               // side buffer of the parallel edge surgery [LB2_MAX_ROWS][LB2_ECAP] + a count per row
	uint32_t *emu;                // order-emulation arrays when they do not fit in shared memory
	lb2_qent *queue; uint32_t *jobs; uint32_t *pstart;
	// --- path processing ---
	char *pathseq; lb2_cov *pcovN; lb2_cov *pcovT; uint32_t *pnodes; uint8_t *pdirs; uint8_t *peidx;
	char *aln_ref; char *aln_path; int32_t *dp; uint8_t *tb; lb2_trans *trans; char *tstr;
};

struct lb2_sizes { size_t total; size_t off[64]; };

// one qualifying tandem repeat of the current path (findTandems event: flagged position i, unit m, matched prefix j, run start off)
struct lb2_tev { uint16_t i, off; uint8_t m, j; uint16_t pad; };
#define LB2_MAX_TEV 48

// shared (smem) scalars of one window
struct lb2_sh {
	lb2_mbar mbar; uint32_t mbar_phase, pad0;      // completion barrier of the bulk-async staging copies (first: 8-byte aligned)
	// window
	uint32_t w, R, L, total_bp, ref_g, has_lowq, lowq_live, bits_live, has_pairs, mapped, status, detail;
	int32_t  ref_start;
	// per k
	int32_t  K, nw;
	uint32_t n_used, n_nodes, n_rows, n_spec, n_jobs, n_eov, err;
	uint32_t totalreadbp;
	uint32_t flag_a, flag_b, flag_c, flag_d; uint32_t scan_emax, scan_wmax, ref_emax, ref_wmax;
	// reference trimming state (Ref_t::seq/trim5/trim3, persists across k: SURVEY B4)
	uint32_t seq_off, seq_len; uint32_t trim5, trim3;
	// order emulation
	uint32_t bkt_count, bkt_cap, elem_count, next_resize, lhead, chead, cm_valid;      // chead/cm_valid: the current component's own node list (d_cnext)
	// anchors
	uint32_t source, sink, anc_src, anc_snk, anc_amb, spec_cap;
	uint32_t arena_used, tstr_used;
	// path
	uint32_t plen, pn, need_align, n_trans, path_found, aln_len, q_smem, bfs_score, bfs_best, bfs_qh, bfs_qt;
	// output
	uint32_t n_var, str_used, n_k_tried, final_k, last_nodes;
	int32_t  numcomp;
	uint32_t stop_k; uint32_t n_dead; uint32_t big; uint32_t n_changed;      // n_changed: nodes removed by the sweeps since the first compaction
	uint32_t maxnk, inst_stride, inst_ref;     // occurrence array layout of this (window,k): see lb2_build.cuh
	uint32_t walk_next, walk_pl, walk_np;      // k-mer walk: work-item counter, pairs per piece, pieces per read
	// non-ACGT ('N') bases of the window reference: one bit per base (two readable words past the end), the N-free work
	// items of the reference walk, the map entries its N-containing k-mers become (src/Graph.cc:534-540: the reference is
	// loaded untrimmed)
	uint32_t ref_hasN, n_refitems, n_nk; uint32_t refn[LB2_MAX_REF / 32 + 2];
	uint32_t n_tev, tev_ovf; lb2_tev tev[LB2_MAX_TEV];     // tandem repeats of the loaded path (n_tev = LB2_NIL: not computed, scan per variant)
	unsigned long long prof[24]; unsigned long long t_last;
	uint32_t scan[40];            // block-scan partials (one per warp, lb2_block_excl)
};

// phase ids for the optional cycle profile (lb2_dev_out::prof)
enum { LB2_PH_STAGE = 0, LB2_PH_PRESCAN, LB2_PH_REFSCAN, LB2_PH_WALK, LB2_PH_COMPACT, LB2_PH_MATES, LB2_PH_LOWQ, LB2_PH_CLEAR,
       LB2_PH_REFCOV, LB2_PH_ORDER, LB2_PH_LOWCOV_CC, LB2_PH_COMP_SEQ, LB2_PH_BFS, LB2_PH_PATHSCAN, LB2_PH_ALIGN, LB2_PH_SCAN, LB2_PH_OTHER,
       LB2_PH_ANCHOR, LB2_PH_CSWEEP, LB2_PH_CMAT, LB2_PH_CCLEAN, LB2_PH_CP_LINK, LB2_PH_CP_FOLD, LB2_PH_BFS_SEQ, LB2_PH_N };

#ifdef __CUDACC__
#define LB2_HD __host__ __device__ inline
#else
#define LB2_HD static inline
#endif

// shared-memory layout: lb2_sh | ref_raw[LB2_MAX_REF] | bits[max_bp/16 + 4] | lowq[max_bp/32 + 4] | region T
LB2_HD size_t lb2_smem_fixed(uint32_t max_bp) {
	return ((sizeof(lb2_sh) + 15) & ~(size_t)15) + LB2_MAX_REF + ((size_t)max_bp / 16 + 4) * 4 + ((size_t)max_bp / 32 + 4) * 4;
}
// region T follows the quality mask: Mer->Node keys (u32) + slot->node ids (u16) during the build; together with the
// mask's bytes it must also hold the graph-stage arrays (cfg.graph_bytes)
LB2_HD size_t lb2_lowq_bytes(uint32_t max_bp) { return ((size_t)max_bp / 32 + 4) * 4; }
LB2_HD size_t lb2_treg_bytes(uint32_t table_slots, uint32_t graph_bytes, uint32_t max_bp) {
	size_t t = (size_t)table_slots * 6, lq = lb2_lowq_bytes(max_bp);
	size_t g = graph_bytes > lq ? graph_bytes - lq : 0;
	return ((t > g ? t : g) + 15) & ~(size_t)15;
}
LB2_HD size_t lb2_smem_bytes(uint32_t max_bp, uint32_t table_slots, uint32_t graph_bytes) { return ((lb2_smem_fixed(max_bp) + 15) & ~(size_t)15) + lb2_treg_bytes(table_slots, graph_bytes, max_bp); }


#endif
