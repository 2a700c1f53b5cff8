// lb2_state.cuh -- per-CTA workspace layout and shared state of the per-window pipeline.
#ifndef LB2_STATE_CUH
#define LB2_STATE_CUH

#include "lb2_common.h"
#include "lb2_kmer.cuh"

#define LB2_NIL 0xFFFFFFFFu
#define LB2_SENT 0xFFFFFFFEu      // "before_begin" of the emulated libstdc++ forward list

// node flag bits
#define LB2_NF_DEAD    0x01
#define LB2_NF_SOURCE  0x02
#define LB2_NF_SINK    0x04
#define LB2_NF_GONE    0x08       // erased from the map (cleanDead)
#define LB2_NF_SPECIAL (LB2_NF_SOURCE | LB2_NF_SINK)

struct lb2_edge { uint32_t to; uint8_t dir; uint8_t flag; uint16_t pad; };

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

// pointers into this CTA's global-memory slab (identical layout for every CTA)
struct lb2_ws {
	// --- build stage, slot indexed ---
	uint32_t *used; uint64_t *sortk; uint32_t *inst; uint32_t *mates; uint32_t *bseq;
	// --- reads ---
	uint32_t *rd_start; uint32_t *rd_len; uint32_t *rd_t5; uint32_t *rd_info; uint32_t *rd_rank; uint32_t *rd_kbase;
	// --- dense nodes ---
	uint32_t *d_rep; uint64_t *d_hash; float *d_cov; uint32_t *d_cnt; uint32_t *d_stn; uint32_t *d_stT;
	int32_t *d_mincov; int32_t *d_mincovqv; uint8_t *d_ne; lb2_edge *d_edge; uint8_t *d_flags;
	int32_t *d_comp; uint8_t *d_color; uint32_t *d_lnext; uint32_t *d_str; uint32_t *d_len; uint32_t *d_cd;
	uint16_t *deficit;            // [node][K][4] low-quality deficits (only when the window has low-qual bases)
	uint32_t *buckets;
	uint32_t *refnode;            // [LB2_MAX_REF] dense node of the reference k-mer at each offset
	uint16_t *refcov;             // [2 samples][LB2_MAX_REF][2] fwd,rev
	uint8_t  *arena;
	lb2_qent *queue;
	uint32_t *stack;
	uint32_t *chain;
	// --- path processing ---
	char *pathseq; lb2_cov *pcovN; lb2_cov *pcovT; uint32_t *pnodes; uint8_t *pdirs; uint8_t *peidx;
	char *aln_ref; char *aln_path; int32_t *dp; uint8_t *tb; lb2_trans *trans; char *tstr;
};

struct lb2_sizes { size_t total; size_t off[64]; };

// shared (smem) scalars of one window
struct lb2_sh {
	// window
	uint32_t w, R, L, total_bp, ref_g, has_lowq, has_pairs, mapped, status, detail;
	int32_t  ref_start;
	// per k
	int32_t  K, nw;
	uint32_t n_used, n_nodes, n_spec, err;
	uint32_t totalreadbp;
	uint32_t flag_a, flag_b, flag_c; uint32_t scan_emax, scan_wmax, ref_emax, ref_wmax;
	// reference trimming state (Ref_t::seq/trim5/trim3, persists across k: SURVEY B4)
	uint32_t seq_off, seq_len; uint32_t trim5, trim3;
	// order emulation
	uint32_t bkt_count, elem_count, next_resize, lhead;
	// anchors
	uint32_t source, sink;
	uint32_t arena_used, tstr_used;
	// path
	uint32_t plen, pn, need_align, n_trans, path_found, aln_len;
	// output
	uint32_t n_var, str_used, n_k_tried, final_k, last_nodes;
	int32_t  numcomp;
	uint32_t stop_k;
	unsigned long long prof[24]; unsigned long long t_last;
};

// phase ids for the optional cycle profile (lb2_dev_out::prof)
enum { LB2_PH_STAGE = 0, LB2_PH_PRESCAN, LB2_PH_REFSCAN, LB2_PH_WALK, LB2_PH_COMPACT, LB2_PH_MATES, LB2_PH_LOWQ, LB2_PH_CLEAR,
       LB2_PH_REFCOV, LB2_PH_ORDER, LB2_PH_LOWCOV_CC, LB2_PH_COMP_SEQ, LB2_PH_BFS, LB2_PH_PATHSCAN, LB2_PH_ALIGN, LB2_PH_SCAN, LB2_PH_OTHER, LB2_PH_N };

#endif
