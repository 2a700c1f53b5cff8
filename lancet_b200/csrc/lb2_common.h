// lb2_common.h -- types shared by the host side and the device pipeline.
#ifndef LB2_COMMON_H
#define LB2_COMMON_H

#include <stdint.h>
#include "../../include/lancet_b200.h"

// ---- compile-time limits -------------------------------------------------------------------
#define LB2_THREADS      128        // threads per CTA (one CTA per window)
#define LB2_ECAP         12         // half-edges per node (8 possible k-mer neighbours + specials)
#define LB2_MAX_REF      1024       // max window reference length
#define LB2_MAX_PATH     2304       // max assembled path length (reflen + MAX_INDEL_LEN + slack)
#define LB2_MAX_TRANS    64         // transcripts per path
#define LB2_MAX_PNODES   512        // nodes per path

// ---- edge directions (reference src/Edge.hh:36) --------------------------------------------
#define LB2_FF 0
#define LB2_FR 1
#define LB2_RF 2
#define LB2_RR 3

// read class = sample*2 + strand : 0 T/fwd, 1 T/rev, 2 N/fwd, 3 N/rev
// per-base coverage record: the four live fields of reference cov_t (src/Ref.hh:41-53)
struct lb2_cov { uint16_t fwd, rev, mqf, mqr; };

// window status detail codes
enum {
	LB2_D_NONE = 0, LB2_D_HASH_FULL, LB2_D_NODES, LB2_D_EDGES, LB2_D_ARENA, LB2_D_QUEUE, LB2_D_PATH,
	LB2_D_VARIANTS, LB2_D_STRINGS, LB2_D_SPECIAL, LB2_D_REFLEN, LB2_D_READS, LB2_D_NREF, LB2_D_ALIGN,
	LB2_D_TRANS, LB2_D_SMEM, LB2_D_KMAX, LB2_D_MOTIF, LB2_D_BUCKETS, LB2_D_STACK
};

// ---- runtime configuration of the device workspace -----------------------------------------
struct lb2_cfg {
	uint32_t table_slots;   // shared-memory Mer->Node table slots (power of two)
	uint32_t max_nodes;     // dense nodes per (window,k)  (<= 3/4 table_slots)
	uint32_t max_reads;     // reads per window (+1 for the reference read)
	uint32_t max_bp;        // staged (trimmed) read bases per window incl. reference (smem)
	uint32_t arena_bytes;   // unitig strings / coverage arrays
	uint32_t deficit_bytes; // low-quality deficit counters [node][k][4] u16
	uint32_t queue_cap;     // BFS path-tree entries
	uint32_t max_inst;      // k-mer occurrences per (window,k) tracked for the overlapping-mate replay
	uint32_t max_var;       // variant records per window
	uint32_t str_bytes;     // string pool per window
	uint32_t bucket_cap;    // buckets for the libstdc++ order emulation (prime >= max_nodes)
	uint32_t max_k;         // largest k supported by the per-k dense arrays
	uint32_t graph_bytes;   // shared memory for the graph-stage arrays (quality-mask bytes included)
	uint32_t n_slots;       // resident CTAs (workspace slabs)
	uint32_t smem_bytes;    // dynamic shared memory per CTA
	uint32_t debug_flags;   // bit 0: sequential first compaction (LB2_DEBUG_FLAGS, debugging aid)
	uint32_t max_special;   // source/sink nodes per (window,k): two per anchored component
};

// one pooled read after the pre-pack pass (lb2_pack.cuh): where its 2-bit words start in the packed pool, the result of
// Graph_t::trim (reference src/Graph.cc:355-384), the read flags the pipeline looks at
struct lb2_pkread {
	uint32_t woff;   // first 16-base word of the read in pk_bits / pk_lowq
	uint16_t t5;     // trm5: bases before the first kept base (the whole length for a junk read)
	uint16_t n;      // kept bases (0: junk read)
	uint16_t nw;     // 16-base words of the untrimmed read
	uint8_t  info;   // bits 0-1 class (sample*2 + strand), bits 2-3 mate order, LB2_PK_*
	uint8_t  lowq;   // some kept base has quality < MIN_QUAL_CALL
};
#define LB2_PK_UNMAPPED 0x10u
#define LB2_PK_TOOLONG  0x20u   /* more than 4095 kept bases: not handled by the device path */

// device view of one uploaded batch
struct lb2_dev_batch {
	uint32_t n_windows;
	const uint32_t *ref_off; const int32_t *ref_start; const uint32_t *wr_off; const uint32_t *wr_idx;
	const uint64_t *base_off; const uint8_t *flags; const uint32_t *name_rank;
	const char *ref_seq; const char *seq; const char *qual;      // seq / qual / base_off / flags: read by the pre-pack pass only
	// the packed pool (written by the pre-pack pass, read by the window pipeline)
	const lb2_pkread *pk;      // [n_reads]
	const uint32_t *pk_bits;   // 2-bit bases, 16 per word, every read on a word boundary
	const uint16_t *pk_lowq;   // one bit per base (quality < MIN_QUAL_CALL), 16 per entry, same indexing as pk_bits
};

// per-window device output slab header (variants + strings follow at fixed strides)
struct lb2_dev_out {
	lb2_window_info *info;      // [n_windows]
	lb2_variant     *variants;  // [n_windows * max_var]
	char            *strings;   // [n_windows * str_bytes]
	uint32_t        *str_used;  // [n_windows]
	unsigned long long *prof;   // [24] cycles per pipeline phase summed over windows (lane 0), may be NULL
	// large output slabs handed to windows of the escalation pass (a window that emits more records than its regular
	// slab holds is one of them): big_slot[w] = index of the large slab window w used, or 0xFFFFFFFF
	lb2_variant     *big_variants; char *big_strings; uint32_t *big_slot; uint32_t *big_count;
	uint32_t         big_cap, big_max_var, big_str_bytes;
};

#endif
