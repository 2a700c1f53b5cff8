/*
 * lancet_b200.h -- C ABI of the B200-native Lancet micro-assembly hot path.
 *
 * The reference (nygenome/lancet) has no FFI/plugin surface; the cut this
 * library replaces is the in-process call
 *
 *     int Microassembler::processGraph(Graph_t& g, const string& refname, int minK, int maxK)
 *                                                   reference src/Microassembler.hh:224, .cc:73-249
 *
 * with, as INPUT, exactly what Graph_t::addAlignment has accumulated for the window
 * (reference src/Graph.cc:487-501, src/ReadInfo.hh:47-66) plus the window's Ref_t
 * (reference src/Lancet.cc:283-300), and, as OUTPUT, the ordered stream of Variant_t
 * constructor argument tuples that Graph_t::processPath hands to VariantDB_t::addVar
 * (reference src/Graph.cc:1184-1188, src/Variant.hh:106-112).
 *
 * Plain pointers and sizes only.  All input pointers are HOST memory owned by the caller until
 * the call returns; result memory is owned by the context until the next lb2_process /
 * lb2_run / lb2_destroy.  One context per GPU; calls on one context are single-threaded
 * (mirrors "one Microassembler per pthread", reference src/Lancet.cc:868).
 * There is NO CPU fallback: every entry point fails with LB2_ERR_CUDA if no sm_100 device.
 */
#ifndef LANCET_B200_H
#define LANCET_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- parameters: Graph_t setters of reference src/Microassembler.cc:726-753,
 *      defaults of reference src/Lancet.hh:33-79 ------------------------------------------- */
typedef struct lb2_params {
	int32_t min_k;             /* minK              = 11   (--min-k)            */
	int32_t max_k;             /* maxK              = 101  (--max-k)            */
	int32_t min_qual_trim;     /* MIN_QUAL_TRIM     = 10+33 (ASCII threshold)   */
	int32_t min_qual_call;     /* MIN_QUAL_CALL     = 17+33 (ASCII threshold)   */
	int32_t cov_threshold;     /* COV_THRESHOLD     = 5                         */
	int32_t low_cov_threshold; /* LOW_COV_THRESHOLD = 1                         */
	int32_t max_tip_len;       /* MAX_TIP_LEN       = 11                        */
	int32_t dfs_limit;         /* DFS_LIMIT         = 1000000                   */
	int32_t max_indel_len;     /* MAX_INDEL_LEN     = 500                       */
	int32_t max_mismatch;      /* MAX_MISMATCH      = 2                         */
	int32_t max_unit_len;      /* MAX_UNIT_LEN      = 4                         */
	int32_t min_report_units;  /* MIN_REPORT_UNITS  = 3                         */
	int32_t min_report_len;    /* MIN_REPORT_LEN    = 7                         */
	int32_t dist_from_str;     /* DIST_FROM_STR     = 1                         */
	double  min_cov_ratio;     /* MIN_COV_RATIO     = 0.01                      */
} lb2_params;

/* fills *p with the reference defaults */
void lb2_default_params(lb2_params *p);

/* ---- read flags (one byte per pooled read) ------------------------------------------------
 * what extractReads passes to addAlignment (reference src/Microassembler.cc:505-512,618-623) */
#define LB2_READ_NORMAL   0x01u  /* label: 0 = TMR (tumor), 1 = NML (normal)                   */
#define LB2_READ_REVERSE  0x02u  /* strand: 0 = FWD, 1 = REV                                   */
#define LB2_READ_MATE_SHIFT 2    /* bits 2-3: mate_order 0 (unpaired) / 1 (first) / 2 (second) */
#define LB2_READ_UNMAPPED 0x10u  /* code: 0 = 'M' (CODE_MAPPED), 1 = 'B' (CODE_BASTARD)        */

/* ---- one batch of windows ------------------------------------------------------------------
 * Reads live once in a pool; every window lists the pool indices of its reads, TUMOR reads
 * first then NORMAL reads, each in BAM order (reference src/Microassembler.cc:833-834).
 * The same layout is the on-disk ".lb2b" file (magic "LB2B", version 2, then the scalars
 * n_windows,n_reads,n_wr (u32) n_ref_bytes,n_base_bytes (u64), then the arrays in the order
 * declared below), written by lancet_b200/batch.py and read by oracle/ref_harness.cc. */
typedef struct lb2_batch {
	uint32_t        n_windows;
	uint32_t        n_reads;      /* reads in the pool                                          */
	uint32_t        n_wr;         /* total window->read references = wr_off[n_windows]          */
	uint64_t        n_ref_bytes;  /* = ref_off[n_windows]                                       */
	uint64_t        n_base_bytes; /* = base_off[n_reads]                                        */
	/* windows */
	const uint32_t *ref_off;      /* [n_windows+1] byte offsets into ref_seq                    */
	const int32_t  *ref_start;    /* [n_windows]   Ref_t::refstart (1-based pos of rawseq[0])   */
	const uint32_t *chr_id;       /* [n_windows]   caller's chromosome id (opaque, echoed)      */
	const uint32_t *wr_off;       /* [n_windows+1] ranges into wr_idx                           */
	const uint32_t *wr_idx;       /* [n_wr]        pool indices                                 */
	/* read pool */
	const uint64_t *base_off;     /* [n_reads+1]   byte offsets into seq / qual                 */
	const uint8_t  *flags;        /* [n_reads]     LB2_READ_* bits                              */
	const uint32_t *name_rank;    /* [n_reads]     order-preserving rank of the query name:
	                                 rank(a) < rank(b) <=> strcmp(name a, name b) < 0, equal
	                                 names <=> equal ranks (see lb2_rank_names)                 */
	const char     *ref_seq;      /* [n_ref_bytes] window rawseq, upper case, IUPAC -> 'N'      */
	const char     *seq;          /* [n_base_bytes] read bases, ASCII (BamAlignment::QueryBases)*/
	const char     *qual;         /* [n_base_bytes] ASCII phred+33 (BamAlignment::Qualities)    */
} lb2_batch;

/* ---- output: one record per Variant_t handed to VariantDB_t::addVar ---------------------------
 * The fields are the constructor ARGUMENTS at reference src/Graph.cc:1184-1188 (before the
 * normalisation done inside the Variant_t ctor, src/Variant.hh:133-153).  Strings live in the
 * result's string pool: ref (ref_len bytes, '-' for inserted columns), then alt (alt_len), then
 * the STR motif (motif_len), starting at str_off. */
typedef struct lb2_variant {
	uint32_t window;        /* index of the window in the batch                                 */
	int32_t  pos;           /* transcript.pos - 1                                               */
	uint32_t str_off;       /* offset into lb2_result::strings                                  */
	uint16_t ref_len, alt_len, motif_len;
	uint16_t str_len;       /* LEN reported by findTandems (0 = no STR)                         */
	uint16_t rcn_fwd, rcn_rev, rct_fwd, rct_rev;   /* RCN, RCT pairs                            */
	uint16_t acn_fwd, acn_rev, act_fwd, act_rev;   /* ACN, ACT pairs                            */
	uint8_t  code;          /* 'x' snv, '^' ins, 'v' del, 'c' complex                            */
	uint8_t  prev_bp_ref, prev_bp_alt;
	uint8_t  kmer;          /* K the window was assembled with                                   */
} lb2_variant;

/* per-window outcome */
#define LB2_WIN_OK          0   /* processed (zero or more variants)                             */
#define LB2_WIN_SKIP_REPEAT 1   /* isRepeat(rawseq,maxK) (reference src/Microassembler.cc:800)   */
#define LB2_WIN_NO_READS    2   /* no 'M' reads (reference src/Microassembler.cc:83)             */
#define LB2_WIN_OVERFLOW    3   /* a device capacity was exceeded; result for this window unset  */
#define LB2_WIN_UNSUPPORTED 4   /* input outside what the device path handles    */
typedef struct lb2_window_info {
	uint8_t  status;        /* LB2_WIN_*                                                         */
	uint8_t  final_k;       /* last k tried (0 if none)                                          */
	uint16_t n_k_tried;     /* graphs built                                                      */
	uint32_t n_variants;
	uint32_t n_nodes;       /* nodes after build at final k (diagnostic, Graph.cc:3688 analogue) */
	uint32_t detail;        /* overflow / unsupported reason code                                */
} lb2_window_info;

typedef struct lb2_result {
	uint32_t               n_windows;
	uint32_t               n_variants;
	const lb2_window_info *windows;   /* [n_windows]                                            */
	const lb2_variant     *variants;  /* [n_variants], grouped by window, emission order inside  */
	const char            *strings;
	uint64_t               n_string_bytes;
	float                  kernel_ms; /* device time of the per-window pipeline kernels           */
} lb2_result;

typedef struct lb2_ctx lb2_ctx;

#define LB2_OK            0
#define LB2_ERR_CUDA     -1   /* no device / CUDA failure (lb2_strerror has the CUDA text)       */
#define LB2_ERR_ARG      -2
#define LB2_ERR_NOMEM    -3
#define LB2_ERR_STATE    -4

/* create a context bound to CUDA device `device` */
int  lb2_create(lb2_ctx **ctx, const lb2_params *params, int device);
void lb2_destroy(lb2_ctx *ctx);
const char *lb2_strerror(const lb2_ctx *ctx, int code);

/* End-to-end call (the drop-in for "addAlignment* ; processGraph" over many windows):
 * host batch in -> H2D -> kernels -> D2H -> host result out.  When the batch arrays are page-locked
 * (lb2_alloc_pinned below, cudaHostAlloc, cudaHostRegister) the copy runs behind the kernels: windows are
 * assembled as soon as their reads have arrived.  Pageable arrays give the same result, copy first. */
int lb2_process(lb2_ctx *ctx, const lb2_batch *batch, lb2_result *result);

/* page-locked host memory for batch arrays, for callers that do not link the CUDA runtime themselves
 * (NULL when there is no device or no memory) */
void *lb2_alloc_pinned(size_t bytes);
void  lb2_free_pinned(void *p);

/* Split-phase variant for callers that keep batches resident in HBM (and for bench.py):
 * upload copies the batch to the device, run executes the pipeline on the resident batch
 * (may be called repeatedly), download fetches the result of the last run. */
int lb2_upload(lb2_ctx *ctx, const lb2_batch *batch);
int lb2_run(lb2_ctx *ctx);
int lb2_download(lb2_ctx *ctx, lb2_result *result);

/* number of kernels launched by this context so far */
uint64_t lb2_kernel_launches(const lb2_ctx *ctx);

/* ---- measurement / diagnostics (used by bench.py and tools/; not needed by an integrator) --------------------
 * The reference has no counterpart: its only timing output is the per-thread "elapsed time" line of
 * src/Microassembler.cc:865. */
int      lb2_wait(lb2_ctx *ctx);                        /* block until the kernels of the last lb2_run have finished      */
float    lb2_last_kernel_ms(lb2_ctx *ctx);              /* device time of the last lb2_run (CUDA events on its stream)     */
float    lb2_last_kernel_ms_of(lb2_ctx *ctx, int which);/* ... of one kernel: 0 pack (pool pre-pack), 1 window pipeline,
                                                           2 escalation pass, 3 result compaction                       */
uint64_t lb2_last_h2d_bytes(const lb2_ctx *ctx);        /* bytes copied host->device by the last upload / process          */
uint64_t lb2_last_d2h_bytes(const lb2_ctx *ctx);        /* bytes copied device->host by the last download / process        */
uint32_t lb2_resident_ctas(const lb2_ctx *ctx);         /* persistent CTAs of the window kernel for the last batch         */
uint32_t lb2_smem_per_cta(const lb2_ctx *ctx);          /* dynamic shared memory per CTA of the window kernel              */
/* cycles per pipeline phase summed over windows (all zero unless the library was built with -DLB2_PROFILE) */
int      lb2_phase_cycles(lb2_ctx *ctx, unsigned long long *out24, int reset);
/* build identity of the kernels: bench.py refuses profile-derived numbers recorded for another version */
const char *lb2_kernel_version(void);

/* ---- several GPUs: one process per GPU, windows are independent, the only exchange is the gather of the variant records on
 * one rank (reference: the per-thread VariantDB_t are merged serially in thread order, src/Lancet.cc:943-959).  NCCL:
 * counts by ncclAllGather, payloads by ncclSend/ncclRecv to `root`.
 *   lb2_comm_unique_id : rank 0 creates the id and hands it to the other ranks by any means (a file, MPI, ...)
 *   lb2_comm_init      : every rank, after lb2_create, with the same id
 *   lb2_comm_gather    : every rank passes its own records (window = caller's global window index, str_off into strs) and
 *                        n_stats <= 8 counters; on root `merged` holds all ranks' records in rank order (string offsets
 *                        rebased; memory owned by the context) and stats[] the sums over ranks; other ranks get n_variants = 0 */
#define LB2_COMM_ID_BYTES 128
int lb2_comm_unique_id(char id_out[LB2_COMM_ID_BYTES]);
int lb2_comm_init(lb2_ctx *ctx, const char id[LB2_COMM_ID_BYTES], int rank, int world);
int lb2_comm_gather(lb2_ctx *ctx, const lb2_variant *vars, uint32_t n_vars, const char *strs, uint64_t n_str_bytes,
                    uint64_t *stats, int n_stats, int root, lb2_result *merged);

/* helper: order-preserving ranks for n NUL-terminated query names (host side, std::sort) */
int lb2_rank_names(const char *const *names, uint32_t n, uint32_t *rank_out);

#ifdef __cplusplus
}
#endif
#endif
