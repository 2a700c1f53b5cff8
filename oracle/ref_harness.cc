// TEST INFRASTRUCTURE ONLY -- never linked into or called by the product path.
//
// ref_windows: drives the UNMODIFIED nygenome/lancet sources (compiled from
// /root/reference/src by oracle/Makefile into oracle/_ref/) at the hot-path
// boundary of SURVEY.md §8(b): per window it replays what
// Microassembler::extractReads would have pushed through Graph_t::addAlignment
// (reference src/Microassembler.cc:618-623, src/Graph.cc:487-501) and then
// calls the reference's own Microassembler::processGraph
// (src/Microassembler.cc:73-249).  Every Variant_t that the reference hands to
// VariantDB_t::addVar (src/Graph.cc:1184-1188) is recorded, in order, through
// a link-time --wrap of that symbol (the reference's VariantDB.cc is linked
// unchanged and still runs).
//
// Input: an .lb2b window-batch file (layout documented in include/lancet_b200.h,
// written by lancet_b200/batch.py).  Output: one TSV line per addVar call.
// Also used as the timed CPU reference arm (bench.py --impl reference).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <string>
#include <vector>
#include <thread>
#include <chrono>
#include <sstream>

#include "Lancet.hh"   // reference globals + Graph_t constants (src/Lancet.hh:33-98); pulls in Microassembler.hh

using namespace std;

struct Rec {
	uint32_t window;
	int pos; char type; int len; string ref, alt; int kmer; string str;
	unsigned short c[8];
	char pbr, pba;
};

static thread_local vector<Rec> * tl_out = NULL;
static thread_local uint32_t tl_window = 0;

extern "C" void __real__ZN11VariantDB_t6addVarERK9Variant_t(VariantDB_t *, const Variant_t &);
extern "C" void __wrap__ZN11VariantDB_t6addVarERK9Variant_t(VariantDB_t * self, const Variant_t & v)
{
	if (tl_out) {
		Rec r;
		r.window = tl_window; r.pos = v.pos; r.type = v.type; r.len = v.len;
		r.ref = v.ref; r.alt = v.alt; r.kmer = v.kmer; r.str = v.str;
		r.c[0] = v.ref_cov_normal_fwd; r.c[1] = v.ref_cov_normal_rev;
		r.c[2] = v.ref_cov_tumor_fwd;  r.c[3] = v.ref_cov_tumor_rev;
		r.c[4] = v.alt_cov_normal_fwd; r.c[5] = v.alt_cov_normal_rev;
		r.c[6] = v.alt_cov_tumor_fwd;  r.c[7] = v.alt_cov_tumor_rev;
		r.pbr = v.prev_bp_ref; r.pba = v.prev_bp_alt;
		tl_out->push_back(r);
	}
	__real__ZN11VariantDB_t6addVarERK9Variant_t(self, v);
}

struct Batch {
	uint32_t n_windows, n_reads, n_wr;
	vector<uint32_t> ref_off, chr_id, wr_off, wr_idx, name_rank;
	vector<int32_t> ref_start;
	vector<uint64_t> base_off;
	vector<uint8_t> flags;
	string ref_seq, seq, qual;
};

template <class T> static void rd(FILE * f, T * p, size_t n) {
	if (n && fread(p, sizeof(T), n, f) != n) { fprintf(stderr, "short read\n"); exit(2); }
}

static void load(const char * fn, Batch & b) {
	FILE * f = fopen(fn, "rb");
	if (!f) { perror(fn); exit(2); }
	char magic[4]; uint32_t ver; uint64_t nref, nbase;
	rd(f, magic, 4); rd(f, &ver, 1); rd(f, &b.n_windows, 1); rd(f, &b.n_reads, 1); rd(f, &b.n_wr, 1); rd(f, &nref, 1); rd(f, &nbase, 1);
	if (memcmp(magic, "LB2B", 4) || ver != 2) { fprintf(stderr, "bad magic/version\n"); exit(2); }
	b.ref_off.resize(b.n_windows + 1); b.ref_start.resize(b.n_windows); b.chr_id.resize(b.n_windows);
	b.wr_off.resize(b.n_windows + 1); b.wr_idx.resize(b.n_wr); b.base_off.resize(b.n_reads + 1); b.flags.resize(b.n_reads);
	b.name_rank.resize(b.n_reads); b.ref_seq.resize(nref); b.seq.resize(nbase); b.qual.resize(nbase);
	rd(f, b.ref_off.data(), b.n_windows + 1); rd(f, b.ref_start.data(), b.n_windows);
	rd(f, b.chr_id.data(), b.n_windows); rd(f, b.wr_off.data(), b.n_windows + 1); rd(f, b.wr_idx.data(), b.n_wr);
	rd(f, b.base_off.data(), b.n_reads + 1); rd(f, b.flags.data(), b.n_reads);
	rd(f, b.name_rank.data(), b.n_reads);
	rd(f, &b.ref_seq[0], nref); rd(f, &b.seq[0], nbase); rd(f, &b.qual[0], nbase);
	fclose(f);
}

struct Opts {
	int minK = 11, maxK = 101, threads = 1, dfs_limit = 1000000, repeat = 1;
	int first = 0, count = -1;
	bool verbose = false;
	const char * out = NULL;
};

static Filters g_filters;

static void worker(const Batch * b, const Opts * o, int tid, int nthreads, int w0, int w1, vector<Rec> * out)
{
	tl_out = out;
	Microassembler ma(false);
	// same field plumbing as reference src/Lancet.cc:871-908 (defaults from src/Lancet.hh:33-79)
	ma.verbose = o->verbose; ma.VERBOSE = false; ma.PRINT_ALL = false; ma.KMER_RECOVERY = false;
	ma.MIN_QV_CALL = 17; ma.MIN_QV_TRIM = 10; ma.QV_RANGE = '!';
	ma.MIN_QUAL_TRIM = 10 + '!'; ma.MIN_QUAL_CALL = 17 + '!';
	ma.minK = o->minK; ma.maxK = o->maxK; ma.MAX_TIP_LEN = 11; ma.MIN_THREAD_READS = 1;
	ma.COV_THRESHOLD = 5; ma.MIN_COV_RATIO = 0.01; ma.LOW_COV_THRESHOLD = 1; ma.MAX_AVG_COV = 10000;
	ma.NODE_STRLEN = 100; ma.DFS_LIMIT = o->dfs_limit; ma.MAX_INDEL_LEN = 500; ma.MAX_MISMATCH = 2;
	ma.MAX_UNIT_LEN = 4; ma.MIN_REPORT_UNITS = 3; ma.MIN_REPORT_LEN = 7; ma.DIST_FROM_STR = 1;
	map<string, Ref_t *> reftable;
	ma.reftable = &reftable;
	ma.setFilters(&g_filters);
	ma.setID(tid + 1);

	Graph_t g;
	// same setters as reference src/Microassembler.cc:726-753
	g.setDB(&ma.vDB); g.setK(ma.minK); g.setVerbose(ma.verbose); g.setMoreVerbose(ma.VERBOSE);
	g.setMinQualTrim(ma.MIN_QUAL_TRIM); g.setMinQualCall(ma.MIN_QUAL_CALL); g.setBufferSize(ma.BUFFER_SIZE);
	g.setDFSLimit(ma.DFS_LIMIT); g.setCovThreshold(ma.COV_THRESHOLD); g.setMinCovRatio(ma.MIN_COV_RATIO);
	g.setLowCovThreshold(ma.LOW_COV_THRESHOLD); g.setPrintDotReads(ma.PRINT_DOT_READS);
	g.setNodeStrlen(ma.NODE_STRLEN); g.setMaxTipLength(ma.MAX_TIP_LEN); g.setMaxIndelLen(ma.MAX_INDEL_LEN);
	g.setMinThreadReads(ma.MIN_THREAD_READS); g.setScaffoldContigs(ma.SCAFFOLD_CONTIGS);
	g.setInsertSize(ma.INSERT_SIZE); g.setInsertStdev(ma.INSERT_STDEV); g.setMaxMismatch(ma.MAX_MISMATCH);
	g.setFilters(&g_filters); g.setLRMode(false);
	g.setMaxUnitLen(ma.MAX_UNIT_LEN); g.setMinReportUnits(ma.MIN_REPORT_UNITS);
	g.setMinReportLen(ma.MIN_REPORT_LEN); g.setDistFromStr(ma.DIST_FROM_STR);

	char namebuf[32];
	for (int w = w0 + tid; w < w1; w += nthreads) {
		tl_window = (uint32_t)w;
		string raw = b->ref_seq.substr(b->ref_off[w], b->ref_off[w + 1] - b->ref_off[w]);
		// window skip rule of reference src/Microassembler.cc:799-800
		if (isNseq(raw)) { continue; }
		if (isRepeat(raw, ma.maxK)) { continue; }
		// Ref_t set-up of reference src/Lancet.cc:283-300
		Ref_t * ref = new Ref_t(ma.minK);
		std::ostringstream chr; chr << "chr" << b->chr_id[w];
		ref->refchr = chr.str();
		ref->refstart = b->ref_start[w];
		ref->refend = ref->refstart + (int)raw.size();
		std::ostringstream hdr; hdr << ref->refchr << ":" << ref->refstart << "-" << ref->refend;
		ref->setHdr(hdr.str()); ref->setSeq(raw); ref->setRawSeq(raw);
		reftable.clear();
		reftable.insert(make_pair(ref->hdr, ref));
		for (uint32_t x = b->wr_off[w]; x < b->wr_off[w + 1]; ++x) {
			uint32_t r = b->wr_idx[x];
			uint8_t fl = b->flags[r];
			int label = (fl & 1) ? NML : TMR;
			unsigned int strand = (fl & 2) ? REV : FWD;
			int mate = (fl >> 2) & 3;
			char code = (fl & 16) ? Graph_t::CODE_BASTARD : Graph_t::CODE_MAPPED;
			snprintf(namebuf, sizeof namebuf, "q%010u", b->name_rank[r]);
			uint64_t o0 = b->base_off[r], o1 = b->base_off[r + 1];
			g.addAlignment(label == TMR ? "tumor" : "normal", namebuf,
				b->seq.substr(o0, o1 - o0), b->qual.substr(o0, o1 - o0), mate, code, label, strand, "", 0);
		}
		if (b->wr_off[w + 1] == b->wr_off[w]) { delete ref; continue; }
		string h = ref->hdr; // processGraph deletes ref (Graph_t::clear(true), reference src/Graph.cc:52-57)
		int n = ma.processGraph(g, h, ma.minK, ma.maxK);
		if (n == 0) { g.clear(true); } // countMappedReads()<=0 early return (reference quirk B15): do not leak into the next window here
	}
	tl_out = NULL;
}

int main(int argc, char ** argv)
{
	Opts o;
	const char * fn = NULL;
	for (int i = 1; i < argc; ++i) {
		string a = argv[i];
		if      (a == "--min-k")   { o.minK = atoi(argv[++i]); }
		else if (a == "--max-k")   { o.maxK = atoi(argv[++i]); }
		else if (a == "--threads") { o.threads = atoi(argv[++i]); }
		else if (a == "--dfs-limit") { o.dfs_limit = atoi(argv[++i]); }
		else if (a == "--repeat")  { o.repeat = atoi(argv[++i]); }
		else if (a == "--first")   { o.first = atoi(argv[++i]); }
		else if (a == "--count")   { o.count = atoi(argv[++i]); }
		else if (a == "--out")     { o.out = argv[++i]; }
		else if (a == "--verbose") { o.verbose = true; }
		else { fn = argv[i]; }
	}
	if (!fn) { fprintf(stderr, "usage: ref_windows batch.lb2b [--threads T] [--min-k a] [--max-k b] [--out f.tsv] [--first i --count n] [--repeat R]\n"); return 2; }
	// filter defaults of reference src/Lancet.cc:627-638 (only used by the VCF printer, not on this path)
	g_filters.minPhredFisherSTR = 25; g_filters.minPhredFisher = 5; g_filters.minCovNormal = 10;
	g_filters.maxCovNormal = 1000000; g_filters.minCovTumor = 4; g_filters.maxCovTumor = 1000000;
	g_filters.minVafTumor = 0.04; g_filters.maxVafNormal = 0; g_filters.minAltCntTumor = 3;
	g_filters.maxAltCntNormal = 0; g_filters.minStrandBias = 1;

	Batch b; load(fn, b);
	int w0 = o.first, w1 = (o.count < 0) ? (int)b.n_windows : min((int)b.n_windows, o.first + o.count);
	int T = max(1, o.threads);
	vector<vector<Rec> > outs(T);
	double best = 1e30, total = 0;
	for (int rep = 0; rep < o.repeat; ++rep) {
		for (int t = 0; t < T; ++t) { outs[t].clear(); }
		auto t0 = chrono::steady_clock::now();
		vector<thread> th;
		for (int t = 0; t < T; ++t) { th.emplace_back(worker, &b, &o, t, T, w0, w1, &outs[t]); }
		for (auto & x : th) { x.join(); }
		double s = chrono::duration<double>(chrono::steady_clock::now() - t0).count();
		best = min(best, s); total += s;
	}
	if (o.out) {
		vector<Rec> all;
		for (int t = 0; t < T; ++t) { all.insert(all.end(), outs[t].begin(), outs[t].end()); }
		// stable by window, emission order inside a window is preserved (one window lives in one thread)
		stable_sort(all.begin(), all.end(), [](const Rec & a, const Rec & c) { return a.window < c.window; });
		FILE * f = strcmp(o.out, "-") ? fopen(o.out, "w") : stdout;
		for (auto & r : all) {
			fprintf(f, "%u\t%d\t%c\t%d\t%s\t%s\t%d\t%s\t%u,%u,%u,%u,%u,%u,%u,%u\t%c\t%c\n",
				r.window, r.pos, r.type, r.len, r.ref.c_str(), r.alt.c_str(), r.kmer,
				r.str.empty() ? "." : r.str.c_str(),
				r.c[0], r.c[1], r.c[2], r.c[3], r.c[4], r.c[5], r.c[6], r.c[7], r.pbr, r.pba);
		}
		if (f != stdout) { fclose(f); }
	}
	printf("{\"windows\": %d, \"threads\": %d, \"repeat\": %d, \"best_s\": %.6f, \"mean_s\": %.6f}\n",
		w1 - w0, T, o.repeat, best, total / o.repeat);
	return 0;
}
