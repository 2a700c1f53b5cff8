// lancet_cli.cc -- the `lancet` command-line surface on top of the C ABI (include/lancet_b200.h).
//
//   lancet_b200 --tumor T.bam --normal N.bam --ref ref.fa --reg chr:a-b | --bed f.bed [options] > out.vcf
//
// Host side of SURVEY.md §8(f): f2 window tiling (reference src/Lancet.cc:189-362), f3 read selection and the
// active-region prefilter (src/Microassembler.cc:255-655, src/util.cc:432-483) on BAM records decoded here
// (BGZF via zlib; the reference uses bamtools), f1 the variant store, filters, Fisher scores and the VCF writer
// (src/VariantDB.cc, src/Variant.{hh,cc}, src/FET.hh).  The micro-assembly itself (processGraph) runs on the GPU
// through lb2_process(); there is no CPU implementation of it in this program.
//
// The VCF is meant to be byte-identical to the reference's for the same --num-threads value (which only decides
// the replay order of VariantDB_t::addVar, reference src/Lancet.cc:305-310,943-959), except for the
// ##fileDate / ##cmdline header lines.
#include <algorithm>
#include <atomic>
#include <thread>
#include <chrono>
#include <mutex>
#include <condition_variable>
#include <deque>
#include <memory>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <fstream>
#include <getopt.h>
#include <iostream>
#include <limits>
#include <map>
#include <sstream>
#include <string>
#include <vector>
#include <zlib.h>
#include <fcntl.h>
#include <unistd.h>
#include <climits>
#include <sys/types.h>
#include <sys/wait.h>

#include "../../include/lancet_b200.h"

using std::string;
using std::vector;

// ------------------------------------------------------------------------------------------------------------
// options (defaults: reference src/Lancet.hh:33-79, filters src/Lancet.cc:627-638)
// ------------------------------------------------------------------------------------------------------------
struct Filters {
	double minPhredFisherSTR = 25, minPhredFisher = 5, maxVafNormal = 0, minVafTumor = 0.04;
	int minCovNormal = 10, maxCovNormal = 1000000, minCovTumor = 4, maxCovTumor = 1000000, minAltCntTumor = 3, maxAltCntNormal = 0, minStrandBias = 1;
};
struct Opts {
	string tumor, normal, ref, bed, reg, cmdline;
	int minK = 11, maxK = 101, tip_len = 11, cov_thr = 5, low_cov = 1, window = 600, padding = 250, max_avg_cov = 10000;
	int min_qv_trim = 10, min_qv_call = 17, qv_range = '!', min_map_qual = 15, dfs_limit = 1000000, num_threads = 1;
	int max_indel_len = 500, max_mismatch = 2, max_unit_len = 4, min_report_units = 3, min_report_len = 7, dist_from_str = 1;
	double cov_ratio = 0.01;
	bool primary_only = false, xa_filter = false, active_regions = true, verbose = false;
	int gpu = 0, gpus = 1, rank = -1, world = 1, batch_windows = 4096, io_threads = 0; string nccl_id_file;
	Filters f;
};

static string itos(int i) { std::stringstream s; s << i; return s.str(); }
static string dtos(double d) { std::stringstream s; s << d; return s.str(); }     // reference src/util.cc:89-94

// ------------------------------------------------------------------------------------------------------------
// SHA-256 (FIPS 180-4): only used as the ordering key of the variant store (reference src/VariantDB.cc:29)
// ------------------------------------------------------------------------------------------------------------
static string sha256_hex(const string &in)
{
	static const uint32_t K[64] = {
		0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01, 0x243185be, 0x550c7dc3,
		0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da,
		0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967, 0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13,
		0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85, 0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070,
		0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208,
		0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2 };
	uint32_t h[8] = { 0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19 };
	vector<uint8_t> m(in.begin(), in.end());
	uint64_t bitlen = (uint64_t)m.size() * 8;
	m.push_back(0x80); while (m.size() % 64 != 56) { m.push_back(0); }
	for (int i = 7; i >= 0; --i) { m.push_back((uint8_t)(bitlen >> (8 * i))); }
	auto rotr = [](uint32_t x, int n) { return (x >> n) | (x << (32 - n)); };
	for (size_t o = 0; o < m.size(); o += 64) {
		uint32_t w[64];
		for (int i = 0; i < 16; ++i) { w[i] = (uint32_t)m[o + 4 * i] << 24 | (uint32_t)m[o + 4 * i + 1] << 16 | (uint32_t)m[o + 4 * i + 2] << 8 | m[o + 4 * i + 3]; }
		for (int i = 16; i < 64; ++i) {
			uint32_t s0 = rotr(w[i - 15], 7) ^ rotr(w[i - 15], 18) ^ (w[i - 15] >> 3), s1 = rotr(w[i - 2], 17) ^ rotr(w[i - 2], 19) ^ (w[i - 2] >> 10);
			w[i] = w[i - 16] + s0 + w[i - 7] + s1;
		}
		uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
		for (int i = 0; i < 64; ++i) {
			uint32_t S1 = rotr(e, 6) ^ rotr(e, 11) ^ rotr(e, 25), ch = (e & f) ^ (~e & g), t1 = hh + S1 + ch + K[i] + w[i];
			uint32_t S0 = rotr(a, 2) ^ rotr(a, 13) ^ rotr(a, 22), mj = (a & b) ^ (a & c) ^ (b & c), t2 = S0 + mj;
			hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
		}
		h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
	}
	char buf[65];
	for (int i = 0; i < 8; ++i) { snprintf(buf + 8 * i, 9, "%08x", h[i]); }
	return string(buf, 64);
}

// ------------------------------------------------------------------------------------------------------------
// Fisher exact test, point probability (reference src/FET.hh:43-127 = htslib's kt_fisher_exact)
// ------------------------------------------------------------------------------------------------------------
struct FET {
	static double lbinom(int n, int k) { if (k == 0 || n == k) return 0; return lgamma(n + 1) - lgamma(k + 1) - lgamma(n - k + 1); }
	static double hypergeo(int n11, int n1_, int n_1, int n) { return exp(lbinom(n1_, n11) + lbinom(n - n1_, n_1 - n11) - lbinom(n, n_1)); }
	// The only number the VCF writer uses is the point probability of the observed table (the reference's routine also
	// accumulates the two tails and then discards them): 1 for a degenerate table, else the hypergeometric term.
	static double point(int n11, int n12, int n21, int n22) {
		const int row1 = n11 + n12, col1 = n11 + n21, total = row1 + n21 + n22;
		const int hi = std::min(row1, col1), lo = std::max(0, row1 + col1 - total);
		if (lo == hi) { return 1.0; }
		return hypergeo(n11, row1, col1, total);
	}
};

// ------------------------------------------------------------------------------------------------------------
// Variant_t / VariantDB_t (reference src/Variant.hh:106-180, src/Variant.cc:39-347, src/VariantDB.cc:28-179)
// ------------------------------------------------------------------------------------------------------------
struct Variant {
	unsigned short kmer = 0; string chr; int pos = 0; char type = '?'; unsigned short len = 0; string ref, alt, str;
	unsigned short rnf = 0, rnr = 0, rtf = 0, rtr = 0, anf = 0, anr = 0, atf = 0, atr = 0;
	char pbr = 0, pba = 0;

	Variant(const string &chr_, int pos_, string ref_, string alt_, const lb2_variant &v, const string &str_) {
		kmer = v.kmer; str = str_; chr = chr_; pos = pos_; char code = (char)v.code; pbr = (char)v.prev_bp_ref; pba = (char)v.prev_bp_alt;
		if (code == '^') { type = 'I'; ref_ = ""; len = alt_.length(); }
		if (code == 'v') { type = 'D'; alt_ = ""; len = ref_.length(); }
		if (code == 'x') { type = 'S'; pos++; }
		if (code == 'c') {
			type = 'C';
			ref_.erase(std::remove(ref_.begin(), ref_.end(), '-'), ref_.end()); alt_.erase(std::remove(alt_.begin(), alt_.end(), '-'), alt_.end());
			unsigned short rl = ref_.length(), al = alt_.length();
			if (rl == al) { len = al; } else if (rl > al) { len = rl - al; } else { len = al - rl; }
		}
		if (type != 'S') { ref = pba + ref_; alt = pba + alt_; } else { alt = alt_; ref = ref_; len = 1; }
		rnf = v.rcn_fwd; rnr = v.rcn_rev; rtf = v.rct_fwd; rtr = v.rct_rev; anf = v.acn_fwd; anr = v.acn_rev; atf = v.act_fwd; atr = v.act_rev;
	}
	string signature() const { return chr + ":" + itos(pos) + ":" + type + ":" + itos(len) + ":" + ref + ":" + alt; }
	static string genotype(int R, int A) { if (R > 0 && A > 0) return "0/1"; if (R > 0 && A == 0) return "0/0"; if (R == 0 && A > 0) return "1/1"; return "."; }
	static double score(double prob, bool fet) {
		if (prob == 1.0) return 0.0;
		if (fet && prob == 0.0) return -10.0 * log10(1 / std::numeric_limits<double>::max());
		return -10.0 * log10(prob);
	}
	string vcf(const Filters &fs) const {
		int trt = rtf + rtr, tat = atf + atr, trn = rnf + rnr, tan = anf + anr;
		double fet = score(FET::point(trn, trt, tan, tat), true);
		double sb = score(FET::point(rtf, rtr, atf, atr), false);
		string status;
		if (tan > 0 && tat > 0) status = "SHARED"; else if (tan == 0 && tat > 0) status = "SOMATIC"; else if (tan > 0 && tat == 0) status = "NORMAL"; else return "";
		string INFO = status + ";FETS=" + dtos(fet);
		if (type == 'I') INFO += ";TYPE=ins"; if (type == 'D') INFO += ";TYPE=del"; if (type == 'S') INFO += ";TYPE=snv"; if (type == 'C') INFO += ";TYPE=complex";
		INFO += ";LEN=" + itos(len) + ";KMERSIZE=" + itos(kmer) + ";SB=" + dtos(sb);
		if (!str.empty()) INFO += ";MS=" + str;
		int tcov = trt + tat, ncov = trn + tan;
		double tvaf = (tcov == 0) ? 0 : ((double)tat / (double)tcov), nvaf = (ncov == 0) ? 0 : ((double)tan / (double)ncov);
		string F;
		auto add = [&](const char *n) { if (F.empty()) F = n; else { F += ";"; F += n; } };
		if (!str.empty()) { if (fet < fs.minPhredFisherSTR) add("LowFisherSTR"); } else if (fet < fs.minPhredFisher) add("LowFisherScore");
		if (ncov < fs.minCovNormal) add("LowCovNormal"); if (ncov > fs.maxCovNormal) add("HighCovNormal");
		if (tcov < fs.minCovTumor) add("LowCovTumor"); if (tcov > fs.maxCovTumor) add("HighCovTumor");
		if (tvaf < fs.minVafTumor) add("LowVafTumor"); if (nvaf > fs.maxVafNormal) add("HighVafNormal");
		if (tat < fs.minAltCntTumor) add("LowAltCntTumor"); if (tan > fs.maxAltCntNormal) add("HighAltCntNormal");
		if (atf < fs.minStrandBias || atr < fs.minStrandBias) add("StrandBias");
		if (F.empty()) F = "PASS";
		string N = genotype(trn, tan) + ":" + itos(trn) + "," + itos(tan) + ":" + itos(rnf) + "," + itos(rnr) + ":" + itos(anf) + "," + itos(anr) + ":" + itos(trn + tan);
		string T = genotype(trt, tat) + ":" + itos(trt) + "," + itos(tat) + ":" + itos(rtf) + "," + itos(rtr) + ":" + itos(atf) + "," + itos(atr) + ":" + itos(trt + tat);
		std::stringstream l;
		l << chr << "\t" << pos << "\t.\t" << ref << "\t" << alt << "\t" << fet << "\t" << F << "\t" << INFO << "\tGT:AD:SR:SA:DP\t" << N << "\t" << T << std::endl;
		return l.str();
	}
};

struct VariantDB {
	std::map<string, Variant> DB;
	void add(const Variant &v) {      // keep the record with the strictly larger total coverage (src/VariantDB.cc:28-91)
		string key = sha256_hex(v.signature());
		auto it = DB.find(key);
		if (it != DB.end()) {
			Variant &o = it->second;
			int oc = o.rnf + o.rnr + o.rtf + o.rtr + o.anf + o.anr + o.atf + o.atr, nc = v.rnf + v.rnr + v.rtf + v.rtr + v.anf + v.anr + v.atf + v.atr;
			if (oc < nc) { o.kmer = v.kmer; o.rnf = v.rnf; o.rnr = v.rnr; o.rtf = v.rtf; o.rtr = v.rtr; o.anf = v.anf; o.anr = v.anr; o.atf = v.atf; o.atr = v.atr; }
		} else { DB.insert(std::make_pair(key, v)); }
	}
};
struct byPos {    // same comparator semantics as reference src/VariantDB.hh:37-52 (the sort is not stable: B13)
	bool operator()(const std::pair<string, Variant> &a, const std::pair<string, Variant> &b) const {
		int c = a.second.chr.compare(b.second.chr);
		if (c == 0) return a.second.pos < b.second.pos;
		return c < 0;
	}
};

static void print_header(const Opts &o, const string &sn, const string &st)
{
	time_t raw; time(&raw); const Filters &fs = o.f;
	std::stringstream h;
	h << "##fileformat=VCFv4.2\n##fileDate=" << ctime(&raw) << "##source=lancet 1.1.0, October 18 2019\n##cmdline=" << o.cmdline << "\n##reference=" << o.ref << "\n"
	  "##INFO=<ID=FETS,Number=1,Type=Float,Description=\"Phred-scaled p-value of the Fisher's exact test for tumor-normal allele counts\">\n"
	  "##INFO=<ID=SOMATIC,Number=0,Type=Flag,Description=\"Somatic mutation\">\n"
	  "##INFO=<ID=SHARED,Number=0,Type=Flag,Description=\"Shared mutation betweem tumor and normal\">\n"
	  "##INFO=<ID=NORMAL,Number=0,Type=Flag,Description=\"Mutation present only in the normal\">\n"
	  "##INFO=<ID=NONE,Number=0,Type=Flag,Description=\"Mutation not supported by data\">\n"
	  "##INFO=<ID=KMERSIZE,Number=1,Type=Integer,Description=\"K-mer size used to assemble the locus\">\n"
	  "##INFO=<ID=SB,Number=1,Type=Float,Description=\"Strand bias score: phred-scaled p-value of the Fisher's exact test for the forward/reverse read counts in the tumor\">\n"
	  "##INFO=<ID=MS,Number=1,Type=String,Description=\"Microsatellite mutation (format: #LEN#MOTIF)\">\n"
	  "##INFO=<ID=LEN,Number=1,Type=Integer,Description=\"Variant size in base pairs\">\n"
	  "##INFO=<ID=TYPE,Number=1,Type=String,Description=\"Variant type (snv, del, ins, complex)\">\n";
	h << "##FILTER=<ID=LowCovNormal,Description=\"Low coverage in the normal (<" << fs.minCovNormal << ")\">\n"
	  "##FILTER=<ID=HighCovNormal,Description=\"High coverage in the normal (>" << fs.maxCovNormal << ")\">\n"
	  "##FILTER=<ID=LowCovTumor,Description=\"Low coverage in the tumor (<" << fs.minCovTumor << ")\">\n"
	  "##FILTER=<ID=HighCovTumor,Description=\"High coverage in the tumor (>" << fs.maxCovTumor << ")\">\n"
	  "##FILTER=<ID=LowVafTumor,Description=\"Low variant allele frequency in the tumor (<" << fs.minVafTumor << ")\">\n"
	  "##FILTER=<ID=HighVafNormal,Description=\"High variant allele frequency in the normal (>" << fs.maxVafNormal << ")\">\n"
	  "##FILTER=<ID=LowAltCntTumor,Description=\"Low alternative allele count in the tumor (<" << fs.minAltCntTumor << ")\">\n"
	  "##FILTER=<ID=HighAltCntNormal,Description=\"High alternative allele count in the normal (>" << fs.maxAltCntNormal << ")\">\n"
	  "##FILTER=<ID=LowFisherScore,Description=\"Low Fisher's exact test score for tumor-normal allele counts (<" << fs.minPhredFisher << ")\">\n"
	  "##FILTER=<ID=LowFisherSTR,Description=\"Low Fisher's exact test score for tumor-normal STR allele counts (<" << fs.minPhredFisherSTR << ")\">\n"
	  "##FILTER=<ID=StrandBias,Description=\"Strand bias: # of non-reference reads in either forward or reverse strand below threshold (<" << fs.minStrandBias << ")\">\n"
	  "##FILTER=<ID=STR,Description=\"Microsatellite mutation\">\n";
	h << "##FORMAT=<ID=GT,Number=1,Type=String,Description=\"Genotype\">\n"
	  "##FORMAT=<ID=DP,Number=1,Type=Integer,Description=\"Depth\">\n"
	  "##FORMAT=<ID=AD,Number=.,Type=Integer,Description=\"Allele depth: # of supporting ref,alt reads at the site\">\n"
	  "##FORMAT=<ID=SR,Number=.,Type=Integer,Description=\"Strand counts for ref: # of supporting forward,reverse reads for reference allele\">\n"
	  "##FORMAT=<ID=SA,Number=.,Type=Integer,Description=\"Strand counts for alt: # of supporting forward,reverse reads for alterantive allele\">\n";
	h << "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t" << sn << "\t" << st << "\n";
	std::cout << h.str();
}

// ------------------------------------------------------------------------------------------------------------
// BAM access (the reference uses bamtools' SetRegion per window, src/Microassembler.cc:802-834; here: BGZF + BAI read
// directly).  A region fetch looks the first block up in the BAI linear index, reads compressed blocks from there,
// inflates them on several threads, cuts the stream into records and decodes those on several threads, and stops at
// the first record that starts at or beyond the region's end.  Without an index the file is scanned from its first record.
// ------------------------------------------------------------------------------------------------------------
struct Aln {
	int32_t pos = 0, end = 0; uint16_t flag = 0; uint8_t mapq = 0; int32_t l_seq = 0;
	string name, seq, qual, md, xt, xa;
	vector<uint32_t> cigar;                // BAM encoding: len << 4 | op
	float as = -1, xs = -1; bool has_md = false;
	mutable vector<int32_t> ev; mutable bool ev_done = false;      // isActiveRegion evidence of this read, (position << 2 | kind), made on first use (a read lies in ~6 windows)
};

static double wall() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
template <class F> static void parallel_for(size_t n, unsigned threads, F f)
{
	if (threads <= 1 || n < 2) { for (size_t i = 0; i < n; ++i) { f(i); } return; }
	std::atomic<size_t> next(0); vector<std::thread> th; const unsigned T = (unsigned)std::min<size_t>(threads, n);
	for (unsigned t = 0; t < T; ++t) { th.emplace_back([&]() { while (true) { size_t i = next.fetch_add(1); if (i >= n) { break; } f(i); } }); }
	for (auto &x : th) { x.join(); }
}

static void decode_record(const uint8_t *r, int32_t bs, Aln &a)
{
	static const char *NT = "=ACMGRSVTWYHKDBN";
	auto i32 = [&](size_t o) { int32_t v; memcpy(&v, r + o, 4); return v; };
	const int32_t pos = i32(4); const uint8_t l_name = r[8], mapq = r[9];
	const uint16_t n_cig = r[12] | (r[13] << 8), flag = r[14] | (r[15] << 8); const int32_t l_seq = i32(16);
	size_t q = 32; const char *name = (const char *)r + q; q += l_name;
	const size_t cig = q; q += 4 * (size_t)n_cig; const size_t sq = q; q += (l_seq + 1) / 2; const size_t ql = q; q += l_seq; size_t t = q; const size_t tend = (size_t)bs;
	a.pos = pos; a.flag = flag; a.mapq = mapq; a.l_seq = l_seq; a.name = name;
	while (t + 3 <= tend) {
		char t0 = r[t], t1 = r[t + 1], ty = r[t + 2]; t += 3; double num = 0; bool isnum = false; string sv;
		switch (ty) {
			case 'A': {      // bamtools GetTag<std::string> does strlen() on the value: a char tag runs on into the next tag's bytes up to a NUL
				size_t e = t; while (e < tend && r[e]) { ++e; } sv = string((const char *)r + t, e - t); t += 1; break; }
			case 'c': num = (int8_t)r[t]; isnum = true; t += 1; break; case 'C': num = r[t]; isnum = true; t += 1; break;
			case 's': { int16_t v; memcpy(&v, r + t, 2); num = v; isnum = true; t += 2; break; }
			case 'S': { uint16_t v; memcpy(&v, r + t, 2); num = v; isnum = true; t += 2; break; }
			case 'i': { int32_t v; memcpy(&v, r + t, 4); num = v; isnum = true; t += 4; break; }
			case 'I': { uint32_t v; memcpy(&v, r + t, 4); num = v; isnum = true; t += 4; break; }
			case 'f': { float v; memcpy(&v, r + t, 4); num = v; isnum = true; t += 4; break; }
			case 'Z': case 'H': { size_t e = t; while (e < tend && r[e]) { ++e; } sv = string((const char *)r + t, e - t); t = e + 1; break; }
			case 'B': { char st = r[t]; int32_t cnt; memcpy(&cnt, r + t + 1, 4); int es = (st == 'c' || st == 'C') ? 1 : (st == 's' || st == 'S') ? 2 : 4; t += 5 + (size_t)es * cnt; break; }
			default: t = tend; break;
		}
		if (t0 == 'M' && t1 == 'D') { a.has_md = true; a.md = sv; }
		else if (t0 == 'A' && t1 == 'S' && isnum) { a.as = (float)num; } else if (t0 == 'X' && t1 == 'S' && isnum) { a.xs = (float)num; }
		else if (t0 == 'X' && t1 == 'T') { a.xt = sv; } else if (t0 == 'X' && t1 == 'A') { a.xa = sv; }
	}
	a.cigar.resize(n_cig); int32_t end = pos;
	for (int c = 0; c < n_cig; ++c) {
		uint32_t v; memcpy(&v, r + cig + 4 * c, 4); a.cigar[c] = v; int op = v & 15;
		if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) { end += (int32_t)(v >> 4); }      // M D N = X (BamAlignment::GetEndPosition)
	}
	a.end = end;
	a.seq.resize(l_seq); a.qual.resize(l_seq);
	for (int i = 0; i < l_seq; ++i) { uint8_t by = r[sq + i / 2]; a.seq[i] = NT[(i & 1) ? (by & 15) : (by >> 4)]; a.qual[i] = (char)(r[ql + i] + 33); }
	if (l_seq && r[ql] == 0xFF) { a.qual.clear(); }
}

struct BamFile {
	string path; FILE *f = nullptr; uint64_t file_size = 0;
	vector<string> ref_names; vector<int32_t> ref_lens; string sample = "NA"; bool first_has_md = true;
	uint64_t first_voff = 0;                       // virtual offset of the first alignment record
	bool has_index = false; vector<vector<uint64_t>> linear; vector<uint64_t> ref_first;      // BAI: linear index and smallest chunk start per reference
	unsigned threads = 1;
	double t_read = 0, t_inflate = 0, t_cut = 0, t_decode = 0; std::mutex t_mu;      // (LB2_CLI_TIMING)
	~BamFile() { if (f) { fclose(f); } }

	// compressed blocks [coff, ...) up to ~want bytes: (file offset, block size, inflated size) of every complete block read
	struct Blk { uint64_t coff; uint32_t bsize, isize; size_t in_off, out_off; };
	bool read_blocks(uint64_t coff, size_t want, vector<uint8_t> &in, vector<Blk> &blks, bool &eof) {
		blks.clear(); eof = false;
		if (coff >= file_size) { eof = true; return true; }
		const size_t n = (size_t)std::min<uint64_t>(want + (1u << 16), file_size - coff);
		in.resize(n);
		{ size_t got = 0; while (got < n) { const ssize_t r_ = pread(fileno(f), in.data() + got, n - got, (off_t)(coff + got)); if (r_ <= 0) { return false; } got += (size_t)r_; } }      // (pread: several batch builders read the file at once)
		size_t p = 0, out = 0;
		while (p + 18 <= n) {
			if (in[p] != 31 || in[p + 1] != 139) { return false; }
			const uint16_t xlen = in[p + 10] | (in[p + 11] << 8); uint32_t bsize = 0; size_t x = p + 12, xe = x + xlen;
			if (xe > n) { break; }
			while (x + 4 <= xe) { uint16_t sl = in[x + 2] | (in[x + 3] << 8); if (in[x] == 'B' && in[x + 1] == 'C') { bsize = (in[x + 4] | (in[x + 5] << 8)) + 1u; } x += 4 + sl; }
			if (!bsize) { return false; }
			if (p + bsize > n) { break; }
			Blk b; b.coff = coff + p; b.bsize = bsize; b.in_off = p; b.out_off = out;
			b.isize = in[p + bsize - 4] | (in[p + bsize - 3] << 8) | (in[p + bsize - 2] << 16) | ((uint32_t)in[p + bsize - 1] << 24);
			blks.push_back(b); out += b.isize; p += bsize;
		}
		if (coff + p >= file_size) { eof = true; }
		return !blks.empty() || eof;
	}
	bool inflate_blocks(const vector<uint8_t> &in, const vector<Blk> &blks, vector<uint8_t> &out, size_t keep) {
		const size_t total = blks.empty() ? 0 : blks.back().out_off + blks.back().isize;
		out.resize(keep + total);
		std::atomic<int> bad(0);
		parallel_for(blks.size(), threads, [&](size_t i) {
			const Blk &b = blks[i]; if (!b.isize) { return; }
			const uint16_t xlen = in[b.in_off + 10] | (in[b.in_off + 11] << 8);
			z_stream zs; memset(&zs, 0, sizeof zs);
			if (inflateInit2(&zs, -15) != Z_OK) { bad = 1; return; }
			zs.next_in = (Bytef *)in.data() + b.in_off + 12 + xlen; zs.avail_in = (uInt)(b.bsize - xlen - 19); zs.next_out = out.data() + keep + b.out_off; zs.avail_out = b.isize;
			const int rc = inflate(&zs, Z_FINISH); inflateEnd(&zs);
			if (rc != Z_STREAM_END) { bad = 1; }
		});
		return !bad;
	}

	bool open(const string &p, unsigned nthreads) {
		path = p; threads = std::max(1u, nthreads);
		f = fopen(p.c_str(), "rb"); if (!f) { return false; }
		fseeko(f, 0, SEEK_END); file_size = (uint64_t)ftello(f);
		// header: text, reference dictionary; then the first record (MD tag check, reference src/Lancet.cc:817-825)
		vector<uint8_t> in, d; vector<Blk> blks; bool eof = false; uint64_t coff = 0; vector<uint64_t> blk_coff, blk_end;      // inflated end offset of every block read
		auto more = [&]() -> bool {
			if (eof) { return false; }
			if (!read_blocks(coff, 1u << 20, in, blks, eof) || blks.empty()) { return false; }
			const size_t keep = d.size();
			if (!inflate_blocks(in, blks, d, keep)) { return false; }
			for (auto &b : blks) { blk_coff.push_back(b.coff); blk_end.push_back(keep + b.out_off + b.isize); }
			coff = blks.back().coff + blks.back().bsize;
			return true;
		};
		auto need = [&](size_t upto) -> bool { while (d.size() < upto) { if (!more()) { return false; } } return true; };
		auto i32 = [&](size_t o) { int32_t v; memcpy(&v, d.data() + o, 4); return v; };
		if (!need(12) || memcmp(d.data(), "BAM\1", 4)) { return false; }
		size_t q = 4; const int32_t l_text = i32(q); q += 4; if (!need(q + l_text + 4)) { return false; }
		const string text((const char *)d.data() + q, l_text); q += l_text;
		{	// first @RG line with an SM: field (reference retriveSampleName, src/Microassembler.cc:52-67)
			std::istringstream is(text); string line;
			while (std::getline(is, line)) {
				if (line.compare(0, 3, "@RG") == 0) {
					size_t x = line.find("\tSM:");
					if (x != string::npos) { size_t e = line.find('\t', x + 1); sample = line.substr(x + 4, e == string::npos ? string::npos : e - x - 4); }
					break;
				}
			}
		}
		const int32_t n_ref = i32(q); q += 4;
		for (int r = 0; r < n_ref; ++r) {
			if (!need(q + 4)) { return false; } const int32_t l = i32(q); q += 4;
			if (!need(q + l + 4)) { return false; }
			ref_names.push_back(string((const char *)d.data() + q, l - 1)); q += l; ref_lens.push_back(i32(q)); q += 4;
		}
		// virtual offset of the first record = (start of the block holding inflated offset q, offset inside it)
		{
			if (!need(q + 1) && d.size() < q) { return false; }
			size_t bi = 0; while (bi < blk_end.size() && blk_end[bi] <= q) { ++bi; }
			if (bi < blk_end.size()) { const size_t bstart = bi ? blk_end[bi - 1] : 0; first_voff = (blk_coff[bi] << 16) | (uint64_t)(q - bstart); }
			else { first_voff = coff << 16; }
		}
		if (need(q + 4)) { const int32_t bs = i32(q); if (bs >= 32 && need(q + 4 + bs)) { Aln a; decode_record(d.data() + q + 4, bs, a); first_has_md = a.has_md; } }
		load_index();
		return true;
	}

	void load_index() {
		string cand[2] = { path + ".bai", path.size() > 4 ? path.substr(0, path.size() - 4) + ".bai" : string() };
		for (auto &ip : cand) {
			if (ip.empty()) { continue; }
			FILE *g = fopen(ip.c_str(), "rb"); if (!g) { continue; }
			vector<uint8_t> b; uint8_t buf[1 << 16]; size_t n; while ((n = fread(buf, 1, sizeof buf, g)) > 0) { b.insert(b.end(), buf, buf + n); } fclose(g);
			size_t q = 0; auto u32 = [&]() { uint32_t v = 0; if (q + 4 <= b.size()) { memcpy(&v, b.data() + q, 4); } q += 4; return v; };
			auto u64 = [&]() { uint64_t v = 0; if (q + 8 <= b.size()) { memcpy(&v, b.data() + q, 8); } q += 8; return v; };
			if (b.size() < 8 || memcmp(b.data(), "BAI\1", 4)) { continue; }
			q = 4; const uint32_t n_ref = u32(); linear.assign(n_ref, {}); ref_first.assign(n_ref, ~0ull);
			for (uint32_t r = 0; r < n_ref && q <= b.size(); ++r) {
				const uint32_t n_bin = u32();
				for (uint32_t i = 0; i < n_bin; ++i) {
					const uint32_t bin = u32(), n_chunk = u32();
					for (uint32_t c = 0; c < n_chunk; ++c) { const uint64_t cb = u64(); u64(); if (bin != 37450 && cb < ref_first[r]) { ref_first[r] = cb; } }      // (bin 37450: metadata pseudo-bin)
				}
				const uint32_t n_intv = u32(); linear[r].resize(n_intv);
				for (uint32_t i = 0; i < n_intv; ++i) { linear[r][i] = u64(); }
			}
			has_index = q <= b.size();
			if (has_index) { return; }
			linear.clear(); ref_first.clear();
		}
	}

	// all records of reference refID whose position lies in [beg, end), in file order
	bool fetch(int refID, int beg, int end, vector<Aln> &out) {
		out.clear();
		if (refID < 0 || refID >= (int)ref_names.size() || end <= beg) { return true; }
		uint64_t voff = first_voff;
		if (has_index && refID < (int)linear.size()) {
			if (ref_first[refID] == ~0ull) { return true; }            // no alignment on this reference
			voff = ref_first[refID];
			const vector<uint64_t> &li = linear[refID];
			if (!li.empty()) { size_t i = std::min<size_t>((size_t)std::max(beg, 0) >> 14, li.size() - 1); while (i > 0 && li[i] == 0) { --i; } if (li[i] > voff) { voff = li[i]; } }
		}
		vector<uint8_t> in, d; vector<Blk> blks; bool eof = false; uint64_t coff = voff >> 16; size_t skip = (size_t)(voff & 0xFFFF);
		vector<uint8_t> carry;
		bool done = false; size_t want = 4u << 20;
		// (the linear index also bounds the far end: the first block with an alignment in the 16 kb bin after `end` --
		// read and inflate up to there, not a fixed 4 MiB; the loop below carries on if that was short)
		if (has_index && refID < (int)linear.size() && !linear[refID].empty()) {
			const vector<uint64_t> &li = linear[refID]; size_t j = ((size_t)std::max(end, 0) >> 14) + 1;
			while (j < li.size() && li[j] == 0) { ++j; }
			if (j < li.size() && (li[j] >> 16) >= (voff >> 16)) { want = (size_t)((li[j] >> 16) - (voff >> 16)) + (1u << 16); }
		}
		while (!done) {
			const double f0 = wall();
			if (!read_blocks(coff, want, in, blks, eof)) { return false; }
			if (blks.empty()) { break; }
			d = carry; const size_t keep = d.size();
			const double f1 = wall();
			if (!inflate_blocks(in, blks, d, keep)) { return false; }
			const double f2 = wall();
			coff = blks.back().coff + blks.back().bsize;
			// cut into records (sequential: every record names its own size), remember the ones of this region
			size_t q = skip; skip = 0; vector<std::pair<size_t, int32_t>> recs;
			while (q + 4 <= d.size()) {
				int32_t bs; memcpy(&bs, d.data() + q, 4);
				if (bs < 32) { return false; }
				if (q + 4 + (size_t)bs > d.size()) { break; }
				int32_t rid, pos; memcpy(&rid, d.data() + q + 4, 4); memcpy(&pos, d.data() + q + 8, 4);
				if (rid < 0 || rid > refID || (rid == refID && pos >= end)) { done = true; break; }      // (coordinate-sorted: unplaced reads come last)
				if (rid == refID && pos >= beg) { recs.push_back(std::make_pair(q + 4, bs)); }
				q += 4 + (size_t)bs;
			}
			const double f3 = wall();
			const size_t o0 = out.size(); out.resize(o0 + recs.size());
			parallel_for((recs.size() + 255) / 256, threads, [&](size_t c) { for (size_t i = c * 256; i < std::min(recs.size(), (c + 1) * 256); ++i) { decode_record(d.data() + recs[i].first, recs[i].second, out[o0 + i]); } });
			{ const double f4 = wall(); std::lock_guard<std::mutex> lk(t_mu); t_read += f1 - f0; t_inflate += f2 - f1; t_cut += f3 - f2; t_decode += f4 - f3; }
			carry.assign(d.begin() + (done ? d.size() : q), d.end());
			if (eof) { break; }
			want = std::min<size_t>(want * 2, 64u << 20);
		}
		return true;
	}
};

// ------------------------------------------------------------------------------------------------------------
// FASTA fetch (reference uses htslib faidx; src/Lancet.cc:245-263): 1-based inclusive region, upper-cased, IUPAC -> N
// ------------------------------------------------------------------------------------------------------------
static bool fasta_fetch(const string &path, const string &chr, int start1, int end1, string &out)
{
	std::ifstream f(path); if (!f) { return false; }
	string line; bool in = false; long posn = 0; out.clear();
	while (std::getline(f, line)) {
		if (!line.empty() && line[0] == '>') { if (in) { break; } string nm = line.substr(1, line.find_first_of(" \t") - 1); in = (nm == chr); posn = 0; continue; }
		if (!in) { continue; }
		if (!line.empty() && line.back() == '\r') { line.pop_back(); }
		long l0 = posn + 1, l1 = posn + (long)line.size();
		if (l1 >= start1 && l0 <= end1) { long a = std::max<long>(start1, l0), b = std::min<long>(end1, l1); out += line.substr(a - l0, b - a + 1); }
		posn = l1; if (posn >= end1) { break; }
	}
	for (auto &c : out) {
		c = (char)toupper(c);
		if (strchr("MRWSYKVHDBX", c)) { c = 'N'; }
	}
	return true;
}

// ------------------------------------------------------------------------------------------------------------
// windows (reference loadRefs / loadBed, src/Lancet.cc:189-362)
// ------------------------------------------------------------------------------------------------------------
struct Window { string chr, hdr, raw; int refstart, refend; int thread; };

static int load_refs(const Opts &o, const BamFile &bam, const string &region, vector<Window> &wins, int thread, int &num_windows)
{
	string CHR, START, END;
	size_t x = region.find_first_of(':');
	if (x == string::npos && region.length() > 0) {
		CHR = region; START = "1"; bool found = false;
		for (size_t i = 0; i < bam.ref_names.size(); ++i) { if (bam.ref_names[i] == CHR) { END = itos(bam.ref_lens[i]); found = true; break; } }
		if (!found) { std::cerr << "ERROR: chromosome label " << CHR << " not found in BAM header!" << std::endl; }
	} else {
		size_t y = region.find_first_of('-', x);
		CHR = region.substr(0, x); START = region.substr(x + 1, y - x - 1); END = region.substr(y + 1);
		int SP = std::stoi(START) - o.padding, EP = std::stoi(END) + o.padding; if (SP < 1) { SP = 1; }
		for (size_t i = 0; i < bam.ref_names.size(); ++i) { if (bam.ref_names[i] == CHR) { if (EP > bam.ref_lens[i]) { EP = bam.ref_lens[i]; } break; } }
		START = itos(SP); END = itos(EP);
	}
	string s;
	if (!fasta_fetch(o.ref, CHR, atoi(START.c_str()), atoi(END.c_str()), s)) { std::cerr << "Could not load " << o.ref << std::endl; exit(1); }
	int end = (int)s.length(), offset = 0, T = thread;
	for (; offset < end; offset += 100) {
		int LEN = o.window;
		if (offset + o.window >= (int)s.length()) { LEN = (int)s.length() - offset - 1; end = offset; }
		Window w; w.chr = CHR; w.raw = s.substr(offset, LEN); w.refstart = atoi(START.c_str()) + offset; w.refend = w.refstart + LEN;
		w.hdr = CHR + ":" + itos(w.refstart) + "-" + itos(w.refend); w.thread = T;
		wins.push_back(w);
		++num_windows; ++T; if ((num_windows % o.num_threads) == 0) { T = 0; }
	}
	return T;
}

// ------------------------------------------------------------------------------------------------------------
// isActiveRegion (src/Microassembler.cc:255-432) and parseMD (src/util.cc:432-483) on decoded records
// ------------------------------------------------------------------------------------------------------------
// The reference counts, per window, the reads that show a mismatch (MD tag, base quality permitting, or an X operation),
// an insertion, a deletion or a soft clip at each position (four std::map<int,int>) and calls the window active when some
// count reaches MIN_EVIDENCE.  What one read contributes does not depend on the window: it is listed once per read
// (kind 0 mismatch, 1 insertion, 2 deletion, 3 soft clip), the window then only merges the lists of its reads.
static void parse_md(const string &md, vector<int32_t> &ev, int start, const string &qual, int min_qv)
{
	const string valid = "acgtumrwsykvhdbxnACGTUMRWSYKVHDBXN^";
	size_t p = md.find_first_of(valid), p_old = (size_t)-1, p2; int pos = start; size_t rpos = 0;
	while (p != string::npos) {
		int step = atoi(md.substr(p_old + 1, p - (p_old + 1)).c_str()); pos += step; rpos += step;
		if (md[p] == '^') {
			p2 = md.find_first_not_of(valid, p + 1);
			string del = md.substr(p + 1, p2 - (p + 1)); pos += (int)del.size();
			p = md.find_first_of(valid, p2); p_old = p2 - 1;
		} else {
			++pos; ++rpos;
			char q = (rpos < qual.length()) ? qual[rpos] : (char)0;          // qual[len] is the terminating NUL of the reference's std::string
			if (q >= min_qv) { ev.push_back(pos * 4 + 0); }
			p_old = p; p = md.find_first_of(valid, p_old + 1);
		}
	}
}
static void read_evidence(const Aln &al, int min_qv)
{
	al.ev.clear(); al.ev_done = true;
	if (al.has_md) { parse_md(al.md, al.ev, al.pos, al.qual, min_qv); }
	int pos = al.pos;
	for (uint32_t c : al.cigar) {
		int op = c & 15; int len = (int)(c >> 4);
		if (op != 1) { pos += len; }                 // every operation except 'I' advances (literal, src/Microassembler.cc:321)
		if (op == 8) { al.ev.push_back(pos * 4 + 0); } if (op == 1) { al.ev.push_back(pos * 4 + 1); } if (op == 2) { al.ev.push_back(pos * 4 + 2); }
	}
	int refp = al.pos;                               // BamAlignment::GetSoftClips genome positions
	for (uint32_t c : al.cigar) {
		int op = c & 15; int len = (int)(c >> 4);
		if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) { refp += len; }
		else if (op == 4) { al.ev.push_back(refp * 4 + 3); }
	}
}

static bool is_active(const vector<Aln> &alns, size_t lo, size_t hi, int left, int right, bool normal, const Opts &o, vector<int32_t> &scratch)
{
	int MQ = normal ? 0 : o.min_map_qual; const int MIN_EVIDENCE = o.f.minAltCntTumor;
	scratch.clear();
	for (size_t i = lo; i < hi; ++i) {
		const Aln &al = alns[i];
		if (al.pos < left || al.end > right) { continue; }
		if (!(al.mapq >= MQ && !(al.flag & 0x400))) { continue; }
		if (al.seq.empty() || al.qual.empty()) { continue; }
		if (!al.ev_done) { read_evidence(al, o.min_qv_call + o.qv_range); }
		scratch.insert(scratch.end(), al.ev.begin(), al.ev.end());
	}
	std::sort(scratch.begin(), scratch.end());       // equal (position, kind) next to each other: a run of MIN_EVIDENCE is a map entry that reached it
	for (size_t i = 0, j = 0; i < scratch.size(); i = j) {
		while (j < scratch.size() && scratch[j] == scratch[i]) { ++j; }
		if ((long)(j - i) >= (long)MIN_EVIDENCE) { return true; }
	}
	return false;
}

// ------------------------------------------------------------------------------------------------------------
int main(int argc, char **argv)
{
	Opts o;
	for (int i = 0; i < argc; ++i) { o.cmdline += string(argv[i]) + " "; }
	static struct option lo[] = {
		{"tumor", 1, 0, 't'}, {"normal", 1, 0, 'n'}, {"ref", 1, 0, 'r'}, {"bed", 1, 0, 'B'}, {"reg", 1, 0, 'p'}, {"rg-file", 1, 0, 'g'}, {"min-k", 1, 0, 'k'}, {"max-k", 1, 0, 'K'},
		{"tip-len", 1, 0, 'l'}, {"cov-thr", 1, 0, 'c'}, {"cov-ratio", 1, 0, 'x'}, {"low-cov", 1, 0, 'd'}, {"window-size", 1, 0, 'w'}, {"padding", 1, 0, 'P'},
		{"max-avg-cov", 1, 0, 'u'}, {"min-map-qual", 1, 0, 'b'}, {"max-as-xs-diff", 1, 0, 'Z'}, {"min-base-qual", 1, 0, 'C'}, {"trim-lowqual", 1, 0, 'q'},
		{"quality-range", 1, 0, 'Q'}, {"node-str-len", 1, 0, 'L'}, {"dfs-limit", 1, 0, 'F'}, {"num-threads", 1, 0, 'X'}, {"max-indel-len", 1, 0, 'T'},
		{"max-mismatch", 1, 0, 'M'}, {"max-unit-length", 1, 0, 'U'}, {"min-report-unit", 1, 0, 'N'}, {"min-report-len", 1, 0, 'Y'}, {"dist-from-str", 1, 0, 'D'},
		{"min-phred-fisher-str", 1, 0, 'E'}, {"min-phred-fisher", 1, 0, 's'}, {"min-strand-bias", 1, 0, 'f'}, {"min-alt-count-tumor", 1, 0, 'a'},
		{"max-alt-count-normal", 1, 0, 'm'}, {"min-vaf-tumor", 1, 0, 'e'}, {"max-vaf-normal", 1, 0, 'i'}, {"min-coverage-tumor", 1, 0, 'o'},
		{"max-coverage-tumor", 1, 0, 'y'}, {"min-coverage-normal", 1, 0, 'z'}, {"max-coverage-normal", 1, 0, 'j'}, {"linked-reads", 0, 0, 'J'},
		{"primary-alignment-only", 0, 0, 'I'}, {"XA-tag-filter", 0, 0, 'O'}, {"active-region-off", 0, 0, 'W'}, {"kmer-recovery-on", 0, 0, 'R'},
		{"verbose", 0, 0, 'v'}, {"more-verbose", 0, 0, 'V'}, {"print-graph", 0, 0, 'A'}, {"print-config-file", 0, 0, 'G'}, {"gpu", 1, 0, 1000}, {"self-test", 0, 0, 1001},
		{"gpus", 1, 0, 1002}, {"batch-windows", 1, 0, 1003}, {"io-threads", 1, 0, 1004}, {"rank", 1, 0, 1005}, {"world", 1, 0, 1006}, {"nccl-id-file", 1, 0, 1007}, {0, 0, 0, 0} };
	int ch, oi = 0;
	while ((ch = getopt_long(argc, argv, "u:n:r:g:k:K:l:f:t:c:C:d:x:GARhSIWJOL:T:P:M:vVF:q:b:B:Q:p:s:E:a:m:e:i:o:y:z:w:j:X:U:N:Y:D:Z:", lo, &oi)) != -1) {
		switch (ch) {
			case 't': o.tumor = optarg; break; case 'n': o.normal = optarg; break; case 'r': o.ref = optarg; break; case 'B': o.bed = optarg; break; case 'p': o.reg = optarg; break;
			case 'k': o.minK = atoi(optarg); break; case 'K': o.maxK = atoi(optarg); break; case 'l': o.tip_len = atoi(optarg); break; case 'c': o.cov_thr = atoi(optarg); break;
			case 'x': o.cov_ratio = atof(optarg); break; case 'd': o.low_cov = atoi(optarg); break; case 'w': o.window = atoi(optarg); break; case 'P': o.padding = atoi(optarg); break;
			case 'u': o.max_avg_cov = atoi(optarg); break; case 'q': o.min_qv_trim = atoi(optarg); break; case 'C': o.min_qv_call = atoi(optarg); break; case 'b': o.min_map_qual = atoi(optarg); break;
			case 'Z': break;   // accepted and ignored, exactly like the reference's main()
			case 'Q': o.qv_range = *optarg; break; case 'L': break; case 'F': o.dfs_limit = atoi(optarg); break; case 'X': o.num_threads = atoi(optarg); break;
			case 'T': o.max_indel_len = atoi(optarg); break; case 'M': o.max_mismatch = atoi(optarg); break; case 'U': o.max_unit_len = atoi(optarg); break;
			case 'N': o.min_report_units = atoi(optarg); break; case 'Y': o.min_report_len = atoi(optarg); break; case 'D': o.dist_from_str = atoi(optarg); break;
			case 'E': o.f.minPhredFisherSTR = atof(optarg); break; case 's': o.f.minPhredFisher = atof(optarg); break; case 'f': o.f.minStrandBias = (int)atof(optarg); break;
			case 'a': o.f.minAltCntTumor = atoi(optarg); break; case 'm': o.f.maxAltCntNormal = atoi(optarg); break; case 'e': o.f.minVafTumor = atof(optarg); break;
			case 'i': o.f.maxVafNormal = atof(optarg); break; case 'o': o.f.minCovTumor = atoi(optarg); break; case 'y': o.f.maxCovTumor = atoi(optarg); break;
			case 'z': o.f.minCovNormal = atoi(optarg); break; case 'j': o.f.maxCovNormal = atoi(optarg); break;
			case 'I': o.primary_only = true; break; case 'O': o.xa_filter = true; break; case 'W': o.active_regions = false; break;
			case 'v': case 'V': o.verbose = true; break; case 'G': break; case 1000: o.gpu = atoi(optarg); break;
			case 1002: o.gpus = atoi(optarg); break; case 1003: o.batch_windows = atoi(optarg); break; case 1004: o.io_threads = atoi(optarg); break;
			case 1005: o.rank = atoi(optarg); break; case 1006: o.world = atoi(optarg); break; case 1007: o.nccl_id_file = optarg; break;
			case 1001: {      // known-answer hooks for tests/test_cli_host.py (hash, Fisher point probability, number formatting)
				std::cout << "sha256(abc)=" << sha256_hex("abc") << "\nsha256()=" << sha256_hex("") << "\nsha256(chr22:1234:S:1:A:T x3)=" << sha256_hex("chr22:1234:S:1:A:Tchr22:1234:S:1:A:Tchr22:1234:S:1:A:T") << "\n";
				int tabs[6][4] = { {10, 12, 0, 7}, {30, 28, 0, 0}, {3, 1, 1, 3}, {100, 90, 2, 25}, {0, 0, 5, 5}, {1000, 800, 0, 300} };
				for (auto &t : tabs) { double p = FET::point(t[0], t[1], t[2], t[3]); std::cout << "fet " << t[0] << " " << t[1] << " " << t[2] << " " << t[3] << " " << p << " " << Variant::score(p, true) << "\n"; }
				std::cout << "dtos " << dtos(0.0) << " " << dtos(12.3456789) << " " << dtos(1e-7) << " " << dtos(3079.99) << "\n";
				return 0; }
			case 'J': case 'R': case 'A': case 'g': std::cerr << "ERROR: option not supported by lancet_b200 (linked reads / k-mer recovery / DOT dumps / RG file)" << std::endl; return EXIT_FAILURE;
			default: std::cerr << "usage: lancet_b200 --tumor T.bam --normal N.bam --ref ref.fa (--reg chr:a-b | --bed f.bed) [lancet options]" << std::endl; return EXIT_FAILURE;
		}
	}
	int err = 0;
	if (o.tumor.empty()) { std::cerr << "ERROR: Must provide the tumor BAM file (-t)" << std::endl; ++err; }
	if (o.normal.empty()) { std::cerr << "ERROR: Must provide the normal BAM file (-n)" << std::endl; ++err; }
	if (o.ref.empty()) { std::cerr << "ERROR: Must provide a reference genome file (-r)" << std::endl; ++err; }
	if (o.bed.empty() && o.reg.empty()) { std::cerr << "ERROR: Must provide region (-p) or BED file (-B)" << std::endl; ++err; }
	if (err) { return EXIT_FAILURE; }
	if (o.num_threads < 1) { o.num_threads = 1; }

	// which reference sequences are needed
	std::map<string, bool> want; vector<string> regions; size_t n_bed_regions = 0;
	if (!o.bed.empty()) {
		std::ifstream bf(o.bed); if (!bf) { std::cerr << "Couldn't open " << o.bed << std::endl; return 1; }
		string line;
		while (std::getline(bf, line)) {
			if (line.find_first_of('#') == 0) { continue; }
			std::istringstream is(line); string tok; vector<string> t; while (std::getline(is, tok, '\t')) { t.push_back(tok); }
			if (t.size() < 3) { continue; }
			int SP = std::stoi(t[1]) - o.padding, EP = std::stoi(t[2]) + o.padding; if (SP < 1) { SP = 1; }       // padded here and again in load_refs (reference quirk)
			regions.push_back(t[0] + ":" + itos(SP) + "-" + itos(EP)); want[t[0]] = true;
		}
	}
	n_bed_regions = regions.size();
	if (!o.reg.empty()) { regions.push_back(o.reg); want[o.reg.substr(0, o.reg.find_first_of(':'))] = true; }

	if (o.gpus < 1) { o.gpus = 1; } if (o.batch_windows < 1) { o.batch_windows = 1; } if (o.batch_windows > 65536) { o.batch_windows = 65536; }
	if (o.io_threads <= 0) { o.io_threads = (int)std::max(1u, std::min(16u, std::thread::hardware_concurrency() / (unsigned)std::max(1, o.rank >= 0 ? o.world : o.gpus))); }
	// ---- several GPUs: one process per GPU.  The launcher (no --rank) writes an NCCL id, starts ranks 1..N-1 as copies of
	// itself and continues as rank 0; every rank assembles a contiguous range of the windows on its own GPU, the records are
	// gathered on rank 0 over NCCL (lb2_comm_gather), rank 0 replays them into the variant store and writes the VCF
	vector<pid_t> children;
	// stdout carries the VCF and nothing else: NCCL's diagnostics go to stderr (its version banner ignores NCCL_DEBUG_FILE,
	// so stdout is pointed at stderr while the communicator is set up)
	if (o.gpus > 1 || o.world > 1) { setenv("NCCL_DEBUG_FILE", "/dev/stderr", 0); }
	struct StdoutToStderr { int saved = -1; void on() { fflush(stdout); saved = dup(1); dup2(2, 1); } void off() { if (saved >= 0) { fflush(stdout); dup2(saved, 1); close(saved); saved = -1; } } } quiet;
	if (o.gpus > 1 && o.rank < 0) {
		char id[LB2_COMM_ID_BYTES];
		quiet.on(); const int id_rc = lb2_comm_unique_id(id); quiet.off();
		if (id_rc != LB2_OK) { std::cerr << "ERROR: cannot create the NCCL id (no GPU / no NCCL)" << std::endl; return 2; }
		o.nccl_id_file = "/tmp/lancet_b200_nccl_" + itos((int)getpid()) + ".id";
		{ std::ofstream idf(o.nccl_id_file, std::ios::binary); idf.write(id, sizeof id); }
		o.world = o.gpus; o.rank = 0;
		for (int r = 1; r < o.world; ++r) {
			pid_t pid = fork();
			if (pid == 0) {
				vector<string> av(argv, argv + argc); av.push_back("--rank"); av.push_back(itos(r)); av.push_back("--world"); av.push_back(itos(o.world)); av.push_back("--nccl-id-file"); av.push_back(o.nccl_id_file);
				vector<char *> cav; for (auto &x : av) { cav.push_back((char *)x.c_str()); } cav.push_back(nullptr);
				int devnull = open("/dev/null", O_WRONLY); if (devnull >= 0) { dup2(devnull, 1); }      // only rank 0 writes the VCF
				execv("/proc/self/exe", cav.data()); _exit(127);
			}
			if (pid < 0) { std::cerr << "ERROR: fork failed" << std::endl; return 2; }
			children.push_back(pid);
		}
	}
	if (o.rank < 0) { o.rank = 0; o.world = 1; }
	const bool lead = o.rank == 0;

	// LB2_CLI_TIMING=1: where the wall time of this run went (stderr)
	const bool timing = getenv("LB2_CLI_TIMING") != nullptr;
	auto now = []() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
	const double t_start = now(); double t_fetch = 0, t_pool = 0, t_rank = 0, t_select = 0, t_process = 0, t_wait = 0; std::mutex t_mu;
	BamFile T, N;
	if (!T.open(o.tumor, (unsigned)o.io_threads)) { std::cerr << "Could not open tumor BAM file." << std::endl; return -1; }
	if (!N.open(o.normal, (unsigned)o.io_threads)) { std::cerr << "Could not open normal BAM file." << std::endl; return -1; }
	if (!(T.first_has_md || N.first_has_md) && o.active_regions) {
		if (lead) {
			std::cerr << "\n--------WARNING--------\nThe MD tag is required to select the active regions, but is missing from the alignments in the BAM(s) file(s).\n"
			          << "To avoid unpredictable behavior, the active region module has been automatically turned off (--active-region-off)\n-----------------------\n" << std::endl;
		}
		o.active_regions = false;
	}
	const double t_open = now();
	vector<Window> wins; int num_windows = 0, t = 0;
	for (size_t r = 0; r < regions.size(); ++r) { t = load_refs(o, T, regions[r], wins, r < n_bed_regions ? t : 0, num_windows); }      // loadBed carries the thread counter, --reg restarts at 0
	{	// reftable[T] is a std::map keyed by the header: a second window with the same header on the same thread is dropped
		std::map<std::pair<int, string>, bool> seen; vector<Window> uniq;
		for (auto &w : wins) { auto key = std::make_pair(w.thread, w.hdr); if (!seen.count(key)) { seen[key] = true; uniq.push_back(w); } }
		wins.swap(uniq);
	}
	if (lead) { std::cerr << num_windows << " total windows to process" << std::endl; }
	const double t_refs = now();
	// the CUDA context before the batch builders start, alone: created beside the batch builders (16 threads inflating, decoding, allocating) it
	// took 3-4 s instead of 1 s on the B200 box (presumably both want the process's address-space lock)
	lb2_params p; lb2_default_params(&p);
	p.min_k = o.minK; p.max_k = o.maxK; p.min_qual_trim = o.min_qv_trim + o.qv_range; p.min_qual_call = o.min_qv_call + o.qv_range; p.cov_threshold = o.cov_thr;
	p.low_cov_threshold = o.low_cov; p.max_tip_len = o.tip_len; p.dfs_limit = o.dfs_limit; p.max_indel_len = o.max_indel_len; p.max_mismatch = o.max_mismatch;
	p.max_unit_len = o.max_unit_len; p.min_report_units = o.min_report_units; p.min_report_len = o.min_report_len; p.dist_from_str = o.dist_from_str; p.min_cov_ratio = o.cov_ratio;
	lb2_ctx *ctx = nullptr; int rc_create = LB2_OK;
	rc_create = lb2_create(&ctx, &p, o.gpu + o.rank);
	if (rc_create != LB2_OK) { std::cerr << "ERROR: " << lb2_strerror(ctx, rc_create) << std::endl; return 2; }
	const double t_ctx = now();

	// ---- this rank's windows, in batches: fetch the reads of the batch's span once per sample (BAI), select per window ----
	const size_t wlo = wins.size() * (size_t)o.rank / (size_t)o.world, whi = wins.size() * (size_t)(o.rank + 1) / (size_t)o.world;
	struct Pinned {      // a growable page-locked array (the copies to the GPU then overlap the assembly, include/lancet_b200.h)
		void *p = nullptr; size_t cap = 0;
		void *need(size_t bytes, size_t keep = 0) { if (bytes > cap) { size_t nc = std::max(bytes + bytes / 2, (size_t)4096); void *q = lb2_alloc_pinned(nc); if (!q) { std::cerr << "ERROR: out of page-locked memory" << std::endl; exit(2); } if (keep) { memcpy(q, p, keep); } if (p) { lb2_free_pinned(p); } p = q; cap = nc; } return p; }
		~Pinned() { if (p) { lb2_free_pinned(p); } }
	};
	struct HostBatch {
		Pinned ref_off, ref_start, chr_id, wr_off, wr_idx, base_off, flags, name_rank, ref_seq, seq, qual;
		vector<uint32_t> win_of_batch; uint32_t n_windows = 0, n_reads = 0; uint64_t n_wr = 0, n_ref = 0, n_base = 0; int skipped = 0; bool last = false;
	};
	auto ref_id = [](const BamFile &b, const string &chr) { for (size_t i = 0; i < b.ref_names.size(); ++i) { if (b.ref_names[i] == chr) { return (int)i; } } return -1; };
	auto usable = [&](const Aln &al, int sample) -> bool {      // the filters of extractReads that do not depend on the window (src/Microassembler.cc:505-616)
		const int MQ = sample ? 0 : o.min_map_qual; const int MIN_DELTA = sample ? -1 : 5;       // MAX_DELTA_AS_XS is 5 whatever -Z says
		if (o.primary_only && (al.flag & 0x100)) { return false; }
		if (!(al.mapq >= MQ && !(al.flag & 0x400))) { return false; }
		const float delta = std::fabs(al.as - al.xs);
		if (delta <= MIN_DELTA && al.as != -1 && al.xs != -1) { return false; }
		if (al.xt == "R" && !sample) { return false; }
		if (!al.xa.empty() && !sample && o.xa_filter) { return false; }
		return true;
	};
	auto build_batch = [&](size_t a, size_t b, HostBatch &hb) {
		hb.win_of_batch.clear(); hb.skipped = 0;
		int left_all = INT32_MAX, right_all = 0; for (size_t wi = a; wi < b; ++wi) { left_all = std::min(left_all, wins[wi].refstart); right_all = std::max(right_all, wins[wi].refend); }
		const string &chr = wins[a].chr;
		vector<Aln> A[2];
		double tb0 = now();
		if (!T.fetch(ref_id(T, chr), left_all, right_all, A[0]) || !N.fetch(ref_id(N, chr), left_all, right_all, A[1])) { std::cerr << "ERROR: cannot read the BAM files (truncated or corrupt?)" << std::endl; exit(2); }
		double tb1 = now();
		// the pool: every usable alignment of the span once, tumour first
		vector<int64_t> pool_of[2]; uint32_t n_reads = 0; uint64_t n_base = 0;
		for (int s_ = 0; s_ < 2; ++s_) { pool_of[s_].assign(A[s_].size(), -1); for (size_t i = 0; i < A[s_].size(); ++i) { if (usable(A[s_][i], s_)) { pool_of[s_][i] = n_reads++; n_base += A[s_][i].seq.size(); } } }
		uint64_t *base_off = (uint64_t *)hb.base_off.need(8 * ((size_t)n_reads + 1)); uint8_t *flags = (uint8_t *)hb.flags.need(n_reads + 1); uint32_t *rank = (uint32_t *)hb.name_rank.need(4 * ((size_t)n_reads + 1));
		char *seq = (char *)hb.seq.need(n_base + 64), *qual = (char *)hb.qual.need(n_base + 64);
		vector<const char *> names(n_reads); uint64_t bo = 0; base_off[0] = 0;
		for (int s_ = 0; s_ < 2; ++s_) {
			for (size_t i = 0; i < A[s_].size(); ++i) {
				const int64_t r = pool_of[s_][i]; if (r < 0) { continue; }
				const Aln &al = A[s_][i];
				int mate = (al.flag & 0x40) ? 1 : 0; if (al.flag & 0x80) { mate = 2; }
				flags[r] = (uint8_t)((s_ ? LB2_READ_NORMAL : 0) | ((al.flag & 0x10) ? LB2_READ_REVERSE : 0) | (uint8_t)(mate << LB2_READ_MATE_SHIFT) | ((al.flag & 0x4) ? LB2_READ_UNMAPPED : 0));
				names[r] = al.name.c_str();
				memcpy(seq + bo, al.seq.data(), al.seq.size());
				if (al.qual.empty()) { memset(qual + bo, (char)(0xFF + 33), al.seq.size()); } else { memcpy(qual + bo, al.qual.data(), al.seq.size()); }
				bo += al.seq.size(); base_off[r + 1] = bo;
			}
		}
		double tb2 = now();
		lb2_rank_names(names.data(), n_reads, rank);
		double tb3 = now();
		uint32_t *ref_off = (uint32_t *)hb.ref_off.need(4 * (b - a + 1)), *wr_off = (uint32_t *)hb.wr_off.need(4 * (b - a + 1)), *chr_id = (uint32_t *)hb.chr_id.need(4 * (b - a + 1));
		int32_t *ref_start = (int32_t *)hb.ref_start.need(4 * (b - a + 1));
		hb.wr_idx.need(4 * ((size_t)n_reads * 8 + 1024)); hb.ref_seq.need((b - a) * 640 + 64);      // (a read lies in ~6 windows: one page-locked allocation, not a series of growing ones)
		uint32_t nw = 0; uint64_t n_wr = 0, n_ref = 0; ref_off[0] = 0; wr_off[0] = 0;
		vector<int32_t> ev_scratch;
		for (size_t wi = a; wi < b; ++wi) {
			const Window &w = wins[wi];
			if (w.raw.empty()) { continue; }
			const int left = w.refstart, right = w.refend;      // 1-based numbers used as a 0-based half-open BAM region (reference quirk B9)
			size_t lo[2], hi[2];
			for (int s_ = 0; s_ < 2; ++s_) {
				lo[s_] = std::lower_bound(A[s_].begin(), A[s_].end(), left, [](const Aln &x, int v) { return x.pos < v; }) - A[s_].begin();
				hi[s_] = std::lower_bound(A[s_].begin(), A[s_].end(), right, [](const Aln &x, int v) { return x.pos < v; }) - A[s_].begin();
			}
			bool activeT = true, activeN = true;
			if (o.active_regions) { activeT = is_active(A[0], lo[0], hi[0], left, right, false, o, ev_scratch); activeN = is_active(A[1], lo[1], hi[1], left, right, true, o, ev_scratch); }
			if (!(activeT || activeN)) { ++hb.skipped; continue; }
			const uint64_t wr0 = n_wr; bool skip = false;
			uint32_t *wr_idx = (uint32_t *)hb.wr_idx.need(4 * (n_wr + (hi[0] - lo[0]) + (hi[1] - lo[1]) + 1), 4 * n_wr);
			for (int s_ = 0; s_ < 2 && !skip; ++s_) {
				long totalbp = 0;
				for (size_t i = lo[s_]; i < hi[s_]; ++i) {
					const Aln &al = A[s_][i];
					if ((double)totalbp / (double)w.raw.length() > o.max_avg_cov) { skip = true; break; }
					if (al.pos < left || al.end > right) { continue; }
					if (pool_of[s_][i] < 0) { continue; }
					wr_idx[n_wr++] = (uint32_t)pool_of[s_][i];
					totalbp += (long)al.seq.length();
				}
			}
			if (skip) {      // "Too much coverage": the window is dropped
				std::cerr << "WARNING: Skip region " << w.chr << ":" << w.refstart << "-" << w.refend << ". Too much coverage (>" << o.max_avg_cov << "x)." << std::endl;
				n_wr = wr0; ++hb.skipped; continue;
			}
			if (n_wr > 0xFFFFFFF0ull || n_ref + w.raw.size() > 0xFFFFFFF0ull) { std::cerr << "ERROR: batch too large for 32-bit offsets (lower --batch-windows)" << std::endl; exit(2); }
			char *rs = (char *)hb.ref_seq.need(n_ref + w.raw.size() + 64, n_ref); memcpy(rs + n_ref, w.raw.data(), w.raw.size()); n_ref += w.raw.size();
			ref_start[nw] = w.refstart; chr_id[nw] = 0; ++nw; ref_off[nw] = (uint32_t)n_ref; wr_off[nw] = (uint32_t)n_wr;
			hb.win_of_batch.push_back((uint32_t)wi);
		}
		hb.wr_idx.need(4 * (n_wr + 1), 4 * n_wr); hb.ref_seq.need(n_ref + 64, n_ref);
		hb.n_windows = nw; hb.n_reads = n_reads; hb.n_wr = n_wr; hb.n_ref = n_ref; hb.n_base = n_base;
		{ const double tb4 = now(); std::lock_guard<std::mutex> lk(t_mu); t_fetch += tb1 - tb0; t_pool += tb2 - tb1; t_rank += tb3 - tb2; t_select += tb4 - tb3; }
	};
	// batches of consecutive windows on one chromosome.  The host side (BGZF inflate, record decode, filters, window
	// selection) is what a run spends its time on -- the GPU needs ~2.5 ms per 1000 windows -- so several builder threads
	// prepare different batches at once (builder p: batches p, p + P, ...; each with its share of the I/O threads) into a
	// ring of page-locked slots, and the GPU consumes them in order.  Regions too small for --batch-windows are still cut
	// into a few batches per builder.
	const unsigned P = (unsigned)std::max(1, std::min(4, o.io_threads / 4)), NS = 2 * P;
	T.threads = N.threads = std::max(1u, (unsigned)o.io_threads / P);
	const size_t per_batch = std::min<size_t>((size_t)std::max(1, o.batch_windows), std::max<size_t>(256, (whi - wlo + 2 * P - 1) / (2 * P)));
	vector<std::pair<size_t, size_t>> spans;
	for (size_t a = wlo; a < whi; ) { size_t b = a; while (b < whi && b - a < per_batch && wins[b].chr == wins[a].chr) { ++b; } spans.push_back(std::make_pair(a, b)); a = b; }
	vector<HostBatch> slots(NS); std::mutex mu; std::condition_variable cv; vector<int> filled(NS, 0);      // 0 free, 1 ready
	vector<std::thread> builders;
	for (unsigned pr = 0; pr < P; ++pr) {
		builders.emplace_back([&, pr]() {
			for (size_t i = pr; i < spans.size(); i += P) {
				HostBatch &hb = slots[i % NS];
				{ std::unique_lock<std::mutex> lk(mu); cv.wait(lk, [&]() { return filled[i % NS] == 0; }); }
				build_batch(spans[i].first, spans[i].second, hb);
				{ std::lock_guard<std::mutex> lk(mu); filled[i % NS] = 1; } cv.notify_all();
			}
		});
	}
	// ---- the communicator (one process per GPU) -------------------------------------------------------------------------
	int rc = rc_create;
	if (o.world > 1) {
		char id[LB2_COMM_ID_BYTES]; std::ifstream idf(o.nccl_id_file, std::ios::binary);
		bool id_ok = (bool)idf.read(id, sizeof id);
		if (id_ok) { quiet.on(); rc = lb2_comm_init(ctx, id, o.rank, o.world); quiet.off(); }
		if (!id_ok || rc != LB2_OK) { std::cerr << "ERROR: rank " << o.rank << ": NCCL set-up failed: " << lb2_strerror(ctx, rc) << std::endl; fflush(nullptr); _exit(2); }
	}

	vector<lb2_variant> vars; string strs; int tot_skip = 0; uint32_t n_failed = 0;      // this rank's records: window = index into wins, str_off into strs
	for (size_t i = 0; i < spans.size(); ++i) {
		HostBatch &hb = slots[i % NS];
		const double tw0 = now();
		{ std::unique_lock<std::mutex> lk(mu); cv.wait(lk, [&]() { return filled[i % NS] == 1; }); }
		t_wait += now() - tw0;
		tot_skip += hb.skipped;
		if (hb.n_windows) {
			lb2_batch b; memset(&b, 0, sizeof b);
			b.n_windows = hb.n_windows; b.n_reads = hb.n_reads; b.n_wr = (uint32_t)hb.n_wr; b.n_ref_bytes = hb.n_ref; b.n_base_bytes = hb.n_base;
			b.ref_off = (const uint32_t *)hb.ref_off.p; b.ref_start = (const int32_t *)hb.ref_start.p; b.chr_id = (const uint32_t *)hb.chr_id.p; b.wr_off = (const uint32_t *)hb.wr_off.p;
			b.wr_idx = (const uint32_t *)hb.wr_idx.p; b.base_off = (const uint64_t *)hb.base_off.p; b.flags = (const uint8_t *)hb.flags.p; b.name_rank = (const uint32_t *)hb.name_rank.p;
			b.ref_seq = (const char *)hb.ref_seq.p; b.seq = (const char *)hb.seq.p; b.qual = (const char *)hb.qual.p;
			lb2_result res; memset(&res, 0, sizeof res);
			const double tp0 = now();
			rc = lb2_process(ctx, &b, &res);
			t_process += now() - tp0; if (timing) { std::cerr << "[timing] lb2_process of batch " << i << ": " << hb.n_windows << " windows, " << hb.n_reads << " reads, " << now() - tp0 << " s, at " << now() - t_start << " s" << std::endl; }
			if (rc != LB2_OK) { std::cerr << "ERROR: " << lb2_strerror(ctx, rc) << std::endl; _exit(2); }
			for (uint32_t w = 0; w < res.n_windows; ++w) {      // a window the device could not assemble would be missing from the VCF without a trace: an error, not a warning
				if (res.windows[w].status >= LB2_WIN_OVERFLOW) {
					++n_failed;
					std::cerr << "ERROR: window " << wins[hb.win_of_batch[w]].hdr << " not assembled on the device (status " << (int)res.windows[w].status << ", detail " << res.windows[w].detail << ")" << std::endl;
				}
			}
			const size_t s0 = strs.size(); strs.append(res.strings, res.n_string_bytes);
			for (uint32_t k = 0; k < res.n_variants; ++k) { lb2_variant v = res.variants[k]; v.window = hb.win_of_batch[v.window]; v.str_off += (uint32_t)s0; vars.push_back(v); }
		}
		{ std::lock_guard<std::mutex> lk(mu); filled[i % NS] = 0; } cv.notify_all();
	}
	for (auto &th_ : builders) { th_.join(); }
	const double t_batches = now();

	// ---- gather on rank 0 (several GPUs) ------------------------------------------------------------------------------
	const lb2_variant *all_v = vars.data(); const char *all_s = strs.data(); uint32_t all_n = (uint32_t)vars.size();
	if (o.world > 1) {
		uint64_t stats[2] = { (uint64_t)tot_skip, (uint64_t)n_failed }; lb2_result merged; memset(&merged, 0, sizeof merged);
		quiet.on(); rc = lb2_comm_gather(ctx, vars.data(), (uint32_t)vars.size(), strs.data(), (uint64_t)strs.size(), stats, 2, 0, &merged); quiet.off();
		if (rc != LB2_OK) { std::cerr << "ERROR: rank " << o.rank << ": record gather failed: " << lb2_strerror(ctx, rc) << std::endl; return 2; }
		if (!lead) { fflush(nullptr); _exit(n_failed ? 3 : 0); }      // (no communicator teardown here: rank 0 is still busy and tears its side down alone)
		all_v = merged.variants; all_s = merged.strings; all_n = merged.n_variants; tot_skip = (int)stats[0]; n_failed = (uint32_t)stats[1];
	}
	std::cerr << "Total # of skipped windows: " << tot_skip << std::endl;
	int child_bad = 0;
	for (pid_t c : children) { int st = 0; waitpid(c, &st, 0); if (!WIFEXITED(st) || WEXITSTATUS(st) != 0) { ++child_bad; } }
	if (!o.nccl_id_file.empty() && !children.empty()) { unlink(o.nccl_id_file.c_str()); }
	if (n_failed || child_bad) { std::cerr << "ERROR: " << n_failed << " window(s) not assembled" << (child_bad ? ", a rank failed" : "") << "; no VCF written" << std::endl; lb2_destroy(ctx); return 3; }

	// ---- replay addVar in the reference's order: thread, then lexicographic window header, then emission ----------------
	vector<vector<uint32_t>> per_window(wins.size());
	for (uint32_t i = 0; i < all_n; ++i) { per_window[all_v[i].window].push_back(i); }
	vector<VariantDB> tdb(o.num_threads);
	for (int th = 0; th < o.num_threads; ++th) {
		vector<uint32_t> mine; for (size_t wi = 0; wi < wins.size(); ++wi) { if (wins[wi].thread == th) { mine.push_back((uint32_t)wi); } }
		std::sort(mine.begin(), mine.end(), [&](uint32_t a, uint32_t c) { return wins[a].hdr < wins[c].hdr; });      // std::map<string,Ref_t*> order
		for (uint32_t wi : mine) {
			for (uint32_t i : per_window[wi]) {
				const lb2_variant &v = all_v[i]; const char *s_ = all_s + v.str_off;
				string ref(s_, v.ref_len), alt(s_ + v.ref_len, v.alt_len), motif(s_ + v.ref_len + v.alt_len, v.motif_len);
				string str = v.str_len ? itos(v.str_len) + motif : string();
				tdb[th].add(Variant(wins[wi].chr, v.pos, ref, alt, v, str));
			}
		}
	}
	VariantDB all;
	for (int th = 0; th < o.num_threads; ++th) { for (auto &kv : tdb[th].DB) { all.add(kv.second); } }
	std::cerr << "Export variants to VCF file" << std::endl;
	print_header(o, N.sample, T.sample);
	vector<std::pair<string, Variant>> vec(all.DB.begin(), all.DB.end());
	std::sort(vec.begin(), vec.end(), byPos());
	for (auto &kv : vec) { std::cout << kv.second.vcf(o.f); }
	if (timing) {
		std::cerr << "[timing] open+index " << t_open - t_start << " s, reference+tiling " << t_refs - t_open << " s, CUDA context " << t_ctx - t_refs << " s, batches " << t_batches - t_ctx << " s (" << spans.size() << " batches, " << P << " builders; consumer: waiting for a batch "
		          << t_wait << " s, lb2_process " << t_process << " s; builders, summed: BAM fetch " << t_fetch << " s, pool " << t_pool << " s, name ranks " << t_rank << " s, window selection "
		          << t_select << " s; inside the fetches: read " << T.t_read + N.t_read << " s, inflate " << T.t_inflate + N.t_inflate << " s, record cut " << T.t_cut + N.t_cut << " s, decode " << T.t_decode + N.t_decode << " s), gather+replay+VCF " << now() - t_batches << " s" << std::endl;
	}
	lb2_destroy(ctx);
	return 0;
}
