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
	int gpu = 0;
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
// BAM decode (BGZF blocks inflated with zlib; records kept only for the reference sequences asked for)
// ------------------------------------------------------------------------------------------------------------
struct Aln {
	int32_t pos = 0, end = 0; uint16_t flag = 0; uint8_t mapq = 0; int32_t l_seq = 0;
	string name, seq, qual, md, xt, xa;
	vector<uint32_t> cigar;                // BAM encoding: len << 4 | op
	float as = -1, xs = -1; bool has_md = false;
};
struct Bam {
	vector<string> ref_names; vector<int32_t> ref_lens; string sample = "NA"; bool first_has_md = true;
	std::map<int, vector<Aln>> by_ref;
};

static bool inflate_bgzf(const string &path, vector<uint8_t> &out)
{
	FILE *f = fopen(path.c_str(), "rb"); if (!f) { return false; }
	vector<uint8_t> in; uint8_t buf[1 << 16]; size_t n;
	while ((n = fread(buf, 1, sizeof buf, f)) > 0) { in.insert(in.end(), buf, buf + n); }
	fclose(f);
	size_t p = 0;
	while (p + 18 <= in.size()) {
		if (in[p] != 31 || in[p + 1] != 139) { return false; }
		uint16_t xlen = in[p + 10] | (in[p + 11] << 8); uint32_t bsize = 0; size_t x = p + 12, xe = x + xlen;
		while (x + 4 <= xe) { uint16_t sl = in[x + 2] | (in[x + 3] << 8); if (in[x] == 'B' && in[x + 1] == 'C') { bsize = (in[x + 4] | (in[x + 5] << 8)) + 1u; } x += 4 + sl; }
		if (!bsize || p + bsize > in.size()) { return false; }
		size_t cdata = p + 12 + xlen, clen = bsize - xlen - 19;
		uint32_t isize = in[p + bsize - 4] | (in[p + bsize - 3] << 8) | (in[p + bsize - 2] << 16) | ((uint32_t)in[p + bsize - 1] << 24);
		size_t o0 = out.size(); out.resize(o0 + isize);
		if (isize) {
			z_stream zs; memset(&zs, 0, sizeof zs);
			if (inflateInit2(&zs, -15) != Z_OK) { return false; }
			zs.next_in = in.data() + cdata; zs.avail_in = (uInt)clen; zs.next_out = out.data() + o0; zs.avail_out = isize;
			int rc = inflate(&zs, Z_FINISH); inflateEnd(&zs);
			if (rc != Z_STREAM_END) { return false; }
		}
		p += bsize;
	}
	return true;
}

static bool read_bam(const string &path, Bam &b, const std::map<string, bool> &want_refs)
{
	vector<uint8_t> d;
	if (!inflate_bgzf(path, d)) { return false; }
	auto i32 = [&](size_t o) { int32_t v; memcpy(&v, d.data() + o, 4); return v; };
	if (d.size() < 12 || memcmp(d.data(), "BAM\1", 4)) { return false; }
	size_t p = 4; int32_t l_text = i32(p); p += 4; string text((const char *)d.data() + p, l_text); p += l_text;
	{	// first @RG line with an SM: field (reference retriveSampleName, src/Microassembler.cc:52-67)
		std::istringstream is(text); string line;
		while (std::getline(is, line)) {
			if (line.compare(0, 3, "@RG") == 0) {
				size_t q = line.find("\tSM:");
				if (q != string::npos) { size_t e = line.find('\t', q + 1); b.sample = line.substr(q + 4, e == string::npos ? string::npos : e - q - 4); }
				break;
			}
		}
	}
	int32_t n_ref = i32(p); p += 4;
	for (int r = 0; r < n_ref; ++r) { int32_t l = i32(p); p += 4; b.ref_names.push_back(string((const char *)d.data() + p, l - 1)); p += l; b.ref_lens.push_back(i32(p)); p += 4; }
	bool first = true;
	static const char *NT = "=ACMGRSVTWYHKDBN";
	while (p + 4 <= d.size()) {
		int32_t bs = i32(p); p += 4; size_t r0 = p; p += bs; if (p > d.size()) { break; }
		int32_t refID = i32(r0), pos = i32(r0 + 4); uint8_t l_name = d[r0 + 8], mapq = d[r0 + 9];
		uint16_t n_cig = d[r0 + 12] | (d[r0 + 13] << 8), flag = d[r0 + 14] | (d[r0 + 15] << 8); int32_t l_seq = i32(r0 + 16);
		size_t q = r0 + 32; const char *name = (const char *)d.data() + q; q += l_name;
		size_t cig = q; q += 4 * (size_t)n_cig; size_t sq = q; q += (l_seq + 1) / 2; size_t ql = q; q += l_seq; size_t tags = q, tend = r0 + bs;
		bool has_md = false; string md, xt, xa; float as = -1, xs = -1;
		size_t t = tags;
		while (t + 3 <= tend) {
			char t0 = d[t], t1 = d[t + 1], ty = d[t + 2]; t += 3; double num = 0; bool isnum = false; string sv;
			switch (ty) {
				case 'A': {      // bamtools GetTag<std::string> does strlen() on the value: a char tag runs on into the next tag's bytes up to a NUL
					size_t e = t; while (e < tend && d[e]) { ++e; } sv = string((const char *)d.data() + t, e - t); t += 1; break; }
				case 'c': num = (int8_t)d[t]; isnum = true; t += 1; break; case 'C': num = d[t]; isnum = true; t += 1; break;
				case 's': { int16_t v; memcpy(&v, d.data() + t, 2); num = v; isnum = true; t += 2; break; }
				case 'S': { uint16_t v; memcpy(&v, d.data() + t, 2); num = v; isnum = true; t += 2; break; }
				case 'i': { int32_t v; memcpy(&v, d.data() + t, 4); num = v; isnum = true; t += 4; break; }
				case 'I': { uint32_t v; memcpy(&v, d.data() + t, 4); num = v; isnum = true; t += 4; break; }
				case 'f': { float v; memcpy(&v, d.data() + t, 4); num = v; isnum = true; t += 4; break; }
				case 'Z': case 'H': { size_t e = t; while (e < tend && d[e]) { ++e; } sv = string((const char *)d.data() + t, e - t); t = e + 1; break; }
				case 'B': { char st = d[t]; int32_t cnt; memcpy(&cnt, d.data() + t + 1, 4); int es = (st == 'c' || st == 'C') ? 1 : (st == 's' || st == 'S') ? 2 : 4; t += 5 + (size_t)es * cnt; break; }
				default: t = tend; break;
			}
			if (t0 == 'M' && t1 == 'D') { has_md = true; md = sv; }
			else if (t0 == 'A' && t1 == 'S' && isnum) { as = (float)num; } else if (t0 == 'X' && t1 == 'S' && isnum) { xs = (float)num; }
			else if (t0 == 'X' && t1 == 'T') { xt = sv; } else if (t0 == 'X' && t1 == 'A') { xa = sv; }
		}
		if (first) { b.first_has_md = has_md; first = false; }
		if (refID < 0 || refID >= n_ref || !want_refs.count(b.ref_names[refID])) { continue; }
		Aln a; a.pos = pos; a.flag = flag; a.mapq = mapq; a.l_seq = l_seq; a.name = name; a.has_md = has_md; a.md = md; a.xt = xt; a.xa = xa; a.as = as; a.xs = xs;
		a.cigar.resize(n_cig); int32_t end = pos;
		for (int c = 0; c < n_cig; ++c) {
			uint32_t v; memcpy(&v, d.data() + cig + 4 * c, 4); a.cigar[c] = v; int op = v & 15;
			if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) { end += (int32_t)(v >> 4); }      // M D N = X (BamAlignment::GetEndPosition)
		}
		a.end = end;
		a.seq.resize(l_seq); a.qual.resize(l_seq);
		for (int i = 0; i < l_seq; ++i) { uint8_t by = d[sq + i / 2]; a.seq[i] = NT[(i & 1) ? (by & 15) : (by >> 4)]; a.qual[i] = (char)(d[ql + i] + 33); }
		if (l_seq && d[ql] == 0xFF) { a.qual.clear(); }
		b.by_ref[refID].push_back(std::move(a));
	}
	return true;
}

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

static int load_refs(const Opts &o, const Bam &bam, const string &region, vector<Window> &wins, int thread, int &num_windows)
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
static void parse_md(const string &md, std::map<int, int> &M, int start, const string &qual, int min_qv)
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
			if (q >= min_qv) { ++M[pos]; }
			p_old = p; p = md.find_first_of(valid, p_old + 1);
		}
	}
}

static bool is_active(const vector<Aln> &alns, size_t lo, size_t hi, int left, int right, bool normal, const Opts &o)
{
	int MQ = normal ? 0 : o.min_map_qual; const int MIN_EVIDENCE = o.f.minAltCntTumor;
	std::map<int, int> mX, mI, mD, mSC;
	for (size_t i = lo; i < hi; ++i) {
		const Aln &al = alns[i];
		if (al.pos < left || al.end > right) { continue; }
		if (!(al.mapq >= MQ && !(al.flag & 0x400))) { continue; }
		if (al.seq.empty() || al.qual.empty()) { continue; }
		if (al.has_md) { parse_md(al.md, mX, al.pos, al.qual, o.min_qv_call + o.qv_range); }
		int pos = al.pos;
		for (uint32_t c : al.cigar) {
			int op = c & 15; int len = (int)(c >> 4);
			if (op != 1) { pos += len; }                 // every operation except 'I' advances (literal, src/Microassembler.cc:321)
			if (op == 8) { ++mX[pos]; } if (op == 1) { ++mI[pos]; } if (op == 2) { ++mD[pos]; }
		}
		int refp = al.pos; bool firstop = true;          // BamAlignment::GetSoftClips genome positions
		for (uint32_t c : al.cigar) {
			int op = c & 15; int len = (int)(c >> 4);
			if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) { refp += len; }
			else if (op == 4) { ++mSC[refp]; }
			(void)firstop; firstop = false;
		}
	}
	auto any = [&](const std::map<int, int> &m) { for (auto &kv : m) { if (kv.second >= MIN_EVIDENCE) { return true; } } return false; };
	return any(mX) || any(mI) || any(mD) || any(mSC);
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
		{"verbose", 0, 0, 'v'}, {"more-verbose", 0, 0, 'V'}, {"print-graph", 0, 0, 'A'}, {"print-config-file", 0, 0, 'G'}, {"gpu", 1, 0, 1000}, {"self-test", 0, 0, 1001}, {0, 0, 0, 0} };
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

	Bam T, N;
	if (!read_bam(o.tumor, T, want)) { std::cerr << "Could not open tumor BAM file." << std::endl; return -1; }
	if (!read_bam(o.normal, N, want)) { std::cerr << "Could not open normal BAM file." << std::endl; return -1; }
	if (!(T.first_has_md || N.first_has_md) && o.active_regions) {
		std::cerr << "\n--------WARNING--------\nThe MD tag is required to select the active regions, but is missing from the alignments in the BAM(s) file(s).\n"
		          << "To avoid unpredictable behavior, the active region module has been automatically turned off (--active-region-off)\n-----------------------\n" << std::endl;
		o.active_regions = false;
	}
	vector<Window> wins; int num_windows = 0, t = 0;
	for (size_t r = 0; r < regions.size(); ++r) { t = load_refs(o, T, regions[r], wins, r < n_bed_regions ? t : 0, num_windows); }      // loadBed carries the thread counter, --reg restarts at 0
	{	// reftable[T] is a std::map keyed by the header: a second window with the same header on the same thread is dropped
		std::map<std::pair<int, string>, bool> seen; vector<Window> uniq;
		for (auto &w : wins) { auto key = std::make_pair(w.thread, w.hdr); if (!seen.count(key)) { seen[key] = true; uniq.push_back(w); } }
		wins.swap(uniq);
	}
	std::cerr << num_windows << " total windows to process" << std::endl;

	// ---- read selection per window -> batch arrays ------------------------------------------------------------
	vector<uint32_t> ref_off{0}, wr_off{0}, wr_idx, chr_id; vector<int32_t> ref_start; vector<uint64_t> base_off{0}; vector<uint8_t> flags;
	vector<const char *> names; string ref_seq, seq, qual; int tot_skip = 0;
	vector<uint32_t> win_of_batch;      // batch window -> index in wins
	auto ref_id = [](const Bam &b, const string &chr) { for (size_t i = 0; i < b.ref_names.size(); ++i) { if (b.ref_names[i] == chr) { return (int)i; } } return -1; };
	static const vector<Aln> none;
	const int qcall = o.min_qv_call + o.qv_range; (void)qcall;
	for (size_t wi = 0; wi < wins.size(); ++wi) {
		const Window &w = wins[wi];
		if (w.raw.empty()) { continue; }
		int rt = ref_id(T, w.chr), rn = ref_id(N, w.chr);
		const vector<Aln> &AT = (rt >= 0 && T.by_ref.count(rt)) ? T.by_ref[rt] : none, &AN = (rn >= 0 && N.by_ref.count(rn)) ? N.by_ref[rn] : none;
		const int left = w.refstart, right = w.refend;      // 1-based numbers used as a 0-based half-open BAM region (reference quirk B9)
		auto range = [&](const vector<Aln> &A, size_t &lo, size_t &hi) {
			lo = std::lower_bound(A.begin(), A.end(), left, [](const Aln &a, int v) { return a.pos < v; }) - A.begin();
			hi = std::lower_bound(A.begin(), A.end(), right, [](const Aln &a, int v) { return a.pos < v; }) - A.begin();
		};
		size_t tl, th, nl, nh; range(AT, tl, th); range(AN, nl, nh);
		bool activeT = true, activeN = true;
		if (o.active_regions) { activeT = is_active(AT, tl, th, left, right, false, o); activeN = is_active(AN, nl, nh, left, right, true, o); }
		if (!(activeT || activeN)) { ++tot_skip; continue; }
		size_t wr0 = wr_idx.size(); bool skip = false;
		for (int sample = 0; sample < 2 && !skip; ++sample) {
			const vector<Aln> &A = sample ? AN : AT; size_t lo_ = sample ? nl : tl, hi_ = sample ? nh : th;
			const int MQ = sample ? 0 : o.min_map_qual; const int MIN_DELTA = sample ? -1 : 5;       // MAX_DELTA_AS_XS is 5 whatever -Z says
			long totalbp = 0;
			for (size_t i = lo_; i < hi_; ++i) {
				const Aln &al = A[i];
				if ((double)totalbp / (double)w.raw.length() > o.max_avg_cov) { skip = true; break; }
				if (al.pos < left || al.end > right) { continue; }
				if (o.primary_only && (al.flag & 0x100)) { continue; }
				if (!(al.mapq >= MQ && !(al.flag & 0x400))) { continue; }
				float delta = std::fabs(al.as - al.xs);
				if (delta <= MIN_DELTA && al.as != -1 && al.xs != -1) { continue; }
				if (al.xt == "R" && !sample) { continue; }
				if (!al.xa.empty() && !sample && o.xa_filter) { continue; }
				int mate = (al.flag & 0x40) ? 1 : 0; if (al.flag & 0x80) { mate = 2; }
				uint8_t fl = (sample ? LB2_READ_NORMAL : 0) | ((al.flag & 0x10) ? LB2_READ_REVERSE : 0) | (uint8_t)(mate << LB2_READ_MATE_SHIFT) | ((al.flag & 0x4) ? LB2_READ_UNMAPPED : 0);
				wr_idx.push_back((uint32_t)flags.size()); flags.push_back(fl); names.push_back(al.name.c_str());
				seq += al.seq; qual += al.qual.empty() ? string(al.seq.size(), (char)(0xFF + 33)) : al.qual; base_off.push_back(seq.size());
				totalbp += (long)al.seq.length();
			}
		}
		if (skip) {      // "Too much coverage": the window is dropped (reads staged for it are simply not referenced)
			std::cerr << "WARNING: Skip region " << w.chr << ":" << w.refstart << "-" << w.refend << ". Too much coverage (>" << o.max_avg_cov << "x)." << std::endl;
			wr_idx.resize(wr0); ++tot_skip; continue;
		}
		ref_seq += w.raw; ref_off.push_back((uint32_t)ref_seq.size()); ref_start.push_back(w.refstart); chr_id.push_back(0); wr_off.push_back((uint32_t)wr_idx.size());
		win_of_batch.push_back((uint32_t)wi);
	}
	std::cerr << "Total # of skipped windows: " << tot_skip << std::endl;

	// ---- micro-assembly on the GPU ----------------------------------------------------------------------------
	lb2_params p; lb2_default_params(&p);
	p.min_k = o.minK; p.max_k = o.maxK; p.min_qual_trim = o.min_qv_trim + o.qv_range; p.min_qual_call = o.min_qv_call + o.qv_range; p.cov_threshold = o.cov_thr;
	p.low_cov_threshold = o.low_cov; p.max_tip_len = o.tip_len; p.dfs_limit = o.dfs_limit; p.max_indel_len = o.max_indel_len; p.max_mismatch = o.max_mismatch;
	p.max_unit_len = o.max_unit_len; p.min_report_units = o.min_report_units; p.min_report_len = o.min_report_len; p.dist_from_str = o.dist_from_str; p.min_cov_ratio = o.cov_ratio;
	lb2_ctx *ctx = nullptr;
	int rc = lb2_create(&ctx, &p, o.gpu);
	if (rc != LB2_OK) { std::cerr << "ERROR: " << lb2_strerror(ctx, rc) << std::endl; return 2; }
	vector<uint32_t> rank(names.size());
	lb2_rank_names(names.data(), (uint32_t)names.size(), rank.data());
	lb2_batch b; memset(&b, 0, sizeof b);
	b.n_windows = (uint32_t)ref_start.size(); b.n_reads = (uint32_t)flags.size(); b.n_wr = (uint32_t)wr_idx.size(); b.n_ref_bytes = ref_seq.size(); b.n_base_bytes = seq.size();
	b.ref_off = ref_off.data(); b.ref_start = ref_start.data(); b.chr_id = chr_id.data(); b.wr_off = wr_off.data(); b.wr_idx = wr_idx.data(); b.base_off = base_off.data();
	b.flags = flags.data(); b.name_rank = rank.data(); b.ref_seq = ref_seq.data(); b.seq = seq.data(); b.qual = qual.data();
	lb2_result res; memset(&res, 0, sizeof res);
	if (b.n_windows) {
		rc = lb2_process(ctx, &b, &res);
		if (rc != LB2_OK) { std::cerr << "ERROR: " << lb2_strerror(ctx, rc) << std::endl; return 2; }
	}
	{	// a window the device could not assemble would be missing from the VCF without a trace: that is an error, not a warning
		uint32_t n_failed = 0;
		for (uint32_t w = 0; w < res.n_windows; ++w) {
			if (res.windows[w].status >= LB2_WIN_OVERFLOW) {
				++n_failed;
				std::cerr << "ERROR: window " << wins[win_of_batch[w]].hdr << " not assembled on the device (status " << (int)res.windows[w].status << ", detail " << res.windows[w].detail << ")" << std::endl;
			}
		}
		if (n_failed) { std::cerr << "ERROR: " << n_failed << " window(s) not assembled; no VCF written" << std::endl; lb2_destroy(ctx); return 3; }
	}

	// ---- replay addVar in the reference's order: thread, then lexicographic window header, then emission ----------------
	vector<vector<uint32_t>> per_window(wins.size());
	for (uint32_t i = 0; i < res.n_variants; ++i) { per_window[win_of_batch[res.variants[i].window]].push_back(i); }
	vector<VariantDB> tdb(o.num_threads);
	for (int th = 0; th < o.num_threads; ++th) {
		vector<uint32_t> mine; for (size_t wi = 0; wi < wins.size(); ++wi) { if (wins[wi].thread == th) { mine.push_back((uint32_t)wi); } }
		std::sort(mine.begin(), mine.end(), [&](uint32_t a, uint32_t c) { return wins[a].hdr < wins[c].hdr; });      // std::map<string,Ref_t*> order
		for (uint32_t wi : mine) {
			for (uint32_t i : per_window[wi]) {
				const lb2_variant &v = res.variants[i]; const char *s = res.strings + v.str_off;
				string ref(s, v.ref_len), alt(s + v.ref_len, v.alt_len), motif(s + v.ref_len + v.alt_len, v.motif_len);
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
	lb2_destroy(ctx);
	return 0;
}
