// DEBUG-ONLY single-thread simulation of the device pipeline (LB2_HOSTSIM): the same
// lancet_b200/csrc/*.cuh sources compiled by g++ with a ONE-thread "CTA".  It exists because the
// development container has no GPU; it is not shipped, not imported by the package, is not the
// oracle, and no test asserts product behaviour through it (tests/ compare the CUDA path with
// oracle/_ref).  Usage: hostsim batch.lb2b [--min-k a --max-k b] [--out f.tsv] [--first i --count n]
#define LB2_HOSTSIM 1
#include <cstdio>
#include <cstdlib>
#include <stdlib.h>
#include <cstring>
#include <cmath>
#include <string>
#include <vector>
#include <algorithm>
#include "../../lancet_b200/csrc/lb2_pipeline.cuh"
#include "sim_pack.h"

template <class T> static void rd(FILE *f, std::vector<T> &v, size_t n) { v.resize(n); if (n && fread(v.data(), sizeof(T), n, f) != n) { fprintf(stderr, "short read\n"); exit(2); } }

int main(int argc, char **argv)
{
	const char *fn = NULL, *out = NULL; int first = 0, count = -1;
	lb2_params P; P.min_k = 11; P.max_k = 101; P.min_qual_trim = 43; P.min_qual_call = 50; P.cov_threshold = 5;
	P.low_cov_threshold = 1; P.max_tip_len = 11; P.dfs_limit = 1000000; P.max_indel_len = 500; P.max_mismatch = 2;
	P.max_unit_len = 4; P.min_report_units = 3; P.min_report_len = 7; P.dist_from_str = 1; P.min_cov_ratio = 0.01;
	for (int i = 1; i < argc; ++i) {
		std::string a = argv[i];
		if (a == "--min-k") P.min_k = atoi(argv[++i]); else if (a == "--max-k") P.max_k = atoi(argv[++i]);
		else if (a == "--out") out = argv[++i]; else if (a == "--first") first = atoi(argv[++i]);
		else if (a == "--count") count = atoi(argv[++i]); else if (a == "--dfs-limit") P.dfs_limit = atoi(argv[++i]);
		else fn = argv[i];
	}
	if (fn && std::string(fn) == "divtest") {
		// the division of the device's coverage fold (s*r and two fused residual corrections, r = RN(1/n)) against IEEE s/n
		unsigned long long rs = 88172645463325252ull; auto rnd = [&]() { rs ^= rs << 13; rs ^= rs >> 7; rs ^= rs << 17; return rs; };
		unsigned long long bad = 0, total = 0; const int per = getenv("LB2_DIVTEST_PER") ? atoi(getenv("LB2_DIVTEST_PER")) : 4000;
		for (int n = 1; n <= 8192; ++n) {
			const float fnn = (float)n, r = 1.0f / fnn;
			for (int it = 0; it < per; ++it) {
				float s; const unsigned long long x = rnd();
				switch (it & 3) {
					case 0: { uint32_t b = (uint32_t)(x >> 20) & 0x7FFFFFFFu; b = (b % (0x4D000000u - 0x30000000u)) + 0x30000000u; memcpy(&s, &b, 4); break; }
					case 1: { s = (float)(x % 200000000ull) / 16.0f; break; }
					case 2: { float t; uint32_t b = 0x3F800000u + (uint32_t)(x & 0x7FFFFFu) + (((uint32_t)(x >> 40) % 20) << 23); memcpy(&t, &b, 4); s = t * fnn; uint32_t sb; memcpy(&sb, &s, 4); sb += (int)((x >> 60) & 7) - 3; memcpy(&s, &sb, 4); break; }
					default: { s = (float)((x >> 8) % 70000) * (float)(1 + (x & 1023)) + (float)((x >> 30) % 70000); break; }
				}
				if (!(s >= 0.0f) || std::isinf(s)) { continue; }
				volatile float want = s / fnn; const float got = lb2_div_nr2(s, fnn, r); const float w2 = want;
				++total; if (memcmp(&w2, &got, 4) != 0) { if (bad++ < 10) { fprintf(stderr, "MISMATCH s=%a n=%d want=%a got=%a\n", s, n, w2, got); } }
			}
		}
		printf("divtest: %llu cases, %llu mismatches\n", total, bad);
		return bad ? 1 : 0;
	}
	if (fn && std::string(fn) == "tandemtest") {
		// lb2_find_tandems on "T <pos> <seq>" lines of stdin, same output format as oracle/kat_reference.cc
		char op[8]; int pos; static char buf[1 << 16];
		while (scanf("%7s %d %65535s", op, &pos, buf) == 3) {
			const std::string q(buf); int len = 0; char motif[256]; uint32_t ml = 0; bool ov = false;
			const bool ans = lb2_find_tandems([&](uint32_t i) -> char { return q[i]; }, (uint32_t)q.size(), &P, pos, len, motif, ml, 255, ov);
			printf("%d %d %s\n", ans ? 1 : 0, len, ml ? std::string(motif, ml).c_str() : ".");
		}
		return 0;
	}
	if (fn && std::string(fn) == "scantest") {
		// the word filter of lb2_diag_scan against the unfiltered scan (min_k < 11 switches the filter off): for every k >= 11
		// both must answer isRepeat / isAlmostRepeat alike, i.e. max(emax,10) and max(wmax,11) agree
		std::vector<uint8_t> smem(sizeof(lb2_sh) + 64, 0); lb2_win Wn; memset(&Wn, 0, sizeof Wn); Wn.sh = (lb2_sh *)smem.data(); Wn.P = &P;
		std::vector<uint64_t> taskbuf(1 << 16); Wn.ws0.sortk = taskbuf.data(); Wn.ws.sortk = taskbuf.data();
		unsigned long long rs = 88172645463325252ull; auto rnd = [&]() { rs ^= rs << 13; rs ^= rs >> 7; rs ^= rs << 17; return (uint32_t)(rs >> 11); };
		int bad = 0, ntest = 0, nrel = 0;
		for (int t = 0; t < 6000; ++t) {
			const int len = 20 + (int)(rnd() % 900); std::string q(len, 'A');
			for (int i = 0; i < len; ++i) { q[i] = "ACGT"[rnd() & 3]; }
			const int mode = t % 6;
			if (mode >= 1) {      // planted repeats: copies with 0..4 substitutions, tandem units, at the string ends too
				const int nplant = 1 + (int)(rnd() % 3);
				for (int z = 0; z < nplant; ++z) {
					const int rl = 8 + (int)(rnd() % 40); if (rl * 2 + 2 > len) { continue; }
					int a = (int)(rnd() % (len - rl)), b = (int)(rnd() % (len - rl));
					if (mode == 2) { a = 0; } if (mode == 3) { b = len - rl; } if (mode == 4) { b = a + 1 + (int)(rnd() % 6); if (b + rl > len) { continue; } }
					for (int i = 0; i < rl; ++i) { q[b + i] = q[a + i]; }
					const int nsub = (int)(rnd() % 5);
					for (int i = 0; i < nsub; ++i) { q[b + (int)(rnd() % rl)] = "ACGT"[rnd() & 3]; }
				}
			}
			std::vector<uint32_t> pk(len / 16 + 8, 0);
			for (int i = 0; i < len; ++i) { pk[i >> 4] |= (uint32_t)lb2_code(q[i]) << (2 * (i & 15)); }
			for (int mm = 1; mm <= 3; ++mm) {
				P.min_k = 11; lb2_diag_scan(Wn, pk.data(), 0, len, mm); const uint32_t e1 = Wn.sh->scan_emax, w1 = Wn.sh->scan_wmax;
				P.min_k = 5;  lb2_diag_scan(Wn, pk.data(), 0, len, mm); const uint32_t e0 = Wn.sh->scan_emax, w0 = Wn.sh->scan_wmax;
				++ntest; if (e0 >= 11 || w0 >= 12) { ++nrel; }
				if (std::max(e1, 10u) != std::max(e0, 10u) || std::max(w1, 11u) != std::max(w0, 11u)) { if (bad++ < 10) { fprintf(stderr, "MISMATCH t=%d len=%d mm=%d: filtered e=%u w=%u, exact e=%u w=%u\n", t, len, mm, e1, w1, e0, w0); } }
			}
		}
		printf("scantest: %d scans (%d with a result >= the thresholds), %d mismatches\n", ntest, nrel, bad);
		return bad ? 1 : 0;
	}
	FILE *f = fopen(fn, "rb"); if (!f) { perror(fn); return 2; }
	char magic[4]; uint32_t ver, W, R, nwr; uint64_t nref, nbase;
	if (fread(magic, 1, 4, f) != 4 || fread(&ver, 4, 1, f) != 1 || fread(&W, 4, 1, f) != 1 || fread(&R, 4, 1, f) != 1 || fread(&nwr, 4, 1, f) != 1 ||
	    fread(&nref, 8, 1, f) != 1 || fread(&nbase, 8, 1, f) != 1) return 2;
	std::vector<uint32_t> ref_off, chr_id, wr_off, wr_idx, name_rank; std::vector<int32_t> ref_start; std::vector<uint64_t> base_off;
	std::vector<uint8_t> flags; std::vector<char> ref_seq, seq, qual;
	rd(f, ref_off, W + 1); rd(f, ref_start, W); rd(f, chr_id, W); rd(f, wr_off, W + 1); rd(f, wr_idx, nwr); rd(f, base_off, R + 1);
	rd(f, flags, R); rd(f, name_rank, R); rd(f, ref_seq, nref); rd(f, seq, nbase); rd(f, qual, nbase); fclose(f);

	lb2_cfg C; memset(&C, 0, sizeof C);
	C.table_slots = 16384; C.max_nodes = 12000; C.max_reads = 8192; C.max_bp = (1 << 20) - 1024; C.arena_bytes = 1 << 21; C.deficit_bytes = 1 << 23;
	C.queue_cap = 1 << 22; C.graph_bytes = 1 << 20; C.max_inst = 1 << 20; C.max_var = 64; C.str_bytes = 8192; C.bucket_cap = 10273; C.max_k = 127; C.n_slots = 1; C.max_special = 2048;
	if (getenv("LB2_SIM_TS")) { C.table_slots = atoi(getenv("LB2_SIM_TS")); C.max_nodes = C.table_slots - C.table_slots / 4; }
	if (getenv("LB2_SIM_BP")) { C.max_bp = atoi(getenv("LB2_SIM_BP")); }
	if (getenv("LB2_SIM_GB")) { C.graph_bytes = atoi(getenv("LB2_SIM_GB")); }
	if (getenv("LB2_SIM_QUEUE")) { C.queue_cap = atoi(getenv("LB2_SIM_QUEUE")); }
	if (getenv("LB2_SIM_SPECIAL")) { C.max_special = atoi(getenv("LB2_SIM_SPECIAL")); }
	if (getenv("LB2_SIM_ARENA")) { C.arena_bytes = atoi(getenv("LB2_SIM_ARENA")); }
	if (getenv("LB2_SIM_INST")) { C.max_inst = atoi(getenv("LB2_SIM_INST")); }
	if (getenv("LB2_SIM_MAXVAR")) { C.max_var = atoi(getenv("LB2_SIM_MAXVAR")); C.str_bytes = 128u << 10; }
	lb2_dev_batch B; B.n_windows = W; B.ref_off = ref_off.data(); B.ref_start = ref_start.data(); B.wr_off = wr_off.data(); B.wr_idx = wr_idx.data();
	B.base_off = base_off.data(); B.flags = flags.data(); B.name_rank = name_rank.data(); B.ref_seq = ref_seq.data(); seq.resize(seq.size() + 64, 0); qual.resize(qual.size() + 64, 0);      /* the staging reads 16 bytes at a time (the device buffers have the same slack) */
	B.seq = seq.data(); B.qual = qual.data();
	SimPack sp; sim_pack(B, R, P, sp);
	std::vector<lb2_window_info> info(W); std::vector<lb2_variant> vars((size_t)W * C.max_var); std::vector<char> strs((size_t)W * C.str_bytes); std::vector<uint32_t> sused(W);
	lb2_dev_out O; memset(&O, 0, sizeof O); O.info = info.data(); O.variants = vars.data(); O.strings = strs.data(); O.str_used = sused.data();
	size_t wsb = lb2_ws_layout(C, NULL, NULL);
	std::vector<uint8_t> slab(wsb, 0); std::vector<uint8_t> smem(lb2_smem_bytes(C.max_bp, C.table_slots, C.graph_bytes) + 64, 0);
	lb2_win Wn; Wn.P = &P; Wn.C = &C; Wn.B = &B; Wn.O = &O; Wn.escal = false;
	lb2_ws_layout(C, slab.data(), &Wn.ws); Wn.ws0 = Wn.ws;
	Wn.sh = (lb2_sh *)smem.data();
	Wn.ref_raw = (char *)smem.data() + ((sizeof(lb2_sh) + 15) & ~(size_t)15);
	Wn.bits = (uint32_t *)(Wn.ref_raw + LB2_MAX_REF);
	Wn.lowq = Wn.bits + (C.max_bp / 16 + 4);
	Wn.treg = smem.data() + ((lb2_smem_fixed(C.max_bp) + 15) & ~(size_t)15);
	int w1 = count < 0 ? (int)W : std::min((int)W, first + count);
	FILE *fo = out ? fopen(out, "w") : stdout;
	for (int w = first; w < w1; ++w) {
		lb2_process_window(Wn, (uint32_t)w);
		const lb2_window_info &wi = info[w];
		if (wi.status != LB2_WIN_OK) { fprintf(stderr, "window %d status %d detail %u\n", w, wi.status, wi.detail); }
		for (uint32_t v = 0; v < wi.n_variants; ++v) {
			const lb2_variant &x = vars[(size_t)w * C.max_var + v]; const char *sp = strs.data() + (size_t)w * C.str_bytes + x.str_off;
			std::string ref(sp, x.ref_len), alt(sp + x.ref_len, x.alt_len), motif(sp + x.ref_len + x.alt_len, x.motif_len);
			// Variant_t constructor normalisation (reference src/Variant.hh:133-153) -- test-side only
			int pos = x.pos; char type = '?'; int len = 0;
			if (x.code == '^') { type = 'I'; ref = ""; len = (int)alt.size(); }
			if (x.code == 'v') { type = 'D'; alt = ""; len = (int)ref.size(); }
			if (x.code == 'x') { type = 'S'; pos++; }
			if (x.code == 'c') { type = 'C'; ref.erase(std::remove(ref.begin(), ref.end(), '-'), ref.end()); alt.erase(std::remove(alt.begin(), alt.end(), '-'), alt.end());
				int rl = (int)ref.size(), al = (int)alt.size(); len = (rl == al) ? al : (rl > al ? rl - al : al - rl); }
			if (type != 'S') { ref = std::string(1, (char)x.prev_bp_alt) + ref; alt = std::string(1, (char)x.prev_bp_alt) + alt; } else { len = 1; }
			std::string str = x.str_len ? std::to_string(x.str_len) + motif : std::string(".");
			fprintf(fo, "%d\t%d\t%c\t%d\t%s\t%s\t%d\t%s\t%u,%u,%u,%u,%u,%u,%u,%u\t%c\t%c\n", w, pos, type, len, ref.c_str(), alt.c_str(), x.kmer, str.c_str(),
				x.rcn_fwd, x.rcn_rev, x.rct_fwd, x.rct_rev, x.acn_fwd, x.acn_rev, x.act_fwd, x.act_rev, x.prev_bp_ref, x.prev_bp_alt);
		}
	}
	if (getenv("LB2_SIM_DBG")) { fprintf(stderr, "compress_par taken %lu, no scratch %lu, ring %lu\n", lb2_dbg_par[0], lb2_dbg_par[1], lb2_dbg_par[2]); }
	if (fo != stdout) fclose(fo);
	return 0;
}
