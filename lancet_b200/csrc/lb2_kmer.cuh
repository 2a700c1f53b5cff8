// lb2_kmer.cuh -- 2-bit packed k-mers (k <= 128), canonical form, table hash, libstdc++ string hash.
//
// Layout: base i of a k-mer lives at bits [2i,2i+1] of a little-endian 2k-bit integer held in
// NW = ceil(k/32) 64-bit words (word 0 = first 32 bases).  Codes A=0 C=1 G=2 T=3, so that the
// reference's std::string comparison (src/Mer.hh:57-71, 'A'<'C'<'G'<'T') is a comparison of the
// first differing 2-bit group.
#ifndef LB2_KMER_CUH
#define LB2_KMER_CUH

#include "lb2_prims.cuh"

#define LB2_MAXW 4

struct lb2_kmer { uint64_t w[LB2_MAXW]; };

LB2_DEV int lb2_nw(int K) { return (K + 31) >> 5; }

LB2_DEV int lb2_code(char c) {   // ACGT -> 0..3, anything else -> -1
	switch (c) { case 'A': return 0; case 'C': return 1; case 'G': return 2; case 'T': return 3; }
	return -1;
}
LB2_DEV char lb2_base(int code) { return (char)((0x54474341u >> (8 * code)) & 0xFF); }
LB2_DEV char lb2_comp(char c) { // rrc() of reference src/util.cc:246-258 for upper-case input
	switch (c) { case 'A': return 'T'; case 'C': return 'G'; case 'G': return 'C'; case 'T': return 'A'; case 'N': return 'N'; }
	return 0;
}

// read base g of a packed 2-bit array
LB2_DEV int lb2_getbase(const uint32_t *bits, uint32_t g) { return (lb2_lds(&bits[g >> 4]) >> ((g & 15) << 1)) & 3; }
LB2_DEV int lb2_getbit(const uint32_t *mask, uint32_t g) { return (lb2_lds(&mask[g >> 5]) >> (g & 31)) & 1; }

template <int NWT> LB2_DEV void lb2_mask_top(struct lb2_kmer &k, int K);

// extract K bases starting at base g (little-endian words)
template <int NWT = LB2_MAXW> LB2_DEV void lb2_extract(const uint32_t *bits, uint32_t g, int K, lb2_kmer &out) {
	int nw = lb2_nw(K);
	uint32_t bitpos = g << 1;
#pragma unroll
	for (int j = 0; j < NWT; ++j) {
		if (j < nw) {
			uint32_t bp = bitpos + (uint32_t)j * 64;
			uint32_t wi = bp >> 5, sh = bp & 31;
			uint64_t lo = lb2_lds(&bits[wi]), mid = lb2_lds(&bits[wi + 1]), hi = lb2_lds(&bits[wi + 2]);
			uint64_t v = (lo >> sh) | (mid << (32 - sh));
			if (sh) { v |= hi << (64 - sh); }
			out.w[j] = v;
		} else { out.w[j] = 0; }
	}
	lb2_mask_top<NWT>(out, K);
}

// word-wise select (a conditional on whole structs takes addresses and sends both k-mers to local memory)
template <int NWT = LB2_MAXW> LB2_DEV lb2_kmer lb2_pick(bool first, const lb2_kmer &a, const lb2_kmer &b) {
	lb2_kmer r;
#pragma unroll
	for (int j = 0; j < LB2_MAXW; ++j) { r.w[j] = (j < NWT) ? (first ? a.w[j] : b.w[j]) : 0ull; }
	return r;
}

// (all indexing below is static after unrolling, so k-mers stay in registers)
template <int NWT = LB2_MAXW> LB2_DEV void lb2_mask_top(lb2_kmer &k, int K) {
	const int nw = lb2_nw(K), rem = (K & 31);
	if (rem) {
		const uint64_t m = (~0ull) >> (64 - 2 * rem);
#pragma unroll
		for (int j = 0; j < NWT; ++j) { if (j == nw - 1) { k.w[j] &= m; } }
	}
}
// rolling update: drop first base, append code c at position K-1
template <int NWT = LB2_MAXW> LB2_DEV void lb2_roll_fwd(lb2_kmer &f, int K, int c) {
	const int top = (K - 1) >> 5; const uint64_t ins = (uint64_t)c << (((K - 1) & 31) << 1);
#pragma unroll
	for (int j = 0; j < NWT; ++j) {
		uint64_t nxt = (j + 1 < NWT) ? f.w[(j + 1 < NWT) ? j + 1 : j] : 0;
		if (j < top) { f.w[j] = (f.w[j] >> 2) | (nxt << 62); }
		else if (j == top) { f.w[j] = (f.w[j] >> 2) | ins; }
	}
}
// reverse complement rolling update: prepend complement of c, drop last base
template <int NWT = LB2_MAXW> LB2_DEV void lb2_roll_rc(lb2_kmer &r, int K, int c) {
	const int nw = lb2_nw(K);
#pragma unroll
	for (int j = NWT - 1; j > 0; --j) { if (j < nw) { r.w[j] = (r.w[j] << 2) | (r.w[j - 1] >> 62); } }
	r.w[0] = (r.w[0] << 2) | (uint64_t)(3 - c);
	lb2_mask_top<NWT>(r, K);
}

// lexicographic a < b  (std::string operator<, equal -> false)
template <int NWT = LB2_MAXW> LB2_DEV bool lb2_less(const lb2_kmer &a, const lb2_kmer &b, int nw) {
#pragma unroll
	for (int j = 0; j < NWT; ++j) {
		if (j < nw) {
			uint64_t x = a.w[j] ^ b.w[j];
			if (x) {
				int pos = lb2_ctz64(x) & ~1;
				return ((a.w[j] >> pos) & 3) < ((b.w[j] >> pos) & 3);
			}
		}
	}
	return false;
}
template <int NWT = LB2_MAXW> LB2_DEV bool lb2_equal(const lb2_kmer &a, const lb2_kmer &b, int nw) {
	bool eq = true;
#pragma unroll
	for (int j = 0; j < NWT; ++j) { if (j < nw && a.w[j] != b.w[j]) { eq = false; } }
	return eq;
}

// table hash (not semantically significant -- only spreads keys over the open-addressing table); 32-bit ops only
template <int NWT = LB2_MAXW> LB2_DEV uint64_t lb2_table_hash(const lb2_kmer &k, int nw) {
	uint32_t h = 0x9E3779B9u, g = 0x85EBCA6Bu;
#pragma unroll
	for (int j = 0; j < NWT; ++j) {
		if (j < nw) {
			uint32_t lo = (uint32_t)k.w[j], hi = (uint32_t)(k.w[j] >> 32);
			h = (h ^ lo) * 0xCC9E2D51u; h = (h << 15) | (h >> 17);
			g = (g ^ hi) * 0x1B873593u; g = (g << 13) | (g >> 19);
			h += g;
		}
	}
	h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; g ^= h; g *= 0xC2B2AE35u; g ^= g >> 16;
	return ((uint64_t)g << 32) | h;
}

// table hash of a one-word k-mer (K <= 32): slot from the low bits, fingerprint from the top ten
LB2_DEV uint32_t lb2_hash1(uint32_t lo, uint32_t hi) {
	uint32_t h = lo * 0xCC9E2D51u + hi * 0x1B873593u;
	h ^= h >> 15; h *= 0x85EBCA6Bu; h ^= h >> 13;
	return h;
}
// K <= 16 (KT = uint32_t) / K <= 32 (KT = uint64_t) bases starting at base g of the packed array, not yet masked to 2K bits
template <class KT> LB2_DEV KT lb2_extract_small(lb2_sp bits, uint32_t g) {
	const uint32_t bp = g << 1, wi = bp >> 5, sh = bp & 31u;
	const uint32_t w0 = lb2s_ld(lb2_sp_at(bits, wi)), w1 = lb2s_ld(lb2_sp_at(bits, wi + 1));
	const uint32_t lo = lb2_fsr(w0, w1, sh);
	if (sizeof(KT) == 4) { return (KT)lo; }
	const uint32_t w2 = lb2s_ld(lb2_sp_at(bits, wi + 2));
	return (KT)(((uint64_t)lb2_fsr(w1, w2, sh) << 32) | lo);
}

// ---- libstdc++ std::hash<std::string> == _Hash_bytes(p, len, 0xc70f6907) (64-bit murmur2 variant)
// (libstdc++-v3/libsupc++/hash_bytes.cc).  Semantically significant: it
// fixes the iteration order of the reference's unordered_map<string,Node_t*> (src/Graph.hh:68).
LB2_DEV uint64_t lb2_shift_mix(uint64_t v) { return v ^ (v >> 47); }

struct lb2_stdhash {
	uint64_t h; uint64_t acc; int nacc;
};
LB2_DEV void lb2_sh_init(lb2_stdhash &s, uint32_t len) {
	const uint64_t mul = 0xc6a4a7935bd1e995ull;
	s.h = 0xc70f6907ull ^ ((uint64_t)len * mul); s.acc = 0; s.nacc = 0;
}
LB2_DEV void lb2_sh_byte(lb2_stdhash &s, unsigned char c) {
	const uint64_t mul = 0xc6a4a7935bd1e995ull;
	s.acc |= (uint64_t)c << (8 * s.nacc);
	if (++s.nacc == 8) {
		uint64_t d = lb2_shift_mix(s.acc * mul) * mul;
		s.h ^= d; s.h *= mul; s.acc = 0; s.nacc = 0;
	}
}
LB2_DEV uint64_t lb2_sh_final(lb2_stdhash &s) {
	const uint64_t mul = 0xc6a4a7935bd1e995ull;
	if (s.nacc) { s.h ^= s.acc; s.h *= mul; }
	s.h = lb2_shift_mix(s.h) * mul;
	s.h = lb2_shift_mix(s.h);
	return s.h;
}
LB2_DEV uint64_t lb2_stdhash_bytes(const char *p, uint32_t len) {
	lb2_stdhash s; lb2_sh_init(s, len);
	for (uint32_t i = 0; i < len; ++i) { lb2_sh_byte(s, (unsigned char)p[i]); }
	return lb2_sh_final(s);
}
// hash of the ASCII spelling of a packed k-mer
template <int NWT = LB2_MAXW> LB2_DEV uint64_t lb2_stdhash_kmer(const lb2_kmer &k, int K) {
	lb2_stdhash s; lb2_sh_init(s, (uint32_t)K);
#pragma unroll
	for (int j = 0; j < NWT; ++j) {
		const uint64_t w = k.w[j];
		for (int i = 0; i < 32 && j * 32 + i < K; ++i) { lb2_sh_byte(s, (unsigned char)lb2_base((int)((w >> (i << 1)) & 3))); }
	}
	return lb2_sh_final(s);
}

#endif
