// lb2_pack.cuh -- the read-pool pre-pack pass: every pooled read is classified ONCE (a read is used by up to six
// overlapping windows), the windows then stage 2-bit words with bulk-async copies instead of re-reading ASCII.
//
// Per read: Graph_t::trim (reference src/Graph.cc:355-384: 5'/3' trim of bases that are not ACGT or whose quality is
// below MIN_QUAL_TRIM, junk flag for a non-ACGT base inside), bases 2-bit packed 16 per word, one bit per base for
// "quality < MIN_QUAL_CALL" (the only other quality fact the path tests: src/Node.cc:470-497, src/Graph.cc:148-158).
// This pass is HBM-shaped: it reads 2 bytes per base (ASCII base + quality) and writes 3/8 byte per base + 12 bytes per read.
#ifndef LB2_PACK_CUH
#define LB2_PACK_CUH

#include "lb2_prims.cuh"
#include "lb2_common.h"

// ---- 16 read bases at a time (device: five aligned 32-bit loads per 16 bytes, SIMD-in-register classification) ----
// 2-bit codes of four ASCII bases in the byte lanes of w -> 8 bits (A=0 C=1 G=2 T=3; other characters give garbage)
LB2_DEV uint32_t lb2_codes4(uint32_t w) {
	const uint32_t x = (w >> 1) & 0x03030303u, c = x ^ ((x >> 1) & 0x01010101u);
	return (c | (c >> 6) | (c >> 12) | (c >> 18)) & 0xFFu;
}
LB2_DEV uint32_t lb2_bytemask4(uint32_t m) { return (((m & 0x01010101u) * 0x01020408u) >> 24) & 0xFu; }      // 0xFF/0x00 byte lanes -> 4 bits
// 16 bases: 2-bit codes (32 bits) and the mask of bases that are not one of ACGT
LB2_DEV void lb2_classify16(const char *s, uint32_t &codes, uint32_t &nacgt) {
	uint32_t sw[4]; lb2_load16(s, sw); codes = 0; nacgt = 0;
#pragma unroll
	for (int k = 0; k < 4; ++k) {
		const uint32_t ok = lb2_eq4(sw[k], 0x41414141u) | lb2_eq4(sw[k], 0x43434343u) | lb2_eq4(sw[k], 0x47474747u) | lb2_eq4(sw[k], 0x54545454u);
		nacgt |= lb2_bytemask4(~ok) << (4 * k); codes |= lb2_codes4(sw[k]) << (8 * k);
	}
}
// 16 qualities: bit i of the low half = quality i < thr_a, bit i of the high half = quality i < thr_b
LB2_DEV uint32_t lb2_low16x2(const char *q, uint32_t thr_a4, uint32_t thr_b4) {
	uint32_t qw[4]; lb2_load16(q, qw); uint32_t r = 0;
#pragma unroll
	for (int k = 0; k < 4; ++k) { r |= (lb2_bytemask4(lb2_ltu4(qw[k], thr_a4)) << (4 * k)) | (lb2_bytemask4(lb2_ltu4(qw[k], thr_b4)) << (16 + 4 * k)); }
	return r;
}

// words a read takes in the packed pool (reads of a million bases and more are cut: such a read is LB2_PK_TOOLONG anyway)
#ifdef __CUDACC__
__host__ __device__
#endif
static inline uint32_t lb2_pack_nwords(uint64_t len) { if (len > 0xFFFF0u) { len = 0xFFFF0u; } return (uint32_t)((len + 15u) >> 4); }

// One read by a group of LB2_GS lanes (all lanes of a warp call this together; act = false: the group has no read).
// woff: the read's first word in the packed pool.
// o0 / len64 / fl: the read's pool offset, length and flag byte (the caller has them at hand: no dependent loads here).
LB2_DEV void lb2_pack_read(const lb2_dev_batch &B, lb2_pkread *pk, uint32_t *pk_bits, uint16_t *pk_lowq, uint32_t qtrim4, uint32_t qcall4, bool act, uint32_t r, uint32_t woff,
                           uint64_t o0, uint64_t len64, uint8_t fl)
{
	const uint32_t gl = lb2_glane();
	uint32_t len = 0;
	if (act) { len = len64 > 0xFFFF0u ? 0xFFFF0u : (uint32_t)len64; } else { o0 = 0; }
	const char *s = B.seq + o0, *q = B.qual + o0;
	const uint32_t nch = (len + 15u) >> 4;
	uint32_t first = 0xFFFFFFFFu, last = 0, my_nac = 0, my_lowc = 0, my_c = 0xFFFFFFFFu;      // last = index of the last good base + 1
	for (uint32_t c = gl; c < nch; c += LB2_GS) {
		const uint32_t m = len - c * 16, valid = (m < 16) ? ((1u << m) - 1u) : 0xFFFFu;
		uint32_t codes, nac; lb2_classify16(s + c * 16, codes, nac); nac &= valid;
		const uint32_t lw = lb2_low16x2(q + c * 16, qtrim4, qcall4), lowt = lw & valid, lowc = (lw >> 16) & valid;
		if (m < 16) { codes &= (1u << (2 * m)) - 1u; }
		pk_bits[woff + c] = codes; pk_lowq[woff + c] = (uint16_t)lowc;
		const uint32_t good = valid & ~(nac | lowt);
		if (good) { const uint32_t f = c * 16 + (uint32_t)lb2_ctz32(good), l = c * 16 + 32u - (uint32_t)lb2_clz32(good); if (f < first) { first = f; } if (l > last) { last = l; } }
		my_nac = nac; my_lowc = lowc; my_c = c;
	}
	first = lb2_gmin(first); last = lb2_gmax(last);
	uint32_t junk = (first == 0xFFFFFFFFu) ? 1u : 0u, lowq = 0;
	if (!junk) {      // a non-ACGT base strictly inside the kept stretch makes the read junk; a low-quality base inside it is remembered
		for (uint32_t c = gl; c < nch; c += LB2_GS) {
			uint32_t nac = my_nac, lowc = my_lowc;
			if (c != my_c) {      // (reads longer than 16 * LB2_GS bases: the other chunks are classified again)
				const uint32_t m = len - c * 16, valid = (m < 16) ? ((1u << m) - 1u) : 0xFFFFu; uint32_t codes;
				lb2_classify16(s + c * 16, codes, nac); nac &= valid; lowc = (lb2_low16x2(q + c * 16, qtrim4, qcall4) >> 16) & valid;
			}
			for (uint32_t x = nac; x; x &= x - 1) { const uint32_t p = c * 16 + (uint32_t)lb2_ctz32(x); if (p > first && p + 1 < last) { junk = 1; } }
			for (uint32_t x = lowc; x; x &= x - 1) { const uint32_t p = c * 16 + (uint32_t)lb2_ctz32(x); if (p >= first && p < last) { lowq = 1; } }
		}
	}
	junk = lb2_gor(junk); lowq = lb2_gor(lowq);
	if (act && gl == 0) {
		uint32_t n = junk ? 0u : last - first; const uint32_t t5 = junk ? len : first;
		uint32_t info = ((fl & LB2_READ_NORMAL) ? 2u : 0u) | ((fl & LB2_READ_REVERSE) ? 1u : 0u) | (((fl >> LB2_READ_MATE_SHIFT) & 3u) << 2);
		if (fl & LB2_READ_UNMAPPED) { info |= LB2_PK_UNMAPPED; }
		if (n > 4095u) { info |= LB2_PK_TOOLONG; n = 0; }
		lb2_pkread rec; rec.woff = woff; rec.t5 = (uint16_t)(t5 > 0xFFFFu ? 0xFFFFu : t5); rec.n = (uint16_t)n; rec.nw = (uint16_t)(nch > 0xFFFFu ? 0xFFFFu : nch);
		rec.info = (uint8_t)info; rec.lowq = (uint8_t)((lowq && !junk) ? 1 : 0);
		pk[r] = rec;
	}
}

#endif
