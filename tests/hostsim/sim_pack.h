// DEBUG-ONLY (see hostsim.cc): the read-pool pre-pack pass of lancet_b200/csrc/lb2_pack.cuh run serially on the host
#ifndef LB2_SIM_PACK_H
#define LB2_SIM_PACK_H
#include <vector>
struct SimPack { std::vector<lb2_pkread> pk; std::vector<uint32_t> bits; std::vector<uint16_t> lowq; };
static void sim_pack(lb2_dev_batch &B, uint32_t R, const lb2_params &P, SimPack &sp)
{
	uint64_t words = 0;
	for (uint32_t r = 0; r < R; ++r) { words += lb2_pack_nwords(B.base_off[r + 1] - B.base_off[r]); }
	sp.pk.assign((size_t)R + 1, lb2_pkread()); sp.bits.assign(words + 64, 0xA5A5A5A5u); sp.lowq.assign(words + 64, 0xA5A5u);      // (pad words carry junk on the device as well)
	const uint32_t qt = ((uint32_t)P.min_qual_trim & 0xFFu) * 0x01010101u, qc = ((uint32_t)P.min_qual_call & 0xFFu) * 0x01010101u;
	uint32_t woff = 0;
	for (uint32_t r = 0; r < R; ++r) {
		lb2_pack_read(B, sp.pk.data(), sp.bits.data(), sp.lowq.data(), qt, qc, true, r, woff, B.base_off[r], B.base_off[r + 1] - B.base_off[r], B.flags[r]);
		woff += lb2_pack_nwords(B.base_off[r + 1] - B.base_off[r]);
	}
	B.pk = sp.pk.data(); B.pk_bits = sp.bits.data(); B.pk_lowq = sp.lowq.data();
}
#endif
