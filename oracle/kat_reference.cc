// TEST INFRASTRUCTURE ONLY.  Golden-vector generator: calls the reference's OWN functions
// (isRepeat, isAlmostRepeat, findTandems: src/util.cc; global_align_aff: src/align.cc; Graph_t::trim via
// addAlignment: src/Graph.cc) on scripted inputs.  Linked against the objects oracle/Makefile compiles from
// /root/reference.  stdin lines:
//   R <K> <seq>                 -> isRepeat isAlmostRepeat(max=2)
//   A <S> <T>                   -> S_aln T_aln
//   T <pos> <seq>               -> ans LEN MOTIF   (findTandems with the reference defaults 4,3,7,1)
//   Q <seq> <qual>              -> trm5 trm3 isjunk  (Graph_t::trim with MIN_QUAL_TRIM = 43)
#include "Lancet.hh"
int main() {
	std::string op;
	Graph_t g; g.setMinQualTrim(43);
	while (std::cin >> op) {
		if (op == "R") { int K; std::string s; std::cin >> K >> s; std::cout << isRepeat(s, K) << " " << isAlmostRepeat(s, K, 2) << "\n"; }
		else if (op == "A") { std::string S, T, a, b; std::cin >> S >> T; global_align_aff(S, T, a, b, 0, 0); std::cout << a << " " << b << "\n"; }
		else if (op == "T") { int pos; std::string s; std::cin >> pos >> s; int len = 0; std::string motif; bool ans = findTandems(s, "t", 4, 3, 7, 1, pos, len, motif);
			std::cout << ans << " " << len << " " << (motif.empty() ? "." : motif) << "\n"; }
		else if (op == "Q") { std::string s, q; std::cin >> s >> q; g.addAlignment("x", "n", s, q, 0, 'M', TMR, FWD, "", 0);
			ReadInfo_t &ri = g.readid2info.back(); std::cout << ri.trm5 << " " << ri.trm3 << " " << ri.isjunk << "\n"; }
	}
	return 0;
}
