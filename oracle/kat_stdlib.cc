// TEST INFRASTRUCTURE ONLY.  Known answers from the REAL libstdc++ (the library the reference is compiled
// against): std::hash<std::string> values and std::unordered_map<std::string,int> iteration order after a
// scripted sequence of inserts / erases.  stdin:  lines "H <string>" | "I <string>" | "E <string>" | "O"
// stdout: for H the hash (decimal); for O the current iteration order (space separated keys).
#include <iostream>
#include <string>
#include <unordered_map>
#include <functional>
int main() {
	std::unordered_map<std::string, int> m; std::string op, s;
	while (std::cin >> op) {
		if (op == "H") { std::cin >> s; std::cout << std::hash<std::string>()(s) << "\n"; }
		else if (op == "I") { std::cin >> s; m.insert(std::make_pair(s, 0)); }
		else if (op == "E") { std::cin >> s; m.erase(s); }
		else if (op == "O") { bool f = true; for (auto &kv : m) { std::cout << (f ? "" : " ") << kv.first; f = false; } std::cout << "\n"; }
	}
	return 0;
}
