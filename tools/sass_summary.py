"""SASS summary of lancet_b200/_lb2.so (cuobjdump -sass): per kernel, instruction count and the mnemonics that matter for
the design claims -- bulk-async copies (UBLKCP) and their mbarrier (SYNCS), generic vs global vs shared vs local memory
instructions, barriers, atomics/reductions.  usage: python tools/sass_summary.py [lib.so] > profiles/rNN_sass_summary.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "lancet_b200", "_lb2.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
kern = None; counts = collections.OrderedDict()
KEYS = ["UBLKCP", "SYNCS", "UTMALDG", "LDGSTS", "LDG", "STG", "LD", "ST", "LDS", "STS", "LDL", "STL", "ATOMS", "ATOMG", "ATOM", "RED", "BAR", "R2UR", "SHFL", "VOTE", "MATCH", "POPC", "LOP3", "SHF"]
for ln in txt.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        kern = m.group(1); counts[kern] = collections.Counter(); continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
    if m and kern:
        op = m.group(1); base = op.split(".")[0]
        counts[kern]["_total"] += 1
        counts[kern][base] += 1
        if base in ("UBLKCP", "SYNCS"):
            counts[kern][op] += 1
print(f"# cuobjdump -sass {os.path.relpath(so, ROOT)}  (sm_100a; compiled with -lineinfo -O3)")
for k, c in counts.items():
    name = subprocess.run(["c++filt", k], capture_output=True, text=True).stdout.strip().split("(")[0]
    print(f"\n{name}: {c['_total']} instructions")
    print("  " + "  ".join(f"{key}={c[key]}" for key in KEYS if c[key]))
    ext = [f"{op}={n}" for op, n in c.items() if "." in op]
    if ext:
        print("  " + "  ".join(sorted(ext)))
