"""DEBUG TOOLING: per-source-line summary of an ncu report (needs -lineinfo and --import-source on).

usage: python tools/ncu_lines.py report.ncu-rep [top_n]
Prints, per source line: stall samples (all / barrier / long_sb / short_sb / wait), warp instructions executed,
average active lanes; then the same summed per file.
"""
import csv
import subprocess
import sys
from collections import defaultdict

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 60
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
cur = None; hdr = None; lines = []
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]; continue
    if len(r) > 10 and r[0] == "Line No":
        hdr = r; continue
    if hdr and len(r) == len(hdr) and r[0] != "":
        d = dict(zip(hdr[4:], r[4:]))
        def f(k):
            try: return float(d.get(k, 0) or 0)
            except ValueError: return 0.0
        lines.append(dict(file=cur, line=int(r[0]), src=r[1].strip()[:110], samples=f("# Samples"), inst=f("Instructions Executed"),
                          tinst=f("Thread Instructions Executed"), bar=f("stall_barrier"), lsb=f("stall_long_sb"), ssb=f("stall_short_sb"),
                          wait=f("stall_wait"), br=f("stall_branch_resolving")))
tot_s = sum(l["samples"] for l in lines) or 1; tot_i = sum(l["inst"] for l in lines) or 1
print(f"total samples {tot_s:.0f}, warp instructions {tot_i:.0f}")
print("== by samples ==")
for l in sorted(lines, key=lambda l: -l["samples"])[:top]:
    print(f"{l['file']}:{l['line']:<5} smp {100*l['samples']/tot_s:5.2f}%  inst {100*l['inst']/tot_i:5.2f}%  lanes {l['tinst']/max(l['inst'],1):4.1f}  bar {l['bar']:.0f} lsb {l['lsb']:.0f} ssb {l['ssb']:.0f} wait {l['wait']:.0f} br {l['br']:.0f} | {l['src']}")
print("== by instructions ==")
for l in sorted(lines, key=lambda l: -l["inst"])[:top]:
    print(f"{l['file']}:{l['line']:<5} inst {100*l['inst']/tot_i:5.2f}%  smp {100*l['samples']/tot_s:5.2f}%  lanes {l['tinst']/max(l['inst'],1):4.1f} | {l['src']}")
print("== per file ==")
pf = defaultdict(lambda: [0, 0])
for l in lines:
    pf[l["file"]][0] += l["samples"]; pf[l["file"]][1] += l["inst"]
for k, v in sorted(pf.items(), key=lambda kv: -kv[1][0]):
    print(f"{k:24s} samples {100*v[0]/tot_s:5.1f}%  inst {100*v[1]/tot_i:5.1f}%")

# ---- barrier waits: the sampler charges a warp parked at a barrier to the instruction AFTER the BAR.SYNC; walk the
# ---- SASS in address order and charge those samples to the source line of the barrier itself
addr2line = {}
cur = None; hdr = None; curline = None
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]; continue
    if len(r) > 10 and r[0] == "Line No":
        hdr = r; continue
    if hdr and len(r) == len(hdr):
        if r[0] != "":
            curline = (cur, int(r[0]), r[1].strip()[:90])
        elif r[2].startswith("0x"):
            addr2line[r[2]] = curline
txt2 = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows2 = list(csv.reader(txt2.splitlines()))
h2 = rows2[1]; ia = h2.index("Address"); isrc = h2.index("Source"); ib = h2.index("stall_barrier")
# (every BAR.SYNC carries the line of the inlined lb2_sync(), so name the barrier by the lines of the code around it)
SKIP = ("BAR.SYNC", "WARPSYNC", "NOP", "BSYNC", "BSSY", "S2R", "S2UR", "UMOV", "R2UR")
agg = defaultdict(float); prev_real = None; at_bar = None
for r in rows2[2:]:
    if len(r) != len(h2):
        continue
    src_ = r[isrc]
    if "BAR.SYNC" in src_:
        at_bar = prev_real
    elif not any(k in src_ for k in SKIP):
        prev_real = r[ia]
    try:
        v = float(r[ib] or 0)
    except ValueError:
        v = 0
    if v > 50:
        before = addr2line.get(at_bar); after = addr2line.get(r[ia])
        agg[(before[:2] if before else None, after[:2] if after else None)] += v
print("== barrier waits: (code before the barrier, code after it) ==")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:30]:
    print(f"{100*v/tot_s:5.2f}%  before {k[0]}  after {k[1]}")
