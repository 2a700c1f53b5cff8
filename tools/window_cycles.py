"""DEBUG TOOLING: per-window cycle counts (the -DLB2_PROFILE -DLB2_PROFILE_SEQ build, lancet_b200/_lb2_profseq.so):
the slowest windows of a synthetic workload, the distribution, and the phase shares of the slowest ones alone."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as _g
os.environ["LB2_SO"] = _g.build_profile(per_window=True)
from lancet_b200.api import Context
from lancet_b200.synth import make_batch
import numpy as np
region = int(sys.argv[1]); kw = json.loads(sys.argv[2])
kw = {**dict(seed=1000, region_start=1_000_001, var_every=5000), **kw}
b = make_batch(region_len=region, **kw)
ctx = Context(); ctx.upload(b); ctx.run(); ctx.wait(); ctx.run(); ctx.wait()
print("ms", {k: ctx.last_kernel_ms_of(k) for k in ("pack", "windows", "escalation", "compaction")})
res = ctx.download(); w = res.windows
cyc = w["detail"].astype(np.int64) * 256
order = np.argsort(-cyc)[:12]
print("total Mcycles", cyc.sum() / 1e6, "mean", cyc.mean() / 1e6)
for i in order:
    print(i, "Mcyc %.1f" % (cyc[i] / 1e6), "status", w["status"][i], "k", w["final_k"][i], "tried", w["n_k_tried"][i], "nodes", w["n_nodes"][i], "nvar", w["n_variants"][i], "reads", int(b.wr_off[i + 1] - b.wr_off[i]))
print(np.histogram(cyc / 1e6, bins=[0, 0.5, 1, 2, 4, 8, 16, 32, 64, 128, 1e9]))
sub = b.subset(sorted(int(i) for i in order))
c2 = Context(); c2.upload(sub); c2.run(); c2.wait(); c2.phase_cycles(reset=True); c2.run(); c2.wait()
pc = c2.phase_cycles(); tot = sum(pc.values()) or 1
print("slowest windows alone: kernel ms", c2.last_kernel_ms, {k: round(v / tot, 4) for k, v in pc.items() if v})
