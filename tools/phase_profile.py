"""Per-phase cycle shares of the window pipeline (lane-0 clock64 marks), for a synthetic region."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as _g
os.environ["LB2_SO"] = os.environ.get("LB2_PROFILE_SO") or _g.build_profile()      # the instrumented build (the product library has no counters)
from lancet_b200.api import Context
from lancet_b200.synth import make_batch

region = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
kw = json.loads(sys.argv[2]) if len(sys.argv) > 2 else {}
kw = {**dict(seed=1000, region_start=1_000_001, var_every=5000), **kw}
b = make_batch(region_len=region, **kw)
ctx = Context()
ctx.upload(b); ctx.run(); ctx.wait(); ctx.phase_cycles(reset=True)
t0 = time.perf_counter(); ctx.run(); ctx.wait(); dt = time.perf_counter() - t0
ms = ctx.last_kernel_ms
pc = ctx.phase_cycles()
tot = sum(pc.values()) or 1
res = ctx.download()
print(json.dumps({"windows": b.n_windows, "kernel_ms": ms, "windows_per_s": b.n_windows / (ms * 1e-3), "resident_ctas": ctx.resident_ctas,
                  "smem": ctx.smem_per_cta, "failed": int((res.windows["status"] >= 3).sum()),
                  "k_tried_mean": float(res.windows["n_k_tried"].mean()), "nodes_mean": float(res.windows["n_nodes"].mean()),
                  "cycles_per_window": tot / b.n_windows,
                  "share": {k: round(v / tot, 4) for k, v in pc.items() if v}}))
