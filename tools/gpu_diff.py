"""DEBUG TOOLING: run seeded batches through the CUDA path and the compiled reference, print per-window differences."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import run_ref
from lancet_b200.api import Context, Params
from lancet_b200.synth import make_batch

cases = [dict(seed=31, region_len=4000, err=0.005), dict(seed=75, region_len=1500, err=0.01), dict(seed=11, region_len=6000)]
if len(sys.argv) > 1:
    cases = [json.loads(a) for a in sys.argv[1:]]
for kw in cases:
    pk = {k: kw.pop(k) for k in list(kw) if k in ("min_k", "max_k")}
    b = make_batch(**kw)
    want, _ = run_ref.run(b, threads=8, **pk)
    c = Context(Params.default(**pk), device=0)
    res = c.process(b)
    got = res.records()
    bad = sorted({r[0] for r in set(want) ^ set(got)})
    print(kw, pk, "windows", b.n_windows, "records", len(want), len(got), "OK" if got == want else "BAD windows %s" % bad)
    for w in bad[:6]:
        print("  window", w, {k: int(res.windows[k][w]) for k in ("status", "final_k", "n_k_tried", "n_variants", "n_nodes", "detail")})
        for r in [r for r in want if r[0] == w and r not in got][:3]: print("    want", r)
        for r in [r for r in got if r[0] == w and r not in want][:3]: print("    got ", r)
    st = res.windows["status"]
    for w in [int(x) for x in (st >= 3).nonzero()[0][:12]]:
        print("  not assembled: window", w, {k: int(res.windows[k][w]) for k in ("status", "final_k", "n_k_tried", "n_variants", "n_nodes", "detail")})
    print("  status counts", {int(s): int((st == s).sum()) for s in set(st.tolist())})
    c.close()
