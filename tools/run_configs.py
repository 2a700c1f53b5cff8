"""BASELINE.json configs[2..4] at the size one GPU's share has (or a stated fraction of it): throughput resident and end to
end, windows not assembled, and parity of a random sample of windows against the compiled reference.

  python tools/run_configs.py chr22_shard   # configs[2]: whole chr22 (50.8 Mb) on 8 GPUs -> one GPU's 6.35 Mb share
  python tools/run_configs.py str           # configs[3]: STR-heavy, 80x/80x, k 11..101 (1 Mb of the 10 Mb region)
  python tools/run_configs.py wgs30         # configs[4]: 30x/30x (1 Mb sample of one GPU's share)
  python tools/run_configs.py err           # 0.5 % substitution errors, 1 Mb (the bubble-rich regime)
Prints one JSON line per configuration (kept under profiles/)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
from lancet_b200.api import Context
from lancet_b200.synth import make_batch

CONFIGS = {
    "chr22_shard": dict(desc="configs[2] whole chr22 60x/60x on 8 GPUs: one GPU's 6.35 Mb share", kw=dict(seed=2000, region_len=6_350_000, region_start=1_000_001, var_every=5000)),
    "str": dict(desc="configs[3] STR-heavy 80x/80x, k 11..101: 1 Mb of the 10 Mb region", kw=dict(seed=2001, region_len=1_000_000, region_start=1_000_001, str_every=200, cov_t=80, cov_n=80, var_every=5000)),
    "wgs30": dict(desc="configs[4] WGS 30x/30x: 1 Mb sample of one GPU's share", kw=dict(seed=2002, region_len=1_000_000, region_start=1_000_001, cov_t=30, cov_n=30, var_every=5000)),
    "err": dict(desc="0.5 % substitution errors, 60x/60x, 1 Mb", kw=dict(seed=2003, region_len=1_000_000, region_start=1_000_001, err=0.005, var_every=5000)),
    "realistic": dict(desc="true pairs + 4 % low-quality bases + 0.2 % errors, 60x/60x, 1 Mb (what a real BAM looks like to the kernels)", kw=dict(seed=2005, region_len=1_000_000, region_start=1_000_001, paired=True, low_qual_frac=0.04, err=0.002, var_every=5000)),
    "paired": dict(desc="true pairs (insert 300+-30, both mates in the window), 60x/60x, 1 Mb", kw=dict(seed=2004, region_len=1_000_000, region_start=1_000_001, paired=True, var_every=5000)),
}


def main():
    import run_ref
    for name in sys.argv[1:] or ["str", "wgs30", "err"]:
        c = CONFIGS[name]
        t0 = time.time(); b = make_batch(**c["kw"]); t_gen = time.time() - t0
        ctx = Context(device=0)
        ctx.upload(b); ctx.run(); ctx.wait()
        ms = []
        for _ in range(3):
            ctx.run(); ctx.wait(); ms.append(ctx.last_kernel_ms)
        res = ctx.download()
        step_ms = {k: ctx.last_kernel_ms_of(k) for k in ("pack", "windows", "escalation", "compaction")}
        t0 = time.perf_counter(); r2 = ctx.process(b); e2e = time.perf_counter() - t0
        st = res.windows["status"]
        out = {"config": name, "desc": c["desc"], "windows": b.n_windows, "reads": b.n_reads, "gen_s": round(t_gen, 1), "kernel_ms": min(ms),
               "windows_per_s": b.n_windows / (min(ms) * 1e-3), "e2e_pageable_windows_per_s": b.n_windows / e2e,
               "assembled": int((st == 0).sum()), "skipped_repeat": int((st == 1).sum()), "no_reads": int((st == 2).sum()), "not_assembled": int((st >= 3).sum()),
               "k_tried_mean": float(res.windows["n_k_tried"].mean()), "records": len(res.variants),
               "step_ms": step_ms}
        if run_ref.available():
            rng = np.random.default_rng(1); pick = np.sort(rng.choice(b.n_windows, min(400, b.n_windows), replace=False))
            sub = b.subset(pick)
            t0 = time.time(); want, _ = run_ref.run(sub, threads=os.cpu_count() or 8); t_ref = time.time() - t0
            got = ctx.process(sub).records()
            out["parity_sample"] = {"windows": len(pick), "records": len(want), "identical": got == want, "reference_s": round(t_ref, 1),
                                    "reference_windows_per_s": len(pick) / t_ref, "threads": os.cpu_count()}
        print(json.dumps(out), flush=True)
        ctx.close()


if __name__ == "__main__":
    main()
