"""Parity at workload scale: the batches bench.py times (BASELINE.json configs[1]) and scaled-down shapes of
configs[3] (STR-heavy, 80x/80x, k 11..101) and of a high-error run, every Variant_t tuple compared with the compiled
reference (oracle/_ref/ref_windows, all host threads).  No window may come back OVERFLOW / UNSUPPORTED."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

WORKLOADS = {
    # exactly bench.py's make_workload(rank=0)
    "bench_1mb": dict(seed=1000, region_len=1_000_000, region_start=1_000_001, var_every=5000),
    # configs[3] shape: STR blocks (unit 1-6 bp x 5-40 copies) every ~200 bp, 80x/80x, default k sweep 11..101
    "str_100kb_80x": dict(seed=401, region_len=100_000, str_every=200, cov_t=80, cov_n=80),
    # 0.5 % substitution errors: thousands of nodes per window, bubbles everywhere, the BFS-heavy regime
    "err_100kb": dict(seed=402, region_len=100_000, err=0.005),
    # true pairs with overlapping mates at scale (Variant B of SURVEY §8d)
    "paired_100kb": dict(seed=403, region_len=100_000, paired=True, insert_mean=170, insert_sd=30, var_every=900),
}


def _first_diff(got, want):
    for i, (a, b) in enumerate(zip(got, want)):
        if a != b:
            return i, a, b
    return min(len(got), len(want)), None, None


@pytest.mark.timeout(900, method="thread")
@pytest.mark.parametrize("name", sorted(WORKLOADS))
def test_workload_matches_reference(name, ctx, tmp_path):
    import run_ref
    if not run_ref.available():
        pytest.skip("oracle/_ref/ref_windows not built")
    from lancet_b200.synth import make_batch
    b = make_batch(**WORKLOADS[name])
    p = str(tmp_path / "w.lb2b"); b.save(p)
    want, t = run_ref.run(path=p, threads=os.cpu_count() or 8)
    res = ctx.process(b)
    st = res.windows["status"]
    failed = np.nonzero(st >= 3)[0]
    assert len(failed) == 0, f"{len(failed)} windows not assembled: {res.windows[failed][:8]}"
    got = res.records()
    assert len(got) == len(want), (len(got), len(want), _first_diff(got, want))
    assert got == want, _first_diff(got, want)
    print(f"{name}: {b.n_windows} windows, {len(want)} records identical; reference {t['best_s']:.1f} s, device {res.kernel_ms:.1f} ms "
          f"({b.n_windows / (res.kernel_ms * 1e-3):.0f} windows/s), k tried mean {res.windows['n_k_tried'].mean():.2f}")
