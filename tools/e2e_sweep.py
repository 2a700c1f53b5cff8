"""End-to-end (host buffers in, records out) time of lb2_process on the bench workload for a few upload-segment settings."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from lancet_b200.api import Context
from lancet_b200.synth import make_batch
b = make_batch(seed=1000, region_len=int(os.environ.get("LB2_BENCH_REGION", 1_000_000)), region_start=1_000_001, var_every=5000)
pins = []
for name in ("ref_off", "ref_start", "chr_id", "wr_off", "wr_idx", "base_off", "flags", "name_rank", "ref_seq", "seq", "qual"):
    a = getattr(b, name); t = torch.empty(max(a.nbytes, 1), dtype=torch.uint8, pin_memory=True); v = t.numpy()[:a.nbytes].view(a.dtype); v[...] = a; setattr(b, name, v); pins.append(t)
ctx = Context(device=0)
ctx.upload(b); ctx.run(); ctx.wait(); res_ms = ctx.last_kernel_ms
out = {"resident_ms": res_ms}
for first, minw in (("4194304", "512"), ("12582912", "768"), ("33554432", "768"), ("67108864", "2048"), ("1073741824", "768")):
    os.environ["LB2_SEG_FIRST"] = first; os.environ["LB2_SEG_MIN_WINDOWS"] = minw
    ctx.process(b)
    ts = []
    for _ in range(5):
        t0 = time.perf_counter(); ctx.process(b); ts.append((time.perf_counter() - t0) * 1e3)
    out[f"first={int(first) >> 20}MiB,minw={minw}"] = round(min(ts), 2)
print(json.dumps(out))
