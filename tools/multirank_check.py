"""Run under torchrun (one rank per GPU): ONE synthetic workload, every rank assembles its contiguous share of the windows,
the records are gathered on rank 0 over NCCL (lb2_comm_gather) and compared with the compiled reference's records for
the whole workload.  Used by tests/test_gpu_multirank.py; prints one JSON line on rank 0."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from lancet_b200.api import Context, Result
from lancet_b200.shard import window_range, init_comm
from lancet_b200.synth import make_batch
region = int(sys.argv[1]) if len(sys.argv) > 1 else 40000
b = make_batch(seed=77, region_len=region, var_every=600)
lo, hi = window_range(b.n_windows, rank, world)
ctx = Context(device=local)
init_comm(ctx, device=torch.device("cuda", local))
res = ctx.process(b.subset(np.arange(lo, hi)))
failed = int((res.windows["status"] >= 3).sum())
gv, gs, st = ctx.comm_gather(res.variants, res.strings, window_offset=lo, stats=(failed, hi - lo))
if rank == 0:
    import run_ref
    class _R(Result):
        def __init__(self, v, s): self.variants, self.strings = v, s
    got = _R(gv, gs).records()
    want, _ = run_ref.run(b, threads=os.cpu_count() or 8)
    print(json.dumps({"same": got == want, "records": len(want), "windows": st[1], "failed": st[0], "world": world}))
dist.destroy_process_group()
