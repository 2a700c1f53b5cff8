"""Multi-GPU plumbing of the window-sharded path: windows are independent, so each rank owns a
contiguous range of the global window index and the only exchange is the final gather of variant records to rank 0.
On GPUs that gather is the C ABI's lb2_comm_gather (NCCL: ncclAllGather of counts, ncclSend/ncclRecv of the payloads;
`init_comm` hands the NCCL id round through torch.distributed); `gather_records` is the same exchange on a
torch.distributed group (gloo on CPU: the host logic of tests/test_shard_gloo.py).  No collective sits on the
per-window path."""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def window_range(n_windows: int, rank: int, world: int):
    """contiguous shard [lo, hi) of the global window index for `rank`."""
    base, rem = divmod(n_windows, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_records(variants: np.ndarray, strings: bytes, window_offset: int, device=None):
    """Gather (variants, strings) of every rank on rank 0.

    `variants` is the structured array of lancet_b200.api.VARIANT_DTYPE with *local* window indices and string offsets;
    the result on rank 0 has global window indices (local + the rank's `window_offset`), rebased string offsets and is
    ordered by rank (= by global window index, since shards are contiguous ranges).  Other ranks get (None, None)."""
    rank, world = dist.get_rank(), dist.get_world_size()
    dev = device if device is not None else (torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu"))
    v = variants.copy()
    v["window"] += window_offset
    vb = np.frombuffer(v.tobytes(), dtype=np.uint8)
    sb = np.frombuffer(strings, dtype=np.uint8)
    sizes = torch.tensor([len(vb), len(sb)], dtype=torch.int64, device=dev)
    all_sizes = [torch.zeros(2, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(all_sizes, sizes)
    all_sizes = [tuple(int(x) for x in s.tolist()) for s in all_sizes]
    mv, ms = max(s[0] for s in all_sizes), max(s[1] for s in all_sizes)
    pad = torch.zeros(mv + ms + 1, dtype=torch.uint8, device=dev)
    if len(vb):
        pad[:len(vb)] = torch.from_numpy(vb.copy()).to(dev)
    if len(sb):
        pad[mv:mv + len(sb)] = torch.from_numpy(sb.copy()).to(dev)
    out = [torch.zeros_like(pad) for _ in range(world)] if rank == 0 else None
    dist.gather(pad, out, dst=0)
    if rank != 0:
        return None, None
    vs, ss, soff = [], [], 0
    for (nv, ns), t in zip(all_sizes, out):
        h = t.cpu().numpy()
        a = np.frombuffer(h[:nv].tobytes(), dtype=variants.dtype).copy()
        a["str_off"] += soff
        vs.append(a); ss.append(h[mv:mv + ns].tobytes()); soff += ns
    return np.concatenate(vs) if vs else variants[:0], b"".join(ss)


def init_comm(ctx, device=None):
    """bind the context's NCCL communicator to the torch.distributed world (rank 0 creates the id, everybody gets it)"""
    from .api import comm_unique_id
    rank, world = dist.get_rank(), dist.get_world_size()
    box = [comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0, device=device)
    ctx.comm_init(box[0], rank, world)
