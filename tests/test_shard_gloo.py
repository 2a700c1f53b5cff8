"""N>1 host logic on CPU: two gloo ranks shard a window range and gather variant records on rank 0."""
import os
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from lancet_b200.shard import window_range, gather_records
    dtype = np.dtype([("window", "<u4"), ("pos", "<i4"), ("str_off", "<u4"), ("ref_len", "<u2"), ("alt_len", "<u2"),
                      ("motif_len", "<u2"), ("str_len", "<u2"), ("c", "<u2", (8,)), ("code", "u1"), ("pbr", "u1"), ("pba", "u1"), ("kmer", "u1")])
    assert dtype.itemsize == 40
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = window_range(11, rank, world)
    n = hi - lo
    v = np.zeros(n * 2 if rank == 0 else n, dtype=dtype)           # different record counts per rank
    strings = b""
    for i in range(len(v)):
        s = f"r{rank}i{i}".encode()
        v[i]["window"] = i % n; v[i]["pos"] = 1000 * rank + i; v[i]["str_off"] = len(strings); v[i]["ref_len"] = len(s)
        strings += s
    gv, gs = gather_records(v, strings, lo)
    if rank == 0:
        q.put((gv.tobytes(), gs, (lo, hi)))
    dist.destroy_process_group()


def test_gather_two_ranks():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in ps]
    gvb, gs, _ = q.get(timeout=120)
    [p.join(timeout=60) for p in ps]
    assert all(p.exitcode == 0 for p in ps)
    dtype = np.dtype([("window", "<u4"), ("pos", "<i4"), ("str_off", "<u4"), ("ref_len", "<u2"), ("alt_len", "<u2"),
                      ("motif_len", "<u2"), ("str_len", "<u2"), ("c", "<u2", (8,)), ("code", "u1"), ("pbr", "u1"), ("pba", "u1"), ("kmer", "u1")])
    gv = np.frombuffer(gvb, dtype=dtype)
    assert len(gv) == 12 + 5                                     # rank 0: 6 windows x 2 records, rank 1: 5 windows x 1
    assert (gv["window"][:12] < 6).all() and (gv["window"][12:] >= 6).all() and gv["window"].max() == 10
    for r in gv:                                                  # strings were rebased correctly
        s = gs[int(r["str_off"]):int(r["str_off"]) + int(r["ref_len"])].decode()
        assert s.startswith("r0i" if r["pos"] < 1000 else "r1i")


def test_window_range_partitions():
    sys.path.insert(0, ROOT)
    from lancet_b200.shard import window_range
    for n in (0, 1, 7, 8, 9995):
        for w in (1, 2, 3, 8):
            rs = [window_range(n, r, w) for r in range(w)]
            assert rs[0][0] == 0 and rs[-1][1] == n and all(rs[i][1] == rs[i + 1][0] for i in range(w - 1))
