#!/usr/bin/env python
"""bench.py -- micro-assembly windows/sec of the Lancet hot path on B200 (BASELINE.json metric).

A "step" = one pass of the per-window micro-assembly pipeline over one batch of synthetic windows:
the configuration BASELINE.json quotes the metric on (configs[1]: a 1 Mb region, synthetic 60x/60x
tumour/normal, 100 bp reads, default k-sweep 11..101) -> ~10 000 windows of 600 bp, ~1.2 M reads.

  value  : windows/s, whole job, batch already resident in HBM when the timed region starts
  e2e    : windows/s through the C ABI call lb2_process() with HOST (pinned) buffers: H2D + kernels + D2H
  roofline / cpu_baseline : see DESIGN.md
  --impl reference : the reference's own CPU implementation (oracle/_ref/ref_windows = unmodified
                     nygenome/lancet sources behind the same window-batch boundary), all host threads

N > 1: one process per GPU (torchrun), ONE workload of N Mb (N regions of 1 Mb laid end to end) sharded by contiguous
window ranges: rank r assembles the windows of Mb r (weak scaling: fixed work per GPU).  Windows are independent, so
there is no collective on the per-window path; the one exchange is the gather of the variant records on rank 0 over
NCCL (C ABI lb2_comm_gather: ncclAllGather of counts, ncclSend/ncclRecv of the payloads), timed inside `e2e`.
Timing = max over ranks.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import numpy as np  # noqa: E402

REGION = int(os.environ.get("LB2_BENCH_REGION", 1_000_000))
METRIC = "microassembly windows/sec (whole box)"
WORKLOAD = f"configs[1]: chr22 {REGION // 1000} kb region, synthetic 60x/60x T/N, 100 bp reads, err 0.1%, default k-sweep 11..101, 600 bp windows / 100 bp stride"


def make_workload(rank: int):
    from lancet_b200.synth import make_batch
    return make_batch(seed=1000 + rank, region_len=REGION, region_start=1_000_001, var_every=5000)


class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index; self.rows = []; self._stop_ev = threading.Event()

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self._stop_ev.is_set():
            try:
                o = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                   capture_output=True, text=True, timeout=5).stdout.strip()
                if o:
                    self.rows.append([x.strip() for x in o.split(",")])
            except Exception:
                pass
            self._stop_ev.wait(0.2)

    def summary(self):
        self._stop_ev.set(); self.join(timeout=3)
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(int(r[0]) for r in self.rows if r[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in self.rows if len(r) > 2 + i)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": int(self.rows[0][1]) if self.rows[0][1].isdigit() else None,
                "reasons": reasons, "samples": len(self.rows)}


def peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


def bam_vcf_leg(cores: int, device: int):
    """BAM -> VCF wall time of the two command lines on the same synthetic BAM pair: the unmodified reference
    CLI (oracle/_ref/lancet --num-threads <cores>) and lancet_b200_cli (BAI region fetch, multi-threaded inflate/decode,
    batches through lb2_process), second of two runs each (warm page cache), VCFs compared byte for byte."""
    import shutil
    import tempfile
    from lancet_b200 import simbam
    refcli = os.path.join(ROOT, "oracle", "_ref", "lancet"); cli = os.path.join(ROOT, "lancet_b200", "lancet_b200_cli")
    if not (os.path.exists(refcli) and os.path.exists(cli) and os.path.exists(os.path.join(ROOT, "oracle", "_ref", "test_view"))):
        return {"unavailable": "needs oracle/_ref/{lancet,test_view} and lancet_b200/lancet_b200_cli"}
    region = int(os.environ.get("LB2_BAM_REGION", 100_000))
    td = tempfile.mkdtemp(prefix="lb2_bamvcf_")
    try:
        d = simbam.write_dataset(td, seed=500, chroms=(("chr22", region),), var_every=700, som_every=1500)
        base = ["--tumor", d["tumor"], "--normal", d["normal"], "--ref", d["ref"], "--reg", f"chr22:1-{region}", "--num-threads", str(cores)]

        def run(cmd):
            best, out = None, None
            for _ in range(2):
                t0 = time.perf_counter(); r = subprocess.run(cmd, capture_output=True, text=True, timeout=900); dt = time.perf_counter() - t0
                if r.returncode != 0:
                    return None, r.stderr[-400:]
                best, out = dt, r.stdout
            return best, out
        t_ref, v_ref = run([refcli] + base)
        t_our, v_our = run([cli] + base + ["--gpu", str(device)])
        if t_ref is None or t_our is None:
            return {"error": (v_ref if t_ref is None else v_our)}
        same = simbam.normalise_vcf(v_ref) == simbam.normalise_vcf(v_our)
        nwin = len(range(0, region, 100))
        return {"region_bp": region, "windows": nwin, "records": sum(1 for l in v_ref.splitlines() if not l.startswith("#")),
                "reference_s": t_ref, "reference_threads": cores, "ours_s": t_our, "speedup": t_ref / t_our, "vcf_identical": same,
                "windows_per_s_reference": nwin / t_ref, "windows_per_s_ours": nwin / t_our}
    finally:
        shutil.rmtree(td, ignore_errors=True)


def reference_arm(args, rank):
    """CPU reference: unmodified lancet sources (oracle/_ref/ref_windows), all host threads, bounded sample."""
    if rank != 0:
        return
    import run_ref
    cores = os.cpu_count() or 1
    if not run_ref.available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/ref_windows not built (needs /root/reference at build time)"}))
        return
    batch = make_workload(0)
    path = "/tmp/lb2_bench_ref.lb2b"
    # bounded sample: the first `count` windows of the workload (~1 s per step per 16 windows/thread)
    count = min(batch.n_windows, int(os.environ.get("LB2_REF_SAMPLE", 24 * cores)))
    batch.subset(np.arange(count)).save(path)
    times = []
    for i in range(args.warmup + args.steps):
        _, t = run_ref.run(path=path, threads=cores, want_records=False)
        if i >= args.warmup:
            times.append(t["best_s"])
    sec = sum(times) / len(times)
    val = count / sec
    sample = f"first {count} of {batch.n_windows} windows of the workload, {cores} threads, per step"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": "windows/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8/int32", "data": "synthetic", "config": {"workload": WORKLOAD}, "sample": sample,
        "cpu_baseline": {"value": val, "unit": "windows/s", "cores": cores, "kind": "reference", "sample": sample},
        "e2e": {"value": val, "unit": "windows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        reference_arm(args, rank)
        return
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the lancet_b200 hot path has no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")      # stdout carries the one JSON line and nothing else
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from lancet_b200.api import Context

    batch = make_workload(rank)
    # pinned host staging so that the e2e leg measures PCIe, not pageable-memory bounce buffers
    for name in ("ref_off", "ref_start", "chr_id", "wr_off", "wr_idx", "base_off", "flags", "name_rank", "ref_seq", "seq", "qual"):
        a = getattr(batch, name)
        t = torch.empty(max(a.nbytes, 1), dtype=torch.uint8, pin_memory=True)
        v = t.numpy()[:a.nbytes].view(a.dtype); v[...] = a
        setattr(batch, name, v); batch.__dict__.setdefault("_pins", []).append(t)
    ctx = Context(device=local)
    if world > 1:
        from lancet_b200.shard import init_comm
        init_comm(ctx, device=torch.device("cuda", local))
    win_offset = rank * batch.n_windows      # this rank's share of the one N Mb workload (every Mb tiles into the same number of windows)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- resident-in-HBM leg -------------------------------------------------------------------
    ctx.upload(batch)
    for _ in range(args.warmup):
        ctx.run(); ctx.wait()
    res = ctx.download()
    n_var = len(res.variants); st = res.windows["status"]
    n_ok = int((st == 0).sum()); n_fail = int((st >= 3).sum())
    launches0 = ctx.kernel_launches
    sampler = ClockSampler(local); sampler.start()
    barrier()
    dev_ms = 0.0; part_ms = {"pack": 0.0, "windows": 0.0, "escalation": 0.0, "compaction": 0.0}
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ctx.run(); ctx.wait(); dev_ms += ctx.last_kernel_ms       # CUDA events on the launching stream
        for k in part_ms:
            part_ms[k] += ctx.last_kernel_ms_of(k)
    barrier()
    wall = time.perf_counter() - t0
    launches = ctx.kernel_launches - launches0
    ms_step = dev_ms / args.steps
    # ---- end-to-end leg: host buffers -> H2D -> kernels -> D2H, through lb2_process ---------------
    gathered = 0
    for _ in range(1):
        r = ctx.process(batch)
        if world > 1:
            ctx.comm_gather(r.variants, r.strings, window_offset=win_offset)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        r = ctx.process(batch)
        if world > 1:      # the exchange step of the sharded path: every rank's records to rank 0
            gv, gs, _st = ctx.comm_gather(r.variants, r.strings, window_offset=win_offset)
            gathered = len(gv) if gv is not None else 0
    barrier()
    e2e_s = (time.perf_counter() - t0) / args.steps
    h2d, d2h = ctx.last_h2d_bytes, ctx.last_d2h_bytes
    clocks = sampler.summary()

    tms = torch.tensor([ms_step, e2e_s * 1e3, wall * 1e3 / args.steps], device="cuda", dtype=torch.float64)
    tot = torch.tensor([batch.n_windows, n_ok, n_fail, n_var], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX); dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    ms_step, e2e_ms, wall_ms = [float(x) for x in tms.tolist()]
    total_windows, n_ok_all, n_fail_all, n_var_all = [int(x) for x in tot.tolist()]

    if rank == 0:
        peak, how = peak_hbm()
        from lancet_b200.api import KERNEL_VERSION
        alg_bytes = batch.algorithmic_bytes(n_var)                  # this rank's launch (B_win summed)
        win_ms = part_ms["windows"] / args.steps                    # lb2_window_kernel alone (first pass), CUDA events on its stream
        achieved = alg_bytes / (win_ms * 1e-3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tp):
            try:
                tj = json.load(open(tp))            # ncu capture of this kernel version (per-window DRAM bytes x windows of this launch)
                if tj.get("kernel_version") == KERNEL_VERSION:      # a capture of another kernel version says nothing about this one
                    traffic = tj["dram_bytes_per_window"] * batch.n_windows
            except Exception:
                pass
        # the pre-pack pass is the one HBM-shaped kernel of the path: 2 bytes in per base, 3/8 byte + 12 bytes per read out
        nbases = int(batch.seq.nbytes)
        pack_bytes = 2 * nbases + int(((batch.base_off[1:] - batch.base_off[:-1] + 15) // 16).sum()) * 6 + 12 * batch.n_reads + 8 * batch.n_reads
        pack_ms = part_ms["pack"] / args.steps
        out = {
            "metric": METRIC, "value": total_windows / (ms_step * 1e-3), "unit": "windows/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8/int32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "windows_per_gpu": batch.n_windows, "reads_per_gpu": batch.n_reads,
                       "windows_ok": n_ok_all, "windows_failed": n_fail_all, "variant_records": n_var_all,
                       "l2": f"inputs ({(batch.seq.nbytes * 2) >> 20} MiB per step) larger than L2",
                       "parallelism": (f"one {world} Mb workload window-sharded x{world} (rank r = Mb r); no collective per window; records gathered on rank 0 by NCCL "
                                       f"(ncclAllGather counts + ncclSend/ncclRecv payloads) inside e2e: {gathered} records per step") if world > 1 else "1 GPU",
                       "resident_ctas": ctx.resident_ctas, "smem_per_cta": ctx.smem_per_cta, "wall_ms_per_step": wall_ms},
            "e2e": {"value": total_windows / (e2e_ms * 1e-3), "unit": "windows/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "peak_source": how, "algorithmic_bytes_per_launch": alg_bytes, "kernel": "lb2_window_kernel", "kernel_ms": win_ms,
                         "kernel_version": KERNEL_VERSION},
            "roofline_pack": {"bound": "hbm", "kernel": "lb2_pack_kernel (+count, scan)", "achieved": pack_bytes / (pack_ms * 1e-3) / 1e9 if pack_ms > 0 else None,
                              "peak": peak, "unit": "GB/s", "frac": (pack_bytes / (pack_ms * 1e-3) / 1e9 / peak) if pack_ms > 0 else None,
                              "algorithmic_bytes_per_launch": pack_bytes, "kernel_ms": pack_ms},
            "step_ms": {k: v / args.steps for k, v in part_ms.items()},
        }
        if world == 1 and os.environ.get("LB2_SKIP_CPU_BASELINE") is None:
            import run_ref
            cores = os.cpu_count() or 1
            if run_ref.available():
                count = min(batch.n_windows, int(os.environ.get("LB2_REF_SAMPLE", 24 * cores)))
                path = "/tmp/lb2_bench_cpu.lb2b"
                batch.subset(np.arange(count)).save(path)
                _, t = run_ref.run(path=path, threads=cores, want_records=False)
                out["cpu_baseline"] = {"value": count / t["best_s"], "unit": "windows/s", "cores": cores, "kind": "reference",
                                       "sample": f"first {count} of {batch.n_windows} windows, {cores} threads, {t['best_s']:.1f} s"}
            else:
                out["cpu_baseline"] = {"value": None, "unit": "windows/s", "cores": cores, "kind": "reference", "sample": "oracle/_ref not built"}
            if os.environ.get("LB2_SKIP_BAM_VCF") is None:
                try:
                    out["e2e_bam_vcf"] = bam_vcf_leg(cores, local)
                except Exception as e:      # the headline line must not depend on this leg
                    out["e2e_bam_vcf"] = {"error": repr(e)[:300]}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
