"""DEBUG TOOLING (not a test, not shipped): run the one-thread host simulation of the device sources
(tests/hostsim/hostsim.cc) and the compiled reference (oracle/_ref/ref_windows) on seeded synthetic batches and
diff the Variant_t tuples.  Lets control-flow changes to lancet_b200/csrc/*.cuh be checked in the GPU-less
container before they go to the B200 (the GPU parity tests in tests/ remain the gate).

usage: python tools/sim_parity.py [quick|full] [min_k,min_k,...]
"""
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
BUILD = os.path.join(ROOT, "tests", "hostsim", "_build")
SIM = os.path.join(BUILD, "hostsim")

QUICK = [
    dict(seed=11, region_len=3000),
    dict(seed=31, region_len=2500, err=0.005),
    dict(seed=32, region_len=2500, low_qual_frac=0.05, err=0.002),
    dict(seed=62, region_len=2000, paired=True, insert_mean=150, insert_sd=20),
    dict(seed=71, region_len=3000, str_every=200, cov_t=80, cov_n=80),
    dict(seed=34, region_len=3000, var_every=120),
]
FULL = QUICK + [
    dict(seed=23, region_len=6000),
    dict(seed=33, region_len=4000, cov_t=20, cov_n=15),
    dict(seed=35, region_len=3000, cov_t=150, cov_n=150),
    dict(seed=36, region_len=3000, read_len=150),
    dict(seed=37, region_len=1777),
    dict(seed=61, region_len=3000, paired=True),
    dict(seed=77, region_len=3000, paired=True, insert_mean=160, insert_sd=40, err=0.004, low_qual_frac=0.02),
    dict(seed=72, region_len=4000, str_every=500),
    dict(seed=75, region_len=1500, err=0.01),
    dict(seed=1000, region_len=20000, region_start=1_000_001, var_every=5000),
    dict(seed=81, region_len=6000, str_every=120, cov_t=60, cov_n=60, err=0.002),
    dict(seed=82, region_len=6000, var_every=60),
    dict(seed=83, region_len=4000, var_every=200, err=0.003, paired=True, insert_mean=220, insert_sd=30),
]


def build():
    os.makedirs(BUILD, exist_ok=True)
    src = os.path.join(ROOT, "tests", "hostsim", "hostsim.cc")
    csrc = os.path.join(ROOT, "lancet_b200", "csrc")
    deps = [src] + [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith((".cuh", ".h"))]
    if not os.path.exists(SIM) or any(os.path.getmtime(d) > os.path.getmtime(SIM) for d in deps):
        subprocess.run(["g++", "-O2", "-std=c++17", "-w", "-o", SIM, src], check=True)


def main():
    import run_ref
    from lancet_b200.synth import make_batch
    build()
    cases = FULL if (len(sys.argv) > 1 and sys.argv[1] == "full") else QUICK
    bad = 0
    ks = [11]
    if len(sys.argv) > 2:
        ks = [int(x) for x in sys.argv[2].split(",")]        # extra argument: comma-separated --min-k values (exercises the wider k-mer paths)
    for mink, kw in [(k, c) for k in ks for c in cases]:
        b = make_batch(**kw)
        with tempfile.TemporaryDirectory() as td:
            p = os.path.join(td, "b.lb2b"); b.save(p)
            t0 = time.time()
            want, _ = run_ref.run(path=p, threads=8, min_k=mink)
            t1 = time.time()
            out = os.path.join(td, "sim.tsv")
            r = subprocess.run([SIM, p, "--out", out, "--min-k", str(mink)], capture_output=True, text=True)
            t2 = time.time()
            got = run_ref.parse_tsv(open(out).read()) if r.returncode == 0 else None
        ok = got == want
        bad += 0 if ok else 1
        print(f"{'OK ' if ok else 'BAD'} k>={mink} {kw} windows={b.n_windows} records={len(want)} ref={t1 - t0:.1f}s sim={t2 - t1:.1f}s"
              + ("" if not r.stderr.strip() else "  [" + r.stderr.strip().splitlines()[0] + (" ..." if len(r.stderr.strip().splitlines()) > 1 else "") + "]"))
        if not ok and got is not None:
            sw, sg = set(want), set(got)
            for x in sorted(sw - sg)[:5]:
                print("   want-only", x)
            for x in sorted(sg - sw)[:5]:
                print("   got-only ", x)
    print("mismatching cases:", bad)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
