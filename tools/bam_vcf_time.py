"""BAM -> VCF wall time of the two command lines on one BAM pair: the unmodified reference CLI
(oracle/_ref/lancet --num-threads <cores>) and lancet_b200_cli, best of `reps` runs each (warm page cache), VCFs compared.
usage: python tools/bam_vcf_time.py <dir with tumor.bam normal.bam ref.fa> <chr:beg-end> [reps]
(without a directory: python tools/bam_vcf_time.py gen <region_bp> writes a synthetic pair to a temporary directory first)"""
import json, os, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lancet_b200 import simbam

if sys.argv[1] == "gen":
    region = int(sys.argv[2]); d = tempfile.mkdtemp(prefix="lb2_bamvcf_")
    simbam.write_dataset(d, seed=500, chroms=(("chr22", region),), var_every=700, som_every=1500); reg = f"chr22:1-{region}"
else:
    d, reg = sys.argv[1], sys.argv[2]
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
cores = os.cpu_count() or 1
base = ["--tumor", os.path.join(d, "tumor.bam"), "--normal", os.path.join(d, "normal.bam"), "--ref", os.path.join(d, "ref.fa"), "--reg", reg, "--num-threads", str(cores)]


def run(cmd, n, env=None):
    best, out, err = None, None, None
    for _ in range(n):
        t0 = time.perf_counter(); r = subprocess.run(cmd, capture_output=True, text=True, env=env); dt = time.perf_counter() - t0
        if r.returncode != 0:
            raise SystemExit(f"{cmd[0]} failed: {r.stderr[-400:]}")
        if best is None or dt < best:
            best = dt
        out, err = r.stdout, r.stderr
    return best, out, err


t_our, v_our, e_our = run([os.path.join(ROOT, "lancet_b200", "lancet_b200_cli")] + base, reps, dict(os.environ, LB2_CLI_TIMING="1"))
t_ref, v_ref, _ = run([os.path.join(ROOT, "oracle", "_ref", "lancet")] + base, 1)
beg, end = (int(x) for x in reg.split(":")[1].split("-"))
nwin = len(range(0, end - beg + 1, 100))
print(json.dumps({"region": reg, "windows": nwin, "records": sum(1 for l in v_ref.splitlines() if not l.startswith("#")), "host_threads": cores,
                  "reference_s": t_ref, "ours_s": t_our, "speedup": t_ref / t_our, "vcf_identical": simbam.normalise_vcf(v_ref) == simbam.normalise_vcf(v_our),
                  "windows_per_s_reference": nwin / t_ref, "windows_per_s_ours": nwin / t_our,
                  "ours_breakdown": [l for l in e_our.splitlines() if l.startswith("[timing] open")]}))
