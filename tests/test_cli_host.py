"""Host side of the `lancet` command line: window tiling, BAM decoding, read filters, active-region
prefilter, variant store, Fisher scores, filters and the VCF writer -- checked WITHOUT a GPU.

The CLI object is linked here against tests/hostsim/lb2_shim.cc (a one-thread build of the device sources behind the
same C ABI -- debug tooling, never shipped) so that its VCF can be compared byte for byte with the VCF the unmodified
reference CLI wrote for the same BAMs (tests/golden/e2e_*/expected.vcf, made by tests/golden/make_e2e_golden.py).
The shipped binary (lancet_b200/lancet_b200_cli) links the CUDA library instead and is covered by tests/test_gpu_cli.py.
"""
import hashlib
import json
import math
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
BUILD = os.path.join(ROOT, "tests", "hostsim", "_build")
SIM = os.path.join(BUILD, "lancet_sim")


@pytest.fixture(scope="session")
def sim_cli():
    os.makedirs(BUILD, exist_ok=True)
    srcs = [os.path.join(ROOT, "lancet_b200", "csrc", "lancet_cli.cc"), os.path.join(ROOT, "tests", "hostsim", "lb2_shim.cc")]
    deps = srcs + [os.path.join(ROOT, "lancet_b200", "csrc", f) for f in os.listdir(os.path.join(ROOT, "lancet_b200", "csrc")) if f.endswith((".cuh", ".h"))]
    if not os.path.exists(SIM) or any(os.path.getmtime(d) > os.path.getmtime(SIM) for d in deps):
        subprocess.run(["g++", "-O2", "-std=c++17", "-w", "-o", SIM] + srcs + ["-lz"], check=True)
    return SIM


def _normalise(text):
    from lancet_b200.simbam import normalise_vcf
    return normalise_vcf(text)


@pytest.mark.parametrize("name", ["e2e_basic", "e2e_bed"])
def test_vcf_identical_to_reference_cli(sim_cli, name):
    d = os.path.join(GOLD, name)
    args = [a.replace("@DIR@", d) for a in json.load(open(os.path.join(d, "args.json")))["args"]]
    r = subprocess.run([sim_cli, "--tumor", os.path.join(d, "tumor.bam"), "--normal", os.path.join(d, "normal.bam"), "--ref", os.path.join(d, "ref.fa")] + args,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    want = open(os.path.join(d, "expected.vcf")).read().rstrip("\n")
    assert _normalise(r.stdout) == want
    assert sum(1 for l in want.splitlines() if not l.startswith("#")) > 10


def test_thread_count_changes_only_replay_order(sim_cli):
    """--num-threads decides which record wins a tie in the store (reference src/Lancet.cc:943-959), never the set of sites"""
    d = os.path.join(GOLD, "e2e_basic")
    outs = []
    for t in ("1", "4"):
        r = subprocess.run([sim_cli, "--tumor", os.path.join(d, "tumor.bam"), "--normal", os.path.join(d, "normal.bam"), "--ref", os.path.join(d, "ref.fa"),
                            "--reg", "chr22:1-4000", "--num-threads", t], capture_output=True, text=True, timeout=600)
        assert r.returncode == 0
        outs.append({tuple(l.split("\t")[:5]) for l in r.stdout.splitlines() if not l.startswith("#")})
    assert outs[0] == outs[1]


def test_batch_builders_do_not_change_the_vcf(sim_cli):
    """many small batches prepared by four builder threads at once (out of order, consumed in order) == one batch, one builder"""
    d = os.path.join(GOLD, "e2e_basic")
    base = [sim_cli, "--tumor", os.path.join(d, "tumor.bam"), "--normal", os.path.join(d, "normal.bam"), "--ref", os.path.join(d, "ref.fa"), "--reg", "chr22:1-6000", "--num-threads", "2"]
    outs = []
    for extra in (["--batch-windows", "100000", "--io-threads", "1"], ["--batch-windows", "3", "--io-threads", "16"]):
        r = subprocess.run(base + extra, capture_output=True, text=True, timeout=600, env=dict(os.environ, LB2_CLI_TIMING="1"))
        assert r.returncode == 0, r.stderr[-2000:]
        outs.append(_normalise(r.stdout))
        timing = [l for l in r.stderr.splitlines() if l.startswith("[timing] open")]
        assert timing and ("4 builders" in timing[0]) == (extra[-1] == "16")
    assert outs[0] == outs[1] and sum(1 for l in outs[0].splitlines() if not l.startswith("#")) > 3


def test_known_answers(sim_cli):
    out = subprocess.run([sim_cli, "--self-test"], capture_output=True, text=True, check=True).stdout.splitlines()
    kv = dict(l.split("=", 1) for l in out if l.startswith("sha256"))
    assert kv["sha256(abc)"] == hashlib.sha256(b"abc").hexdigest()
    assert kv["sha256()"] == hashlib.sha256(b"").hexdigest()
    assert kv["sha256(chr22:1234:S:1:A:T x3)"] == hashlib.sha256(b"chr22:1234:S:1:A:T" * 3).hexdigest()      # 54 bytes: padding crosses a block
    for l in out:
        if not l.startswith("fet "):
            continue
        _, a, b, c, d, p, s = l.split()
        a, b, c, d = int(a), int(b), int(c), int(d)
        # hypergeometric point probability of the table (what reference src/FET.hh:113-126 returns)
        lg = math.lgamma
        lb = lambda n, k: lg(n + 1) - lg(k + 1) - lg(n - k + 1)
        n1_, n_1, n = a + b, a + c, a + b + c + d
        want = 1.0 if max(0, n1_ + n_1 - n) == min(n1_, n_1) else math.exp(lb(n1_, a) + lb(n - n1_, n_1 - a) - lb(n, n_1))
        assert abs(float(p) - want) <= 1e-5 * want + 1e-300
    assert [l for l in out if l.startswith("dtos")][0] == "dtos 0 12.3457 1e-07 3079.99"


def test_requires_inputs(sim_cli):
    r = subprocess.run([sim_cli, "--tumor", "x.bam"], capture_output=True, text=True)
    assert r.returncode != 0 and "normal BAM" in r.stderr
