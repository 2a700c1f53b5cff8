"""Control flow of the device sources, checked on the CPU: tests/hostsim compiles lancet_b200/csrc/*.cuh for ONE thread
(g++, LB2_HOSTSIM) and the result must reproduce the committed golden vectors of the compiled reference
(tests/golden/*.ref.tsv, made by tests/golden/make_golden.py).  A logic regression in the k-sweep, the graph sweeps, the
path enumeration or the transcripts shows up here without a GPU; warp/CTA cooperation does not (one thread) -- the GPU
parity tests remain the gate for the product path, and nothing shipped runs through this build."""
import gzip
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
GOLD = os.path.join(ROOT, "tests", "golden")
BUILD = os.path.join(ROOT, "tests", "hostsim", "_build")
SIM = os.path.join(BUILD, "hostsim")
CASES = {"config1_k25": ["--min-k", "25", "--max-k", "25"], "small_s7": [], "errors_s5": [], "lowqual_s3": [], "dense_s9": [], "paired_s62": [], "str_k99": []}


@pytest.fixture(scope="module")
def sim():
    os.makedirs(BUILD, exist_ok=True)
    subprocess.run(["g++", "-O2", "-std=c++17", "-w", "-o", SIM, os.path.join(ROOT, "tests", "hostsim", "hostsim.cc")], check=True)
    return SIM


@pytest.mark.parametrize("name", sorted(CASES))
def test_single_thread_build_reproduces_golden(sim, name, tmp_path):
    import run_ref
    p = tmp_path / (name + ".lb2b")
    with gzip.open(os.path.join(GOLD, name + ".lb2b.gz"), "rb") as g:
        p.write_bytes(g.read())
    out = tmp_path / "sim.tsv"
    r = subprocess.run([sim, str(p), "--out", str(out)] + CASES[name], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-400:]
    assert not [l for l in r.stderr.splitlines() if "status 3" in l or "status 4" in l], r.stderr[-400:]      # no window left unassembled
    want = run_ref.parse_tsv(open(os.path.join(GOLD, name + ".ref.tsv")).read())
    assert run_ref.parse_tsv(out.read_text()) == want
