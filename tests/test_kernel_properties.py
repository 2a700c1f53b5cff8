"""Algorithm properties of two device-side shortcuts, checked on the CPU by compiling the kernel sources' single-thread
debug build (tests/hostsim, g++): (1) the word filter of the repeat scan answers isRepeat / isAlmostRepeat for every
k >= 11 exactly like the unfiltered scan; (2) the division used by the coverage fold (reciprocal + two fused residual
corrections) equals IEEE single-precision division on the fold's operand range.  The GPU parity tests remain the gate
for the product path; nothing here runs the assembler."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "tests", "hostsim", "_build")
SIM = os.path.join(BUILD, "hostsim")


@pytest.fixture(scope="module")
def sim():
    os.makedirs(BUILD, exist_ok=True)
    subprocess.run(["g++", "-O2", "-std=c++17", "-w", "-o", SIM, os.path.join(ROOT, "tests", "hostsim", "hostsim.cc")], check=True)
    return SIM


def test_repeat_scan_filter_equals_unfiltered(sim):
    r = subprocess.run([sim, "scantest"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert " 0 mismatches" in r.stdout


def test_fold_division_equals_ieee(sim):
    r = subprocess.run([sim, "divtest"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert " 0 mismatches" in r.stdout


def test_scalar_find_tandems_matches_reference_goldens(sim):
    """the device's scalar findTandems (run-event formulation, lb2_graph.cuh) against the outputs of the reference's own
    findTandems (tests/golden/kat_reference.json, generated from /root/reference by tests/golden/make_kat.py)"""
    import json
    cases = [c for c in json.load(open(os.path.join(ROOT, "tests", "golden", "kat_reference.json"))) if c["in"][0] == "T"]
    assert len(cases) >= 40
    inp = "".join(f"T {c['in'][1]} {c['in'][2]}\n" for c in cases)
    r = subprocess.run([sim, "tandemtest"], input=inp, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    got = r.stdout.splitlines()
    assert len(got) == len(cases)
    for c, g in zip(cases, got):
        ans, ln, motif = g.split()
        assert ans == c["out"][0], c
        if ans == "1":
            assert ln == c["out"][1] and motif == c["out"][2], (c, g)
