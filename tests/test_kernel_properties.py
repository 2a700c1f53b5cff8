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
