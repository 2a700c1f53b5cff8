"""TEST INFRASTRUCTURE ONLY: run the compiled reference (oracle/_ref/ref_windows) on a window batch.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
from __future__ import annotations

import json
import os
import subprocess
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF_WINDOWS = os.path.join(HERE, "_ref", "ref_windows")


def available() -> bool:
    return os.path.exists(REF_WINDOWS)


def parse_tsv(text: str):
    out = []
    for ln in text.splitlines():
        f = ln.split("\t")
        out.append((int(f[0]), int(f[1]), f[2], int(f[3]), f[4], f[5], int(f[6]), f[7],
                    tuple(int(x) for x in f[8].split(",")), f[9], f[10]))
    return out


def run(batch=None, path=None, threads=1, min_k=11, max_k=101, want_records=True, first=None, count=None, repeat=1):
    """returns (records or None, timing dict)."""
    tmp = None
    if path is None:
        tmp = tempfile.NamedTemporaryFile(suffix=".lb2b", delete=False); tmp.close(); path = tmp.name
        batch.save(path)
    out = tempfile.NamedTemporaryFile(suffix=".tsv", delete=False); out.close()
    cmd = [REF_WINDOWS, path, "--threads", str(threads), "--min-k", str(min_k), "--max-k", str(max_k), "--repeat", str(repeat)]
    if want_records:
        cmd += ["--out", out.name]
    if first is not None:
        cmd += ["--first", str(first), "--count", str(count)]
    try:
        p = subprocess.run(cmd, check=True, capture_output=True, text=True)
        timing = json.loads(p.stdout.strip().splitlines()[-1])
        recs = parse_tsv(open(out.name).read()) if want_records else None
    finally:
        os.unlink(out.name)
        if tmp is not None:
            os.unlink(path)
    return recs, timing
