"""Regenerates tests/golden/kat_reference.json: inputs + the outputs of the reference's OWN functions
(oracle/_ref/kat_reference, linked from /root/reference sources).  Run in the build container."""
import json
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
rng = np.random.default_rng(12345)


def rs(n, alpha="ACGT"):
    return "".join(rng.choice(list(alpha), n))


cases = []
# repeats: random, planted exact repeat, planted near repeat (1-3 mismatches), short strings
for i in range(60):
    L = int(rng.integers(30, 400)); s = list(rs(L)); K = int(rng.choice([11, 13, 15, 21, 25, 31]))
    mode = i % 4
    if mode and L > 3 * K:
        a = int(rng.integers(0, L // 2 - K)); b = int(rng.integers(L // 2, L - K - 2)); ln = K + int(rng.integers(0, 3))
        s[b:b + ln] = s[a:a + ln]
        for _ in range(mode - 1):
            p = b + int(rng.integers(0, ln)); s[p] = "ACGT"[("ACGT".index(s[p]) + 1) % 4]
    cases.append(["R", str(K), "".join(s)])
# alignments: ref vs ref with indels / snvs
for i in range(40):
    L = int(rng.integers(40, 260)); S = rs(L); T = list(S)
    for _ in range(int(rng.integers(1, 4))):
        p = int(rng.integers(12, len(T) - 12)); kind = int(rng.integers(0, 3))
        if kind == 0: T[p] = "ACGT"[("ACGT".index(T[p]) + 1) % 4]
        elif kind == 1: T[p:p] = list(rs(int(rng.integers(1, 12))))
        else: del T[p:p + int(rng.integers(1, 12))]
    cases.append(["A", S, "".join(T)])
for hp in (3, 6, 9):   # homopolymer / STR contexts where gap placement ties matter
    S = rs(40) + "A" * hp + rs(40); T = S[:40] + "A" * (hp + 2) + S[40 + hp:]
    cases.append(["A", S, T]); cases.append(["A", T, S])
    S = rs(30) + "CAG" * hp + rs(30); T = S[:30] + "CAG" * (hp - 1) + S[30 + 3 * hp:]
    cases.append(["A", S, T])
# tandems
for i in range(40):
    unit = rs(int(rng.integers(1, 5))); n = int(rng.integers(2, 12)); pre = rs(int(rng.integers(0, 30))); post = rs(int(rng.integers(0, 30)))
    s = pre + unit * n + post
    cases.append(["T", str(int(rng.integers(0, len(s)))), s])
for i in range(10):
    s = rs(int(rng.integers(20, 120))); cases.append(["T", str(int(rng.integers(0, len(s)))), s])
# trimming
for i in range(40):
    L = int(rng.integers(5, 120)); s = list(rs(L)); q = [chr(int(x)) for x in rng.choice([35, 40, 43, 45, 60, 70, 73], L)]
    if i % 3 == 0: s[int(rng.integers(0, L))] = "N"
    if i % 7 == 0: q = [chr(35)] * L
    cases.append(["Q", "".join(s), "".join(q)])

inp = "\n".join(" ".join(c) for c in cases) + "\n"
out = subprocess.run([os.path.join(ROOT, "oracle", "_ref", "kat_reference")], input=inp, capture_output=True, text=True, check=True).stdout.splitlines()
assert len(out) == len(cases)
json.dump([{"in": c, "out": o.split()} for c, o in zip(cases, out)], open(os.path.join(HERE, "kat_reference.json"), "w"), indent=0)
print(len(cases), "cases")
