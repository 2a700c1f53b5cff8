"""Regenerates tests/golden/*.lb2b.gz and *.ref.tsv by running the compiled reference
(oracle/_ref/ref_windows, built from /root/reference by oracle/Makefile) on seeded synthetic batches.
Run in the build container:  python tests/golden/make_golden.py"""
import gzip
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
from lancet_b200.synth import make_batch  # noqa: E402
import run_ref  # noqa: E402

CASES = {
    # name: (generator kwargs, harness kwargs)
    "config1_k25": (dict(seed=1, region_len=600, region_start=1201, cov_t=17, cov_n=17, var_every=250), dict(min_k=25, max_k=25)),
    "small_s7": (dict(seed=7, region_len=1300), {}),
    "errors_s5": (dict(seed=5, region_len=1200, err=0.005), {}),
    "lowqual_s3": (dict(seed=3, region_len=1200, low_qual_frac=0.03), {}),
    "dense_s9": (dict(seed=9, region_len=1200, var_every=150), {}),
    # round 2: overlapping mates (the unsorted binary_search quirk), and STR-rich windows that are only assembled at k = 83..99
    # (at k = 99 one k-mer pair per 100 bp read: ~650 nodes in dozens of small components -- the per-component sweeps)
    "paired_s62": (dict(seed=62, region_len=1200, paired=True, insert_mean=150, insert_sd=20), {}),
    "str_k99": (dict(seed=1000, region_start=1_000_001, var_every=5000, region_len=100000, str_every=200, cov_t=80, cov_n=80, _windows=[289, 309, 337, 443, 715]), {}),
}

if __name__ == "__main__":
    only = sys.argv[1:]
    for name, (gk, hk) in CASES.items():
        if only and name not in only:
            continue
        gk = dict(gk); pick = gk.pop("_windows", None)
        b = make_batch(**gk)
        if pick is not None:
            b = b.subset(pick)
        path = os.path.join(HERE, name + ".lb2b")
        b.save(path)
        recs, _ = run_ref.run(path=path, **hk)
        with open(path, "rb") as f, gzip.open(path + ".gz", "wb", compresslevel=9) as g:
            g.write(f.read())
        os.unlink(path)
        with open(os.path.join(HERE, name + ".ref.tsv"), "w") as f:
            for r in recs:
                f.write("\t".join([str(r[0]), str(r[1]), r[2], str(r[3]), r[4], r[5], str(r[6]), r[7],
                                   ",".join(map(str, r[8])), r[9], r[10]]) + "\n")
        print(name, b.n_windows, "windows", b.n_reads, "reads", len(recs), "records")
