"""Makes the end-to-end (BAM -> VCF) golden fixtures: synthetic tumour/normal BAMs (lancet_b200/simbam.py) and the VCF the
UNMODIFIED reference CLI (oracle/_ref/lancet, built from /root/reference by oracle/Makefile) writes for them.
Run in the development container:  python tests/golden/make_e2e_golden.py"""
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from lancet_b200 import simbam  # noqa: E402

CASES = {
    "e2e_basic": (dict(seed=21, chroms=(("chr22", 4000),)), ["--reg", "chr22:1-4000", "--num-threads", "2"]),
    "e2e_bed": (dict(seed=22, chroms=(("chr1", 3000), ("chr22", 3000)), str_every=170, odd_frac=0.25, cov_t=50, cov_n=40),
                ["--bed", "@DIR@/regions.bed", "--reg", "chr1:1200-1900", "--num-threads", "3", "--min-vaf-tumor", "0.05"]),
}

if __name__ == "__main__":
    for name, (kw, args) in CASES.items():
        d = os.path.join(HERE, name)
        ds = simbam.write_dataset(d, **kw)
        with open(os.path.join(d, "regions.bed"), "w") as f:
            f.write("#chrom\tstart\tend\nchr22\t300\t1200\nchr22\t300\t1200\nchr22\t1500\t2600\tx\nchr1\t500\t900\n")
        a = [x.replace("@DIR@", d) for x in args]
        r = subprocess.run([os.path.join(ROOT, "oracle", "_ref", "lancet"), "--tumor", ds["tumor"], "--normal", ds["normal"], "--ref", ds["ref"]] + a,
                           capture_output=True, text=True, check=True)
        with open(os.path.join(d, "expected.vcf"), "w") as f:
            f.write(simbam.normalise_vcf(r.stdout) + "\n")
        json.dump({"args": args}, open(os.path.join(d, "args.json"), "w"))
        for junk in ("tumor.sam", "normal.sam", "ref.fa.fai"):
            p = os.path.join(d, junk)
            if os.path.exists(p):
                os.remove(p)
        print(name, sum(1 for l in r.stdout.splitlines() if not l.startswith("#")), "records")
