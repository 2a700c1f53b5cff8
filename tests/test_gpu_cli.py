"""End-to-end parity of the shipped command line (lancet_b200/lancet_b200_cli -> _lb2.so -> B200) with the reference CLI:
same BAMs in, byte-identical VCF out (except ##fileDate/##cmdline/##reference)."""
import json
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
CLI = os.path.join(ROOT, "lancet_b200", "lancet_b200_cli")
REFCLI = os.path.join(ROOT, "oracle", "_ref", "lancet")


def _run(binp, d, args, timeout=240):
    from lancet_b200.simbam import normalise_vcf
    r = subprocess.run([binp, "--tumor", d["tumor"], "--normal", d["normal"], "--ref", d["ref"]] + args, capture_output=True, text=True, timeout=timeout)
    assert r.returncode == 0, r.stderr[-2000:]
    return normalise_vcf(r.stdout), r.stderr


@pytest.mark.parametrize("name", ["e2e_basic", "e2e_bed"])
def test_cli_golden(name):
    assert os.path.exists(CLI), "lancet_b200_cli not built (python -c 'import __graft_entry__ as g; g.build()')"
    d = os.path.join(GOLD, name)
    args = [a.replace("@DIR@", d) for a in json.load(open(os.path.join(d, "args.json")))["args"]]
    got, err = _run(CLI, {"tumor": os.path.join(d, "tumor.bam"), "normal": os.path.join(d, "normal.bam"), "ref": os.path.join(d, "ref.fa")}, args)
    assert "not assembled" not in err
    assert got == open(os.path.join(d, "expected.vcf")).read().rstrip("\n")


LIVE = {
    "thr4": (dict(seed=102, chroms=(("chr22", 12000),)), ["--reg", "chr22:1-12000", "--num-threads", "4"]),
    "sparse": (dict(seed=103, chroms=(("chr22", 20000),), var_every=1500, som_every=2500), ["--reg", "chr22:200-19800", "--num-threads", "3"]),
    "str": (dict(seed=105, chroms=(("chr22", 8000),), str_every=150), ["--reg", "chr22:1-8000", "--num-threads", "2"]),
    "noactive": (dict(seed=106, chroms=(("chr22", 6000),)), ["--reg", "chr22:1-6000", "--active-region-off", "--num-threads", "2"]),
    "opts": (dict(seed=108, chroms=(("chr22", 6000),)), ["--reg", "chr22:500-5500", "--num-threads", "2", "--min-k", "15", "--max-k", "61", "--window-size", "500",
                                                          "--padding", "100", "--primary-alignment-only", "--XA-tag-filter", "--min-vaf-tumor", "0.1",
                                                          "--min-alt-count-tumor", "4", "--min-base-qual", "20", "--min-map-qual", "20"]),
    "err1": (dict(seed=109, chroms=(("chr22", 6000),), err=0.01), ["--reg", "chr22:1-6000", "--num-threads", "2"]),
    "odd": (dict(seed=111, chroms=(("chr22", 6000),), odd_frac=0.5), ["--reg", "chr22", "--num-threads", "2"]),
}


@pytest.mark.parametrize("name", sorted(LIVE))
def test_cli_live_reference(name, tmp_path):
    if not (os.path.exists(REFCLI) and os.path.exists(os.path.join(ROOT, "oracle", "_ref", "test_view"))):
        pytest.skip("compiled reference not present")
    from lancet_b200 import simbam
    kw, args = LIVE[name]
    d = simbam.write_dataset(str(tmp_path / name), **kw)
    want, _ = _run(REFCLI, d, args)
    got, err = _run(CLI, d, args)
    assert "not assembled" not in err
    assert got == want
    assert sum(1 for l in want.splitlines() if not l.startswith("#")) > 5
