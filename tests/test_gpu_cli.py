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
    # a run of N in the reference: the windows over it are assembled on the device like any other (no "unsupported" exit)
    "nref": (dict(seed=112, chroms=(("chr22", 6000),), n_in_ref=True), ["--reg", "chr22:1-6000", "--num-threads", "2"]),
    # many small batches (one lb2_process call each, the reads of a batch's span fetched through the BAI), two chromosomes
    "batches": (dict(seed=113, chroms=(("chr21", 5000), ("chr22", 9000))), ["--reg", "chr22:1-9000", "--num-threads", "3"]),
}
OURS_ONLY = {"batches": ["--batch-windows", "11", "--io-threads", "3"]}      # options the reference does not have


@pytest.mark.parametrize("name", sorted(LIVE))
def test_cli_live_reference(name, tmp_path):
    if not (os.path.exists(REFCLI) and os.path.exists(os.path.join(ROOT, "oracle", "_ref", "test_view"))):
        pytest.skip("compiled reference not present")
    from lancet_b200 import simbam
    kw, args = LIVE[name]
    d = simbam.write_dataset(str(tmp_path / name), **kw)
    want, _ = _run(REFCLI, d, args)
    got, err = _run(CLI, d, args + OURS_ONLY.get(name, []))
    assert "not assembled" not in err
    assert got == want
    assert sum(1 for l in want.splitlines() if not l.startswith("#")) > 5


def _gpu_count():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.parametrize("ngpu", [2, 4])
def test_cli_multi_gpu(ngpu, tmp_path):
    """--gpus N: one process per GPU, each assembles a contiguous range of the windows, the records are gathered on rank 0
    over NCCL and replayed in the reference's order -- the VCF is the one the reference writes for the same --num-threads."""
    if _gpu_count() < ngpu:
        pytest.skip(f"needs {ngpu} GPUs")
    if not (os.path.exists(REFCLI) and os.path.exists(os.path.join(ROOT, "oracle", "_ref", "test_view"))):
        pytest.skip("compiled reference not present")
    from lancet_b200 import simbam
    d = simbam.write_dataset(str(tmp_path / "mg"), seed=131, chroms=(("chr22", 30000),), var_every=400, som_every=900)
    args = ["--reg", "chr22:1-30000", "--num-threads", "4"]
    want, _ = _run(REFCLI, d, args, timeout=300)
    got, err = _run(CLI, d, args + ["--gpus", str(ngpu), "--batch-windows", "64"], timeout=150)
    assert "not assembled" not in err
    assert got == want
    assert sum(1 for l in want.splitlines() if not l.startswith("#")) > 50
