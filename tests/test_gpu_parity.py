"""Parity tests proper: the CUDA path, called through the C ABI (lancet_b200/_lb2.so via ctypes), against
(a) the committed golden vectors produced by the compiled reference, (b) the compiled reference run live on
seeded inputs (oracle/_ref/ref_windows travels to the GPU box as a binary), (c) size-independent properties
on larger batches.  Bit-exact: every field of every Variant_t tuple, in emission order."""
import gzip
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")
CASES = {"config1_k25": dict(min_k=25, max_k=25), "small_s7": {}, "errors_s5": {}, "lowqual_s3": {}, "dense_s9": {}, "paired_s62": {}, "str_k99": {}}


def _load_gz(name, tmp_path):
    from lancet_b200.batch import Batch
    p = tmp_path / (name + ".lb2b")
    with gzip.open(os.path.join(GOLD, name + ".lb2b.gz"), "rb") as g:
        p.write_bytes(g.read())
    return Batch.load(str(p))


def _ctx(**over):
    from lancet_b200.api import Context, Params
    return Context(Params.default(**over), device=0)


@pytest.mark.parametrize("name", sorted(CASES))
def test_golden(name, tmp_path):
    import run_ref
    b = _load_gz(name, tmp_path)
    want = run_ref.parse_tsv(open(os.path.join(GOLD, name + ".ref.tsv")).read())
    c = _ctx(**CASES[name])
    res = c.process(b)
    assert (res.windows["status"] != 3).all() and (res.windows["status"] != 4).all(), res.windows
    assert res.records() == want
    assert c.kernel_launches >= 1
    c.close()


@pytest.mark.parametrize("kw", [
    dict(seed=11, region_len=6000),
    dict(seed=23, region_len=6000),
    dict(seed=31, region_len=4000, err=0.005),
    dict(seed=32, region_len=4000, low_qual_frac=0.05, err=0.002),
    dict(seed=33, region_len=4000, cov_t=20, cov_n=15),
    dict(seed=34, region_len=4000, var_every=120),
    dict(seed=35, region_len=3000, cov_t=150, cov_n=150),
    dict(seed=36, region_len=3000, read_len=150),
    dict(seed=37, region_len=1777),              # short last window (len-offset-1 rule)
    dict(seed=61, region_len=3000, paired=True),                                   # Variant B: true pairs, names unsorted
    dict(seed=62, region_len=3000, paired=True, insert_mean=150, insert_sd=20),    # overlapping mates (unsorted binary_search quirk)
    dict(seed=77, region_len=3000, paired=True, insert_mean=160, insert_sd=40, err=0.004, low_qual_frac=0.02),
    dict(seed=71, region_len=4000, str_every=200, cov_t=80, cov_n=80),             # STR-rich reference: long k sweeps, skipped windows
    dict(seed=72, region_len=4000, str_every=500),
    dict(seed=75, region_len=1500, err=0.01),                                      # heavy BFS: goes through the escalation pass
])
def test_live_reference(kw, ctx):
    import run_ref
    if not run_ref.available():
        pytest.skip("oracle/_ref/ref_windows not built")
    from lancet_b200.synth import make_batch
    b = make_batch(**kw)
    want, _ = run_ref.run(b, threads=8)
    res = ctx.process(b)
    assert (res.windows["status"] < 3).all(), res.windows[res.windows["status"] >= 3]
    assert res.records() == want


@pytest.mark.parametrize("k", [31, 33, 63, 65, 101])
def test_large_k(k):
    """multi-word k-mers: k-mer sizes around the 32/64-base word boundaries and the reference's maximum"""
    import run_ref
    if not run_ref.available():
        pytest.skip("oracle/_ref/ref_windows not built")
    from lancet_b200.synth import make_batch
    b = make_batch(seed=81, region_len=2500, read_len=150)
    want, _ = run_ref.run(b, threads=8, min_k=k, max_k=k)
    c = _ctx(min_k=k, max_k=k)
    res = c.process(b)
    assert (res.windows["status"] < 3).all()
    assert res.records() == want
    c.close()


def test_edge_cases(ctx):
    """empty window, window without reads of one sample, all-junk reads, window shorter than k."""
    import run_ref
    from lancet_b200.synth import make_batch
    from lancet_b200.batch import Batch
    b = make_batch(seed=41, region_len=1500)
    # drop all reads of window 1, keep only normal reads in window 2, junk qualities for window 3
    wr = [b.wr_idx[b.wr_off[w]:b.wr_off[w + 1]].copy() for w in range(b.n_windows)]
    wr[1] = wr[1][:0]
    wr[2] = wr[2][(b.flags[wr[2]] & 1) == 1]
    qual = b.qual.copy()
    for r in wr[3]:
        qual[int(b.base_off[r]):int(b.base_off[r + 1])] = 33 + 2
    wr_off = np.zeros(b.n_windows + 1, np.uint32); wr_off[1:] = np.cumsum([len(x) for x in wr])
    b2 = Batch(ref_off=b.ref_off, ref_start=b.ref_start, chr_id=b.chr_id, wr_off=wr_off, wr_idx=np.concatenate(wr),
               base_off=b.base_off, flags=b.flags, name_rank=b.name_rank, ref_seq=b.ref_seq, seq=b.seq, qual=qual)
    res = ctx.process(b2)
    assert res.windows["status"][1] == 2          # LB2_WIN_NO_READS
    if run_ref.available():
        want, _ = run_ref.run(b2)
        assert res.records() == want


def test_properties_large(ctx):
    """size-independent properties on a batch the reference would need minutes for:
    determinism, independence of windows from their batch neighbours and from batch order."""
    from lancet_b200.synth import make_batch
    b = make_batch(seed=51, region_len=60000, var_every=900)
    r1 = ctx.process(b); rec1 = r1.records()
    r2 = ctx.process(b); assert r2.records() == rec1
    assert (r1.windows["status"] < 3).all()
    rng = np.random.default_rng(0)
    pick = np.sort(rng.choice(b.n_windows, 40, replace=False))
    sub = ctx.process(b.subset(pick)).records()
    remap = {int(w): i for i, w in enumerate(pick)}
    want = [(remap[r[0]],) + r[1:] for r in rec1 if r[0] in remap]
    assert sub == want
    perm = rng.permutation(pick)
    subp = ctx.process(b.subset(perm)).records()
    remap = {int(w): i for i, w in enumerate(perm)}
    want = sorted([(remap[r[0]],) + r[1:] for r in rec1 if r[0] in remap], key=lambda r: r[0])
    assert sorted(subp, key=lambda r: r[0]) == want
    # planted variants are recovered: every planted SNV position appears in some record
    assert len(rec1) > 50


def _pinned(b):
    """the batch with its arrays in page-locked host memory (what a caller that wants the copy overlap provides)"""
    import copy
    import torch
    b = copy.copy(b); b._pins = []
    for name in ("ref_off", "ref_start", "chr_id", "wr_off", "wr_idx", "base_off", "flags", "name_rank", "ref_seq", "seq", "qual"):
        a = getattr(b, name)
        t = torch.empty(max(a.nbytes, 1), dtype=torch.uint8, pin_memory=True)
        v = t.numpy()[:a.nbytes].view(a.dtype); v[...] = a
        setattr(b, name, v); b._pins.append(t)
    return b


@pytest.mark.timeout(240, method="thread")
def test_segmented_equals_resident(ctx, monkeypatch):
    """From page-locked buffers lb2_process uploads the batch in segments of consecutive windows and assembles a segment
    while the next one is on its way (one launch per segment, two compute streams); lb2_upload + lb2_run +
    lb2_download work on a resident batch, and so does lb2_process from pageable memory (one segment).  Same records
    every way, also with many tiny segments and with a permuted window order (windows that need late reads come first)."""
    from lancet_b200.synth import make_batch
    b = make_batch(seed=52, region_len=30000, var_every=700)
    ctx.upload(b); ctx.run(); ctx.wait(); res0 = ctx.download(); rec0 = res0.records()
    assert len(rec0) > 20
    assert ctx.process(b).records() == rec0                    # pageable: one segment
    bp = _pinned(b)
    for first, minw in (("65536", "8"), ("262144", "40"), ("1048576", "64"), (str(1 << 30), "768")):
        monkeypatch.setenv("LB2_SEG_FIRST", first); monkeypatch.setenv("LB2_SEG_MIN_WINDOWS", minw)
        assert ctx.process(bp).records() == rec0
    monkeypatch.setenv("LB2_SEG_FIRST", "65536"); monkeypatch.setenv("LB2_SEG_MIN_WINDOWS", "16")
    perm = np.random.default_rng(1).permutation(b.n_windows)
    bq = b.subset(perm)
    ctx.upload(bq); ctx.run(); ctx.wait(); recp = ctx.download().records()
    assert ctx.process(_pinned(bq)).records() == recp
    # a second context on the same device with another shared-memory configuration must not disturb the first one
    from lancet_b200.api import Context
    c2 = Context(device=0)
    small = make_batch(seed=53, region_len=2000)
    r_small = c2.process(small).records()
    ctx.upload(b); ctx.run(); ctx.wait(); assert ctx.download().records() == rec0
    assert c2.process(small).records() == r_small
    c2.close()


@pytest.mark.parametrize("mink,maxk", [(7, 9), (9, 15), (15, 17)])
def test_small_k(mink, maxk):
    """k below 11 switches the repeat scan's word filter off (every word takes the exact rule); k <= 16 is the
    32-bit k-mer walk, 17 the first 64-bit one."""
    import run_ref
    if not run_ref.available():
        pytest.skip("oracle/_ref/ref_windows not built")
    from lancet_b200.synth import make_batch
    b = make_batch(seed=91, region_len=2500, var_every=300)
    want, _ = run_ref.run(b, threads=8, min_k=mink, max_k=maxk)
    c = _ctx(min_k=mink, max_k=maxk)
    res = c.process(b)
    assert (res.windows["status"] < 3).all()
    assert res.records() == want
    c.close()


def test_long_reads_and_ragged_trims(ctx):
    """250 bp reads (two rounds of 16-base chunks per 8-lane group), N bases and low-quality runs at both ends and
    inside reads (Graph_t::trim: 5'/3' trims, junk reads), reads shorter than k after trimming."""
    import run_ref
    if not run_ref.available():
        pytest.skip("oracle/_ref/ref_windows not built")
    from lancet_b200.synth import make_batch
    from lancet_b200.batch import Batch
    b = make_batch(seed=92, region_len=3000, read_len=250, cov_t=50, cov_n=50, var_every=400)
    rng = np.random.default_rng(5)
    seq = b.seq.copy(); qual = b.qual.copy()
    for r in range(b.n_reads):
        o0, o1 = int(b.base_off[r]), int(b.base_off[r + 1]); n = o1 - o0
        u = rng.random()
        if u < 0.15:      # low-quality 5' and 3' tails of random length
            a, z = int(rng.integers(0, 40)), int(rng.integers(0, 40))
            qual[o0:o0 + a] = 33 + 5; qual[o1 - z:o1] = 33 + 5 if z else qual[o1 - z:o1]
        elif u < 0.22:    # N at the ends (trimmed) ...
            seq[o0:o0 + int(rng.integers(1, 5))] = ord("N"); seq[o1 - int(rng.integers(1, 5)):o1] = ord("N")
        elif u < 0.26:    # ... or in the middle (junk read)
            seq[o0 + n // 2] = ord("N")
        elif u < 0.29:    # almost everything low quality: shorter than k after the trim
            qual[o0:o1 - 8] = 33 + 3
        elif u < 0.31:    # nothing left
            qual[o0:o1] = 33 + 2
    b2 = Batch(ref_off=b.ref_off, ref_start=b.ref_start, chr_id=b.chr_id, wr_off=b.wr_off, wr_idx=b.wr_idx,
               base_off=b.base_off, flags=b.flags, name_rank=b.name_rank, ref_seq=b.ref_seq, seq=seq, qual=qual)
    want, _ = run_ref.run(b2, threads=8)
    res = ctx.process(b2)
    assert (res.windows["status"] < 3).all(), res.windows[res.windows["status"] >= 3]
    assert res.records() == want
    assert len(want) > 5


def test_many_records_per_window(ctx):
    """windows with more records than a first-pass output slab holds (32) are redone by the escalation pass, which emits
    into the large slabs: dense variants + 0.8 % errors at 100x give up to ~130 records in one window"""
    import run_ref
    if not run_ref.available():
        pytest.skip("oracle/_ref/ref_windows not built")
    from collections import Counter
    from lancet_b200.api import Context, Params
    from lancet_b200.synth import make_batch
    b = make_batch(seed=1002, region_len=1200, cov_t=100, cov_n=60, read_len=76, err=0.008, var_every=80)
    want, _ = run_ref.run(b, threads=8, min_k=17)
    assert max(Counter(r[0] for r in want).values()) > 64
    c = Context(Params.default(min_k=17), device=0)
    res = c.process(b)
    assert (res.windows["status"] < 3).all(), res.windows[res.windows["status"] >= 3]
    assert res.records() == want
    c.close()


def _with_n_runs(b, seed):
    """window references with runs of N, N at both ends, periodic N (many distinct N k-mers, short N-free stretches), all N"""
    from lancet_b200.batch import Batch
    rng = np.random.default_rng(seed); ref = b.ref_seq.copy()
    for w in range(b.n_windows):
        o0, o1 = int(b.ref_off[w]), int(b.ref_off[w + 1]); L = o1 - o0
        u = rng.random()
        if u < 0.3:
            a = int(rng.integers(0, L - 40)); ref[o0 + a:o0 + a + int(rng.integers(2, 40))] = ord("N")
        elif u < 0.5:
            ref[o0:o0 + int(rng.integers(1, 20))] = ord("N"); ref[o1 - int(rng.integers(1, 20)):o1] = ord("N")
        elif u < 0.6:
            ref[o0 + int(rng.integers(0, 15)):o1:int(rng.integers(12, 30))] = ord("N")
        elif u < 0.65:
            ref[o0:o1] = ord("N")
    return Batch(ref_off=b.ref_off, ref_start=b.ref_start, chr_id=b.chr_id, wr_off=b.wr_off, wr_idx=b.wr_idx, base_off=b.base_off,
                 flags=b.flags, name_rank=b.name_rank, ref_seq=ref, seq=b.seq, qual=b.qual)


@pytest.mark.parametrize("case,mink", [("single", 11), ("dense", 11), ("runs", 11), ("runs", 31), ("runs_err", 11), ("runs_str", 63)])
def test_reference_with_n(case, mink):
    """A window reference with non-ACGT bases: the reference loads it untrimmed (src/Graph.cc:534-540), its N-containing
    k-mers become map entries that die in the first low-coverage sweep but shift the iteration order of everything else,
    'N' columns come out as SNV records.  Assembled on the device like any other window."""
    import run_ref
    if not run_ref.available():
        pytest.skip("oracle/_ref/ref_windows not built")
    from lancet_b200.synth import make_batch
    b = {"single": lambda: make_batch(seed=201, region_len=4000, n_in_ref=12),
         "dense": lambda: make_batch(seed=202, region_len=4000, n_in_ref=40, var_every=300),
         "runs": lambda: _with_n_runs(make_batch(seed=203, region_len=5000, var_every=300), 1),
         "runs_err": lambda: _with_n_runs(make_batch(seed=204, region_len=4000, err=0.004), 2),
         "runs_str": lambda: _with_n_runs(make_batch(seed=205, region_len=4000, str_every=300), 3)}[case]()
    want, _ = run_ref.run(b, threads=8, min_k=mink)
    c = _ctx(min_k=mink)
    res = c.process(b)
    assert (res.windows["status"] < 3).all(), res.windows[res.windows["status"] >= 3]
    assert res.records() == want
    assert any("N" in r[4] for r in want) or case == "runs_str"
    c.close()


def _concat(b1, b2):
    """two batches as one (pools and windows appended)"""
    from lancet_b200.batch import Batch
    return Batch(ref_off=np.concatenate([b1.ref_off, b2.ref_off[1:] + b1.ref_off[-1]]), ref_start=np.concatenate([b1.ref_start, b2.ref_start]),
                 chr_id=np.concatenate([b1.chr_id, b2.chr_id]), wr_off=np.concatenate([b1.wr_off, b2.wr_off[1:] + b1.wr_off[-1]]),
                 wr_idx=np.concatenate([b1.wr_idx, b2.wr_idx + b1.n_reads]), base_off=np.concatenate([b1.base_off, b2.base_off[1:] + b1.base_off[-1]]),
                 flags=np.concatenate([b1.flags, b2.flags]), name_rank=np.concatenate([b1.name_rank, b2.name_rank + (int(b1.name_rank.max()) + 1 if b1.n_reads else 0)]),
                 ref_seq=np.concatenate([b1.ref_seq, b2.ref_seq]), seq=np.concatenate([b1.seq, b2.seq]), qual=np.concatenate([b1.qual, b2.qual]))


def test_deep_windows_do_not_set_the_launch_configuration():
    """A few deep windows (160x + 160x) in a batch of ordinary ones: the first pass stays sized for the ordinary
    windows (three CTAs per SM), the deep ones go through the escalation pass with its own, larger staging area --
    and every window comes out identical to the reference."""
    import run_ref
    if not run_ref.available():
        pytest.skip("oracle/_ref/ref_windows not built")
    from lancet_b200.synth import make_batch
    normal = make_batch(seed=301, region_len=12000, var_every=600)
    deep = make_batch(seed=302, region_len=900, cov_t=160, cov_n=160, var_every=300)
    b = _concat(normal, deep)
    want, _ = run_ref.run(b, threads=8)
    c = _ctx()
    res = c.process(b)
    assert (res.windows["status"] < 3).all(), res.windows[res.windows["status"] >= 3]
    assert res.records() == want
    assert c.resident_ctas >= 3 * 148 or c.resident_ctas >= b.n_windows       # the deep windows did not cost the batch its occupancy
    c.close()
