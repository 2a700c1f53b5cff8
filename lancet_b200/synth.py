"""Seeded synthetic tumour/normal window batches (SURVEY.md §8d "Configs -> concrete synthetic inputs").

Reference = i.i.d. uniform ACGT; reads ``read_len`` bp, Phred in {30,30,30,35,37,40}, per-base
substitution error ``err``; tumour = fraction ``vaf`` of reads from a haplotype with planted
SNV / 1-8 bp insertion / 1-8 bp deletion (alternating) every ~``var_every`` bp, normal = reference
haplotype.  Windows are tiled exactly like the reference's loadRefs (src/Lancet.cc:266-311:
step 100, length 600, last window ``len-offset-1``) and reads are attached to a window with the
reference's containment rule and its 1-based/0-based quirk (src/Microassembler.cc:802-805, :505:
keep iff ``Position >= refstart && GetEndPosition() <= refend``).

Variant A (default): every read has a unique query name (mate logic inert).
Variant B (``paired=True``): true pairs sharing a query name, names in random order.
"""
from __future__ import annotations

import numpy as np

from .batch import Batch, READ_NORMAL, READ_REVERSE, READ_MATE_SHIFT

ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
QUALS = np.array([30, 30, 30, 35, 37, 40], dtype=np.uint8) + 33


def tile_windows(seq_len: int, start_1based: int, window: int = 600, delta: int = 100):
    """(offset, LEN, refstart) per window, as reference src/Lancet.cc:266-311."""
    out = []
    end = seq_len
    offset = 0
    while offset < end:
        ln = window
        if offset + window >= seq_len:
            ln = seq_len - offset - 1
            end = offset
        out.append((offset, ln, start_1based + offset))
        offset += delta
    return out


def _make_alt(ref_codes: np.ndarray, rng, var_every: int, first: int, kinds=("snv", "ins", "del")):
    """Plant variants; returns alt codes, start/end coordinate maps and the truth list."""
    G = len(ref_codes)
    pieces, smap, emap, truth = [], [], [], []
    cur = 0
    pos = first
    k = 0
    while pos < G - 50:
        kind = kinds[k % len(kinds)]
        k += 1
        pieces.append(ref_codes[cur:pos]); smap.append(np.arange(cur, pos)); emap.append(np.arange(cur, pos) + 1)
        if kind == "snv":
            b = (ref_codes[pos] + rng.integers(1, 4)) % 4
            pieces.append(np.array([b], dtype=np.uint8)); smap.append(np.array([pos])); emap.append(np.array([pos + 1]))
            truth.append(("snv", pos, 1))
            cur = pos + 1
        elif kind == "ins":
            L = int(rng.integers(1, 9))
            ins = rng.integers(0, 4, L).astype(np.uint8)
            pieces.append(ins); smap.append(np.full(L, pos)); emap.append(np.full(L, pos))
            truth.append(("ins", pos, L))
            cur = pos
        else:
            L = int(rng.integers(1, 9))
            truth.append(("del", pos, L))
            cur = pos + L
        pos += var_every + int(rng.integers(-var_every // 10, var_every // 10 + 1))
    pieces.append(ref_codes[cur:]); smap.append(np.arange(cur, G)); emap.append(np.arange(cur, G) + 1)
    return (np.concatenate(pieces).astype(np.uint8), np.concatenate(smap).astype(np.int64),
            np.concatenate(emap).astype(np.int64), truth)


def _sample_reads(hap, smap, emap, starts, read_len):
    idx = starts[:, None] + np.arange(read_len, dtype=np.int64)[None, :]
    bases = hap[idx]
    pos = smap[starts]
    end = emap[starts + read_len - 1]
    return bases, pos, end


def make_batch(seed: int = 7, region_len: int = 3000, region_start: int = 1001, cov_t: float = 60.0,
               cov_n: float = 60.0, read_len: int = 100, err: float = 0.001, vaf: float = 0.5,
               var_every: int = 700, window: int = 600, delta: int = 100, paired: bool = False,
               insert_mean: float = 300.0, insert_sd: float = 30.0, chr_id: int = 22, n_in_ref: int = 0,
               low_qual_frac: float = 0.0, kinds=("snv", "ins", "del"), str_every: int = 0,
               return_truth: bool = False):
    rng = np.random.default_rng(seed)
    margin = 1000
    g0 = max(0, region_start - 1 - margin)           # 0-based genome coordinate of local index 0
    G = region_len + 2 * margin
    ref_codes = rng.integers(0, 4, G).astype(np.uint8)
    if str_every:   # config 4: STR blocks (unit 1-6 bp x 5-40 copies) every ~str_every bp
        p = int(rng.integers(0, str_every))
        while p < G - 300:
            unit = rng.integers(0, 4, int(rng.integers(1, 7))).astype(np.uint8)
            blk = np.tile(unit, int(rng.integers(5, 41)))[:250]
            ref_codes[p:p + len(blk)] = blk
            p += len(blk) + str_every + int(rng.integers(-str_every // 4, str_every // 4 + 1))
    alt, a_s, a_e, truth = _make_alt(ref_codes, rng, var_every, first=margin // 2 + int(rng.integers(0, 200)), kinds=kinds)
    r_s = np.arange(G, dtype=np.int64)
    r_e = r_s + 1

    def sample(cov, is_tumor):
        n = int(round(cov * G / read_len))
        if paired:
            n //= 2
            ins = np.clip(rng.normal(insert_mean, insert_sd, n).round().astype(np.int64), read_len, None)
        from_alt = (rng.random(n) < vaf) if is_tumor else np.zeros(n, bool)
        out = []
        for hap, sm, em, sel in ((ref_codes, r_s, r_e, ~from_alt), (alt, a_s, a_e, from_alt)):
            m = int(sel.sum())
            if m == 0:
                continue
            if paired:
                isz = ins[sel]
                fs = (rng.random(m) * (len(hap) - isz)).astype(np.int64)
                b1, p1, e1 = _sample_reads(hap, sm, em, fs, read_len)
                b2, p2, e2 = _sample_reads(hap, sm, em, fs + isz - read_len, read_len)
                out.append((b1, p1, e1, np.zeros(m, bool), np.full(m, 1), np.arange(m)))
                out.append((b2, p2, e2, np.ones(m, bool), np.full(m, 2), np.arange(m)))
            else:
                st = (rng.random(m) * (len(hap) - read_len)).astype(np.int64)
                b, p, e = _sample_reads(hap, sm, em, st, read_len)
                out.append((b, p, e, rng.random(m) < 0.5, rng.integers(1, 3, m), np.arange(m)))
        # names: unique per fragment within (sample, haplotype) group
        bases = np.concatenate([o[0] for o in out])
        pos = np.concatenate([o[1] for o in out]); end = np.concatenate([o[2] for o in out])
        rev = np.concatenate([o[3] for o in out]); mate = np.concatenate([o[4] for o in out])
        grp_base = np.cumsum([0] + [len(o[5]) for o in out])
        if paired:  # the two mate blocks of one haplotype share fragment ids
            frag = np.concatenate([o[5] + (grp_base[(i // 2) * 2]) for i, o in enumerate(out)])
        else:
            frag = np.concatenate([o[5] + grp_base[i] for i, o in enumerate(out)])
        # sequencing errors
        e_mask = rng.random(bases.shape) < err
        bases = np.where(e_mask, (bases + rng.integers(1, 4, bases.shape)) % 4, bases).astype(np.uint8)
        q = QUALS[rng.integers(0, len(QUALS), bases.shape)]
        if low_qual_frac > 0:
            lq = rng.random(bases.shape) < low_qual_frac
            q = np.where(lq, (rng.integers(2, 20, bases.shape) + 33).astype(np.uint8), q)
        order = np.argsort(pos, kind="stable")       # coordinate-sorted like a BAM
        return bases[order], q[order], pos[order], end[order], rev[order], mate[order], frag[order]

    tb, tq, tpos, tend, trev, tmate, tfrag = sample(cov_t, True)
    nb, nq, npos, nend, nrev, nmate, nfrag = sample(cov_n, False)
    nT, nN = len(tpos), len(npos)
    # name ranks: a random permutation of fragment ids (so names are not in coordinate order)
    n_names = int(tfrag.max(initial=-1)) + 1 + int(nfrag.max(initial=-1)) + 1
    perm = rng.permutation(n_names).astype(np.uint32)
    t_rank = perm[tfrag]
    n_rank = perm[nfrag + int(tfrag.max(initial=-1)) + 1]

    ref_ascii = ACGT[ref_codes].copy()
    if n_in_ref:
        lo = margin + 50
        ref_ascii[rng.integers(lo, G - lo, n_in_ref)] = ord("N")

    # windows over the region (already "padded" by the caller's choice of region)
    s0 = region_start - 1 - g0                        # local index of the first region base
    region = ref_ascii[s0:s0 + region_len]
    tiles = tile_windows(region_len, region_start, window, delta)
    ref_parts, ref_start, wr_lists = [], [], []
    # global 0-based coordinates of reads
    tpos_g, tend_g, npos_g, nend_g = tpos + g0, tend + g0, npos + g0, nend + g0
    for off, ln, rstart in tiles:
        ref_parts.append(region[off:off + ln])
        ref_start.append(rstart)
        rend = rstart + ln
        lo, hi = np.searchsorted(tpos_g, rstart, "left"), np.searchsorted(tpos_g, rend, "left")
        ti = np.arange(lo, hi)[tend_g[lo:hi] <= rend]
        lo, hi = np.searchsorted(npos_g, rstart, "left"), np.searchsorted(npos_g, rend, "left")
        ni = np.arange(lo, hi)[nend_g[lo:hi] <= rend] + nT
        wr_lists.append(np.concatenate([ti, ni]))
    flags = np.concatenate([
        (trev.astype(np.uint8) * READ_REVERSE) | (tmate.astype(np.uint8) << READ_MATE_SHIFT),
        (nrev.astype(np.uint8) * READ_REVERSE) | (nmate.astype(np.uint8) << READ_MATE_SHIFT) | READ_NORMAL])
    R = nT + nN
    base_off = np.arange(R + 1, dtype=np.uint64) * read_len
    ref_off = np.zeros(len(tiles) + 1, dtype=np.uint32)
    ref_off[1:] = np.cumsum([len(p) for p in ref_parts])
    wr_off = np.zeros(len(tiles) + 1, dtype=np.uint32)
    wr_off[1:] = np.cumsum([len(x) for x in wr_lists])
    batch = Batch(ref_off=ref_off, ref_start=np.array(ref_start), chr_id=np.full(len(tiles), chr_id),
                  wr_off=wr_off, wr_idx=np.concatenate(wr_lists) if wr_lists else np.zeros(0, np.uint32),
                  base_off=base_off, flags=flags, name_rank=np.concatenate([t_rank, n_rank]),
                  ref_seq=np.concatenate(ref_parts), seq=ACGT[np.concatenate([tb, nb])].reshape(-1),
                  qual=np.concatenate([tq, nq]).reshape(-1))
    if return_truth:
        return batch, [(k, p + g0 + 1, L) for k, p, L in truth]
    return batch
