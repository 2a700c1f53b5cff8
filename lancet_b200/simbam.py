"""Synthetic tumour/normal SAM + FASTA generator for the end-to-end (BAM -> VCF) parity tests.

Test tooling only.  Reads are sampled from explicit haplotypes so that CIGAR, MD and NM are exact (what an aligner
would report for the true alignment); the SAM text is turned into BAM + BAI by htslib's own `test_view` / `test_index`
built under oracle/_ref (oracle/Makefile).  Records exercise every read filter of the reference's extractReads
(src/Microassembler.cc:434-655): duplicates, low MAPQ, secondary alignments, small AS-XS deltas, XT:A:R, XA:Z,
soft clips, overlapping mates, unmapped mates placed at the mate position.
"""
import os
import subprocess
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFBIN = os.path.join(ROOT, "oracle", "_ref")
BASES = np.frombuffer(b"ACGT", dtype=np.uint8)


def random_ref(rng, n, str_every=0):
    s = BASES[rng.integers(0, 4, n)].copy()
    if str_every:
        for p in range(str_every, n - 60, str_every):
            unit = BASES[rng.integers(0, 4, int(rng.integers(1, 5)))]
            reps = int(rng.integers(4, 12))
            t = np.tile(unit, reps)
            s[p:p + len(t)] = t[: max(0, min(len(t), n - p))]
    return s.tobytes().decode()


def make_variants(rng, ref, every, start=300):
    """list of (pos0, kind, payload): 'S' base, 'I' inserted string (after pos0), 'D' deleted length (starting at pos0+1)"""
    out = []
    p = start
    while p < len(ref) - 300:
        r = rng.random()
        if r < 0.55:
            alt = "ACGT"[(("ACGT".index(ref[p]) if ref[p] in "ACGT" else 0) + int(rng.integers(1, 4))) % 4]
            out.append((p, "S", alt))
        elif r < 0.78:
            out.append((p, "I", "".join("ACGT"[i] for i in rng.integers(0, 4, int(rng.integers(1, 12))))))
        else:
            out.append((p, "D", int(rng.integers(1, 15))))
        p += int(every * (0.6 + 0.8 * rng.random()))
    return out


def haplotype(ref, variants):
    """alignment columns (ref_pos0 or -1, hap_base or '')"""
    cols = []
    vs = {v[0]: v for v in variants}
    i = 0
    n = len(ref)
    while i < n:
        v = vs.get(i)
        if v is None:
            cols.append((i, ref[i])); i += 1
        elif v[1] == "S":
            cols.append((i, v[2])); i += 1
        elif v[1] == "I":
            cols.append((i, ref[i]))
            for b in v[2]:
                cols.append((-1, b))
            i += 1
        else:
            cols.append((i, ref[i]))
            for j in range(1, v[2] + 1):
                if i + j < n:
                    cols.append((i + j, ""))
            i += v[2] + 1
    return cols


def _read_from_cols(rng, ref, cols, hap_idx, a, L, err, q_hi, q_lo, low_frac):
    """read covering hap bases a..a+L-1 -> (pos1, cigar, md, nm, seq, qual)"""
    c0, c1 = hap_idx[a], hap_idx[a + L - 1]
    seg = cols[c0:c1 + 1]
    seq = []; ops = []; md = []; nm = 0; run = 0; pos1 = None; in_del = False
    # leading / trailing insertions become soft clips
    lead = 0
    while lead < len(seg) and seg[lead][0] < 0:
        lead += 1
    trail = 0
    while trail < len(seg) and seg[len(seg) - 1 - trail][0] < 0:
        trail += 1
    for k, (rp, hb) in enumerate(seg):
        if hb == "":
            ops.append("D"); nm += 1
            if not in_del:
                md.append(str(run)); run = 0; md.append("^"); in_del = True
            md.append(ref[rp]); continue
        b = hb
        if rng.random() < err:
            b = "ACGT"[(("ACGT".index(hb) if hb in "ACGT" else 0) + int(rng.integers(1, 4))) % 4]
        seq.append(b)
        if rp < 0:
            ops.append("S" if (k < lead or k >= len(seg) - trail) else "I")
            if ops[-1] == "I":
                nm += 1
            continue
        if pos1 is None:
            pos1 = rp + 1
        ops.append("M")
        if b == ref[rp]:
            run += 1; in_del = False
        else:
            md.append(str(run)); run = 0; md.append(ref[rp]); nm += 1; in_del = False
    md.append(str(run))          # MD grammar: a number (possibly 0) separates every mismatch / deletion
    md_s = "".join(md)
    cig = []
    for o in ops:
        if cig and cig[-1][1] == o:
            cig[-1][0] += 1
        else:
            cig.append([1, o])
    cigar = "".join(f"{n}{o}" for n, o in cig)
    q = np.where(rng.random(len(seq)) < low_frac, q_lo, q_hi).astype(np.uint8)
    return pos1, cigar, md_s, nm, "".join(seq), (q + 33).tobytes().decode()


def sample_reads(rng, chrom, ref, haps, weights, cov, L=100, err=0.001, insert_mean=320, insert_sd=30, rg="rg1", prefix="r",
                 q_hi=35, q_lo=8, low_frac=0.02, odd_frac=0.04):
    """-> list of SAM lines (unsorted) for one sample on one chromosome"""
    n_pairs = int(cov * len(ref) / (2 * L))
    hcols = [haplotype(ref, h) for h in haps]
    hidx = [[i for i, c in enumerate(cols) if c[1] != ""] for cols in hcols]
    lines = []
    for p in range(n_pairs):
        h = int(rng.choice(len(haps), p=weights))
        cols, idx = hcols[h], hidx[h]
        ins = max(L + 5, int(rng.normal(insert_mean, insert_sd)))
        if rng.random() < 0.05:
            ins = int(rng.integers(L + 5, 2 * L - 10))          # overlapping mates
        if ins >= len(idx) - 2:
            continue
        a = int(rng.integers(0, len(idx) - ins))
        r1 = _read_from_cols(rng, ref, cols, idx, a, L, err, q_hi, q_lo, low_frac)
        r2 = _read_from_cols(rng, ref, cols, idx, a + ins - L, L, err, q_hi, q_lo, low_frac)
        if r1[0] is None or r2[0] is None:
            continue
        first_fwd = rng.random() < 0.5
        name = f"{prefix}{p}"
        mapq = 60; AS = L - 5 * 0; XS = int(rng.integers(0, 60)); extra1 = ""; extra2 = ""
        fl_extra = 0
        u = rng.random()
        if u < odd_frac * 0.2: fl_extra = 0x400                              # duplicate
        elif u < odd_frac * 0.4: mapq = int(rng.integers(0, 15))             # low MAPQ
        elif u < odd_frac * 0.55: fl_extra = 0x100                           # secondary
        elif u < odd_frac * 0.7: XS = L - int(rng.integers(0, 6))            # |AS-XS| <= 5
        elif u < odd_frac * 0.75: extra1 = "\tXT:A:R"                       # last tag: seen as "R" by the reference
        elif u < odd_frac * 0.8: extra1 = "XT:A:R\t"                        # first tag (bwa aln order): the reference's strlen() runs into the next tag
        elif u < odd_frac * 0.9: extra2 = "\tXA:Z:chrZ,+100,100M,1;"
        tlen = (r2[0] + L) - r1[0]
        unmapped2 = u >= odd_frac * 0.9 and u < odd_frac * 0.95                # mate 2 unmapped, placed at mate 1's position
        for k, r in enumerate((r1, r2)):
            pos1, cigar, md, nm, seq, qual = r
            other = r2 if k == 0 else r1
            flag = 0x1 | 0x2 | fl_extra
            is_first = (k == 0) == first_fwd
            flag |= 0x40 if is_first else 0x80
            if k == 1: flag |= 0x10
            else: flag |= 0x20
            tl = tlen if k == 0 else -tlen
            asv = L - 6 * nm
            tags = f"NM:i:{nm}\tMD:Z:{md}\tAS:i:{asv}\tXS:i:{min(XS, 255)}\tRG:Z:{rg}"
            ex = extra1 if k == 0 else extra2
            tags = ex + tags if ex.endswith("\t") else tags + ex
            if unmapped2:
                if k == 0:
                    flag = (flag & ~0x2) | 0x8
                    lines.append((pos1, f"{name}\t{flag}\t{chrom}\t{pos1}\t{mapq}\t{cigar}\t=\t{pos1}\t0\t{seq}\t{qual}\t{tags}"))
                else:
                    flag = (flag & ~(0x2 | 0x10)) | 0x4
                    lines.append((r1[0], f"{name}\t{flag}\t{chrom}\t{r1[0]}\t0\t*\t=\t{r1[0]}\t0\t{seq}\t{qual}\tRG:Z:{rg}"))
                continue
            lines.append((pos1, f"{name}\t{flag}\t{chrom}\t{pos1}\t{mapq}\t{cigar}\t=\t{other[0]}\t{tl}\t{seq}\t{qual}\t{tags}"))
    return lines


def write_dataset(outdir, seed=1, chroms=(("chr22", 4000),), cov_t=60, cov_n=60, var_every=180, som_every=400, vaf=0.35, err=0.001,
                  str_every=0, L=100, n_in_ref=False, odd_frac=0.04, low_frac=0.02):
    """writes ref.fa, tumor.bam(.bai), normal.bam(.bai) under outdir; returns dict of paths and truth"""
    os.makedirs(outdir, exist_ok=True)
    rng = np.random.default_rng(seed)
    refs = {}
    samT, samN = [], []
    truth = {}
    for ci, (name, n) in enumerate(chroms):
        ref = random_ref(rng, n, str_every)
        if n_in_ref:
            q = n // 2
            ref = ref[:q] + "N" * 7 + ref[q + 7:]
        refs[name] = ref
        germ = make_variants(rng, ref, var_every)
        som = [v for v in make_variants(rng, ref, som_every, start=360) if all(abs(v[0] - g[0]) > 25 for g in germ)]
        hom = [g for i, g in enumerate(germ) if i % 3 == 0]
        het = [g for i, g in enumerate(germ) if i % 3 != 0]
        hA, hB = sorted(hom + het), sorted(hom)
        hS = sorted(hom + het + som)
        samN += [(ci, p, l) for p, l in sample_reads(rng, name, ref, [hA, hB], [0.5, 0.5], cov_n, L, err, rg="rgN", prefix=f"n{ci}_", odd_frac=odd_frac, low_frac=low_frac)]
        samT += [(ci, p, l) for p, l in sample_reads(rng, name, ref, [hA, hB, hS], [0.5 - vaf / 2, 0.5 - vaf / 2, vaf], cov_t, L, err, rg="rgT", prefix=f"t{ci}_", odd_frac=odd_frac, low_frac=low_frac)]
        truth[name] = {"germline": germ, "somatic": som}
    fa = os.path.join(outdir, "ref.fa")
    with open(fa, "w") as f:
        for name, _ in chroms:
            f.write(f">{name}\n")
            s = refs[name]
            for i in range(0, len(s), 60):
                f.write(s[i:i + 60] + "\n")
    out = {"ref": fa, "truth": truth, "refs": refs}
    for tag, sam, sm, rg in (("tumor", samT, "TUMOR_S", "rgT"), ("normal", samN, "NORMAL_S", "rgN")):
        sam.sort(key=lambda t: (t[0], t[1]))
        sp = os.path.join(outdir, tag + ".sam"); bp = os.path.join(outdir, tag + ".bam")
        with open(sp, "w") as f:
            f.write("@HD\tVN:1.6\tSO:coordinate\n")
            for name, n in chroms:
                f.write(f"@SQ\tSN:{name}\tLN:{n}\n")
            f.write(f"@RG\tID:{rg}\tSM:{sm}\tPL:ILLUMINA\n")
            for _, _, l in sam:
                f.write(l + "\n")
        subprocess.run([os.path.join(REFBIN, "test_view"), "-S", "-b", "-p", bp, sp], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        subprocess.run([os.path.join(REFBIN, "test_index"), "-b", bp], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        out[tag] = bp
    return out


def normalise_vcf(text):
    """drop the two header lines that legitimately differ between runs"""
    return "\n".join(l for l in text.splitlines() if not l.startswith("##fileDate=") and not l.startswith("##cmdline=") and not l.startswith("##reference="))
