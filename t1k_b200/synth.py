"""Seeded synthetic allele references and read sets.

No HLA/KIR allele FASTA ships with the reference and there is no network, so every
workload that is not the tiny in-tree CYP2D6 set is generated here.  Formats follow the
reference's inputs: reference FASTA header ``>GENE*fields exonCnt s1 e1 ...`` (0-based
inclusive exon coordinates, parsed by SeqSet::InputRefSeq, /root/reference/SeqSet.hpp:933-976),
single ``N`` between truncated introns (SeqSet.hpp:924-928), reads upper-case ACGTN in FR
orientation (SeqSet.hpp:2369-2380).

Everything is deterministic in (parameters, seed): numpy Generator(PCG64) streams only.
"""
from __future__ import annotations

import numpy as np

_ALPHA = np.frombuffer(b"ACGT", dtype=np.uint8)
_COMP = np.zeros(256, dtype=np.uint8)
for _a, _b in zip(b"ACGTN", b"TGCAN"):
    _COMP[_a] = _b


def _rand_seq(rng, n):
    return _ALPHA[rng.integers(0, 4, size=n)]


def _mutate(rng, base, sites, alts, lo, hi):
    m = int(rng.integers(lo, hi + 1))
    pick = rng.choice(len(sites), size=m, replace=False)
    s = base.copy()
    s[sites[pick]] = alts[pick]
    return s


def make_hla_rna_ref(genes=None, length=1100, n_sites=160, min_sub=3, max_sub=14, seed=11):
    """HLA-RNA-like reference: per gene a random base sequence, each allele = base with
    3..14 substitutions drawn from a fixed per-gene pool of biallelic sites.
    Returns list of (name, comment, seq_bytes)."""
    if genes is None:
        genes = [("HLA-A", 8000), ("HLA-B", 9000), ("HLA-C", 8000),
                 ("HLA-DRB1", 3000), ("HLA-DQB1", 1500), ("HLA-DPB1", 500)]
    rng = np.random.default_rng(seed)
    out = []
    for gname, cnt in genes:
        base = _rand_seq(rng, length)
        sites = np.sort(rng.choice(length, size=min(n_sites, length), replace=False))
        alts = _ALPHA[(np.searchsorted(_ALPHA, base[sites]) + rng.integers(1, 4, size=len(sites))) % 4]
        comment = "1 0 %d" % (length - 1)
        for i in range(cnt):
            s = _mutate(rng, base, sites, alts, min_sub, max_sub)
            name = "%s*%02d:%02d:01" % (gname, i // 99 + 1, i % 99 + 1)
            out.append((name, comment, s.tobytes()))
    return out


def make_dna_ref(n_genes=17, alleles_per_gene=90, n_exons=9, exon_mean=300, pad=200,
                 n_sites=120, min_sub=2, max_sub=10, seed=23, prefix="KIR", family_div=0.0):
    """KIR/HLA-DNA-like reference: per gene n_exons exons, each kept with +-pad bp of intron,
    the padded blocks joined by a single 'N'; the header comment lists the exon coordinates.
    family_div > 0 derives every gene from one ancestor with that substitution rate (gene family)."""
    rng = np.random.default_rng(seed)
    out = []
    ancestor = None
    for g in range(n_genes):
        exon_len = np.maximum(40, rng.normal(exon_mean, exon_mean * 0.25, size=n_exons).astype(int))
        if ancestor is None or family_div <= 0:
            blocks = [_rand_seq(rng, int(l) + 2 * pad) for l in exon_len]
            if family_div > 0:
                ancestor = ([b.copy() for b in blocks], exon_len.copy())
        else:
            blocks = []
            exon_len = ancestor[1]
            for b in ancestor[0]:
                b2 = b.copy()
                mask = rng.random(len(b2)) < family_div
                b2[mask] = _rand_seq(rng, int(mask.sum()))
                blocks.append(b2)
        parts = []
        exons = []
        pos = 0
        for bi, b in enumerate(blocks):
            if bi > 0:
                parts.append(np.frombuffer(b"N", dtype=np.uint8))
                pos += 1
            exons.append((pos + pad, pos + pad + int(exon_len[bi]) - 1))
            parts.append(b)
            pos += len(b)
        base = np.concatenate(parts)
        valid = np.nonzero(base != ord("N"))[0]
        sites = np.sort(rng.choice(valid, size=min(n_sites, len(valid)), replace=False))
        alts = _ALPHA[(np.searchsorted(_ALPHA, base[sites]) + rng.integers(1, 4, size=len(sites))) % 4]
        comment = "%d %s" % (len(exons), " ".join("%d %d" % e for e in exons))
        gname = "%s%dDL%d" % (prefix, 2 + g % 2, g + 1)
        for i in range(alleles_per_gene):
            s = _mutate(rng, base, sites, alts, min_sub, max_sub)
            name = "%s*%03d%02d" % (gname, i // 20 + 1, i % 20 + 1)
            out.append((name, comment, s.tobytes()))
    return out


def write_fasta(path, records):
    with open(path, "wb") as f:
        for name, comment, seq in records:
            f.write(b">" + name.encode() + (b" " + comment.encode() if comment else b"") + b"\n")
            f.write(seq + b"\n")


def read_fasta(path):
    out = []
    name = None
    with open(path, "rb") as f:
        for line in f:
            line = line.rstrip(b"\r\n")
            if line.startswith(b">"):
                if name is not None:
                    out.append((name, comment, b"".join(chunks)))
                head = line[1:].split(None, 1)
                name = head[0].decode()
                comment = head[1].decode() if len(head) > 1 else ""
                chunks = []
            elif name is not None:
                chunks.append(line)
    if name is not None:
        out.append((name, comment, b"".join(chunks)))
    return out


def revcomp(a: np.ndarray) -> np.ndarray:
    return _COMP[a[..., ::-1]]


def simulate_pairs(records, n_pairs, read_len=150, insert=(300, 450), err=0.002, n_rate=0.0,
                   alleles_per_gene=2, seed=1, gene_of=None, single_end=False, indel_rate=0.0, src_seed=None,
                   start_frac=(0.0, 1.0)):
    """Draw n_pairs FR fragments uniformly from `alleles_per_gene` alleles of every gene.
    Returns (reads1, reads2) as uint8 arrays [n, read_len] (reads2 None if single_end),
    plus the list of source allele indices.  Substitution errors at rate `err`, 'N' at
    `n_rate`, single-base indels (read-level) at `indel_rate` per read.  `src_seed` draws the source alleles from
    their own stream, so that shards of one sample (different `seed`) come from the same alleles.  `start_frac` restricts
    the fragment starts to that fraction of every source allele (a small sample with the duplicate rate of a deep one)."""
    rng = np.random.default_rng(seed)
    src_rng = rng if src_seed is None else np.random.default_rng(src_seed)
    if gene_of is None:
        gene_of = [r[0].split("*")[0] for r in records]
    genes = {}
    for i, g in enumerate(gene_of):
        genes.setdefault(g, []).append(i)
    src = []
    for g in sorted(genes):
        idx = genes[g]
        k = min(alleles_per_gene, len(idx))
        src.extend(int(x) for x in src_rng.choice(idx, size=k, replace=False))
    seqs = [np.frombuffer(records[i][2], dtype=np.uint8) for i in src]
    which = rng.integers(0, len(src), size=n_pairs)
    ins = rng.integers(insert[0], insert[1] + 1, size=n_pairs)
    flip = rng.integers(0, 2, size=n_pairs).astype(bool)
    u = start_frac[0] + rng.random(n_pairs) * (start_frac[1] - start_frac[0])
    r1 = np.empty((n_pairs, read_len), dtype=np.uint8)
    r2 = None if single_end else np.empty((n_pairs, read_len), dtype=np.uint8)
    ar = np.arange(read_len)
    for si, s in enumerate(seqs):
        sel = np.nonzero(which == si)[0]
        if len(sel) == 0:
            continue
        L = len(s)
        fl = np.minimum(ins[sel], L)
        fl = np.maximum(fl, read_len)
        if L < read_len:
            raise ValueError("allele shorter than read")
        start = np.floor(u[sel] * (L - fl + 1)).astype(np.int64)
        left = s[start[:, None] + ar[None, :]]
        right = revcomp(s[(start + fl - read_len)[:, None] + ar[None, :]])
        f = flip[sel]
        a = np.where(f[:, None], right, left)
        b = np.where(f[:, None], left, right)
        r1[sel] = a
        if r2 is not None:
            r2[sel] = b
    for arr in ([r1] if r2 is None else [r1, r2]):
        if err > 0:
            m = rng.random(arr.shape) < err
            cnt = int(m.sum())
            cur = np.searchsorted(_ALPHA, np.where(arr[m] == ord("N"), ord("A"), arr[m]))
            arr[m] = _ALPHA[(cur + rng.integers(1, 4, size=cnt)) % 4]
        if indel_rate > 0:
            rows = np.nonzero(rng.random(arr.shape[0]) < indel_rate)[0]
            for r in rows:
                p = int(rng.integers(20, read_len - 20))
                if rng.integers(0, 2):
                    arr[r, p + 1:] = arr[r, p:-1].copy()        # insertion (duplicate base p)
                    arr[r, p] = _ALPHA[rng.integers(0, 4)]
                else:
                    arr[r, p:-1] = arr[r, p + 1:].copy()        # deletion, pad the tail
                    arr[r, -1] = _ALPHA[rng.integers(0, 4)]
        if n_rate > 0:
            m = rng.random(arr.shape) < n_rate
            arr[m] = ord("N")
    return r1, r2, src


def write_fastq(path, reads, prefix="r"):
    n, L = reads.shape
    qual = b"I" * L
    with open(path, "wb") as f:
        for i in range(n):
            f.write(b"@%s%d\n" % (prefix.encode(), i))
            f.write(reads[i].tobytes())
            f.write(b"\n+\n" + qual + b"\n")
