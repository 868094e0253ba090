"""Seeded workloads whose alleles repeat a segment inside themselves (intra-allele duplications, tandem repeats): several
seed overlaps of one read on ONE allele tie in (matchCnt, similarity, read span) and only the seqStart / seqEnd tail of
`_overlap::operator<` (SeqSet.hpp:103-127) orders them.  Shared by the golden generator (tests/golden/make_golden_dup.py),
the CPU emulation tests and the GPU parity tests."""
from __future__ import annotations

import numpy as np

_COMP = bytes.maketrans(b"ACGTN", b"TGCAN")


def _rnd(rng, n):
    return bytes(rng.choice(list(b"ACGT"), size=n).astype(np.uint8))


def _flip(c):
    return {65: 67, 67: 71, 71: 84, 84: 65}[c]


def _mutate(rng, s, n_sub):
    s = bytearray(s)
    for _ in range(n_sub):
        p = int(rng.integers(0, len(s)))
        s[p] = _flip(s[p])
    return bytes(s)


def _rc(s):
    return s.translate(_COMP)[::-1]


def triple_segment():
    """One allele that holds the same 120-bp segment three times behind flanks of decreasing identity, and reads that
    start in the first flank and run into the segment: three seed overlaps on one allele with equal readStart / readEnd."""
    rng = np.random.default_rng(1)
    F1, S, X, Y, tail = _rnd(rng, 100), _rnd(rng, 120), _rnd(rng, 50), _rnd(rng, 50), _rnd(rng, 80)
    F2 = bytearray(F1[-40:])
    for p in (3, 9, 15, 21, 27, 33, 39):
        F2[p] = _flip(F2[p])
    F3 = bytearray(_flip(c) for c in F1[-40:])
    allele = F1 + S + X + bytes(F2) + S + Y + bytes(F3) + S + tail
    reads = [F1[-40:] + S[:60], _rc(F1[-40:] + S[:60]), S[:100], S[10:110], F1[-20:] + S[:80], bytes(F2[-30:]) + S[:70],
             S[60:] + X[:40], S[60:] + Y[:40], S[40:] + tail[:40]]
    other = _rnd(rng, 500)
    recs = [("D*01:01", "", allele), ("D*02:01", "", other)]
    return recs, reads


def duplication_set(n_alleles=30, dup=140, seed=5, n_reads=360, read_len=100):
    """n_alleles variants of a gene with a `dup`-bp segment present twice (a few substitutions apart), reads drawn from
    the alleles (both strands, substitution errors)."""
    rng = np.random.default_rng(seed)
    A, D, B, C = _rnd(rng, 220), _rnd(rng, dup), _rnd(rng, 180), _rnd(rng, 240)
    recs = []
    for i in range(n_alleles):
        d2 = _mutate(rng, D, int(rng.integers(0, 3)))
        s = _mutate(rng, A + D + B + d2 + C, int(rng.integers(0, 7)))
        recs.append(("DUP*%02d:01" % i, "", s))
    reads = []
    for _ in range(n_reads):
        a = recs[int(rng.integers(0, n_alleles))][2]
        st = int(rng.integers(0, len(a) - read_len))
        r = _mutate(rng, a[st:st + read_len], int(rng.integers(0, 3)))
        reads.append(_rc(r) if rng.integers(0, 2) else r)
    return recs, reads


def tandem_set(seed=9, n_reads=300):
    """Alleles with tandem repeats of unit length 11 ... 40 (k = 11: the unit is at least one k-mer long), 3-6 copies,
    variants with one copy more or fewer and point substitutions; reads of 60-150 bases across the repeats."""
    rng = np.random.default_rng(seed)
    recs = []
    for g, unit_len in enumerate((11, 12, 13, 17, 23, 31, 40)):
        unit = _rnd(rng, unit_len)
        left, right = _rnd(rng, 150), _rnd(rng, 150)
        for v in range(4):
            copies = int(rng.integers(3, 7))
            body = unit * copies
            s = _mutate(rng, left + body + right, v)
            recs.append(("TR%d*%02d:01" % (g, v), "", s))
    reads = []
    for _ in range(n_reads):
        a = recs[int(rng.integers(0, len(recs)))][2]
        L = int(rng.integers(60, 151))
        if L >= len(a):
            L = len(a) - 1
        st = int(rng.integers(0, len(a) - L))
        r = _mutate(rng, a[st:st + L], int(rng.integers(0, 3)))
        reads.append(_rc(r) if rng.integers(0, 2) else r)
    return recs, reads


def cases():
    """(name, records, reads, similarity, relax)"""
    recs, reads = triple_segment()
    yield "triple_s080", recs, reads, 0.8, False
    yield "triple_s085", recs, reads, 0.85, False
    yield "triple_s090", recs, reads, 0.9, False
    recs, reads = duplication_set()
    yield "dup140_s080", recs, reads, 0.8, False
    yield "dup140_s090", recs, reads, 0.9, False
    recs, reads = tandem_set()
    yield "tandem_s080", recs, reads, 0.8, False
    yield "tandem_s090", recs, reads, 0.9, False
