"""Reads longer than 255 bases (2 x 300 MiSeq, merged pairs): the reference handles any length (AlignAlgo.hpp:251-253 heap
DP; README "any read length").  Seeded workloads shared by the golden generator (tests/golden/make_golden_long.py), the CPU
emulation tests and the GPU parity tests."""
from __future__ import annotations

import numpy as np

import workloads as W
from t1k_b200 import synth


def _reads(records, n, read_len, seed, err, n_rate, indel_rate):
    r1, _, _ = synth.simulate_pairs(records, n, read_len=read_len, insert=(read_len, read_len), err=err, n_rate=n_rate, alleles_per_gene=3,
                                    seed=seed, single_end=True, indel_rate=indel_rate)
    rng = np.random.default_rng(seed + 1)
    out = []
    comp = bytes.maketrans(b"ACGTN", b"TGCAN")
    for r in r1:
        s = r.tobytes()
        out.append(s.translate(comp)[::-1] if rng.integers(0, 2) else s)
    return out


def cases():
    """(name, records, reads, similarity, relax)"""
    rna = synth.make_hla_rna_ref(genes=[("HLA-A", 40), ("HLA-B", 30)], length=900, n_sites=80, min_sub=1, max_sub=10, seed=61)
    yield "rna_300", rna, _reads(rna, 60, 300, 62, 0.004, 0.0005, 0.05), 0.9, False
    yield "rna_600", rna, _reads(rna, 40, 600, 63, 0.01, 0.001, 0.1), 0.8, False
    dna = synth.make_dna_ref(n_genes=3, alleles_per_gene=20, n_exons=4, exon_mean=200, pad=150, n_sites=60, min_sub=1, max_sub=6, seed=64, family_div=0.04)
    yield "dna_320_relax", dna, _reads(dna, 60, 320, 65, 0.006, 0.001, 0.05), 0.9, True
    yield "dna_256", dna, _reads(dna, 40, 256, 66, 0.004, 0.0, 0.02), 0.9, False
    rna2 = synth.make_hla_rna_ref(genes=[("HLA-C", 24)], length=1500, n_sites=90, min_sub=1, max_sub=12, seed=68)
    yield "rna_1000", rna2, _reads(rna2, 20, 1000, 67, 0.005, 0.0005, 0.1), 0.85, False
