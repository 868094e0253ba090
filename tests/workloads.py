"""Seeded small workloads shared by the parity tests (reference sets + read sets)."""
from __future__ import annotations

import os

import numpy as np

from t1k_b200 import synth

CYP_DIR = "/root/reference/vcf_database/cyp2d6_idx"


def small_rna_ref(seed=5):
    return synth.make_hla_rna_ref(genes=[("HLA-A", 60), ("HLA-B", 50), ("HLA-C", 30)], length=700, n_sites=60,
                                  min_sub=1, max_sub=8, seed=seed)


def small_dna_ref(seed=7):
    return synth.make_dna_ref(n_genes=4, alleles_per_gene=25, n_exons=4, exon_mean=160, pad=120, n_sites=50,
                              min_sub=1, max_sub=6, seed=seed, family_div=0.04)


def reads_for(records, n_pairs, read_len=100, seed=1, err=0.01, n_rate=0.002, indel_rate=0.05, insert=(180, 320),
              single_end=False):
    r1, r2, _ = synth.simulate_pairs(records, n_pairs, read_len=read_len, insert=insert, err=err, n_rate=n_rate,
                                     alleles_per_gene=2, seed=seed, single_end=single_end, indel_rate=indel_rate)
    return r1, r2


def cyp2d6(kind):
    p = os.path.join(CYP_DIR, "cyp2d6_%s_seq.fa" % kind)
    return synth.read_fasta(p) if os.path.exists(p) else None


def write_lines(path, reads, weights=None):
    with open(path, "wb") as f:
        for i in range(reads.shape[0]):
            f.write(reads[i].tobytes())
            if weights is not None:
                f.write(b" %d" % weights[i])
            f.write(b"\n")
