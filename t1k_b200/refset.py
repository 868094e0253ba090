"""Host-side loading of the allele reference: what Genotyper::InitRefSet + SeqSet::InputRefSeq +
SeqSet::UpdateDnaSeqWeight + Genotyper::InitAlleleInfo do before the hot path starts
(/root/reference/Genotyper.hpp:707-730,559-682; SeqSet.hpp:906-982,1008-1029).  Pure host bookkeeping
(names, exon coordinates, weights); the sequences themselves go to the device through t1k_ref_create."""
from __future__ import annotations

import gzip

import numpy as np


def read_fasta(path):
    """(name, comment, sequence bytes) per record; multi-line sequences are joined (kseq semantics)."""
    out = []
    name = None
    op = gzip.open if str(path).endswith(".gz") else open
    with op(path, "rb") as f:
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


def parse_exons(comment, length):
    """SeqSet.hpp:933-976: digits separated by anything; first number is the exon count, the rest are
    (start, end) pairs, 0-based inclusive; no comment => one exon spanning the sequence."""
    nums, n = [], 0
    for ch in comment:
        if "0" <= ch <= "9":
            n = n * 10 + ord(ch) - 48
        else:
            nums.append(n)
            n = 0
    if n:
        nums.append(n)
    if not comment or not nums:
        return [(0, length - 1)]
    return [(nums[i], nums[i + 1]) for i in range(1, len(nums) - 1, 2)]


def parse_allele_name(allele, digit_units=-1, delimiter=""):
    """Genotyper::ParseAlleleName (Genotyper.hpp:63-131), fieldsType 0 -> (gene, majorAllele)."""
    parse_type, fields, delim = 1, digit_units, ""
    if fields == -1:
        fields = 3
        if ":" in allele:
            delim, parse_type = ":", 2
    if delimiter:
        delim, parse_type = delimiter, 2
    i = allele.find("*")
    if i < 0:
        i = len(allele)
    if parse_type == 1:
        return allele[:i], allele[:min(len(allele), i + fields + 1)]
    k, j = 0, i
    while j < len(allele):
        if allele[j] == delim:
            k += 1
            if k >= fields:
                break
        j += 1
    return allele[:i], allele[:j]


class RefSet:
    """The de-duplicated allele list with everything the hot path and the EM need on the host."""

    def __init__(self, records, digit_units=-1, delimiter=""):
        seen = {}
        self.names, self.seqs, self.comments, weight = [], [], [], []
        for name, comment, seq in records:          # Genotyper.hpp:717-725: identical sequences collapse, weight++
            k = seen.get(seq)
            if k is None:
                seen[seq] = len(self.seqs)
                self.names.append(name)
                self.comments.append(comment)
                self.seqs.append(seq)
                weight.append(1)
            else:
                weight[k] += 1
        self.n = len(self.seqs)
        self.exons = [parse_exons(c, len(s)) for c, s in zip(self.comments, self.seqs)]
        # rnaData turns false as soon as one allele has a gap between consecutive exons (SeqSet.hpp:705-713)
        self.rna = not any(ex[i][0] > ex[i - 1][1] + 1 for ex in self.exons for i in range(1, len(ex)))
        if not self.rna:                             # SeqSet::UpdateDnaSeqWeight
            keys, tot = [], {}
            for s, ex in zip(self.seqs, self.exons):
                m = np.zeros(len(s), dtype=bool)
                for a, b in ex:
                    m[a:min(b, len(s) - 1) + 1] = True
                keys.append(np.frombuffer(s, dtype=np.uint8)[m].tobytes())
            for k, w in zip(keys, weight):
                tot[k] = tot.get(k, 0) + w
            weight = [tot[k] for k in keys]
        self.seq_weight = np.asarray(weight, dtype=np.int32)
        # SeqSet::ComputeEffectiveLen (SeqSet.hpp:747-758): a run of N counts once
        eff = []
        for s in self.seqs:
            a = np.frombuffer(s, dtype=np.uint8)
            isn = a == ord("N")
            prev_n = np.concatenate([[True], isn[:-1]])     # position 0 counts only if it is not N
            eff.append(int(np.count_nonzero(~isn | ~prev_n)))
        # Genotyper::InitAlleleInfo: gene / major-allele ids in first-appearance order, large-deletion length fix
        genes, majors, gi, mi = {}, {}, [], []
        for nme in self.names:
            g, m = parse_allele_name(nme, digit_units, delimiter)
            gi.append(genes.setdefault(g, len(genes)))
            mi.append(majors.setdefault(m, len(majors)))
        self.gene_names = list(genes)
        self.major_names = list(majors)
        self.allele_gene = np.asarray(gi, dtype=np.int32)
        self.allele_major = np.asarray(mi, dtype=np.int32)
        eff = np.asarray(eff, dtype=np.int32)
        adj = eff.copy()
        for g in range(len(genes)):
            ids = np.nonzero(self.allele_gene == g)[0]
            vals, cnts = np.unique(eff[ids], return_counts=True)     # ascending; first maximum wins (Genotyper.hpp:659-671)
            mode = int(vals[int(np.argmax(cnts))])
            adj[ids[eff[ids] < mode - 500]] = mode
        self.effective_len = adj

    @classmethod
    def from_fasta(cls, path, **kw):
        return cls(read_fasta(path), **kw)

    def packed(self):
        bases = b"".join(self.seqs)
        off = np.zeros(self.n + 1, dtype=np.int64)
        np.cumsum([len(s) for s in self.seqs], out=off[1:])
        ptr, se = [0], []
        for ex in self.exons:
            for a, b in ex:
                se.extend((a, b))
            ptr.append(len(se) // 2)
        return bases, off, np.asarray(ptr, dtype=np.int32), np.asarray(se if se else [0, 0], dtype=np.int32)
