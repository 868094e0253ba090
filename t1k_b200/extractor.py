"""Host-side mirror of fastq-extractor's candidate filter over the C ABI: the parameter set-up of
FastqExtractor.cpp main() (:381-418) and IsGoodCandidate (:113-118) per read, a pair is kept if either mate is good
(:199-212).  All compute happens in the CUDA library (k_filter); nothing here falls back to the CPU."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L
from .genotyper import _reads_to_batch


def infer_kmer_length(total_len: int) -> int:
    """SeqSet::InferKmerLength (SeqSet.hpp:2830-2845): base-4 digits of the total length + 1."""
    d = 0
    while total_len:
        d += 1
        total_len //= 4
    return d + 1


class CandidateFilter:
    def __init__(self, records, reads1, paired, similarity=0.8, device=-1):
        """records: (name, comment, sequence) of the extraction reference; reads1: the first mates (their first 1000 give the
        mean read length as in FastqExtractor.cpp:381-407)."""
        seqs = [r[2] for r in records]
        bases = b"".join(seqs)
        off = np.zeros(len(seqs) + 1, dtype=np.int64)
        np.cumsum([len(x) for x in seqs], out=off[1:])
        first = reads1[:1000]
        hit_len = 27 if paired else 23
        mean5 = sum(len(r) for r in first) // (max(1, len(first)) * 5)
        if mean5 > hit_len:
            hit_len = mean5
        k = 9
        inferred = infer_kmer_length(int(off[-1]))
        if inferred > k:
            k = inferred
            if k > hit_len:
                hit_len = k
        self.k, self.hit_len = k, hit_len
        self._keep = (bases, off)
        d = L.FilterDesc(len(seqs), bases, L.ptr(off), k, hit_len, float(similarity), device)
        h = C.c_void_p()
        L.check(L.lib().t1k_filter_create(C.byref(d), C.byref(h)))
        self.h = h

    def __del__(self):
        if getattr(self, "h", None):
            L.lib().t1k_filter_destroy(self.h)
            self.h = None

    def IsGoodCandidate(self, reads, with_stats=False):
        """-> uint8 per read (batched)"""
        bases, off, lens = _reads_to_batch(reads)
        good = np.zeros(len(lens), dtype=np.uint8)
        st = L.FilterStats()
        L.check(L.lib().t1k_filter_batch(self.h, L.ptr(bases), L.ptr(off), L.ptr(lens), len(lens), L.ptr(good), C.byref(st)))
        if with_stats:
            return good, {k: getattr(st, k) for k, _ in st._fields_}
        return good

    def keep_pairs(self, reads1, reads2=None):
        g = self.IsGoodCandidate(reads1)
        if reads2 is not None:
            g = g | self.IsGoodCandidate(reads2)
        return g
