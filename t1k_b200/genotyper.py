"""Host-side mirror of the reference's interface for the hot path, over the C ABI.

`SeqSet` and `Genotyper` keep the reference's method names and argument meaning
(/root/reference/SeqSet.hpp, Genotyper.hpp) so that parity tests read like calls into the reference:

    SeqSet.AssignRead             SeqSet.hpp:2119       (batched: one call per batch of unique read-ends)
    SeqSet.ReadAssignmentToFragmentAssignment + Genotyper.SetReadAssignments   SeqSet.hpp:2310, Genotyper.hpp:778
    SeqSet.GetSeqMissingBaseCoverage            SeqSet.hpp:2717
    Genotyper.QuantifyAlleleEquivalentClass     Genotyper.hpp:1142
    Genotyper.Genotype            the Genotyper.cpp:450-646 flow in one call

All compute happens in the CUDA library; nothing here falls back to the CPU.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L
from .refset import RefSet


def _reads_to_batch(reads):
    """list of bytes / uint8 [n, L] array -> (bases, off u64, len u32)."""
    if isinstance(reads, np.ndarray) and reads.ndim == 2:
        n, ln = reads.shape
        bases = np.ascontiguousarray(reads).reshape(-1)
        off = (np.arange(n, dtype=np.uint64) * np.uint64(ln))
        lens = np.full(n, ln, dtype=np.uint32)
        # trailing NULs (padding of ragged sets) are not bases
        if n and (bases == 0).any():
            lens = (reads != 0).sum(axis=1).astype(np.uint32)
        return bases, off, lens
    lens = np.asarray([len(r) for r in reads], dtype=np.uint32)
    off = np.zeros(len(reads), dtype=np.uint64)
    if len(reads):
        off[1:] = np.cumsum(lens[:-1], dtype=np.uint64)
    bases = np.frombuffer(b"".join(reads) + b"\0", dtype=np.uint8)
    return bases, off, lens


class Assignment:
    """Per-read-end overlap lists resident on the device (result of SeqSet.AssignRead)."""

    def __init__(self, handle, n, owner=None):
        self.h = handle
        self.n = n
        self._owner = owner       # the SeqSet whose reference the lists point into stays alive as long as they do

    def __del__(self):
        if getattr(self, "h", None):
            L.lib().t1k_assignment_destroy(self.h)
            self.h = None

    def fetch(self):
        """-> (row_ptr[n+1], ret[n], records) in the reference's output order."""
        row = np.zeros(self.n + 1, dtype=np.uint64)
        ret = np.zeros(self.n, dtype=np.int32)
        tot = C.c_uint64(0)
        L.check(L.lib().t1k_assignment_fetch(self.h, L.ptr(row), L.ptr(ret), None, C.byref(tot)))
        rec = np.zeros(tot.value, dtype=L.OVERLAP_DT)
        if tot.value:
            L.check(L.lib().t1k_assignment_fetch(self.h, L.ptr(row), L.ptr(ret), L.ptr(rec), C.byref(tot)))
        return row, ret, rec

    def stats(self):
        s = L.AssignStats()
        L.check(L.lib().t1k_assignment_stats(self.h, C.byref(s)))
        return {k: getattr(s, k) for k, _ in s._fields_}


class AssignJob:
    """A batch in flight (t1k_assign_batch_async); the input arrays are kept alive until wait()."""

    def __init__(self, handle, n, owner, keep):
        self.h, self.n, self._owner, self._keep = handle, n, owner, keep

    def wait(self) -> Assignment:
        h = C.c_void_p()
        job, self.h = self.h, None
        try:
            L.check(L.lib().t1k_assign_wait(job, C.byref(h)))
        finally:
            self._keep = None
        return Assignment(h, self.n, self._owner)

    def __del__(self):
        if getattr(self, "h", None):       # never waited for: join and drop the result
            h = C.c_void_p()
            L.lib().t1k_assign_wait(self.h, C.byref(h))
            if h:
                L.lib().t1k_assignment_destroy(h)
            self.h = None


class SeqSet:
    """Allele reference + k-mer index + base coverage on one GPU (SeqSet.hpp)."""

    def __init__(self, refset: RefSet, similarity=0.8, relax_intron=False, device=-1):
        self.ref = refset
        bases, off, ptr, se = refset.packed()
        self._keep = (bases, off, ptr, se)
        d = L.RefDesc(refset.n, bases, L.ptr(off), L.ptr(ptr), L.ptr(se), float(similarity), int(bool(relax_intron)), device)
        h = C.c_void_p()
        L.check(L.lib().t1k_ref_create(C.byref(d), C.byref(h)))
        self.h = h
        self.offset = off
        self.similarity = similarity
        self.relax_intron = bool(relax_intron)

    def __del__(self):
        if getattr(self, "h", None):
            L.lib().t1k_ref_destroy(self.h)
            self.h = None

    def Size(self):
        return L.lib().t1k_ref_n_alleles(self.h)

    def AssignRead(self, reads, weights=None) -> Assignment:
        """SeqSet::AssignRead for a batch of unique read-ends; weights[i] = number of duplicates (0: analyzer mode)."""
        bases, off, lens = _reads_to_batch(reads)
        n = len(lens)
        w = np.ones(n, dtype=np.int32) if weights is None else np.ascontiguousarray(weights, dtype=np.int32)
        h = C.c_void_p()
        L.check(L.lib().t1k_assign_batch(self.h, L.ptr(bases), L.ptr(off), L.ptr(lens), L.ptr(w), n, C.byref(h)))
        return Assignment(h, n, self)

    def AssignReadAsync(self, reads, weights=None):
        """t1k_assign_batch_async: returns a job at once; job.wait() -> Assignment.  Jobs of one SeqSet run in submission order."""
        bases, off, lens = _reads_to_batch(reads)
        n = len(lens)
        w = np.ones(n, dtype=np.int32) if weights is None else np.ascontiguousarray(weights, dtype=np.int32)
        j = C.c_void_p()
        L.check(L.lib().t1k_assign_batch_async(self.h, L.ptr(bases), L.ptr(off), L.ptr(lens), L.ptr(w), n, C.byref(j)))
        return AssignJob(j, n, self, (bases, off, lens, w))

    def ReadAssignmentToFragmentAssignment(self, assignment: Assignment, end1, end2=None, has_n=None, max_assign=2000,
                                           with_assigned=False):
        """Fragment pairing + Genotyper::SetReadAssignments -> (row_ptr[n_frag+1], entries) in the reference's order
        (+ the `fragmentAssigned` flags of Genotyper.cpp:564 when with_assigned)."""
        e1 = np.ascontiguousarray(end1, dtype=np.uint32)
        e2 = None if end2 is None else np.ascontiguousarray(end2, dtype=np.uint32)
        hn = None if has_n is None else np.ascontiguousarray(has_n, dtype=np.uint8)
        rp, en = C.c_void_p(), C.c_void_p()
        fa = np.zeros(len(e1), dtype=np.uint8)
        L.check(L.lib().t1k_pair_batch(self.h, assignment.h, L.ptr(e1), L.ptr(e2), L.ptr(hn), len(e1), int(max_assign),
                                       C.byref(rp), C.byref(en), L.ptr(fa)))
        try:
            row = np.ctypeslib.as_array(C.cast(rp, C.POINTER(C.c_uint64)), shape=(len(e1) + 1,)).copy()
            tot = int(row[-1])
            ent = np.zeros(tot, dtype=L.ASSIGN_DT)
            if tot:
                C.memmove(ent.ctypes.data, en, tot * L.ASSIGN_DT.itemsize)
        finally:
            L.lib().t1k_free(rp)
            L.lib().t1k_free(en)
        return (row, ent, fa) if with_assigned else (row, ent)

    def AddOverlapAlignmentInfo(self, reads, read_idx, overlaps, force_dp=False, with_stats=False, raw=False):
        """SeqSet::AddOverlapAlignmentInfo (SeqSet.hpp:2657-2680) for a batch of (read, overlap) items: overlaps are records of
        Assignment.fetch() (OVERLAP_DT), read_idx[i] the read each belongs to -> list of int8 edit strings (0 M, 1 X, 2 I, 3 D;
        None where seqIdx == -1).  force_dp: the band DP for every item (A/B switch).  raw: (blob, align_ptr) instead of the list."""
        bases, off, lens = _reads_to_batch(reads)
        ov = np.ascontiguousarray(overlaps, dtype=L.OVERLAP_DT)
        idx = np.ascontiguousarray(read_idx, dtype=np.uint32)
        n = len(ov)
        aptr = np.zeros(n, dtype=np.uint64)
        blob, nbytes, st = C.c_void_p(), C.c_uint64(0), L.AlignInfoStats()
        L.check(L.lib().t1k_align_info_batch(self.h, L.ptr(bases), L.ptr(off), L.ptr(lens), len(lens), L.ptr(idx), L.ptr(ov), n,
                                             1 if force_dp else 0, L.ptr(aptr), C.byref(blob), C.byref(nbytes), C.byref(st)))
        try:
            buf = np.ctypeslib.as_array(C.cast(blob, C.POINTER(C.c_int8)), shape=(max(1, nbytes.value),)).copy()
        finally:
            L.lib().t1k_free(blob)
        if raw:
            out = (buf, aptr)
            return (out, {k: getattr(st, k) for k, _ in st._fields_}) if with_stats else out
        out = []
        for i in range(n):
            if aptr[i] == np.uint64(0xFFFFFFFFFFFFFFFF):
                out.append(None)
                continue
            a = int(aptr[i])
            cap = int(ov["seqEnd"][i] - ov["seqStart"][i] + 1 + ov["readEnd"][i] - ov["readStart"][i] + 1) + 2
            s = buf[a:a + cap]
            out.append(s[:int(np.argmax(s == -1))])
        if with_stats:
            return out, {k: getattr(st, k) for k, _ in st._fields_}
        return out

    def GetBaseCoverage(self):
        """posWeight[].count[consensus base] of every allele, concatenated (Q11)."""
        out = np.zeros(int(self.offset[-1]), dtype=np.int32)
        L.check(L.lib().t1k_coverage_fetch(self.h, L.ptr(out)))
        return out

    def ResetBaseCoverage(self):
        L.check(L.lib().t1k_coverage_reset(self.h))

    def GetSeqMissingBaseCoverage(self):
        out = np.zeros(self.ref.n, dtype=np.int32)
        L.check(L.lib().t1k_missing_coverage(self.h, L.ptr(out)))
        return out


def QuantifyAlleleEquivalentClass(row_ptr, col, count, ec_len, x0, min_squarem_alpha=0.0, filter_frac=0.15,
                                  ec_allele_ptr=None, ec_alleles=None, allele_major=None, allele_gene=None, device=-1,
                                  fast_sums=False, comm=None):
    """The EM loop of Genotyper::QuantifyAlleleEquivalentClass on the device -> (iterations, x, ecReadCount).
    comm: a dist_em.Comm — every rank passes the same problem and runs the E-step over its own row range."""
    row_ptr = np.ascontiguousarray(row_ptr, dtype=np.int64)
    col = np.ascontiguousarray(col, dtype=np.int32)
    count = np.ascontiguousarray(count, dtype=np.float64)
    ec_len = np.ascontiguousarray(ec_len, dtype=np.int32)
    x0 = np.ascontiguousarray(x0, dtype=np.float64)
    p = L.EmProblem()
    p.n_groups, p.n_ec = len(row_ptr) - 1, len(ec_len)
    p.row_ptr, p.col, p.count, p.ec_len, p.x0 = L.ptr(row_ptr), L.ptr(col), L.ptr(count), L.ptr(ec_len), L.ptr(x0)
    p.min_squarem_alpha, p.filter_frac = float(min_squarem_alpha), float(filter_frac)
    p.fast_sums = int(bool(fast_sums))
    p.comm = comm.h if comm is not None else None
    keep = []
    if allele_major is not None:
        am = np.ascontiguousarray(allele_major, dtype=np.int32)
        ag = np.ascontiguousarray(allele_gene, dtype=np.int32)
        ep = np.ascontiguousarray(ec_allele_ptr, dtype=np.int32)
        ea = np.ascontiguousarray(ec_alleles, dtype=np.int32)
        keep = [am, ag, ep, ea]
        p.n_alleles, p.n_major, p.n_gene = len(am), int(am.max()) + 1, int(ag.max()) + 1
        p.ec_allele_ptr, p.ec_alleles, p.allele_major, p.allele_gene = L.ptr(ep), L.ptr(ea), L.ptr(am), L.ptr(ag)
    x = np.zeros(p.n_ec)
    rc = np.zeros(p.n_ec)
    r = L.EmResult(L.ptr(x), L.ptr(rc), 0, 0.0, 0)
    L.check(L.lib().t1k_em_run(C.byref(p), C.byref(r), device))
    del keep
    return r.iterations, x, rc, {"ms_kernel": r.ms_kernel, "n_launches": r.n_launches}


class Genotyper:
    """Genotyper.cpp:450-646 in one call: de-duplicate read-ends, align, pair, coalesce, equivalence classes, EM."""

    def __init__(self, refset: RefSet, similarity=0.8, relax_intron=False, max_assign=2000, filter_frac=0.15,
                 min_squarem_alpha=0.0, device=-1, em_fast_sums=False):
        self.ref = refset
        self.em_fast_sums = em_fast_sums
        self.refSet = SeqSet(refset, similarity, relax_intron, device)
        self.max_assign = max_assign
        self.filter_frac = filter_frac
        self.min_squarem_alpha = min_squarem_alpha

    def Genotype(self, reads1, reads2=None, comm=None):
        """reads: uint8 arrays [n, stride] (NUL padded for shorter reads).  Returns a dict of per-allele results.
        comm: a dist_em.Comm — reads are THIS rank's shard; per-allele results are the whole job's on every rank."""
        r1 = np.ascontiguousarray(reads1, dtype=np.uint8)
        r2 = None if reads2 is None else np.ascontiguousarray(reads2, dtype=np.uint8)
        n, stride = r1.shape
        ref = self.ref
        prm = L.GenotypeParams(self.max_assign, self.min_squarem_alpha, self.filter_frac, L.ptr(ref.seq_weight),
                               L.ptr(ref.effective_len), L.ptr(ref.allele_major), L.ptr(ref.allele_gene),
                               len(ref.major_names), len(ref.gene_names), int(bool(self.em_fast_sums)),
                               comm.h if comm is not None else None, None)
        out = dict(abundance=np.zeros(ref.n), ec_abundance=np.zeros(ref.n),
                   equivalent_class=np.zeros(ref.n, dtype=np.int32), missing_coverage=np.zeros(ref.n, dtype=np.int32),
                   fragment_assigned=np.zeros(n, dtype=np.uint8),
                   # Genotyper::RemoveLowLikelihoodAlleleInEquivalentClass (Genotyper.cpp:647): members that stay in their class
                   allele_kept=np.zeros(ref.n, dtype=np.uint8), allele_span=np.zeros(2 * ref.n, dtype=np.int32))
        res = L.GenotypeResult()
        res.abundance, res.ec_abundance = L.ptr(out["abundance"]), L.ptr(out["ec_abundance"])
        res.equivalent_class, res.missing_coverage = L.ptr(out["equivalent_class"]), L.ptr(out["missing_coverage"])
        res.fragment_assigned = L.ptr(out["fragment_assigned"])
        res.allele_kept, res.allele_span = L.ptr(out["allele_kept"]), L.ptr(out["allele_span"])
        L.check(L.lib().t1k_genotype(self.refSet.h, L.ptr(r1), L.ptr(r2), stride, n, C.byref(prm), C.byref(res)))
        for k, _ in res._fields_:
            v = getattr(res, k)
            if k not in out and not isinstance(v, C.c_void_p):
                out[k] = v
        return out
