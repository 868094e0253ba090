"""ctypes binding of oracle/liboracle.so (TEST INFRASTRUCTURE: the CPU checker, never the product)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
REF_HARNESS = os.path.join(ORACLE_DIR, "_ref", "ref_harness")

OVERLAP_DT = np.dtype([(n, "<i4") for n in ("seqIdx", "readStart", "readEnd", "seqStart", "seqEnd", "strand",
                                             "matchCnt", "relaxedMatchCnt", "leftClip", "rightClip")])
ASSIGN_DT = np.dtype([("alleleIdx", "<i4"), ("start", "<i4"), ("end", "<i4"),
                      ("weight", "<f4"), ("qual", "<f4"), ("adjustWeight", "<f4")])

_lib = None


def lib():
    global _lib
    if _lib is None:
        so = os.path.join(ORACLE_DIR, "liboracle.so")
        src = os.path.join(ORACLE_DIR, "t1k_oracle.cpp")
        if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
            subprocess.check_call(["make", "-C", ORACLE_DIR, "liboracle.so"])
        L = C.CDLL(so)
        L.t1ko_create.restype = C.c_void_p
        L.t1ko_create.argtypes = [C.c_int32, C.c_char_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.c_double, C.c_int32]
        L.t1ko_destroy.argtypes = [C.c_void_p]
        L.t1ko_global_alignment.restype = C.c_int32
        L.t1ko_global_alignment.argtypes = [C.c_char_p, C.c_int32, C.c_char_p, C.c_int32, C.c_void_p, C.c_void_p]
        L.t1ko_assign_read.restype = C.c_int32
        L.t1ko_assign_read.argtypes = [C.c_void_p, C.c_char_p, C.c_int32, C.c_void_p, C.c_int32]
        L.t1ko_coverage.argtypes = [C.c_void_p, C.c_int32, C.c_void_p]
        L.t1ko_coverage_reset.argtypes = [C.c_void_p]
        L.t1ko_allele_len.restype = C.c_int32
        L.t1ko_allele_len.argtypes = [C.c_void_p, C.c_int32]
        L.t1ko_effective_len.restype = C.c_int32
        L.t1ko_effective_len.argtypes = [C.c_void_p, C.c_int32]
        L.t1ko_missing_coverage.restype = C.c_int32
        L.t1ko_missing_coverage.argtypes = [C.c_void_p, C.c_int32]
        L.t1ko_fragment_assign.restype = C.c_int32
        L.t1ko_fragment_assign.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int32,
                                           C.c_int32, C.c_void_p, C.c_int32]
        L.t1ko_em.restype = C.c_int32
        L.t1ko_em.argtypes = [C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                              C.c_double, C.c_double, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                              C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
        L.t1ko_infer_kmer_length.restype = C.c_int32
        L.t1ko_infer_kmer_length.argtypes = [C.c_int64]
        L.t1ko_filter_create.restype = C.c_void_p
        L.t1ko_filter_create.argtypes = [C.c_int32, C.c_char_p, C.c_void_p, C.c_int32, C.c_int32, C.c_double]
        for f in (L.t1ko_has_hit_in_set, L.t1ko_is_good_candidate):
            f.restype = C.c_int32
            f.argtypes = [C.c_void_p, C.c_char_p]
        L.t1ko_is_low_complexity.restype = C.c_int32
        L.t1ko_is_low_complexity.argtypes = [C.c_char_p]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def parse_exons(comment: str, length: int):
    """SeqSet::InputRefSeq comment parsing (SeqSet.hpp:933-976): first number = exon count, then pairs."""
    nums = []
    n = 0
    for ch in comment:
        if ch.isdigit():
            n = n * 10 + ord(ch) - 48
        else:
            nums.append(n)
            n = 0
    if n:
        nums.append(n)
    if not comment or not nums:
        return [(0, length - 1)]
    return [(nums[i], nums[i + 1]) for i in range(1, len(nums) - 1, 2)]


def collapse_reference(records):
    """Genotyper::InitRefSet (Genotyper.hpp:707-730): identical sequences collapse, weight++.
    Returns (kept records, weights)."""
    seen = {}
    kept = []
    w = []
    for name, comment, seq in records:
        if seq in seen:
            w[seen[seq]] += 1
        else:
            seen[seq] = len(kept)
            kept.append((name, comment, seq))
            w.append(1)
    return kept, w


def pack_reference(records):
    bases = b"".join(r[2] for r in records)
    off = np.zeros(len(records) + 1, dtype=np.int64)
    np.cumsum([len(r[2]) for r in records], out=off[1:])
    ptr = [0]
    se = []
    for name, comment, seq in records:
        ex = parse_exons(comment, len(seq))
        for s, e in ex:
            se.extend((s, e))
        ptr.append(len(se) // 2)
    return bases, off, np.asarray(ptr, dtype=np.int32), np.asarray(se if se else [0, 0], dtype=np.int32)


class Oracle:
    def __init__(self, records, similarity=0.8, relax=False, weights=None):
        self.records = records
        bases, off, ptr, se = pack_reference(records)
        self._keep = (bases, off, ptr, se)
        w = np.asarray(weights if weights is not None else [1] * len(records), dtype=np.int32)
        self._w = w
        self.h = lib().t1ko_create(len(records), bases, _p(off), _p(ptr), _p(se), _p(w), similarity, int(relax))
        self.n = len(records)

    def __del__(self):
        if getattr(self, "h", None):
            lib().t1ko_destroy(self.h)
            self.h = None

    def assign(self, read: bytes, weight=1, cap=1 << 16):
        buf = np.zeros(cap, dtype=OVERLAP_DT)
        ret = lib().t1ko_assign_read(self.h, read, weight, _p(buf), cap)
        return ret, buf[:max(ret, 0)].copy()

    def coverage(self, allele):
        n = lib().t1ko_allele_len(self.h, allele)
        out = np.zeros(n, dtype=np.int32)
        lib().t1ko_coverage(self.h, allele, _p(out))
        return out

    def coverage_reset(self):
        lib().t1ko_coverage_reset(self.h)

    def missing_coverage(self, allele):
        return lib().t1ko_missing_coverage(self.h, allele)

    def effective_len(self, allele):
        return lib().t1ko_effective_len(self.h, allele)

    def fragment_assign(self, o1, o2, hasN=False, max_assign=2000, cap=1 << 16):
        out = np.zeros(cap, dtype=ASSIGN_DT)
        o1 = np.ascontiguousarray(o1, dtype=OVERLAP_DT)
        if o2 is not None:
            o2 = np.ascontiguousarray(o2, dtype=OVERLAP_DT)
        n = lib().t1ko_fragment_assign(self.h, _p(o1), len(o1), _p(o2) if o2 is not None else None,
                                       len(o2) if o2 is not None else 0, int(hasN), max_assign, _p(out), cap)
        return out[:n].copy()


def global_alignment(t: bytes, p: bytes):
    ops = np.zeros(len(t) + len(p) + 4, dtype=np.int8)
    n = C.c_int32(0)
    s = lib().t1ko_global_alignment(t, len(t), p, len(p), _p(ops), C.byref(n))
    return s, ops[:n.value].copy()


def em(rowptr, col, count, eclen, x0, min_alpha=0.0, filter_frac=0.15, ec_allele_ptr=None, ec_alleles=None,
       allele_major=None, allele_gene=None):
    G = len(rowptr) - 1
    E = len(eclen)
    rowptr = np.ascontiguousarray(rowptr, dtype=np.int64)
    col = np.ascontiguousarray(col, dtype=np.int32)
    count = np.ascontiguousarray(count, dtype=np.float64)
    eclen = np.ascontiguousarray(eclen, dtype=np.int32)
    x0 = np.ascontiguousarray(x0, dtype=np.float64)
    x = np.zeros(E)
    rc = np.zeros(E)
    if allele_major is not None:
        am = np.ascontiguousarray(allele_major, dtype=np.int32)
        ag = np.ascontiguousarray(allele_gene, dtype=np.int32)
        ep = np.ascontiguousarray(ec_allele_ptr, dtype=np.int32)
        ea = np.ascontiguousarray(ec_alleles, dtype=np.int32)
        it = lib().t1ko_em(G, E, _p(rowptr), _p(col), _p(count), _p(eclen), _p(x0), min_alpha, filter_frac,
                           len(am), _p(ep), _p(ea), _p(am), _p(ag), int(am.max()) + 1, int(ag.max()) + 1, _p(x), _p(rc))
    else:
        it = lib().t1ko_em(G, E, _p(rowptr), _p(col), _p(count), _p(eclen), _p(x0), min_alpha, filter_frac,
                           0, None, None, None, None, 0, 0, _p(x), _p(rc))
    return it, x, rc


def parse_harness(path):
    """Parse ref_harness assign/genotype output into python structures."""
    out = {"uniq": [], "frag": [], "groups": [], "ecs": [], "q": [], "cov": {}}
    cur = None
    with open(path) as f:
        for line in f:
            t = line.split()
            if not t:
                continue
            k = t[0]
            if k == "A":
                out["nAlleles"] = int(t[1])
            elif k == "U":
                cur = {"seq": t[2].encode(), "weight": int(t[3]), "ov": []}
                out["uniq"].append(cur)
            elif k == "R":
                cur = {"ret": int(t[2]), "ov": []}
                out["uniq"].append(cur)
            elif k == "F":
                cur = {"fa": [], "as": []}
                out["frag"].append(cur)
            elif k == "f":
                cur["fa"].append((int(t[1]), int(t[2]), int(t[3]), int(t[4]), int(t[5]), float.fromhex(t[6]), int(t[7])))
            elif k == "a":
                cur["as"].append((int(t[1]), int(t[2]), int(t[3]), float.fromhex(t[4]), float.fromhex(t[5]),
                                  float.fromhex(t[6])))
            elif k == "G":
                out["nGroups"] = int(t[1])
                out["aligned"] = int(t[2])
            elif k == "g":
                out["groups"].append([(int(x.split(":")[0]), float.fromhex(x.split(":")[1])) for x in t[3:]])
            elif k == "E":
                out["nEc"] = int(t[1])
            elif k == "e":
                out["ecs"].append([int(x) for x in t[3:]])
            elif k == "M":
                out["missing"] = [int(x) for x in t[1:]]
            elif k == "Q":
                out["iters"] = int(t[1])
            elif k == "q":
                out["q"].append((int(t[2]), float.fromhex(t[3]), float.fromhex(t[4]), int(t[5]), int(t[6])))
            elif k == "C":
                out["cov"][int(t[1])] = np.asarray(t[3:], dtype=np.int32)
            elif k == "K":
                out["kept"] = [int(x) for x in t[1:]]
            else:
                cur["ov"].append(tuple(int(x) for x in t[:10]) + (float.fromhex(t[10]),))
    return out


# ---------------------------------------------------------------------------------------------
# Host-side model steps of the reference, restated for the checker (small inputs only).

def parse_allele_name(allele: str, digit_units=-1, delimiter=""):
    """Genotyper::ParseAlleleName, Genotyper.hpp:63-131 (fieldsType 0).  Returns (gene, majorAllele)."""
    parse_type = 1
    fields = digit_units
    delim = ""
    if fields == -1:
        fields = 3
        if ":" in allele:
            delim = ":"
            parse_type = 2
    if delimiter:
        delim = delimiter
        parse_type = 2
    i = allele.find("*")
    if i < 0:
        i = len(allele)
    gene = allele[:i]
    if parse_type == 1:
        j = 0
        while j <= fields and i + j < len(allele):
            j += 1
        return gene, allele[:i + j]
    k = 0
    j = i
    while j < len(allele):
        if allele[j] == delim:
            k += 1
            if k >= fields:
                break
        j += 1
    return gene, allele[:j]


def allele_info(names, eff_len, digit_units=-1, delimiter=""):
    """Genotyper::InitAlleleInfo, Genotyper.hpp:559-682: gene / major-allele ids in first-appearance order and
    the large-deletion effective-length adjustment (:641-681).  Returns (gene_idx, major_idx, eff_len')."""
    genes, majors = {}, {}
    gi, mi = [], []
    for n in names:
        g, m = parse_allele_name(n, digit_units, delimiter)
        gi.append(genes.setdefault(g, len(genes)))
        mi.append(majors.setdefault(m, len(majors)))
    eff = list(eff_len)
    for g in range(len(genes)):
        ids = [i for i in range(len(names)) if gi[i] == g]
        lens = sorted(eff_len[i] for i in ids)
        mode, best, j = 0, 0, 0
        while j < len(lens):
            k = j
            while k < len(lens) and lens[k] == lens[j]:
                k += 1
            if k - j > best:
                best, mode = k - j, lens[j]
            j = k
        for i in ids:
            if eff_len[i] < mode - 500:
                eff[i] = mode
    return np.asarray(gi, dtype=np.int32), np.asarray(mi, dtype=np.int32), np.asarray(eff, dtype=np.int32)


def coalesce(frag_assignments):
    """Genotyper::CoalesceReadAssignments (Genotyper.hpp:841-908) over all fragments in order.
    frag_assignments: list of ASSIGN_DT arrays.  Returns (groups: list of ASSIGN_DT arrays, assigned count)."""
    groups, index = [], {}
    assigned = 0
    for a in frag_assignments:
        if len(a) == 0:
            continue
        assigned += 1
        a = a[np.argsort(a["alleleIdx"], kind="stable")].copy()
        key = (a["alleleIdx"].tobytes(), a["qual"].tobytes())
        g = index.get(key)
        if g is None:
            index[key] = len(groups)
            groups.append(a)
        else:
            t = groups[g]
            q1 = a["qual"] == 1
            lower = q1 & (a["start"] < t["start"])
            t["start"][lower] = a["start"][lower]
            lowere = q1 & (a["end"] < t["end"])          # Q10: end is overwritten with start
            t["end"][lowere] = a["start"][lowere]
            t["weight"] += a["weight"]                     # float32 adds, fragment order
            t["adjustWeight"] += a["adjustWeight"]
    return groups, assigned


def build_ecs(groups, n_alleles):
    """Genotyper::FinalizeReadAssignments + BuildAlleleEquivalentClass (Genotyper.hpp:912-939,1072-1139).
    Returns (ecs: list of allele lists, allele_ec[n_alleles])."""
    G = len(groups)
    reads_in = [[] for _ in range(n_alleles)]
    for g, a in enumerate(groups):
        for al in a["alleleIdx"]:
            reads_in[int(al)].append(g)
    fp = []
    for i in range(n_alleles):
        b = -1
        if reads_in[i]:
            b = 0
            for g in reads_in[i]:
                b = ((((b & 0xFFFFFFFF) * (G & 0xFFFFFFFF)) & 0xFFFFFFFF) + g) & 0xFFFFFFFF
                b %= 1000003
        fp.append((i, b))
    fp.sort(key=lambda p: (-p[1], p[0]))
    ecs = []
    allele_ec = np.full(n_alleles, -1, dtype=np.int32)
    for i, (a, b) in enumerate(fp):
        if b == -1:
            break
        found = -1
        j = i - 1
        while j >= 0 and fp[j][1] == b:
            if reads_in[fp[j][0]] == reads_in[a]:
                found = fp[j][0]
                break
            j -= 1
        if found < 0:
            allele_ec[a] = len(ecs)
            ecs.append([a])
        else:
            allele_ec[a] = allele_ec[found]
            ecs[allele_ec[found]].append(a)
    return ecs, allele_ec


def em_problem(groups, ecs, allele_ec, eff_len, seq_weight):
    """EM inputs as Genotyper::QuantifyAlleleEquivalentClass builds them (Genotyper.hpp:1155-1232)."""
    rowptr = [0]
    col = []
    count = []
    for a in groups:
        count.append(float(a["weight"].max()))
        seen = []
        for al in a["alleleIdx"]:
            e = int(allele_ec[int(al)])
            if e not in seen:
                seen.append(e)
        col.extend(seen)
        rowptr.append(len(col))
    eclen = [min(int(eff_len[a]) for a in ec) for ec in ecs]
    x0 = [float(sum(int(seq_weight[a]) for a in ec)) for ec in ecs]
    ptr = [0]
    flat = []
    for ec in ecs:
        flat.extend(ec)
        ptr.append(len(flat))
    return dict(rowptr=np.asarray(rowptr, dtype=np.int64), col=np.asarray(col, dtype=np.int32),
                count=np.asarray(count, dtype=np.float64), eclen=np.asarray(eclen, dtype=np.int32),
                x0=np.asarray(x0, dtype=np.float64), ec_allele_ptr=np.asarray(ptr, dtype=np.int32),
                ec_alleles=np.asarray(flat, dtype=np.int32))


def set_allele_abundance(rc, eclen, ecs, n_alleles):
    """Genotyper::SetAlleleAbundance (Genotyper.hpp:957-987): per allele abundance / ecAbundance."""
    ab = np.zeros(n_alleles)
    ecab = np.zeros(n_alleles)
    for e, ec in enumerate(ecs):
        a = rc[e] / eclen[e] * 1000.0
        for k in ec:
            ab[k] = a / len(ec)
            ecab[k] = a
    return ab, ecab


def genotype_pipeline(orc: "Oracle", reads1, reads2, names, seq_weight, max_assign=2000, min_alpha=0.0,
                      filter_frac=0.15, digit_units=-1, delimiter=""):
    """The Genotyper.cpp:450-646 flow on the oracle: dedup read-ends, AssignRead per unique sequence,
    fragment pairing, coalescing, equivalence classes, EM.  reads: uint8 arrays [n, L] or lists of bytes."""
    def as_list(r):
        return [bytes(x) if isinstance(x, (bytes, bytearray)) else x.tobytes() for x in r]
    s1 = as_list(reads1)
    s2 = as_list(reads2) if reads2 is not None else None
    all_seq = s1 + (s2 if s2 is not None else [])
    uniq = {}
    for s in all_seq:
        uniq[s] = uniq.get(s, 0) + 1
    orc.coverage_reset()
    ov = {}
    for s in sorted(uniq):
        ret, o = orc.assign(s, uniq[s])
        ov[s] = o
    frags = []
    for i in range(len(s1)):
        has_n = b"N" in s1[i] or (s2 is not None and b"N" in s2[i])
        frags.append(orc.fragment_assign(ov[s1[i]], ov[s2[i]] if s2 is not None else None, has_n, max_assign))
    groups, assigned = coalesce(frags)
    n = orc.n
    ecs, allele_ec = build_ecs(groups, n)
    eff = [orc.effective_len(a) for a in range(n)]
    gi, mi, eff = allele_info(names, eff, digit_units, delimiter)
    missing = np.asarray([orc.missing_coverage(a) for a in range(n)], dtype=np.int32)
    res = dict(uniq=ov, frags=frags, groups=groups, assigned=assigned, ecs=ecs, allele_ec=allele_ec, missing=missing,
               eff_len=eff, gene=gi, major=mi)
    if ecs:
        P = em_problem(groups, ecs, allele_ec, eff, seq_weight)
        it, x, rc = em(P["rowptr"], P["col"], P["count"], P["eclen"], P["x0"], min_alpha, filter_frac,
                       P["ec_allele_ptr"], P["ec_alleles"], mi, gi)
        ab, ecab = set_allele_abundance(rc, P["eclen"], ecs, n)
        res.update(problem=P, iters=it, x=x, rc=rc, abundance=ab, ec_abundance=ecab)
    else:
        res.update(problem=None, iters=0, abundance=np.zeros(n), ec_abundance=np.zeros(n))
    return res


def seq_weights(records, weights):
    """SeqSet::UpdateDnaSeqWeight (SeqSet.hpp:1008-1029): when any allele has non-adjacent exons (rnaData false,
    SeqSet.hpp:705-713) every allele's weight becomes the summed weight of all alleles with its exon sequence."""
    exons = [parse_exons(c, len(s)) for _, c, s in records]
    dna = any(ex[i][0] > ex[i - 1][1] + 1 for ex in exons for i in range(1, len(ex)))
    if not dna:
        return list(weights)
    keys = []
    for (_, _, s), ex in zip(records, exons):
        mask = np.zeros(len(s), dtype=bool)
        for a, b in ex:
            mask[a:min(b, len(s) - 1) + 1] = True
        keys.append(np.frombuffer(s, dtype=np.uint8)[mask].tobytes())
    tot = {}
    for k, w in zip(keys, weights):
        tot[k] = tot.get(k, 0) + w
    return [tot[k] for k in keys]


class CandidateFilter:
    """The candidate filter of fastq-extractor as the oracle restates it: FastqExtractor.cpp main()'s
    set-up (k = max(9, InferKmerLength), hitLenRequired = max(27 | 23, mean length of the first 1000 reads // 5, k)),
    IsGoodCandidate per read, a pair is kept if either mate is good (FastqExtractor.cpp:199-212)."""

    def __init__(self, records, reads1, paired, similarity=0.8):
        seqs = [r[2] for r in records]
        bases = b"".join(seqs)
        off = np.zeros(len(seqs) + 1, dtype=np.int64)
        np.cumsum([len(x) for x in seqs], out=off[1:])
        first = reads1[:1000]
        hit_len = 27 if paired else 23
        mean5 = sum(len(r) for r in first) // (len(first) * 5)
        if mean5 > hit_len:
            hit_len = mean5
        k = 9
        inferred = lib().t1ko_infer_kmer_length(int(off[-1]))
        if inferred > k:
            k = inferred
            if k > hit_len:
                hit_len = k
        self.k, self.hit_len = k, hit_len
        self.h = lib().t1ko_filter_create(len(seqs), bases, _p(off), k, hit_len, similarity)
        assert self.h

    def good(self, read):
        return bool(lib().t1ko_is_good_candidate(self.h, read))

    def keep_pair(self, r1, r2=None):
        return self.good(r1) or (r2 is not None and self.good(r2))

    def __del__(self):
        if getattr(self, "h", None):
            lib().t1ko_destroy(self.h)
            self.h = None
