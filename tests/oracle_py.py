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
            else:
                cur["ov"].append(tuple(int(x) for x in t[:10]) + (float.fromhex(t[10]),))
    return out
