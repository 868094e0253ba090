// Fragment pairing on the device: SeqSet::ReadAssignmentToFragmentAssignment (SeqSet.hpp:2310-2655) followed by
// Genotyper::SetReadAssignments (Genotyper.hpp:778-832) and Genotyper::ReadAssignmentWeight (Genotyper.hpp:205-230),
// one warp per fragment, reading the HBM-resident per-read-end record lists that the AssignRead kernels (k_passes) left behind.
//
// The record lists are stored in allele order (ties in candidate order); `key` + the record's position give the
// reference's list order (the order AssignRead returned them in).  Nothing is materialised per (fragment, allele):
// every pass re-derives the allele's best mate pair from the two sorted runs, so the only state is a handful of
// warp-reduced scalars.  Output rows are written in allele order, which is the order CoalesceReadAssignments
// sorts them into anyway (Genotyper.hpp:851); `posKey/posIdx` reproduce the reference's own order on request.
#pragma once
#include "t1k_kernels.cuh"

namespace t1k {

struct PairEntry {          // == T1KReadAssignment == struct _readAssignment (Genotyper.hpp:44-56)
  int32_t alleleIdx, start, end;
  float weight, qual, adjustWeight;
};

struct PairParams {
  RefView R;
  const Rec *store;
  const u64 *readOff;
  const u32 *readCnt;
  const u32 *readTop;       // per read-end: max matchCnt << 16 | (65535 - denominator) over its records (k_passes)
  const u32 *end1, *end2;   // end2 == NULL: single-end data
  const u8 *hasN;
  u32 fragBase, nFrag;      // fragments [fragBase, fragBase + nFrag) of the caller's arrays
  int maxAssign;
  // rows are appended to `out` through one atomic per fragment (dense, in no particular fragment order); if the
  // buffer is too small nothing is written and *outCtr tells the host the exact capacity for the re-run
  PairEntry *out;
  u64 outCap;
  unsigned long long *outCtr;
  u64 *rowOff;              // per fragment of this launch: first entry of its row in `out`
  u64 *ordKey;              // optional: list position of the allele's first candidate (reference order)
  u32 *ordIdx;
  u32 *rowCnt;              // per fragment of this launch; bit 31 = the fragment had assignments before the
                            // SetReadAssignments cuts (Genotyper.cpp:564 `fragmentAssigned`)
  u64 *rowHash;             // optional, 2 per fragment: order-free hash of the allele set
  unsigned int *workCtr;
  u32 *b0;                  // per warp, b0Stride entries: start of each first-list allele run in the second list
  u32 b0Stride;
  PairEntry *stage;         // per warp, stageCap entries: the fragment's row until the fragment-level cuts are decided
  u64 *stageKey; u32 *stageIdx;   // with ordKey
  u32 stageCap;
};

struct RV { int seqIdx, ss, se, rs, re, lc, rc, mc, st, relaxed; u64 key; };

__device__ __forceinline__ RV load_rv(const Rec *p) {
  const uint4 a = *reinterpret_cast<const uint4 *>(p);
  const uint4 b = *(reinterpret_cast<const uint4 *>(p) + 1);
  RV r;
  r.seqIdx = (int)a.x; r.ss = (int)a.y; r.se = (int)a.z;
  r.rs = a.w & 0xffff; r.re = a.w >> 16; r.lc = b.x & 0xffff; r.rc = b.x >> 16;
  r.mc = rec_mc(b.y); r.st = rec_strand01(b.y);
  r.relaxed = rec_relaxed(b.y);
  r.key = (u64)b.z | ((u64)b.w << 32);
  return r;
}
__device__ __forceinline__ int rv_denom(const RV &r) { return r.re - r.rs + 1 + r.se - r.ss + 1 + 2 * r.lc + 2 * r.rc; }

// struct _overlap::operator< (SeqSet.hpp:103-127).  similarity = matchCnt / denom, so for equal matchCnt a larger
// similarity is a smaller denominator.
__device__ __forceinline__ bool rv_less(const RV &a, const RV &b) {
  if (a.mc != b.mc) return a.mc > b.mc;
  const int da = rv_denom(a), db = rv_denom(b);
  if (da != db) return da < db;
  if (a.re - a.rs != b.re - b.rs) return a.re - a.rs > b.re - b.rs;
  if (a.seqIdx != b.seqIdx) return a.seqIdx < b.seqIdx;
  if (a.st != b.st) return a.st < b.st;
  if (a.rs != b.rs) return a.rs < b.rs;
  if (a.re != b.re) return a.re < b.re;
  if (a.ss != b.ss) return a.ss < b.ss;
  return a.se < b.se;
}

// IsSeparatorInRange (SeqSet.hpp:487-498) with the -1 / len sentinels of SeqSet.hpp:922-928
__device__ __forceinline__ bool sep_exact(const RefView &R, int seqIdx, int s, int e) {
  const int len = R.len[seqIdx];
  if (s <= -1 && e >= -1) return true;
  if (s <= len && e >= len) return true;
  if (s < 0) s = 0;
  if (e > len - 1) e = len - 1;
  if (s > e) return false;
  return n_in_range(R, R.wordOff[seqIdx], s, e);
}

__device__ __forceinline__ int lower_bound_allele(const Rec *L, int n, int seqIdx, int lo = 0, int hi = -1) {
  if (hi < 0) hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (L[mid].seqIdx < seqIdx) lo = mid + 1; else hi = mid;
  }
  return lo;
}
// The same lower bound, galloping outwards from a guess.  Both mates' lists are in allele order and mostly hold the same
// alleles, so the run of allele A[i] in B sits near i + (offset of the previous run): two or three probes of records the
// neighbouring lanes load anyway, instead of log2(n) dependent loads spread over the whole list.
__device__ __forceinline__ int lower_bound_allele_near(const Rec *L, int n, int seqIdx, int guess) {
  int g = guess < 0 ? 0 : guess > n ? n : guess;
  if (g >= n || L[g].seqIdx >= seqIdx) {        // answer <= g: walk down
    int hi = g, step = 1, lo = g - 1;
    while (lo >= 0 && L[lo].seqIdx >= seqIdx) { hi = lo; step <<= 1; lo = hi - step; }
    return lower_bound_allele(L, n, seqIdx, lo < 0 ? 0 : lo + 1, hi);
  }
  int lo = g + 1, step = 1, p = g + 1;          // L[g] < seqIdx: walk up
  while (p < n && L[p].seqIdx < seqIdx) { lo = p + 1; step <<= 1; p = lo - 1 + step; }
  return lower_bound_allele(L, n, seqIdx, lo, p < n ? p : n);
}

// best fragment of one allele (the per-seqIdx slot of SeqSet.hpp:2440-2455)
struct AlleleBest {
  bool valid;
  int mc, denom, relaxed, ss, se;
  int ia, jb;               // positions in the lists (jb = -1: no mate)
  int b0;                   // paired: lower_bound of the allele in the second list
  u64 posKey; int posIdx;   // list position of the allele's first candidate = its rank in `assign`
  int relaxBy;
  RV o1;
};

__device__ __forceinline__ bool pos_less(u64 k, int i, u64 k2, int i2) { return k < k2 || (k == k2 && i < i2); }
// list order of two records of ONE allele run of one list (rec_before, t1k_core.cuh): key, then — in a list the > 1000 cut
// re-sorted — the extended coordinates, else the store position.  `b` is only consulted on a key tie of a re-sorted list.
__device__ __forceinline__ bool run_pos_less(const RV &a, int ia, u64 kb, int ib, const Rec *L) {
  if (a.key != kb) return a.key < kb;
  if (!(a.key & 1) || ib < 0 || ib == 0x7fffffff) return ia < ib;
  const RV b = load_rv(L + ib);
  if (a.rs != b.rs) return a.rs < b.rs;
  if (a.re != b.re) return a.re < b.re;
  if (a.ss != b.ss) return a.ss < b.ss;
  if (a.se != b.se) return a.se < b.se;
  return ia < ib;
}

// mate compatibility (SeqSet.hpp:2366-2380)
__device__ __forceinline__ bool mates_ok(const RV &a, const RV &b) {
  if (a.st == b.st) return false;
  return a.st == 1 ? a.ss < b.ss : a.ss > b.ss;
}

// A[a0,a1) is the run of one allele in the first list; paired mode looks the allele up in B.
// b0Hint >= 0: the position found by an earlier pass (lower_bound of the allele in B); out.b0 returns it.
// b0Guess (used when b0Hint < 0): where to start looking for the allele in B.
__device__ void eval_allele(const RefView &R, bool paired, const Rec *A, int a0, int a1, const Rec *B, int nB, AlleleBest &out,
                            int b0Hint = -1, int b0Guess = 0) {
  out.valid = false;
  out.posKey = ~0ull; out.posIdx = 0x7fffffff;
  RV bo1, bo2;
  u64 bKeyA = 0, bKeyB = 0;
  if (!paired) {
    for (int ia = a0; ia < a1; ++ia) {
      const RV a = load_rv(A + ia);
      if (run_pos_less(a, ia, out.posKey, out.posIdx, A)) { out.posKey = a.key; out.posIdx = ia; }
      bool better;
      if (!out.valid) better = true;
      else if (rv_less(a, bo1)) better = true;
      else if (rv_less(bo1, a)) better = false;
      else better = run_pos_less(a, ia, bKeyA, out.ia, A);
      if (better) { out.valid = true; bo1 = a; bKeyA = a.key; out.ia = ia; out.jb = -1; }
    }
    if (out.valid) {
      out.mc = bo1.mc; out.denom = rv_denom(bo1); out.relaxed = bo1.relaxed; out.ss = bo1.ss; out.se = bo1.se;
      out.relaxBy = 2; out.o1 = bo1;
    }
    return;
  }
  const int seqIdx = A[a0].seqIdx;
  const int b0 = b0Hint >= 0 ? b0Hint : lower_bound_allele_near(B, nB, seqIdx, b0Guess);
  out.b0 = b0;
  if (b0 >= nB || B[b0].seqIdx != seqIdx) return;
  int bmc = 0, bden = 0;
  for (int ia = a0; ia < a1; ++ia) {
    const RV a = load_rv(A + ia);
    const int dA = rv_denom(a);
    for (int jb = b0; jb < nB; ++jb) {
      const RV b = load_rv(B + jb);
      if (b.seqIdx != seqIdx) break;
      if (!mates_ok(a, b)) continue;
      if (run_pos_less(a, ia, out.posKey, out.posIdx, A)) { out.posKey = a.key; out.posIdx = ia; }
      const int mc = a.mc + b.mc, den = dA + rv_denom(b);
      bool better;
      if (!out.valid) better = true;
      else if (mc != bmc) better = mc > bmc;
      else if (den != bden) better = den < bden;
      else if (rv_less(a, bo1)) better = true;
      else if (rv_less(bo1, a)) better = false;
      else if (ia != out.ia) better = run_pos_less(a, ia, bKeyA, out.ia, A);
      else better = run_pos_less(b, jb, bKeyB, out.jb, B);
      if (better) { out.valid = true; bmc = mc; bden = den; bo1 = a; bo2 = b; bKeyA = a.key; bKeyB = b.key; out.ia = ia; out.jb = jb; }
    }
  }
  if (!out.valid) return;
  out.mc = bmc; out.denom = bden; out.relaxed = bo1.relaxed + bo2.relaxed;
  if (bo1.st == 1) { out.ss = bo1.ss; out.se = bo2.se; } else { out.ss = bo2.ss; out.se = bo1.se; }
  out.o1 = bo1;
  out.relaxBy = 2;
  // IsOverlapIntersect (SeqSet.hpp:317-324) + intronic mismatches on both mates (SeqSet.hpp:2489-2500)
  if (R.relax && ((bo1.ss <= bo2.ss && bo1.se >= bo2.ss) || (bo2.ss <= bo1.ss && bo2.se >= bo1.ss)) &&
      bo1.mc < bo1.relaxed && bo2.mc < bo2.relaxed)
    out.relaxBy = 4;
}

__device__ bool allele_has_pair(const Rec *A, int nA, const Rec *B, int nB, int seqIdx) {
  const int a0 = lower_bound_allele(A, nA, seqIdx), b0 = lower_bound_allele(B, nB, seqIdx);
  for (int ia = a0; ia < nA && A[ia].seqIdx == seqIdx; ++ia) {
    const RV a = load_rv(A + ia);
    for (int jb = b0; jb < nB && B[jb].seqIdx == seqIdx; ++jb)
      if (mates_ok(a, load_rv(B + jb))) return true;
  }
  return false;
}

// TruncatedMatePairOverlap (SeqSet.hpp:502-523)
__device__ bool truncated_mate(const RefView &R, const RV &x, const RV &m1, const RV &m2) {
  if (x.st == 1) {
    const int far = x.se + m2.se - m1.se;
    if (R.len[x.seqIdx] - 1 < far || sep_exact(R, x.seqIdx, x.se, far + 1)) return true;
  } else {
    const int far = x.ss - (m1.ss - m2.ss);
    if (far < 0 || sep_exact(R, x.seqIdx, far - 1, x.ss)) return true;
  }
  return false;
}

__device__ __forceinline__ u64 mix64(u64 x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
  return x;
}

__device__ void pair_one(const PairParams &P, u32 fLocal, int lane) {
  const RefView &R = P.R;
  const u32 f = P.fragBase + fLocal;
  const u32 e1 = P.end1[f];
  const Rec *L1 = P.store + P.readOff[e1];
  const int n1 = (int)P.readCnt[e1];
  const bool pe = P.end2 != NULL;
  const Rec *L2 = NULL; int n2 = 0;
  if (pe) { const u32 e2 = P.end2[f]; L2 = P.store + P.readOff[e2]; n2 = (int)P.readCnt[e2]; }
  const bool paired = pe && n1 > 0 && n2 > 0;
  // the list whose allele runs drive the scan
  const Rec *A = L1; int nA = n1;
  if (pe && n1 == 0) { A = L2; nA = n2; }
  const Rec *B = paired ? L2 : NULL; const int nB = paired ? n2 : 0;
  const size_t gw = (size_t)(blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5));
  u32 *b0s = P.b0 + gw * P.b0Stride * 2;    // b0[b0Stride] then keys[b0Stride]
  PairEntry *stage = P.stage + gw * P.stageCap;
  u64 *stageKey = P.ordKey ? P.stageKey + gw * P.stageCap : NULL;
  u32 *stageIdx = P.ordKey ? P.stageIdx + gw * P.stageCap : NULL;
  u32 cnt = 0;
  bool assigned = false;
  u64 h0 = 0, h1 = 0;
  if (nA > 0) {
    // pass 1: best (matchCnt, similarity) over the alleles (SeqSet.hpp:2477-2487)
    // Every run start i keeps its (matchCnt, denominator) key in the warp's scratch: the later passes re-derive only the
    // alleles that can survive.  `delta` = offset of the previous chunk's last run in B, the guess for this chunk.
    u32 k1 = 0;
    u32 *keys = b0s + P.b0Stride;
    int delta = 0;
    for (int b = 0; b < nA; b += 32) {
      const int i = b + lane;
      {   // the records the next two rounds will touch: into L2 now (both lists are walked front to back)
        const int ip = i + 64;
        if (ip < nA) prefetch_l2(A + ip);
        if (paired) { const int g = ip + delta; if (g >= 0 && g < nB) prefetch_l2(B + g); }
      }
      int myDelta = 0x7fffffff;
      const bool runStart = i < nA && (i == 0 || A[i - 1].seqIdx != A[i].seqIdx);
      if (i < nA && !runStart) keys[i] = 0xffffffffu;        // later passes find the run starts without touching the records
      if (runStart) {
        int a1 = i + 1;
        while (a1 < nA && A[a1].seqIdx == A[i].seqIdx) ++a1;
        AlleleBest ab; eval_allele(R, paired, A, i, a1, B, nB, ab, -1, i + delta);
        if (paired) { b0s[i] = (u32)ab.b0; myDelta = ab.b0 - i; }
        const u32 key = ab.valid ? (((u32)ab.mc << 16) | (u32)(65535 - ab.denom)) : 0u;
        keys[i] = key;
        k1 = max(k1, key);
      }
      if (paired) {
        const unsigned has = __ballot_sync(FULL, myDelta != 0x7fffffff);
        if (has) delta = __shfl_sync(FULL, myDelta, 31 - __clz(has));
      }
    }
    k1 = __reduce_max_sync(FULL, k1);
    if (k1 != 0) {
      const int bestMc = (int)(k1 >> 16), bestDen = 65535 - (int)(k1 & 65535);
      // pass 2: relaxedMatchCnt of the first allele (assign order) that reaches the best
      u64 pk = ~0ull; int pi = 0x7fffffff; int bestRelax = 0;
      if (R.relax) {
        for (int b = 0; b < nA; b += 32) {
          const int i = b + lane;
          if (i < nA && keys[i] == k1) {
            int a1 = i + 1;
            while (a1 < nA && A[a1].seqIdx == A[i].seqIdx) ++a1;
            AlleleBest ab; eval_allele(R, paired, A, i, a1, B, nB, ab, paired ? (int)b0s[i] : -1);
            if (ab.valid && ab.mc == bestMc && ab.denom == bestDen && pos_less(ab.posKey, ab.posIdx, pk, pi)) {
              pk = ab.posKey; pi = ab.posIdx; bestRelax = ab.relaxed;
            }
          }
        }
        u64 k = pk; int ii = pi;
        warp_min_pair(k, ii);
        const unsigned who = __ballot_sync(FULL, pk == k && pi == ii);
        bestRelax = __shfl_sync(FULL, bestRelax, __ffs(who) - 1);
      }
      // pass 3: survivors, the representative (first survivor in assign order), the per-survivor tests, and the
      // rows themselves, staged in the warp's scratch row (allele order) until the fragment-level decisions are known
      double seg = (1.0 - R.sim) / 4.0;
      if (seg < 0.01) seg = 0.01;
      const bool hasN = P.hasN && P.hasN[f];
      bool anySep = false, anyFull = false, dangleFail = false;
      u64 rk = ~0ull; int ri = 0x7fffffff; int repIa = -1, repJb = -1;
      int nKeep = 0;
      for (int b = 0; b < nA; b += 32) {
        const int i = b + lane;
        bool keep = false;
        AlleleBest ab;
        // candidates only: the best key itself, or (relaxed mode) a matchCnt within relaxBy <= 4 of the best
        bool cand = false;
        if (i < nA) {
          const u32 key = keys[i];                       // 0xffffffff: not the first record of its allele; 0: no valid pair
          cand = key == k1 || (R.relax && key != 0 && key != 0xffffffffu && (int)(key >> 16) >= bestMc - 4);
        }
        if (cand) {
          int a1 = i + 1;
          while (a1 < nA && A[a1].seqIdx == A[i].seqIdx) ++a1;
          eval_allele(R, paired, A, i, a1, B, nB, ab, paired ? (int)b0s[i] : -1);
          keep = ab.valid && ((ab.mc == bestMc && ab.denom == bestDen) ||
                              (R.relax && ab.mc >= bestMc - ab.relaxBy && ab.relaxed == bestRelax));
        }
        const unsigned bal = __ballot_sync(FULL, keep);
        if (keep) {
          const int seqIdx = A[i].seqIdx;
          const bool sep = sep_exact(R, seqIdx, ab.ss, ab.se);
          anySep |= sep;
          anyFull |= ab.mc >= ab.denom;
          if (pos_less(ab.posKey, ab.posIdx, rk, ri)) { rk = ab.posKey; ri = ab.posIdx; repIa = ab.ia; repJb = ab.jb; }
          if (pe && !paired) {   // dangling mates (SeqSet.hpp:2554-2578)
            if (ab.mc < ab.denom || sep || ab.se - ab.ss + 1 + ab.o1.re - ab.o1.rs + 1 < 3 * HIT_LEN_REQ) dangleFail = true;
            else if ((ab.o1.st == 1 && ab.se + 100 < R.len[seqIdx]) || (ab.o1.st == 0 && ab.ss - 100 >= 0)) dangleFail = true;
          }
          const u32 slot = (u32)nKeep + __popc(bal & ((1u << lane) - 1));
          if (slot < P.stageCap) {         // a longer row is dropped by the -n cut anyway
            const double sim = (double)ab.mc / (double)ab.denom;
            double w = 1.0;
            if (sim < 1 - 3 * seg) w = 0.01;
            else if (sim < 1 - 2 * seg) w = 0.1;
            else if (sim < 1 - seg) w = 0.5;
            if (hasN) w /= 10.0;
            PairEntry e;
            e.alleleIdx = seqIdx; e.start = ab.ss; e.end = ab.se;
            e.weight = (float)w; e.qual = 1.0f; e.adjustWeight = 0.0f;
            stage[slot] = e;
            if (P.ordKey) { stageKey[slot] = ab.posKey; stageIdx[slot] = (u32)ab.posIdx; }
          }
        }
        nKeep += __popc(bal);
      }
      anySep = __any_sync(FULL, anySep); anyFull = __any_sync(FULL, anyFull); dangleFail = __any_sync(FULL, dangleFail);
      {
        u64 k = rk; int ii = ri;
        warp_min_pair(k, ii);
        const unsigned who = __ballot_sync(FULL, rk == k && ri == ii && rk != ~0ull);
        const int src = who ? __ffs(who) - 1 : 0;
        repIa = __shfl_sync(FULL, repIa, src); repJb = __shfl_sync(FULL, repJb, src);
      }
      bool drop = nKeep == 0 || dangleFail;
      // truncated reference (SeqSet.hpp:2581-2653): a better single-end hit whose mate falls off the allele
      if (!drop && paired) {
        const RV o1 = load_rv(A + repIa), o2 = load_rv(B + repJb);
        const int d1 = rv_denom(o1), d2 = rv_denom(o2);
        const double s1 = (double)o1.mc / (double)d1, s2 = (double)o2.mc / (double)d2;
        bool filter = false;
        // nothing in a list can beat the chosen mate unless the list's best (matchCnt, denominator) key exceeds the mate's
        const u32 key1 = ((u32)o1.mc << 16) | (u32)(65535 - d1), key2 = ((u32)o2.mc << 16) | (u32)(65535 - d2);
        const bool scanA = P.readTop[A == L1 ? e1 : P.end2[f]] > key1, scanB = P.readTop[B == L1 ? e1 : P.end2[f]] > key2;
        if (scanA)
        for (int i = lane; i < nA; i += 32) {
          const RV x = load_rv(A + i);
          bool better = x.mc > o1.mc;
          if (!better && x.mc == o1.mc && rv_denom(x) < d1) better = !allele_has_pair(A, nA, B, nB, x.seqIdx);
          if (!better) continue;
          if (truncated_mate(R, x, o1, o2)) filter = true;
          else if ((double)x.mc / (double)rv_denom(x) > s2 + 0.1) filter = true;
        }
        if (scanB)
        for (int j = lane; j < nB; j += 32) {
          const RV x = load_rv(B + j);
          bool better = x.mc > o2.mc;
          if (!better && x.mc == o2.mc && rv_denom(x) < d2) better = !allele_has_pair(A, nA, B, nB, x.seqIdx);
          if (!better) continue;
          if (truncated_mate(R, x, o2, o1)) filter = true;
          else if ((double)x.mc / (double)rv_denom(x) > s1 + 0.1) filter = true;
        }
        drop = __any_sync(FULL, filter);
      }
      assigned = !drop;
      // SetReadAssignments (Genotyper.hpp:778-832)
      if (!drop && P.maxAssign > 0 && nKeep > P.maxAssign) drop = true;
      if (!drop && anySep) drop = true;
      if (!drop) {
        const double adjust = anyFull ? 1.0 : 0.25;
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(P.outCtr, (unsigned long long)nKeep);
        base = __shfl_sync(FULL, base, 0);
        const bool fits = base + (unsigned long long)nKeep <= P.outCap;
        if (lane == 0) P.rowOff[fLocal] = base;
        __syncwarp();
        for (int k = lane; k < nKeep; k += 32) {     // staged row -> output (coalesced)
          PairEntry e = stage[k];
          e.adjustWeight = (float)(adjust * (double)e.weight);
          if (fits) {
            P.out[base + k] = e;
            if (P.ordKey) { P.ordKey[base + k] = stageKey[k]; P.ordIdx[base + k] = stageIdx[k]; }
          }
          const u64 m = mix64((u64)(u32)e.alleleIdx + 0x9e3779b97f4a7c15ull);
          h0 += m; h1 += mix64(m ^ 0xd6e8feb86659fd93ull);
        }
        cnt = (u32)nKeep;
      }
    }
  }
  if (P.rowHash) {
#pragma unroll
    for (int o = 16; o; o >>= 1) { h0 += __shfl_xor_sync(FULL, h0, o); h1 += __shfl_xor_sync(FULL, h1, o); }
  }
  if (lane == 0) {
    P.rowCnt[fLocal] = cnt | (assigned ? 0x80000000u : 0u);
    if (P.rowHash) { P.rowHash[2 * (size_t)fLocal] = h0 ^ ((u64)cnt << 40); P.rowHash[2 * (size_t)fLocal + 1] = h1; }
  }
}

// MINB = resident blocks per SM the register budget is compiled for (3: no spills; 4: 128 registers, 16 warps/SM —
// the kernel waits on dependent loads of the record lists, so the extra warps pay)
template <int MINB>
__global__ void __launch_bounds__(128, MINB) k_pair(PairParams P) {
  const int lane = threadIdx.x & 31;
  for (;;) {
    u32 w = 0;
    if (lane == 0) w = atomicAdd(P.workCtr, 1u);
    w = __shfl_sync(FULL, w, 0);
    if (w >= P.nFrag) break;
    pair_one(P, w, lane);
  }
}

}  // namespace t1k
