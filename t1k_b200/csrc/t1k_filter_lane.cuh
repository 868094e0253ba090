// SURVEY.md §8f N1 (candidate filter of fastq-extractor) — lane-level building blocks with a RUNTIME k-mer length.  The
// kernel around them is k_filter (t1k_filter.cuh), the C-ABI entry point t1k_filter_batch; tests/filter_emu.cpp runs the
// same code sequentially on the CPU against what the reference binary keeps.
//
// Reference: IsLowComplexity FastqExtractor.cpp:89-112, SeqSet::HasHitInSet SeqSet.hpp:1915-1990 (GetHitsFromRead
// :1071-1229 with kmerLength = max(9, InferKmerLength), GetOverlapsFromHits :1232-1556 with filter = 0).
#pragma once
#include "t1k_core.cuh"

namespace t1k {
namespace filt {

struct IndexView {
  const u32 *kstart;       // direct-address table, 4^k + 1 entries (k <= 15: at most 4 GB, resident in HBM)
  const Posting *post;     // postings in KmerIndex::BuildIndexFromRead order
  int k;                   // k-mer length (FastqExtractor.cpp:272,411-418)
  int hitLenReq;           // hitLenRequired (FastqExtractor.cpp:381-407,415-416)
};

// IsLowComplexity on the packed forward strand: a base holding at least half of the read, >= 10 % N, or two bases with
// at most two occurrences
T1K_HD bool read_low_complexity(const u64 *seq2, const u64 *n2, int len) {
  int cnt[5] = {0, 0, 0, 0, 0};
  T1K_NOUNROLL
  for (int kk = 0; kk < len; kk += 32) {
    const u64 w = seq2[kk >> 5], nm = n2[kk >> 5];
    const u64 keep = ~nm & M55 & lowmask2(len - kk);
    const u64 lo = w & M55, hi = (w >> 1) & M55;
    cnt[0] += popc64(~lo & ~hi & keep);
    cnt[1] += popc64(lo & ~hi & keep);
    cnt[2] += popc64(~lo & hi & keep);
    cnt[3] += popc64(lo & hi & keep);
    cnt[4] += popc64(nm & M55 & lowmask2(len - kk));
  }
  if (cnt[0] >= len / 2 || cnt[1] >= len / 2 || cnt[2] >= len / 2 || cnt[3] >= len / 2 || cnt[4] >= len / 10) return true;
  int low = 0;
  for (int i = 0; i < 4; ++i) low += cnt[i] <= 2;
  return low >= 2;
}

// The k-mers GetHitsFromRead looks up on one strand (skip rule Q2 with skipLimit = k / 2; `prev` is the previous looked-at
// code and survives from the forward into the reverse pass as in the reference).  Returns the number of seeds;
// seedA / lo / hi receive read offset and posting range of each.
T1K_HDN inline int seed_list(const IndexView &I, const u64 *seq2, const u64 *n2, int len, u32 &prev, u16 *seedA, u32 *lo, u32 *hi) {
  const int k = I.k, NP = len - k + 1;
  const u64 codeMask = (1ull << (2 * k)) - 1, nMask = codeMask & M55;
  int nS = 0, skip = 0;
  T1K_NOUNROLL
  for (int a = 0; a < NP; ++a) {
    const u32 code = (u32)(fetch32(seq2, a) & codeMask);
    const bool valid = (fetch32(n2, a) & nMask) == 0;
    if (a == 0 || prev != code) {
      u32 l = 0, h = 0;
      if (valid) { l = I.kstart[code]; h = I.kstart[code + 1]; }
      const int size = (int)(h - l);
      if (size >= 100 && a != 0 && a != NP - 1 && skip < k / 2) { ++skip; continue; }
      skip = 0;
      if (size > 0) { seedA[nS] = (u16)a; lo[nS] = l; hi[nS] = h; ++nS; }
    }
    prev = code;
  }
  return nS;
}

// GetTotalHitLengthOnRead / OnSeq (SeqSet.hpp:1032-1069) of a chain with runtime k
T1K_HD int chain_span(const u32 *c, int n, int k, bool onRead) {
  int ret = 0;
  T1K_NOUNROLL
  for (int i = 0; i < n;) {
    int j = i + 1;
    T1K_NOUNROLL
    for (; j < n; ++j) {
      const int cur = onRead ? hit_a(c[j]) : hit_b(c[j]), pre = onRead ? hit_a(c[j - 1]) : hit_b(c[j - 1]);
      if (cur > pre + k - 1) break;
    }
    ret += (onRead ? hit_a(c[j - 1]) - hit_a(c[i]) : hit_b(c[j - 1]) - hit_b(c[i])) + k;
    i = j;
  }
  return ret;
}

// GetOverlapsFromHits (filter = 0, reference sequences) on the hits of ONE (strand, sequence) bucket: the largest hit
// length (= matchCnt / 2) over the overlaps it would emit, 0 if none.  h: n encoded hits (hit_make(readOffset, seqOffset)) in
// (readOffset, seqOffset) order, sorted in place by diagonal.  scratch: 12 * n + FILTER_USED_BYTES bytes.
constexpr int FILTER_USED_BYTES = 2 * 1024;      // min |diagonal - dominant| per read offset (read offsets < 1024)
T1K_HDN inline int bucket_best_hit_len(u32 *h, int n, int k, int hitLenReq, u8 *scratch) {
  if (n < 3) return 0;
  T1K_NOUNROLL
  for (int i = 1; i < n; ++i) {                 // insertion sort by (diagonal, seqOffset, readOffset)
    const u32 v = h[i];
    if (!hit_diag_less(v, h[i - 1])) continue;
    int j = i - 1;
    T1K_NOUNROLL
    while (j >= 0 && hit_diag_less(v, h[j])) { h[j + 1] = h[j]; --j; }
    h[j + 1] = v;
  }
  int best = 0, dom = 0;
  T1K_NOUNROLL
  for (int s = 0; s < n;) {
    int e, cur = hit_a(h[s]) - hit_b(h[s]), curCnt = 1, domCnt = 0, prevC = cur;
    T1K_NOUNROLL
    for (e = s + 1; e < n; ++e) {
      const int c = hit_a(h[e]) - hit_b(h[e]);
      const int diff = c - prevC;
      if (diff > RADIUS) break;
      if (diff == 0) ++curCnt;
      else { if (curCnt > domCnt) { dom = cur; domCnt = curCnt; } cur = c; curCnt = 1; }
      prevC = c;
    }
    if (curCnt > domCnt) dom = cur;
    const int m = e - s;
    if (m < 3 || m * k < hitLenReq) { s = e; continue; }
    // per read offset the hits closest to the dominant diagonal (SeqSet.hpp:1437-1456), in (seqOffset, readOffset) order
    u16 *used = (u16 *)scratch;
    u32 *conc = (u32 *)(scratch + FILTER_USED_BYTES);
    u32 *chain = conc + m;
    u16 *top = (u16 *)(chain + m);
    u16 *link = top + m;
    T1K_NOUNROLL
    for (int q = s; q < e; ++q) used[hit_a(h[q])] = 0xFFFF;
    T1K_NOUNROLL
    for (int q = s; q < e; ++q) {
      int d = iabs(hit_a(h[q]) - hit_b(h[q]) - dom);
      if (d > 0xFFFE) d = 0xFFFE;
      if (used[hit_a(h[q])] > d) used[hit_a(h[q])] = (u16)d;
    }
    int cn = 0;
    T1K_NOUNROLL
    for (int q = s; q < e; ++q) {
      const u32 v = h[q];
      int d = iabs(hit_a(v) - hit_b(v) - dom);
      if (d > 0xFFFE) d = 0xFFFE;
      if (d == used[hit_a(v)]) {
        int j = cn - 1;
        T1K_NOUNROLL
        while (j >= 0 && conc[j] > v) { conc[j + 1] = conc[j]; --j; }
        conc[j + 1] = v; ++cn;
      }
    }
    // LIS over read offsets (non-strict probe, strict extend; Q4), then drop equal seq offsets
    int ret = 1;
    top[0] = 0; link[0] = 0xFFFF;
    T1K_NOUNROLL
    for (int i = 1; i < cn; ++i) {
      const int ai = hit_a(conc[i]);
      int tag;
      if (hit_a(conc[top[ret - 1]]) <= ai) tag = ret - 1;
      else {
        int l = 0, r = ret - 1; tag = -2;
        T1K_NOUNROLL
        while (l <= r) {
          const int mid = (l + r) / 2, am = hit_a(conc[top[mid]]);
          if (ai == am) { tag = mid; break; }
          if (ai < am) r = mid - 1; else l = mid + 1;
        }
        if (tag == -2) tag = l - 1;
      }
      if (tag == -1) { top[0] = (u16)i; link[i] = 0xFFFF; }
      else if (ai > hit_a(conc[top[tag]])) {
        if (tag == ret - 1) { top[ret] = (u16)i; ++ret; link[i] = top[tag]; }
        else if (ai < hit_a(conc[top[tag + 1]])) { top[tag + 1] = (u16)i; link[i] = top[tag]; }
      }
    }
    {
      int q = top[ret - 1];
      T1K_NOUNROLL
      for (int i = ret - 1; i >= 0; --i) { chain[i] = conc[q]; q = link[q]; }
    }
    int sz = 0;
    T1K_NOUNROLL
    for (int i = 0; i < ret; ++i)
      if (i == 0 || hit_b(chain[i]) != hit_b(chain[sz - 1])) chain[sz++] = chain[i];
    s = e;
    if (sz * k < hitLenReq) continue;
    const int hitLen = chain_span(chain, sz, k, true);
    if (hitLen < hitLenReq) continue;
    if (chain_span(chain, sz, k, false) < hitLenReq) continue;
    if (hitLen > best) best = hitLen;
  }
  return best;
}

// the verdict of HasHitInSet given the best bucket's largest hit length (SeqSet.hpp:1969-1977)
T1K_HD bool hit_length_passes(int len, int bestHitLen, int k, double similarity) {
  const int mismatchThreshold = (int)(len * (1 - similarity)) * k;
  return bestHitLen > 0 && len - bestHitLen <= mismatchThreshold;
}

}  // namespace filt
}  // namespace t1k
