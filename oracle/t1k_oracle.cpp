// TEST INFRASTRUCTURE ONLY — see t1k_oracle.h.  Sequential CPU restatement of the T1K hot path.
// Every function cites the reference lines whose behaviour it restates (paths under /root/reference).
// Written for clarity, not speed: full DP matrices, std::sort everywhere, one read at a time.
#include "t1k_oracle.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

namespace {

// k-mer length and required hit length of the oracle in use.  The genotyper fixes them (Genotyper.cpp:207, SeqSet.hpp:764);
// the candidate filter of fastq-extractor derives its own (FastqExtractor.cpp:381-418).  Single-threaded test code: every
// extern entry point loads them from its oracle object before doing anything.
int K = 11;                  // Genotyper.cpp:207
const int RADIUS = 10;       // SeqSet.hpp:763
int HIT_LEN_REQ = 31;        // SeqSet.hpp:764
const int BAND = 5;          // AlignAlgo.hpp:215

// nucToNum & 3 (Genotyper.cpp:37-40, KmerCode.hpp:99): A0 C1 G2 T3, everything else -1&3 = 3
inline int code2(char c) { return c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : 3; }
inline bool base_eq(char t, char p) { return t == p || t == 'N' || p == 'N'; }  // AlignAlgo.hpp:304-305

struct Posting { uint32_t idx, off; };

struct Allele {
  std::string seq;
  std::vector<int> sep;          // -1, N positions, len (SeqSet.hpp:924-928)
  std::vector<uint8_t> exon;     // isValidDiff[].exon (SeqSet.hpp:651-671)
  std::vector<int> cov;          // posWeight[i].count[consensus base] (Q11)
  int effLen;                    // ComputeEffectiveLen, SeqSet.hpp:747-758
  int weight;
};

struct Ov {
  int seqIdx, readStart, readEnd, seqStart, seqEnd, strand, matchCnt;
  double similarity;
  int leftClip, rightClip, relaxedMatchCnt;
  std::vector<std::pair<int, int> > coords;  // (readOff, seqOff) chain
};

// struct _overlap::operator<  SeqSet.hpp:103-127
bool ov_less(const Ov &a, const Ov &b) {
  if (a.matchCnt != b.matchCnt) return a.matchCnt > b.matchCnt;
  if (a.similarity != b.similarity) return a.similarity > b.similarity;
  if (a.readEnd - a.readStart != b.readEnd - b.readStart) return a.readEnd - a.readStart > b.readEnd - b.readStart;
  if (a.seqIdx != b.seqIdx) return a.seqIdx < b.seqIdx;
  if (a.strand != b.strand) return a.strand < b.strand;
  if (a.readStart != b.readStart) return a.readStart < b.readStart;
  if (a.readEnd != b.readEnd) return a.readEnd < b.readEnd;
  if (a.seqStart != b.seqStart) return a.seqStart < b.seqStart;
  return a.seqEnd < b.seqEnd;
}

}  // namespace

struct T1KOracle {
  std::vector<Allele> al;
  std::vector<uint32_t> kstart;   // 4^K + 1 (k <= 12); longer k-mers: `codes` (sorted, distinct) + cstart
  std::vector<uint32_t> codes, cstart;
  std::vector<Posting> post;      // per k-mer: (allele asc, offset asc) = insertion order of KmerIndex.hpp:58-71
  double sim;
  bool relax;
  int k, hitLenReq;
  // posting range of one k-mer (KmerIndex::Search, KmerIndex.hpp:93-105)
  void range(uint32_t code, uint32_t &lo, uint32_t &hi) const {
    if (!kstart.empty()) { lo = kstart[code]; hi = kstart[code + 1]; return; }
    std::vector<uint32_t>::const_iterator it = std::lower_bound(codes.begin(), codes.end(), code);
    if (it == codes.end() || *it != code) { lo = hi = 0; return; }
    const size_t i = (size_t)(it - codes.begin());
    lo = cstart[i]; hi = cstart[i + 1];
  }
};

namespace {

// ---------------------------------------------------------------------------------------------
// k-mer stream (KmerCode::Append, KmerCode.hpp:93-108): valid iff no N in the last K chars.
struct KStream {
  uint32_t code; int bad;  // bad = chars since last N (-1 none)
  KStream() : code(0), bad(-1) {}
  void push(char c) {
    if (bad != -1) ++bad;
    code = ((code << 2) & (uint32_t)((1ull << (2 * K)) - 1)) | (uint32_t)code2(c);
    if (c == 'N') bad = 0;
    if (bad >= K) bad = -1;
  }
  bool valid() const { return bad == -1; }
};

// KmerIndex::BuildIndexFromRead, KmerIndex.hpp:107-130 (incl. the i==kl quirk, Q1)
struct Triple { uint32_t code; Posting p; };
void index_allele(std::vector<Triple> &lists, const std::string &s, int id) {
  int len = (int)s.size();
  if (len < K) return;
  KStream ks; uint32_t prev = 0;
  int i;
  for (i = 0; i < K - 1; ++i) ks.push(s[i]);
  for (; i < len; ++i) {
    ks.push(s[i]);
    if (ks.valid() && (i == K || ks.code != prev)) {
      Triple t; t.code = ks.code; t.p.idx = (uint32_t)id; t.p.off = (uint32_t)(i - K + 1);
      lists.push_back(t);
    }
    prev = ks.code;
  }
}

std::string revcomp(const std::string &s) {  // SeqSet::ReverseComplement, SeqSet.hpp:2103-2114
  std::string r(s.size(), 'N');
  for (size_t i = 0; i < s.size(); ++i) {
    char c = s[s.size() - 1 - i];
    r[i] = c == 'N' ? 'N' : "ACGT"[3 - code2(c)];
  }
  return r;
}

// ---------------------------------------------------------------------------------------------
// AlignAlgo::GlobalAlignment, AlignAlgo.hpp:215-421.  Full (lenp+1)x(lent+1) matrices, band 5 (+ length
// difference on the long side), sentinels at the band edge, stale-index init of e[0][j] (Q5),
// traceback tie order diag > (f>=e ? D : I).
int global_alignment(const char *t, int lent, const char *p, int lenp, std::vector<int8_t> &ops) {
  ops.clear();
  if (lent == 0 || lenp == 0) return 0;
  if (lent == 1 && lenp == 1) {
    if (base_eq(t[0], p[0])) { ops.push_back(0); return 2; }
    ops.push_back(1); return -2;
  }
  int lb = BAND, rb = BAND;
  if (lent > lenp) rb += lent - lenp; else if (lent < lenp) lb += lenp - lent;
  const int W = lent + 1;
  const int negInf = (lent + 1) * (lenp + 1) * -4;
  std::vector<int> m((size_t)(lenp + 1) * W), e(m.size()), f(m.size());
  m[0] = e[0] = f[0] = 0;
  for (int i = 1; i <= lenp; ++i) { e[i * W] = -4 - i; f[i * W] = -4 - 4 * i; m[i * W] = -4 - 4 * i; }
  for (int j = 1; j <= lent; ++j) { f[j] = -4 - j; e[j] = -4 + (lenp + 1) * -4; m[j] = -4 - 4 * j; }
  for (int i = 1; i <= lenp; ++i) {
    int start = i - lb < 1 ? 1 : i - lb;
    int end = i + rb > lent ? lent : i + rb;
    if (start > 1) e[i * W + start - 1] = f[i * W + start - 1] = m[i * W + start - 1] = negInf;
    if (end < lent) e[i * W + end + 1] = f[i * W + end + 1] = m[i * W + end + 1] = negInf;
    for (int j = start; j <= end; ++j) {
      int ev = std::max(e[(i - 1) * W + j] - 1, m[(i - 1) * W + j] - 5);
      int fv = std::max(f[i * W + j - 1] - 1, m[i * W + j - 1] - 5);
      int mv = m[(i - 1) * W + j - 1] + (base_eq(t[j - 1], p[i - 1]) ? 2 : -2);
      e[i * W + j] = ev; f[i * W + j] = fv;
      m[i * W + j] = std::max(mv, std::max(ev, fv));
    }
  }
  int ret = m[lenp * W + lent];
  int ti = lenp, tj = lent, mat = 0;
  while (ti > 0 || tj > 0) {
    if (mat == 0) {
      int a = f[ti * W + tj] >= e[ti * W + tj] ? 3 : 2;
      if (ti > 0 && tj > 0) {
        bool eq = base_eq(t[tj - 1], p[ti - 1]);
        if (m[(ti - 1) * W + tj - 1] + (eq ? 2 : -2) == m[ti * W + tj]) a = eq ? 0 : 1;
      }
      if (a <= 1) { ops.push_back((int8_t)a); --ti; --tj; }
      else mat = a == 2 ? 1 : 2;
    } else if (mat == 1) {
      ops.push_back(2);
      if (ti > 0) { mat = (m[(ti - 1) * W + tj] - 5 == e[ti * W + tj]) ? 0 : 1; --ti; }
      else mat = 2;
    } else {
      ops.push_back(3);
      if (tj > 0) { mat = (m[ti * W + tj - 1] - 5 == f[ti * W + tj]) ? 0 : 2; --tj; }
      else mat = 1;
    }
  }
  std::reverse(ops.begin(), ops.end());
  return ret;
}

int count_match(const std::vector<int8_t> &ops) {  // GetAlignStats, SeqSet.hpp:438-455 (matches only)
  int c = 0;
  for (size_t i = 0; i < ops.size(); ++i) c += ops[i] == 0;
  return c;
}

// IsSeparatorInRange, SeqSet.hpp:487-498 (sentinels -1 and len are part of the list)
bool sep_in_range(const Allele &a, int s, int e) {
  for (size_t i = 0; i < a.sep.size(); ++i)
    if (a.sep[i] >= s && a.sep[i] <= e) return true;
  return false;
}

// ---------------------------------------------------------------------------------------------
// SeqSet::GetHitsFromRead (SeqSet.hpp:1071-1229) for one strand: which k-mers are looked up.
// The skip rule (list >= 100, not first/last k-mer, up to K/2 in a row) bypasses the prev update (Q2).
struct Hit { int strand; uint32_t idx; int a; int b; };

void collect_hits(const T1KOracle &o, const std::string &r, int strand, uint32_t &prev, std::vector<Hit> &hits) {
  int len = (int)r.size();
  KStream ks;
  int i, skip = 0;
  for (i = 0; i < K - 1; ++i) ks.push(r[i]);
  for (; i < len; ++i) {
    ks.push(r[i]);
    if (i == K - 1 || prev != ks.code) {
      uint32_t lo = 0, hi = 0;
      if (ks.valid()) o.range(ks.code, lo, hi);
      int size = (int)(hi - lo);
      if (size >= 100 && i != K - 1 && i != len - 1 && skip < K / 2) { ++skip; continue; }
      skip = 0;
      for (uint32_t j = lo; j < hi; ++j) {
        Hit h; h.strand = strand; h.idx = o.post[j].idx; h.a = i - K + 1; h.b = (int)o.post[j].off;
        hits.push_back(h);
      }
    }
    prev = ks.code;
  }
}

bool hit_less(const Hit &x, const Hit &y) {  // struct _hit::operator<, SeqSet.hpp:74-86
  if (x.strand != y.strand) return x.strand < y.strand;
  if (x.idx != y.idx) return x.idx < y.idx;
  if (x.a != y.a) return x.a < y.a;
  return x.b < y.b;
}

// LongestIncreasingSubsequence over .first, SeqSet.hpp:352-436 (non-strict probe, strict extend,
// then drop equal .second, Q4)
std::vector<std::pair<int, int> > lis(const std::vector<std::pair<int, int> > &h) {
  int n = (int)h.size();
  std::vector<int> top(n), link(n, -1);
  int ret = 1;
  top[0] = 0;
  for (int i = 1; i < n; ++i) {
    int tag;
    if (h[top[ret - 1]].first <= h[i].first) tag = ret - 1;
    else {
      int l = 0, r = ret - 1; tag = -2;
      while (l <= r) {
        int mid = (l + r) / 2;
        if (h[i].first == h[top[mid]].first) { tag = mid; break; }
        if (h[i].first < h[top[mid]].first) r = mid - 1; else l = mid + 1;
      }
      if (tag == -2) tag = l - 1;
    }
    if (tag == -1) { top[0] = i; link[i] = -1; }
    else if (h[i].first > h[top[tag]].first) {
      if (tag == ret - 1) { top[ret] = i; ++ret; link[i] = top[tag]; }
      else if (h[i].first < h[top[tag + 1]].first) { top[tag + 1] = i; link[i] = top[tag]; }
    }
  }
  std::vector<std::pair<int, int> > out;
  int k = top[ret - 1];
  for (int i = ret - 1; i >= 0; --i) { out.push_back(h[k]); k = link[k]; }
  std::reverse(out.begin(), out.end());
  std::vector<std::pair<int, int> > ded;
  for (size_t i = 0; i < out.size(); ++i)
    if (i == 0 || out[i].second != ded.back().second) ded.push_back(out[i]);
  return ded;
}

// GetTotalHitLengthOnRead / OnSeq, SeqSet.hpp:1032-1069
int hit_span(const std::vector<std::pair<int, int> > &c, bool onRead) {
  int n = (int)c.size(), ret = 0;
  for (int i = 0; i < n;) {
    int j;
    for (j = i + 1; j < n; ++j) {
      int cur = onRead ? c[j].first : c[j].second, pre = onRead ? c[j - 1].first : c[j - 1].second;
      if (cur > pre + K - 1) break;
    }
    ret += (onRead ? c[j - 1].first - c[i].first : c[j - 1].second - c[i].second) + K;
    i = j;
  }
  return ret;
}

// SeqSet::GetOverlapsFromHits for one (strand, allele) group, SeqSet.hpp:1303-1553 with filter=0, isRef=true.
void chain_group(const Hit *g, int n, std::vector<Ov> &out) {
  if (n < 3) return;
  struct T { int a, b, c; };
  std::vector<T> d(n);
  for (int i = 0; i < n; ++i) { d[i].a = g[i].a; d[i].b = g[i].b; d[i].c = g[i].a - g[i].b; }
  std::sort(d.begin(), d.end(), [](const T &x, const T &y) {  // CompSortHitCoordDiff, SeqSet.hpp:266-274
    if (x.c != y.c) return x.c < y.c;
    if (x.b != y.b) return x.b < y.b;
    return x.a < y.a;
  });
  int dom = 0;
  for (int s = 0; s < n;) {
    int e, cur = d[s].c, curCnt = 1, domCnt = 0;
    for (e = s + 1; e < n; ++e) {
      int diff = std::abs(d[e].c - d[e - 1].c);
      if (diff > RADIUS) break;
      if (diff == 0) ++curCnt;
      else {
        if (curCnt > domCnt) { dom = cur; domCnt = curCnt; }
        cur = d[e].c; curCnt = 1;
      }
    }
    if (curCnt > domCnt) dom = cur;      // SeqSet.hpp:1393-1397 (count not updated, Q3)
    if (e - s < 3 || (e - s) * K < HIT_LEN_REQ) { s = e; continue; }
    // keep, per read offset, the hits closest to the dominant diagonal (SeqSet.hpp:1437-1456)
    std::map<int, int> best;
    for (int k = s; k < e; ++k) {
      int dist = std::abs(d[k].a - d[k].b - dom);
      std::map<int, int>::iterator it = best.find(d[k].a);
      if (it == best.end() || it->second > dist) best[d[k].a] = dist;
    }
    std::vector<std::pair<int, int> > conc;
    for (int k = s; k < e; ++k)
      if (std::abs(d[k].a - d[k].b - dom) == best[d[k].a]) conc.push_back(std::make_pair(d[k].a, d[k].b));
    std::sort(conc.begin(), conc.end(), [](const std::pair<int, int> &x, const std::pair<int, int> &y) {
      if (x.second != y.second) return x.second < y.second;   // CompSortPairBInc, SeqSet.hpp:233-239
      return x.first < y.first;
    });
    std::vector<std::pair<int, int> > chain = lis(conc);
    int sz = (int)chain.size();
    s = e;
    if (sz * K < HIT_LEN_REQ) continue;
    int hitLen = hit_span(chain, true);
    if (hitLen < HIT_LEN_REQ) continue;
    if (hit_span(chain, false) < HIT_LEN_REQ) continue;
    Ov o;
    o.seqIdx = (int)g[0].idx; o.strand = g[0].strand;
    o.readStart = chain[0].first; o.readEnd = chain[sz - 1].first + K - 1;
    o.seqStart = chain[0].second; o.seqEnd = chain[sz - 1].second + K - 1;
    o.matchCnt = 2 * hitLen; o.similarity = 0;
    o.leftClip = o.rightClip = 0; o.relaxedMatchCnt = 0;
    o.coords = chain;
    out.push_back(o);
  }
}

// IsOverlapLowComplex, SeqSet.hpp:458-485
bool low_complex(const std::string &r, int s, int e) {
  int cnt[4] = {0, 0, 0, 0};
  for (int i = s; i <= e; ++i) if (r[i] != 'N') ++cnt[code2(r[i])];
  int low = 0, lowTotal = 0;
  for (int i = 0; i < 4; ++i) if (cnt[i] <= 2) { ++low; lowTotal += cnt[i]; }
  if (lowTotal * 7 >= e - s + 1) return false;
  return low >= 2;
}

// SeqSet::GetOverlapsFromRead, SeqSet.hpp:1594-1912 (strand=0, barcode=-1, all seqs isRef, radius 10)
int overlaps_from_read(const T1KOracle &o, const std::string &read, std::vector<Ov> &ovs) {
  int len = (int)read.size();
  if (len < K) return -1;
  std::string rc = revcomp(read);
  std::vector<Hit> hits;
  uint32_t prev = 0;
  collect_hits(o, read, 1, prev, hits);
  collect_hits(o, rc, -1, prev, hits);       // prev k-mer state survives into the second pass
  std::sort(hits.begin(), hits.end(), hit_less);  // SortHits, SeqSet.hpp:1558-1590 (same order both branches)
  size_t n = hits.size();
  for (size_t i = 0; i < n;) {
    size_t j = i + 1;
    while (j < n && hits[j].strand == hits[i].strand && hits[j].idx == hits[i].idx) ++j;
    chain_group(&hits[i], (int)(j - i), ovs);
    i = j;
  }
  if (ovs.empty()) return 0;
  // keep the strand of the best seed overlap only (SeqSet.hpp:1619-1648)
  size_t best = 0;
  for (size_t i = 1; i < ovs.size(); ++i) if (ov_less(ovs[i], ovs[best])) best = i;
  int strand = ovs[best].strand;
  std::vector<Ov> kept;
  for (size_t i = 0; i < ovs.size(); ++i) if (ovs[i].strand == strand) kept.push_back(ovs[i]);
  ovs.swap(kept);
  std::vector<int8_t> ops;
  for (size_t i = 0; i < ovs.size(); ++i) {
    Ov &v = ovs[i];
    const std::string &r = v.strand == 1 ? read : rc;
    const std::string &cons = o.al[v.seqIdx].seq;
    const std::vector<std::pair<int, int> > &c = v.coords;
    int mc = 2 * K;
    for (size_t j = 1; j < c.size(); ++j) {     // SeqSet.hpp:1700-1833
      int pa = c[j - 1].first, pb = c[j - 1].second, a = c[j].first, b = c[j].second;
      bool aOv = pa + K - 1 >= a, bOv = pb + K - 1 >= b;
      if (pb - pa == b - a) {
        if (aOv) mc += 2 * (a - pa);
        else {
          mc += 2 * K;
          global_alignment(cons.c_str() + pb + K, b - (pb + K), r.c_str() + pa + K, a - (pa + K), ops);
          mc += 2 * count_match(ops);
        }
      } else if (aOv && !bOv) mc += 2 * (a - pa);
      else if (!aOv && bOv) mc += 2 * (b - pb);
      else if (aOv && bOv) mc += 2 * std::min(a - pa, b - pb);
      else {
        mc += 2 * K;
        global_alignment(cons.c_str() + pb + K, b - (pb + K), r.c_str() + pa + K, a - (pa + K), ops);
        mc += 2 * count_match(ops);
      }
    }
    v.matchCnt = mc;
    v.similarity = (double)mc / (v.seqEnd - v.seqStart + 1 + v.readEnd - v.readStart + 1);
    if (low_complex(r, v.readStart, v.readEnd)) v.similarity = 0;
  }
  std::vector<Ov> pass;
  for (size_t i = 0; i < ovs.size(); ++i) if (!(ovs[i].similarity < o.sim)) pass.push_back(ovs[i]);
  ovs.swap(pass);
  return (int)ovs.size();
}

// SeqSet::ExtendOverlap, SeqSet.hpp:1994-2100
int extend_overlap(const T1KOracle &o, const std::string &r, const Ov &ov, Ov &ex) {
  const Allele &al = o.al[ov.seqIdx];
  int len = (int)r.size(), clen = (int)al.seq.size();
  int lo = std::min(ov.readStart, ov.seqStart), leftClip = 0, rightClip = 0;
  if (ov.readStart > ov.seqStart) leftClip = ov.readStart - ov.seqStart;
  for (int i = 0; i < lo; ++i)
    if (al.seq[ov.seqStart - i - 1] == 'N') { leftClip = lo - i; lo = i; break; }
  std::vector<int8_t> ops;
  global_alignment(al.seq.c_str() + ov.seqStart - lo, lo, r.c_str() + ov.readStart - lo, lo, ops);
  int matchCnt = count_match(ops);
  int ro = std::min(len - 1 - ov.readEnd, clen - 1 - ov.seqEnd);
  if (len - 1 - ov.readEnd > clen - 1 - ov.seqEnd) rightClip = len - 1 - ov.readEnd - (clen - 1 - ov.seqEnd);
  for (int i = 0; i < ro; ++i)
    if (al.seq[ov.seqEnd + 1 + i] == 'N') { rightClip = ro - i; ro = i; break; }
  global_alignment(al.seq.c_str() + ov.seqEnd + 1, ro, r.c_str() + ov.readEnd + 1, ro, ops);
  matchCnt += count_match(ops);
  ex = ov; ex.coords.clear();
  ex.readStart = ov.readStart - lo; ex.readEnd = ov.readEnd + ro;
  ex.seqStart = ov.seqStart - lo; ex.seqEnd = ov.seqEnd + ro;
  ex.matchCnt = 2 * matchCnt + ov.matchCnt;
  ex.similarity = (double)ex.matchCnt / (ex.readEnd - ex.readStart + 1 + ex.seqEnd - ex.seqStart + 1);
  ex.relaxedMatchCnt = ex.matchCnt;
  ex.leftClip = leftClip; ex.rightClip = rightClip;
  int ret = ex.similarity < o.sim ? 0 : 1;
  if (leftClip > 0 || rightClip > 0) {
    ex.matchCnt += 2 * leftClip + 2 * rightClip;
    ex.similarity = (double)ex.matchCnt /
        (ex.readEnd - ex.readStart + 1 + ex.seqEnd - ex.seqStart + 1 + 2 * leftClip + 2 * rightClip);
  }
  return ret;
}

// SeqSet::AssignRead, SeqSet.hpp:2119-2303
int assign_read(T1KOracle &o, const std::string &read, int weight, std::vector<Ov> &assign) {
  assign.clear();
  std::vector<Ov> ovs;
  int cnt = overlaps_from_read(o, read, ovs);
  if (cnt <= 0 || o.al.empty()) return -1;
  std::sort(ovs.begin(), ovs.end(), ov_less);
  int len = (int)read.size();
  std::string rc = revcomp(read);
  const std::string &r = ovs[0].strand == -1 ? rc : read;
  std::vector<Ov> ext;
  bool onlyClip = false; int good = -1;
  for (int i = 0; i < cnt; ++i) {               // order-dependent scan, Q7
    const Allele &al = o.al[ovs[i].seqIdx];
    if (sep_in_range(al, ovs[i].seqStart, ovs[i].seqEnd)) continue;
    bool needClip = sep_in_range(al, ovs[i].seqStart - ovs[i].readStart, ovs[i].seqEnd + (len - ovs[i].readEnd - 1));
    if (onlyClip && ovs[i].matchCnt < good && (!needClip || ovs[i].similarity < 0.95)) continue;
    Ov ex;
    if (extend_overlap(o, r, ovs[i], ex) == 1) {
      ext.push_back(ex);
      if (!onlyClip && (good == -1 || ovs[i].matchCnt > good)) good = ovs[i].matchCnt;
    } else onlyClip = true;
  }
  if (!ext.empty() && weight >= 0) {
    int bestMc = ext[0].matchCnt;
    for (size_t i = 0; i < ext.size(); ++i) bestMc = std::max(bestMc, ext[i].matchCnt);
    std::vector<int8_t> ops;
    for (size_t i = 0; i < ext.size(); ++i) {
      Ov &e = ext[i];
      if (e.matchCnt < bestMc - 10) { e.relaxedMatchCnt = 0; continue; }   // Q8
      Allele &al = o.al[e.seqIdx];
      global_alignment(al.seq.c_str() + e.seqStart, e.seqEnd - e.seqStart + 1, r.c_str() + e.readStart,
                       e.readEnd - e.readStart + 1, ops);
      if (o.relax) {                                 // SeqSet.hpp:2215-2246
        int m = 0, refPos = e.seqStart;
        for (size_t k = 0; k < ops.size(); ++k) {
          if (al.exon[refPos]) { if (ops[k] == 0) ++m; } else ++m;
          if (ops[k] != 2) ++refPos;
        }
        e.relaxedMatchCnt = 2 * m;
      } else e.relaxedMatchCnt = e.matchCnt;
      if (weight > 0) {                              // SeqSet.hpp:2253-2274, Q6/Q11
        int refPos = e.seqStart, readPos = e.readStart;
        for (size_t k = 0; k < ops.size(); ++k) {
          if (ops[k] == 0 && r[readPos] != 'N' && al.seq[refPos] == r[readPos]) al.cov[refPos] += weight;
          if (ops[k] != 2) ++refPos;
          if (ops[k] != 3) ++readPos;
        }
      }
    }
  }
  if (ext.size() > 1000) {                           // SeqSet.hpp:2290-2298
    std::sort(ext.begin(), ext.end(), ov_less);
    size_t j;
    for (j = 1; j < ext.size(); ++j) if (ext[j].similarity < ext[0].similarity - 0.1) break;
    ext.resize(j);
  }
  assign.swap(ext);
  return (int)assign.size();
}

Ov from_pod(const OracleOverlap &p) {
  Ov v;
  v.seqIdx = p.seqIdx; v.readStart = p.readStart; v.readEnd = p.readEnd; v.seqStart = p.seqStart; v.seqEnd = p.seqEnd;
  v.strand = p.strand; v.matchCnt = p.matchCnt; v.relaxedMatchCnt = p.relaxedMatchCnt;
  v.leftClip = p.leftClip; v.rightClip = p.rightClip;
  v.similarity = (double)p.matchCnt /
      (p.readEnd - p.readStart + 1 + p.seqEnd - p.seqStart + 1 + 2 * p.leftClip + 2 * p.rightClip);
  return v;
}

struct Frag {   // struct _fragmentOverlap, SeqSet.hpp:146-173
  int seqIdx, seqStart, seqEnd, matchCnt, relaxedMatchCnt;
  double similarity;
  bool hasMate;
  Ov o1, o2;
};

bool frag_less(const Frag &a, const Frag &b) {
  if (a.matchCnt != b.matchCnt) return a.matchCnt > b.matchCnt;
  if (a.similarity != b.similarity) return a.similarity > b.similarity;
  return ov_less(a.o1, b.o1);
}

// TruncatedMatePairOverlap, SeqSet.hpp:502-523
bool truncated_mate(const T1KOracle &o, const Ov &x, const Ov &m1, const Ov &m2) {
  const Allele &al = o.al[x.seqIdx];
  if (x.strand == 1) {
    int far = x.seqEnd + m2.seqEnd - m1.seqEnd;
    if ((int)al.seq.size() - 1 < far || sep_in_range(al, x.seqEnd, far + 1)) return true;
  } else if (x.strand == -1) {
    int far = x.seqStart - (m1.seqStart - m2.seqStart);
    if (far < 0 || sep_in_range(al, far - 1, x.seqStart)) return true;
  }
  return false;
}

// SeqSet::ReadAssignmentToFragmentAssignment, SeqSet.hpp:2310-2655
void pair_fragment(const T1KOracle &o, const std::vector<Ov> &ov1, const std::vector<Ov> *pov2, std::vector<Frag> &assign) {
  assign.clear();
  std::vector<std::pair<int, int> > cand;
  int n1 = (int)ov1.size();
  if (!pov2) {
    for (int i = 0; i < n1; ++i) cand.push_back(std::make_pair(i, -1));
  } else if (n1 == 0 || pov2->empty()) {
    for (int i = 0; i < n1; ++i) cand.push_back(std::make_pair(i, -1));
    for (int i = 0; i < (int)pov2->size(); ++i) cand.push_back(std::make_pair(-1, i));
  } else {
    const std::vector<Ov> &ov2 = *pov2;
    std::map<int, std::vector<int> > byAllele;
    for (int i = 0; i < (int)ov2.size(); ++i) byAllele[ov2[i].seqIdx].push_back(i);
    for (int i = 0; i < n1; ++i) {
      std::map<int, std::vector<int> >::iterator it = byAllele.find(ov1[i].seqIdx);
      if (it == byAllele.end()) continue;
      for (size_t k = 0; k < it->second.size(); ++k) {
        int j = it->second[k];
        if (ov1[i].strand == ov2[j].strand) continue;
        if ((ov1[i].strand == 1 && ov1[i].seqStart < ov2[j].seqStart) ||
            (ov1[i].strand == -1 && ov1[i].seqStart > ov2[j].seqStart))
          cand.push_back(std::make_pair(i, j));
      }
    }
  }
  std::map<int, int> slot;
  for (size_t c = 0; c < cand.size(); ++c) {
    Frag f;
    if (cand[c].first >= 0) {
      const Ov &a = ov1[cand[c].first];
      f.matchCnt = a.matchCnt; f.similarity = a.similarity; f.seqIdx = a.seqIdx;
      f.seqStart = a.seqStart; f.seqEnd = a.seqEnd; f.hasMate = false; f.o1 = a;
      f.relaxedMatchCnt = a.relaxedMatchCnt;
      if (cand[c].second >= 0) {
        const Ov &b = (*pov2)[cand[c].second];
        f.matchCnt += b.matchCnt; f.relaxedMatchCnt += b.relaxedMatchCnt;
        if (a.strand == 1) f.seqEnd = b.seqEnd; else f.seqStart = b.seqStart;
        f.similarity = (double)f.matchCnt /
            (a.readEnd - a.readStart + 1 + b.readEnd - b.readStart + 1 + a.seqEnd - a.seqStart + 1 +
             b.seqEnd - b.seqStart + 1 + 2 * a.leftClip + 2 * a.rightClip + 2 * b.leftClip + 2 * b.rightClip);
        f.hasMate = true; f.o2 = b;
      }
    } else {
      const Ov &a = (*pov2)[cand[c].second];
      f.matchCnt = a.matchCnt; f.similarity = a.similarity; f.seqIdx = a.seqIdx;
      f.seqStart = a.seqStart; f.seqEnd = a.seqEnd; f.hasMate = false; f.o1 = a;
      f.relaxedMatchCnt = a.relaxedMatchCnt;
    }
    std::map<int, int>::iterator it = slot.find(f.seqIdx);
    if (it != slot.end()) { if (frag_less(f, assign[it->second])) assign[it->second] = f; }
    else { slot[f.seqIdx] = (int)assign.size(); assign.push_back(f); }
  }
  int bestMc = -1; double bestSim = 0; int bestRelax = 0;
  for (size_t i = 0; i < assign.size(); ++i)
    if (assign[i].matchCnt > bestMc || (assign[i].matchCnt == bestMc && assign[i].similarity > bestSim)) {
      bestMc = assign[i].matchCnt; bestSim = assign[i].similarity; bestRelax = assign[i].relaxedMatchCnt;
    }
  std::vector<Frag> kept;
  for (size_t i = 0; i < assign.size(); ++i) {
    const Frag &f = assign[i];
    int relaxBy = 2;
    if (o.relax && f.hasMate && f.o1.seqIdx == f.o2.seqIdx &&
        ((f.o1.seqStart <= f.o2.seqStart && f.o1.seqEnd >= f.o2.seqStart) ||
         (f.o2.seqStart <= f.o1.seqStart && f.o2.seqEnd >= f.o1.seqStart)) &&   // IsOverlapIntersect, :317-324
        f.o1.matchCnt < f.o1.relaxedMatchCnt && f.o2.matchCnt < f.o2.relaxedMatchCnt)
      relaxBy = 4;
    if (f.matchCnt == bestMc && f.similarity == bestSim) kept.push_back(f);
    else if (o.relax && f.matchCnt >= bestMc - relaxBy && f.relaxedMatchCnt == bestRelax) kept.push_back(f);
  }
  assign.swap(kept);
  if (!assign.empty() && pov2 && !assign[0].hasMate) {           // dangling mates, :2554-2578
    bool drop = false;
    for (size_t i = 0; i < assign.size() && !drop; ++i) {
      const Frag &f = assign[i];
      const Allele &al = o.al[f.seqIdx];
      if (f.similarity < 1 || sep_in_range(al, f.seqStart, f.seqEnd) ||
          f.seqEnd - f.seqStart + 1 + f.o1.readEnd - f.o1.readStart + 1 < 3 * HIT_LEN_REQ) drop = true;
      else if ((f.o1.strand == 1 && f.seqEnd + 100 < (int)al.seq.size()) || (f.o1.strand == -1 && f.seqStart - 100 >= 0))
        drop = true;
    }
    if (drop) assign.clear();
  }
  if (!assign.empty() && pov2 && assign[0].hasMate) {             // truncated reference, :2581-2653
    const std::vector<Ov> &ov2 = *pov2;
    const Frag rep = assign[0];
    bool filter = false;
    for (int pass = 0; pass < 2 && !filter; ++pass) {
      const std::vector<Ov> &list = pass == 0 ? ov1 : ov2;
      const Ov &mine = pass == 0 ? rep.o1 : rep.o2;
      const Ov &other = pass == 0 ? rep.o2 : rep.o1;
      for (size_t i = 0; i < list.size() && !filter; ++i) {
        const Ov &x = list[i];
        bool better = x.matchCnt > mine.matchCnt ||
            (x.matchCnt == mine.matchCnt && x.similarity > mine.similarity && slot.find(x.seqIdx) == slot.end());
        if (!better) continue;
        if (truncated_mate(o, x, mine, other)) filter = true;
        else if (x.similarity > other.similarity + 0.1) filter = true;
      }
    }
    if (filter) assign.clear();
  }
}

// Genotyper::ReadAssignmentWeight, Genotyper.hpp:205-230
double assignment_weight(const T1KOracle &o, double similarity, bool hasN) {
  double ret = 1, seg = (1 - o.sim) / 4.0;
  if (seg < 0.01) seg = 0.01;
  if (similarity < 1 - 3 * seg) ret = 0.01;
  else if (similarity < 1 - 2 * seg) ret = 0.1;
  else if (similarity < 1 - seg) ret = 0.5;
  if (hasN) ret /= 10.0;
  return ret;
}

}  // namespace

extern "C" {

static void build_index(T1KOracle *o, std::vector<Triple> &lists) {
  // stable by code => inside a k-mer the postings keep their insertion order (allele asc, offset asc)
  std::stable_sort(lists.begin(), lists.end(), [](const Triple &a, const Triple &b) { return a.code < b.code; });
  o->post.resize(lists.size());
  for (size_t i = 0; i < lists.size(); ++i) o->post[i] = lists[i].p;
  o->kstart.clear(); o->codes.clear(); o->cstart.clear();
  if (K <= 12) {
    o->kstart.assign(((size_t)1 << (2 * K)) + 1, 0);
    for (size_t i = 0; i < lists.size(); ++i) ++o->kstart[lists[i].code + 1];
    for (size_t c = 0; c + 1 < o->kstart.size(); ++c) o->kstart[c + 1] += o->kstart[c];
  } else {
    for (size_t i = 0; i < lists.size(); ++i)
      if (i == 0 || lists[i].code != lists[i - 1].code) { o->codes.push_back(lists[i].code); o->cstart.push_back((uint32_t)i); }
    o->cstart.push_back((uint32_t)lists.size());
  }
}
static inline void use(const T1KOracle *o) { K = o->k; HIT_LEN_REQ = o->hitLenReq; }

T1KOracle *t1ko_create(int32_t n, const char *bases, const int64_t *off, const int32_t *exonPtr, const int32_t *exonSE,
                       const int32_t *seqWeight, double similarity, int32_t relaxIntron) {
  T1KOracle *o = new T1KOracle;
  o->sim = similarity; o->relax = relaxIntron != 0;
  o->k = 11; o->hitLenReq = 31;
  use(o);
  o->al.resize(n);
  std::vector<Triple> lists;
  for (int i = 0; i < n; ++i) {
    Allele &a = o->al[i];
    a.seq.assign(bases + off[i], bases + off[i + 1]);
    int len = (int)a.seq.size();
    a.sep.push_back(-1);
    a.effLen = 0;
    for (int j = 0; j < len; ++j) {
      if (a.seq[j] == 'N') a.sep.push_back(j);
      if (a.seq[j] != 'N' || (j > 0 && a.seq[j - 1] != 'N')) ++a.effLen;
    }
    a.sep.push_back(len);
    a.exon.assign(len, 0);
    for (int e = exonPtr[i]; e < exonPtr[i + 1]; ++e)
      for (int j = exonSE[2 * e]; j <= exonSE[2 * e + 1] && j < len; ++j) a.exon[j] = 1;
    a.cov.assign(len, 0);
    a.weight = seqWeight ? seqWeight[i] : 1;
    index_allele(lists, a.seq, i);
  }
  build_index(o, lists);
  return o;
}

// ---------------------------------------------------------------------------------------------
// The candidate filter of fastq-extractor.  Set-up as FastqExtractor.cpp:272-273,381-418:
// k = max(9, SeqSet::InferKmerLength) over ALL reference sequences as loaded by InputRefFa (no collapsing of identical
// sequences, SeqSet.hpp:872-904), hitLenRequired = max(27 paired / 23 single, mean read length / 5, k).
int32_t t1ko_infer_kmer_length(int64_t totalLength) {   // SeqSet::InferKmerLength, SeqSet.hpp:2830-2845
  int ret = 0;
  while (totalLength) { ++ret; totalLength /= 4; }
  return ret + 1;
}
T1KOracle *t1ko_filter_create(int32_t n, const char *bases, const int64_t *off, int32_t k, int32_t hitLenRequired, double similarity) {
  if (k < 1 || k > 15) return NULL;
  T1KOracle *o = new T1KOracle;
  o->sim = similarity; o->relax = false;
  o->k = k; o->hitLenReq = hitLenRequired;
  use(o);
  o->al.resize(n);
  std::vector<Triple> lists;
  for (int i = 0; i < n; ++i) {
    Allele &a = o->al[i];
    a.seq.assign(bases + off[i], bases + off[i + 1]);
    a.effLen = (int)a.seq.size(); a.weight = 1;
    index_allele(lists, a.seq, i);
  }
  build_index(o, lists);
  return o;
}
// IsLowComplexity, FastqExtractor.cpp:89-112
int32_t t1ko_is_low_complexity(const char *seq) {
  int cnt[5] = {0, 0, 0, 0, 0};
  int i;
  for (i = 0; seq[i]; ++i) { if (seq[i] == 'N') ++cnt[4]; else ++cnt[code2(seq[i])]; }
  if (cnt[0] >= i / 2 || cnt[1] >= i / 2 || cnt[2] >= i / 2 || cnt[3] >= i / 2 || cnt[4] >= i / 10) return 1;
  int lowCnt = 0;
  for (i = 0; i < 4; ++i) if (cnt[i] <= 2) ++lowCnt;
  return lowCnt >= 2;
}
// SeqSet::HasHitInSet, SeqSet.hpp:1915-1990
int32_t t1ko_has_hit_in_set(T1KOracle *o, const char *readC) {
  use(o);
  const std::string read(readC);
  const int len = (int)read.size();
  if (len < K) return 0;
  std::vector<Hit> hits;
  uint32_t prev = 0;
  collect_hits(*o, read, 1, prev, hits);
  collect_hits(*o, revcomp(read), -1, prev, hits);
  if (hits.empty()) return 0;
  // buckets per (strand tag, sequence) in arrival order; the first largest one, strand -1 first (:1929-1957)
  std::map<std::pair<int, uint32_t>, std::vector<Hit> > buckets;
  for (size_t i = 0; i < hits.size(); ++i) buckets[std::make_pair(hits[i].strand == 1 ? 1 : 0, hits[i].idx)].push_back(hits[i]);
  int mx = -1;
  const std::vector<Hit> *best = NULL;
  for (std::map<std::pair<int, uint32_t>, std::vector<Hit> >::const_iterator it = buckets.begin(); it != buckets.end(); ++it)
    if ((int)it->second.size() > mx) { mx = (int)it->second.size(); best = &it->second; }     // map order = (tag, idx) ascending
  if (K * mx < HIT_LEN_REQ) return 0;
  std::vector<Ov> ovs;
  chain_group(best->data(), (int)best->size(), ovs);
  const int mismatchThreshold = (int)(len * (1 - o->sim)) * K;
  for (size_t i = 0; i < ovs.size(); ++i) if (len - ovs[i].matchCnt / 2 <= mismatchThreshold) return 1;
  return 0;
}
// IsGoodCandidate, FastqExtractor.cpp:114-119
int32_t t1ko_is_good_candidate(T1KOracle *o, const char *read) {
  return !t1ko_is_low_complexity(read) && t1ko_has_hit_in_set(o, read);
}

void t1ko_destroy(T1KOracle *o) { delete o; }

int32_t t1ko_global_alignment(const char *t, int32_t lent, const char *p, int32_t lenp, int8_t *ops, int32_t *nOps) {
  std::vector<int8_t> v;
  int s = global_alignment(t, lent, p, lenp, v);
  if (ops) std::copy(v.begin(), v.end(), ops);
  if (nOps) *nOps = (int32_t)v.size();
  return s;
}

int32_t t1ko_assign_read(T1KOracle *o, const char *read, int32_t weight, OracleOverlap *out, int32_t cap) {
  use(o);
  std::vector<Ov> a;
  int ret = assign_read(*o, std::string(read), weight, a);
  for (size_t i = 0; i < a.size() && (int)i < cap; ++i) {
    OracleOverlap &p = out[i];
    p.seqIdx = a[i].seqIdx; p.readStart = a[i].readStart; p.readEnd = a[i].readEnd;
    p.seqStart = a[i].seqStart; p.seqEnd = a[i].seqEnd; p.strand = a[i].strand;
    p.matchCnt = a[i].matchCnt; p.relaxedMatchCnt = a[i].relaxedMatchCnt;
    p.leftClip = a[i].leftClip; p.rightClip = a[i].rightClip;
  }
  return ret;
}

void t1ko_coverage(T1KOracle *o, int32_t allele, int32_t *out) {
  std::copy(o->al[allele].cov.begin(), o->al[allele].cov.end(), out);
}
void t1ko_coverage_reset(T1KOracle *o) {
  for (size_t i = 0; i < o->al.size(); ++i) std::fill(o->al[i].cov.begin(), o->al[i].cov.end(), 0);
}
int32_t t1ko_allele_len(T1KOracle *o, int32_t allele) { return (int32_t)o->al[allele].seq.size(); }
int32_t t1ko_effective_len(T1KOracle *o, int32_t allele) { return o->al[allele].effLen; }

// SeqSet::GetSeqMissingBaseCoverage, SeqSet.hpp:2717-2755
int32_t t1ko_missing_coverage(T1KOracle *o, int32_t allele) {
  const Allele &a = o->al[allele];
  std::vector<int> c;
  for (size_t i = 0; i < a.seq.size(); ++i) if (a.exon[i]) c.push_back(a.seq[i] == 'N' ? 0 : a.cov[i]);
  std::sort(c.begin(), c.end());
  if (c.empty()) return 0;
  double cutoff = c[c.size() / 2] * 0.01;
  if (cutoff < 1) cutoff = 1;
  size_t i;
  for (i = 0; i < c.size(); ++i) if (c[i] >= cutoff) break;
  return (int32_t)i;
}

// ReadAssignmentToFragmentAssignment + Genotyper::SetReadAssignments (Genotyper.hpp:778-832)
int32_t t1ko_fragment_assign(T1KOracle *o, const OracleOverlap *p1, int32_t n1, const OracleOverlap *p2, int32_t n2,
                             int32_t hasN, int32_t maxAssign, OracleAssignment *out, int32_t cap) {
  use(o);
  std::vector<Ov> a(n1), b;
  for (int i = 0; i < n1; ++i) a[i] = from_pod(p1[i]);
  if (p2) { b.resize(n2); for (int i = 0; i < n2; ++i) b[i] = from_pod(p2[i]); }
  std::vector<Frag> fr;
  pair_fragment(*o, a, p2 ? &b : NULL, fr);
  int cnt = (int)fr.size();
  if (maxAssign > 0 && cnt > maxAssign) return 0;
  for (int i = 0; i < cnt; ++i) if (sep_in_range(o->al[fr[i].seqIdx], fr[i].seqStart, fr[i].seqEnd)) return 0;
  double maxSim = 0;
  for (int i = 0; i < cnt; ++i) maxSim = std::max(maxSim, fr[i].similarity);
  double adjust = maxSim < 1 ? 0.25 : 1.0;
  int w = 0;
  for (int i = 0; i < cnt && w < cap; ++i, ++w) {
    OracleAssignment &x = out[w];
    x.alleleIdx = fr[i].seqIdx; x.start = fr[i].seqStart; x.end = fr[i].seqEnd;
    x.weight = (float)assignment_weight(*o, fr[i].similarity, hasN != 0);
    x.qual = 1.0f;
    x.adjustWeight = (float)(adjust * x.weight);
  }
  return w;
}

// Genotyper::EMupdate (Genotyper.hpp:372-421)
static void em_update(int G, int E, const int64_t *rowPtr, const int32_t *col, const double *count, const int32_t *ecLen,
                      const double *x, double *xNext, double *rc) {
  std::fill(rc, rc + E, 0.0);
  for (int g = 0; g < G; ++g) {
    double psum = 0;
    for (int64_t k = rowPtr[g]; k < rowPtr[g + 1]; ++k) psum += x[col[k]];
    if (psum == 0) psum = 1;
    for (int64_t k = rowPtr[g]; k < rowPtr[g + 1]; ++k) rc[col[k]] += count[g] * (x[col[k]] / psum);
  }
  double norm = 0;
  for (int e = 0; e < E; ++e) norm += rc[e] / ecLen[e];
  for (int e = 0; e < E; ++e) xNext[e] = rc[e] / ecLen[e] / norm;
}

// Genotyper::QuantifyAlleleEquivalentClass main loop (Genotyper.hpp:1234-1316) + SetAlleleAbundance mask (:957-1014)
int32_t t1ko_em(int32_t G, int32_t E, const int64_t *rowPtr, const int32_t *col, const double *count,
                const int32_t *ecLen, const double *x0in, double minAlpha, double filterFrac,
                int32_t nAlleles, const int32_t *ecAllelePtr, const int32_t *ecAlleles,
                const int32_t *alleleMajor, const int32_t *alleleGene, int32_t nMajor, int32_t nGene,
                double *xOut, double *rcOut) {
  std::vector<double> x0(x0in, x0in + E), x1(E), x2(E), x3(E), rc(E);
  int ret = 0;
  const int maxIter = 1000;
  for (int t = 0; t < maxIter; ++t) {
    ++ret;
    em_update(G, E, rowPtr, col, count, ecLen, &x0[0], &x1[0], &rc[0]);
    em_update(G, E, rowPtr, col, count, ecLen, &x1[0], &x2[0], &rc[0]);
    double sr = 0, sv = 0;                                   // SQUAREMalpha, :424-437
    for (int e = 0; e < E; ++e) {
      sr += (x1[e] - x0[e]) * (x1[e] - x0[e]);
      double v = x2[e] - 2 * x1[e] + x0[e];
      sv += v * v;
    }
    double alpha = sv == 0 ? -1 : -std::sqrt(sr) / std::sqrt(sv);
    if (minAlpha < 0 && alpha < minAlpha) alpha = minAlpha;
    for (int e = 0; e < E; ++e)
      x3[e] = x0[e] - 2 * alpha * (x1[e] - x0[e]) + alpha * alpha * (x2[e] - 2 * x1[e] + x0[e]);
    em_update(G, E, rowPtr, col, count, ecLen, &x3[0], &x1[0], &rc[0]);
    double diff = 0;
    for (int e = 0; e < E; ++e) { diff += std::fabs(x1[e] - x0[e]); x0[e] = x1[e]; }
    if (diff < 1e-5 && t < maxIter - 2) t = maxIter - 2;
    if (t > 0 && t % 10 == 0 && nAlleles > 0) {
      std::vector<double> ab(nAlleles, 0.0), ecAb(nAlleles, 0.0), major(nMajor, 0.0), gmax(nGene, 0.0);
      for (int e = 0; e < E; ++e) {
        int size = ecAllelePtr[e + 1] - ecAllelePtr[e];
        double a = rc[e] / ecLen[e] * 1000.0;
        for (int k = ecAllelePtr[e]; k < ecAllelePtr[e + 1]; ++k) { ab[ecAlleles[k]] = a / size; ecAb[ecAlleles[k]] = a; }
      }
      for (int i = 0; i < nAlleles; ++i) major[alleleMajor[i]] += ab[i];
      for (int i = 0; i < nAlleles; ++i) gmax[alleleGene[i]] = std::max(gmax[alleleGene[i]], major[alleleMajor[i]]);
      for (int i = 0; i < nAlleles; ++i)
        if (major[alleleMajor[i]] < filterFrac * 0.5 * gmax[alleleGene[i]]) ecAb[i] = 0;
      for (int e = 0; e < E; ++e) x0[e] = ecAb[ecAlleles[ecAllelePtr[e]]];
    }
  }
  std::copy(x0.begin(), x0.end(), xOut);
  std::copy(rc.begin(), rc.end(), rcOut);
  return ret;
}

}  // extern "C"
