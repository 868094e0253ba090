// Host-side packing of alleles / reads into the 2-bit planes described in t1k_core.cuh and the
// direct-address k-mer table.  Plain C++ (no CUDA), shared by the ABI implementation and the test emulator.
#pragma once
#include <stdint.h>
#include <string.h>
#include <algorithm>
#include <utility>
#include <vector>

#include "t1k_core.cuh"

namespace t1k {

inline int code_of(char c) { return c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : 3; }   // nucToNum & 3
inline bool valid_base(char c) { return c == 'A' || c == 'C' || c == 'G' || c == 'T' || c == 'N'; }

struct PackedRef {
  int32_t nAlleles = 0;
  std::vector<u64> seq2, n2, ex2;
  std::vector<u64> wordOff;
  std::vector<int32_t> len;
  std::vector<u8> hasN;
  std::vector<AlleleMeta> meta;
  std::vector<u32> kstart;
  std::vector<Posting> post;
  // the device's form of the index (see KmerEntry in t1k_core.cuh): per k-mer the postings regrouped by
  // (tile of 32 consecutive alleles, offset) with one allele bit mask per entry
  std::vector<u64> covOff;            // [nAlleles] first coverage entry of the allele (tile-interleaved layout, see RefView)
  size_t covEntries = 0;
  std::vector<KmerInfo> kinfo;        // [4^K + 1]
  std::vector<KmerEntry> entries;
  size_t totalWords = 0;
};

inline int code_of(char c);
inline bool valid_base(char c);
// postings (sorted by k-mer, allele, offset) -> tile entries sorted by (k-mer, tile, offset).  An allele's hits of one
// k-mer stay in ascending offset order, which is the order GetHitsFromRead appends them in (SeqSet.hpp:1124-1150).
inline void build_tile_index(size_t nK, const std::vector<u32> &kstart, const std::vector<Posting> &post, std::vector<KmerInfo> &kinfo,
                             std::vector<KmerEntry> &entries) {
  kinfo.assign(nK + 1, KmerInfo{0, 0});
  entries.clear();
  std::vector<std::pair<u32, u32> > cur;      // (offset, allele bit) of the tile being collected
  for (size_t c = 0; c < nK; ++c) {
    kinfo[c].estart = (u32)entries.size(); kinfo[c].pstart = kstart[c];
    const u32 lo = kstart[c], hi = kstart[c + 1];
    for (u32 j = lo; j < hi;) {
      const u32 tile = post[j].idx >> 5;
      cur.clear();
      for (; j < hi && (post[j].idx >> 5) == tile; ++j) cur.push_back(std::make_pair(post[j].off, post[j].idx & 31u));
      std::stable_sort(cur.begin(), cur.end(), [](const std::pair<u32, u32> &x, const std::pair<u32, u32> &y) { return x.first < y.first; });
      const size_t first = entries.size();
      for (size_t q = 0; q < cur.size(); ++q) {
        if (q == 0 || cur[q].first != cur[q - 1].first) { KmerEntry e; e.tile = tile; e.off = cur[q].first; e.mask = 0; e.more = 0; entries.push_back(e); }
        entries.back().mask |= 1u << cur[q].second;
      }
      const size_t cnt = entries.size() - first;
      for (size_t q = 0; q < cnt; ++q) entries[first + q].more = (u32)(cnt - 1 - q);
    }
  }
  kinfo[nK].estart = (u32)entries.size(); kinfo[nK].pstart = kstart[nK];
}
inline void build_tile_index(PackedRef &P) { build_tile_index((size_t)1 << (2 * KMER), P.kstart, P.post, P.kinfo, P.entries); }

// Postings of a sequence set for a RUNTIME k-mer length (the candidate filter of fastq-extractor: k = max(9,
// SeqSet::InferKmerLength), FastqExtractor.cpp:272,411-418), in KmerIndex::BuildIndexFromRead order incl. the i == kl quirk
// (KmerIndex.hpp:107-130, Q1); then the tile form.  Returns false on a character outside ACGTN.
inline bool build_filter_index(int32_t n, const char *bases, const int64_t *off, int k, std::vector<KmerInfo> &kinfo, std::vector<KmerEntry> &entries) {
  const size_t nK = (size_t)1 << (2 * k);
  std::vector<u32> cnt(nK + 1, 0), kstart;
  std::vector<Posting> post;
  for (int pass = 0; pass < 2; ++pass) {
    for (int i = 0; i < n; ++i) {
      const char *s = bases + off[i];
      const int len = (int)(off[i + 1] - off[i]);
      if (len < k) continue;
      u32 code = 0, prev = 0; int bad = -1;
      const u32 mask = (u32)(nK - 1);
      for (int j = 0; j < len; ++j) {
        if (pass == 0 && !valid_base(s[j])) return false;
        if (bad != -1) ++bad;
        code = (code >> 2) | ((u32)code_of(s[j]) << (2 * (k - 1)));      // first base of the window in the low bits
        if (s[j] == 'N') bad = 0;
        if (bad >= k) bad = -1;
        if (j < k - 1) continue;
        code &= mask;
        if (bad == -1 && (j == k || code != prev)) {
          if (pass == 0) ++cnt[code + 1];
          else { Posting p; p.idx = (u32)i; p.off = (u32)(j - k + 1); post[cnt[code]++] = p; }
        }
        prev = code;
      }
    }
    if (pass == 0) {
      for (size_t c = 0; c < nK; ++c) cnt[c + 1] += cnt[c];
      kstart = cnt;
      post.resize(cnt[nK]);
    }
  }
  build_tile_index(nK, kstart, post, kinfo, entries);
  return true;
}

inline void set2(std::vector<u64> &plane, u64 w0, int pos, u64 v) { plane[w0 + (pos >> 5)] |= v << ((pos & 31) * 2); }

// returns false on a character outside ACGTN
inline bool pack_reference(int32_t n, const char *bases, const int64_t *off, const int32_t *exonPtr, const int32_t *exonSE,
                           PackedRef &P) {
  P.nAlleles = n;
  P.wordOff.resize(n); P.len.resize(n); P.hasN.assign(n, 0);
  size_t words = 8;                           // pad words before the first allele: diag_hot reads read-aligned windows
  for (int i = 0; i < n; ++i) {
    int len = (int)(off[i + 1] - off[i]);
    P.wordOff[i] = words; P.len[i] = len;
    words += (size_t)(len + 31) / 32 + 2;     // >= 1 pad word after every allele (coverage diff writes at len)
  }
  words += 8;                                 // ... and after the last one
  P.totalWords = words;
  P.seq2.assign(words, 0); P.n2.assign(words, 0); P.ex2.assign(words, 0);
  const size_t nK = (size_t)1 << (2 * KMER);
  std::vector<u32> cnt(nK + 1, 0);
  // two passes over the k-mers (count, fill) in allele order => postings sorted by (k-mer, allele, offset),
  // the insertion order of KmerIndex::BuildIndexFromRead (KmerIndex.hpp:107-130) including its i==kl quirk.
  for (int pass = 0; pass < 2; ++pass) {
    for (int i = 0; i < n; ++i) {
      const char *s = bases + off[i];
      int len = P.len[i];
      u64 w0 = P.wordOff[i];
      if (pass == 0) {
        for (int j = 0; j < len; ++j) {
          if (!valid_base(s[j])) return false;
          set2(P.seq2, w0, j, (u64)code_of(s[j]));
          if (s[j] == 'N') { set2(P.n2, w0, j, 1); P.hasN[i] = 1; }
        }
        for (int e = exonPtr[i]; e < exonPtr[i + 1]; ++e)
          for (int j = exonSE[2 * e]; j <= exonSE[2 * e + 1] && j < len; ++j)
            if (j >= 0) P.ex2[w0 + (j >> 5)] |= 1ull << ((j & 31) * 2);
      }
      if (len < KMER) continue;
      u32 code = 0, prev = 0; int bad = -1;
      const u32 mask = (u32)(nK - 1);
      for (int j = 0; j < len; ++j) {
        if (bad != -1) ++bad;
        // same code layout as the device: base j of the window at bits 2*(position in window)
        code = (code >> 2) | ((u32)code_of(s[j]) << (2 * (KMER - 1)));
        if (s[j] == 'N') bad = 0;
        if (bad >= KMER) bad = -1;
        if (j < KMER - 1) continue;
        code &= mask;
        if (bad == -1 && (j == KMER || code != prev)) {
          if (pass == 0) ++cnt[code + 1];
          else { Posting p; p.idx = (u32)i; p.off = (u32)(j - KMER + 1); P.post[cnt[code]++] = p; }
        }
        prev = code;
      }
    }
    if (pass == 0) {
      for (size_t c = 0; c < nK; ++c) cnt[c + 1] += cnt[c];
      P.kstart = cnt;
      P.post.resize(cnt[nK]);
      // cnt[c] now = start of bucket c; fill pass advances it
    }
  }
  P.meta.resize(n);
  for (int i = 0; i < n; ++i) { P.meta[i].wordOff = P.wordOff[i]; P.meta[i].len = P.len[i]; P.meta[i].hasN = P.hasN[i]; }
  // coverage layout: per tile of 32 alleles, 32 x (longest allele of the tile + 2) interleaved entries
  P.covOff.resize(n);
  size_t cov = 0;
  for (int t0 = 0; t0 < n; t0 += 32) {
    int longest = 0;
    for (int i = t0; i < n && i < t0 + 32; ++i) longest = std::max(longest, P.len[i]);
    for (int i = t0; i < n && i < t0 + 32; ++i) P.covOff[i] = cov + (size_t)(i - t0);
    cov += (size_t)32 * ((size_t)longest + 2);
  }
  P.covEntries = cov;
  build_tile_index(P);
  return true;
}

// pack one read (both strands) into planes of RWORDS words each; returns false on invalid characters
inline bool pack_read(const char *s, int len, u64 *fseq, u64 *fn, u64 *rseq, u64 *rn, int rwords = RWORDS) {
  for (int w = 0; w < rwords; ++w) fseq[w] = fn[w] = rseq[w] = rn[w] = 0;
  for (int j = 0; j < len; ++j) {
    char c = s[j];
    if (!valid_base(c)) return false;
    int sh = (j & 31) * 2, w = j >> 5;
    int rj = len - 1 - j, rsh = (rj & 31) * 2, rw = rj >> 5;
    if (c == 'N') {
      fseq[w] |= 3ull << sh; fn[w] |= 1ull << sh;
      rseq[rw] |= 3ull << rsh; rn[rw] |= 1ull << rsh;
    } else {
      u64 v = (u64)code_of(c);
      fseq[w] |= v << sh;
      rseq[rw] |= (3 - v) << rsh;
    }
  }
  return true;
}

}  // namespace t1k
