// Lane-level building blocks of the B200 alignment path.  Everything here is per-(read-end, allele)
// work that one CUDA lane executes; the warp orchestration (tile gather, ordered emission, reductions)
// lives in t1k_kernels.cuh.  The functions are __host__ __device__ so that tests/host_emu.cpp can run
// the very same code sequentially on the CPU box (test harness only — the product never calls it).
//
// Data layout (HBM):
//   sequences are 2 bits/base, 32 bases per uint64 word, base i at bits [2i,2i+1] of word i>>5;
//   A0 C1 G2 T3, N stored as 3 (what nucToNum&3 gives, KmerCode.hpp:99) with a parallel "n2" plane
//   that holds 01 at N bases (so masks combine with the XOR plane without bit spreading) and, for
//   alleles, an "ex2" plane with 01 at exonic bases (isValidDiff[].exon, SeqSet.hpp:651-671).
//   Every allele starts on a word boundary and is followed by >= 1 zero pad word.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define T1K_HD __host__ __device__ __forceinline__
#define T1K_HDN __host__ __device__
#define T1K_NOINLINE __noinline__
#define T1K_NOUNROLL _Pragma("unroll 1")
#else
#define T1K_HD inline
#define T1K_HDN
#define T1K_NOINLINE
#define T1K_NOUNROLL
#endif

// host-emulation-only call-site counters (tests/host_emu.cpp); compiled out of the device build
#if !defined(__CUDACC__) && defined(T1K_EMU_COUNTERS)
extern long long t1k_emu_counters[128];
#define T1K_COUNT(i, v) (t1k_emu_counters[i] += (v))
#else
#define T1K_COUNT(i, v) ((void)0)
#endif

namespace t1k {

typedef uint64_t u64;
typedef uint32_t u32;
typedef uint16_t u16;
typedef uint8_t u8;

constexpr int KMER = 11;
constexpr int RADIUS = 10;          // SeqSet.hpp:763
constexpr int HIT_LEN_REQ = 31;     // SeqSet.hpp:764
constexpr int BAND = 5;             // AlignAlgo.hpp:215
constexpr int MAX_READ_LEN = 1000;  // == T1K_MAX_READ_LEN (positions are 10 bits in the hit encoding)
constexpr int RWORDS = 9;           // packed words per read strand plane of a batch of reads up to 255 bases (+ fetch slack)
constexpr int MAX_RWORDS = 33;      // ... up to MAX_READ_LEN bases
T1K_HD int read_words(int maxLen) { return maxLen <= 255 ? RWORDS : MAX_RWORDS; }
constexpr int MAX_BAND_W = 64;      // widest DP band kept (band = 11 + |lent-lenp| + 2 sentinels)
constexpr int MAX_EMIT = 48;        // seed overlaps one (strand, allele) group may emit
constexpr u64 M55 = 0x5555555555555555ull;

struct Posting { u32 idx, off; };
// The k-mer index as the device reads it.  The postings of one k-mer (allele, offset), which the reference keeps as one
// list in allele order (KmerIndex.hpp:58-71), are regrouped by TILE (32 consecutive allele ids) and offset: one entry
// says "this k-mer occurs at offset `off` in the alleles tile*32 + {bits of mask}".  Alleles of one gene are neighbours in
// the reference file and share most k-mers at the same offset, so an entry stands for up to 32 postings: the index
// shrinks from 8 B per posting towards 0.5 B, and a warp that sweeps the allele axis tile by tile with one lane per
// allele gets its hits from one broadcast word per seed instead of a posting scatter.
//   more = number of further entries of the same (k-mer, tile) that follow this one (other offsets: a k-mer repeated
//   inside an allele, or alleles of the tile that carry an indel before it), in ascending offset order.
struct alignas(16) KmerEntry { u32 tile, off, mask, more; };
struct KmerInfo { u32 estart, pstart; };   // per k-mer code: first entry, first posting (posting COUNT drives the skip rule, SeqSet.hpp:1109)

struct alignas(16) AlleleMeta { u64 wordOff; int32_t len; u32 hasN; };   // one 16-byte load instead of three dependent-free ones

struct RefView {
  const AlleleMeta *meta;
  const u64 *seq2, *n2, *ex2;
  const u64 *wordOff;    // [nAlleles] first word of allele
  const int32_t *len;    // [nAlleles]
  const u8 *hasN;        // [nAlleles] the allele holds at least one N (separator)
  const u32 *kstart;     // [4^K + 1]   (host emulation / tests; the device reads kinfo + entries)
  const Posting *post;
  const KmerInfo *kinfo; // [4^K + 1]
  const KmerEntry *entries;
  // coverage: range-add difference array + point corrections.  Layout: the 32 alleles of a tile are interleaved position by
  // position (entry of allele a, base p = covOff[a] + 32 * p, covOff[a] = first entry of the tile + (a & 31)), so the lanes of
  // a warp, which hold neighbouring alleles at the same read-relative position, hit neighbouring words: their atomics
  // coalesce into a few sectors instead of 32 sectors kilobytes apart.
  int32_t *covDiff, *covPoint;
  const u64 *covOff;     // [nAlleles]
  int32_t nAlleles;
  double sim;
  int32_t relax;
  // simThr[w * SIM_DEN + den] = smallest matchCnt whose (double)matchCnt / (double)den is NOT below the threshold
  // (w = 0: refSeqSimilarity, w = 1: the 0.95 of SeqSet.hpp:2178), built on the host with the same double division, so
  // `sim < threshold` becomes an integer compare (the inlined double divisions were a tenth of the kernel's code)
  const u16 *simThr;
};
constexpr int SIM_DEN = 4096;
T1K_HDN inline void sim_threshold_table(double sim, u16 *thr) {      // host side: 2 * SIM_DEN entries
  const double lim[2] = {sim, 0.95};
  for (int w = 0; w < 2; ++w)
    for (int den = 0; den < SIM_DEN; ++den) {
      int mc = 0;
      while (mc < 65535 && den > 0 && (double)mc / (double)den < lim[w]) ++mc;
      thr[w * SIM_DEN + den] = (u16)mc;
    }
}
T1K_HDN T1K_NOINLINE inline bool sim_below_slow(double lim, int mc, int den) { return (double)mc / (double)den < lim; }
// (double)mc / (double)den < (w == 0 ? R.sim : 0.95)
T1K_HD bool sim_below(const RefView &R, int mc, int den, int w) {
  if (den > 0 && den < SIM_DEN && mc >= 0) return mc < (int)R.simThr[w * SIM_DEN + den];
  return sim_below_slow(w == 0 ? R.sim : 0.95, mc, den);
}

struct ReadView {        // one strand of one read-end
  const u64 *seq2, *n2;  // read_words(longest read of the batch) words each
  int len;
  bool anyN;             // the read holds at least one N
};

// one allele as the lane code sees it: plane pointers at the allele's first word.  Most alleles (every RNA allele)
// and most reads hold no N, and then the two N-plane fetches of every 32-column comparison are skipped.
struct AlleleView {
  const u64 *seq, *n2, *ex2;
  int len;
  bool hasN;             // allele holds an N
  bool useN;             // allele or read holds an N: the N planes take part in base comparisons
};

// candidate = seed overlap that passed the similarity filter of GetOverlapsFromRead (SeqSet.hpp:1893-1908)
struct alignas(16) Cand {                  // 48 bytes: three 16-byte loads
  int32_t seqIdx, seqStart, seqEnd;
  u16 readStart, readEnd;
  u8 strand01, flags;                      // strand01: 1 = same strand (+1), 0 = reverse (-1)
  u16 matchCnt;
  // filled by the extension stage (SeqSet::ExtendOverlap, SeqSet.hpp:1994-2100)
  int32_t eSeqStart, eSeqEnd;
  u16 eReadStart, eReadEnd, leftClip, rightClip;
  int32_t eMatchCnt, relaxed;
  // CF_FA: the full-read alignment is the pure diagonal with <= 3 mismatches, known already from the seeding stage:
  // read positions of the mismatches (bytes 0-2), their number (bits 24-25) and the exonic ones among them (bits 26-27)
  u32 mmPos;
};
enum { CF_SEP = 1, CF_NEEDCLIP = 2, CF_RET = 4, CF_INCLUDE = 8,
       CF_PRE = 16,       // extension fields already filled by diag_fast (ExtendOverlap need not run)
       CF_FA = 32 };      // full-read alignment already known (mmPos)

// final record kept resident in HBM for pairing: 32 B.  Positions are 16 bits wide (reads of any supported length).
struct Rec {
  int32_t seqIdx, seqStart, seqEnd;
  u16 readStart, readEnd, leftClip, rightClip;
  u32 mcx;               // matchCnt | relaxedMatchCnt << 14 | strand01 << 31
  u64 key;               // list-order key (order_key) | bit 0: the list is in post-extension order (the > 1000 cut ran)
};
T1K_HD u32 rec_mcx(int matchCnt, int relaxed, int strand01) { return (u32)matchCnt | ((u32)relaxed << 14) | ((u32)strand01 << 31); }
T1K_HD int rec_mc(u32 mcx) { return (int)(mcx & 0x3FFFu); }
T1K_HD int rec_relaxed(u32 mcx) { return (int)((mcx >> 14) & 0x3FFFu); }
T1K_HD int rec_strand01(u32 mcx) { return (int)(mcx >> 31); }
// The reference's output order of one read-end's records (`assign`, SeqSet.hpp:2300) from the stored fields.  The store keeps
// a list in allele order and, inside one allele, in (readStart, readEnd, seqStart, seqEnd) order of the SEED overlaps, so
// for a list in pre-extension order `key` and the store position reproduce _overlap::operator< (SeqSet.hpp:103-127) down to
// its last field.  A list the > 1000 cut re-sorted (SeqSet.hpp:2290-2298; bit 0 of key) is ordered on the extended fields,
// which the record itself holds.
T1K_HD bool rec_before(const Rec &a, long long ia, const Rec &b, long long ib) {
  if (a.key != b.key) return a.key < b.key;
  if (!(a.key & 1) || a.seqIdx != b.seqIdx) return ia < ib;
  if (a.readStart != b.readStart) return a.readStart < b.readStart;
  if (a.readEnd != b.readEnd) return a.readEnd < b.readEnd;
  if (a.seqStart != b.seqStart) return a.seqStart < b.seqStart;
  if (a.seqEnd != b.seqEnd) return a.seqEnd < b.seqEnd;
  return ia < ib;
}

// per-lane scratch in global memory, sized for the longest read of the batch:
//   ops    edit string of one alignment (<= lent + lenp + 8 entries)
//   rows   rolling DP rows of dp_align
//   dir    direction nibbles of the band ((lenp + 1) x band width), also the chaining scratch of chain_cluster_general
//   emit   seed overlaps one (strand, allele) group emitted
//   chain  a LIS chain (strictly increasing read offsets)
constexpr int SCR_ROWS = 4 * (MAX_BAND_W + 2) * 4;
constexpr int SCR_EMIT = MAX_EMIT * (int)sizeof(Cand);
T1K_HD int scr_ops(int maxLen) { return (2 * maxLen + 128 + 15) & ~15; }
T1K_HD int scr_dir(int maxLen) { return ((maxLen + 3) * MAX_BAND_W + 15) & ~15; }
T1K_HD int scr_chain(int maxLen) { return (maxLen + 8 + 3) & ~3; }              // entries
T1K_HD size_t scr_bytes(int maxLen) { return (((size_t)scr_ops(maxLen) + SCR_ROWS + scr_dir(maxLen) + SCR_EMIT + 4 * (size_t)scr_chain(maxLen)) + 255) & ~(size_t)255; }

struct LaneScratch {
  u8 *base;
  int opsCap, dirBytes, chainCap;
  T1K_HD u8 *ops() const { return base; }
  T1K_HD int *rows() const { return (int *)(base + opsCap); }
  T1K_HD u8 *dir() const { return base + opsCap + SCR_ROWS; }
  T1K_HD Cand *emit() const { return (Cand *)(base + opsCap + SCR_ROWS + dirBytes); }
  T1K_HD u32 *chain() const { return (u32 *)(base + opsCap + SCR_ROWS + dirBytes + SCR_EMIT); }
};
T1K_HD LaneScratch lane_scratch(u8 *base, int maxLen) {
  LaneScratch S; S.base = base; S.opsCap = scr_ops(maxLen); S.dirBytes = scr_dir(maxLen); S.chainCap = scr_chain(maxLen);
  return S;
}

enum { ERR_BAND = 1, ERR_SCRATCH = 2, ERR_EMIT = 4, ERR_CAND = 8, ERR_STORE = 16, ERR_HITS = 32, ERR_READ_LEN = 64, ERR_READ_CHAR = 128 };

T1K_HD int popc64(u64 x) {
#ifdef __CUDA_ARCH__
  return __popcll(x);
#else
  return __builtin_popcountll(x);
#endif
}
T1K_HD int ctz64(u64 x) {
#ifdef __CUDA_ARCH__
  return __ffsll((long long)x) - 1;
#else
  return __builtin_ctzll(x);
#endif
}
T1K_HD int imin(int a, int b) { return a < b ? a : b; }
T1K_HD int imax(int a, int b) { return a > b ? a : b; }
T1K_HD int iabs(int a) { return a < 0 ? -a : a; }
// DPX (sm_90+): max(a + b, c) and the three-way max as one instruction each (VIADDMNMX / VIMNMX3) — the two shapes the
// affine-gap recurrences of GlobalAlignment consist of.  Plain integer arithmetic on the host (same values).
T1K_HD int addmax(int a, int b, int c) {
#ifdef __CUDA_ARCH__
  return __viaddmax_s32(a, b, c);
#else
  const int s = a + b; return s > c ? s : c;
#endif
}
T1K_HD int max3(int a, int b, int c) {
#ifdef __CUDA_ARCH__
  return __vimax3_s32(a, b, c);
#else
  const int m = a > b ? a : b; return m > c ? m : c;
#endif
}

constexpr int COV_STRIDE = 32;
T1K_HD void cov_add(int32_t *p, int v) {
#ifdef __CUDA_ARCH__
  atomicAdd(p, v);
#else
  *p += v;
#endif
}

// 32 bases starting at base `pos` (>= 0) of a plane
T1K_HD u64 fetch32(const u64 *plane, int pos) {
  const u64 *p = plane + (pos >> 5);
  int sh = (pos & 31) * 2;
  return (p[0] >> sh) | ((p[1] << 1) << (63 - sh));      // branch-free funnel shift (sh in 0..62)
}
T1K_HD u64 fetch32(const u64 *plane, u64 w0, int pos) { return fetch32(plane + w0, pos); }
// 128-bit funnel shifts of a two-word value, 0 < s < 64
T1K_HD u64 shr2w(u64 lo, u64 hi, int s) { return (lo >> s) | (hi << (64 - s)); }
T1K_HD u64 shl2w(u64 hi, u64 lo, int s) { return (hi << s) | (lo >> (64 - s)); }
T1K_HD int base2(const u64 *plane, int pos) { return (int)((plane[pos >> 5] >> ((pos & 31) * 2)) & 3); }
T1K_HD int base2(const u64 *plane, u64 w0, int pos) { return base2(plane + w0, pos); }
T1K_HD u64 lowmask2(int nBases) { return nBases >= 32 ? ~0ull : ((1ull << (2 * nBases)) - 1); }

T1K_HD AlleleView allele_view(const RefView &R, int seqIdx, const ReadView &Q) {
  AlleleView T;
  const u64 w0 = R.wordOff[seqIdx];
  T.seq = R.seq2 + w0; T.n2 = R.n2 + w0; T.ex2 = R.ex2 + w0;
  T.len = R.len[seqIdx];
  T.hasN = R.hasN[seqIdx] != 0;
  T.useN = T.hasN || Q.anyN;
  return T;
}

// AlignAlgo.hpp:304-305: equal, or either side N
T1K_HD bool base_eq(const AlleleView &T, int tpos, const ReadView &Q, int ppos) {
  if (T.useN && (base2(T.n2, tpos) | base2(Q.n2, ppos))) return true;
  return base2(T.seq, tpos) == base2(Q.seq2, ppos);
}

// mismatch plane (01 per mismatching column) of 32 columns starting at (tpos+k, ppos+k)
T1K_HD u64 mm_chunk(const AlleleView &T, int tpos, const ReadView &Q, int ppos, int nLeft) {
  u64 x = fetch32(T.seq, tpos) ^ fetch32(Q.seq2, ppos);
  u64 d = (x | (x >> 1)) & M55;
  if (T.useN) d &= ~(fetch32(T.n2, tpos) | fetch32(Q.n2, ppos));
  return d & lowmask2(nLeft);
}

// mismatching columns among rows [lo, hi] of the window when row r of the read is paired with allele column r + d.
// The caller only asks whether the count stays <= limit: the scan stops at the first chunk that exceeds it (a shifted
// comparison of unrelated sequence mismatches within a few columns, so this is almost always the first chunk).
T1K_HDN T1K_NOINLINE inline int shifted_mm(const AlleleView &T, int tpos, const ReadView &Q, int ppos, int n, int d, int lo, int hi, int limit) {
  if (lo < 0) lo = 0;
  if (lo < -d) lo = -d;
  if (hi > n - 1) hi = n - 1;
  if (hi > n - 1 - d) hi = n - 1 - d;
  int c = 0;
  T1K_NOUNROLL
  for (int k = lo; k <= hi && c <= limit; k += 32) c += popc64(mm_chunk(T, tpos + k + d, Q, ppos + k, hi - k + 1));
  return c;
}

// Equal-length global alignment without the DP (DESIGN.md "Diagonal certificate").  For lent == lenp == n the
// reference's traceback returns the pure diagonal iff no gapped path reaches a diagonal cell (b,b) with a strictly
// higher score than the diagonal prefix (ties go to the diagonal, AlignAlgo.hpp:332-345).  An excursion that leaves
// the diagonal at row A and returns after row B with m gap events of lengths L_g gains
//     4*(mm0[A..B] - mmShifted) - 4*m - 2*sum(L_g)
// over the diagonal (mm0 = diagonal mismatches in the segment, mmShifted = mismatches of its shifted pairs).
//   * m = 2 (one shift d): gain = 4*delta - 8 - 4|d|, positive only if delta >= 3 + |d|;
//   * m >= 3 needs sum(L_g) >= 4: gain <= 4*delta - 20, positive only if delta >= 6.
// So <= 3 mismatches are always diagonal; with 4 or 5 only single-shift excursions with |d| <= mm - 3 can win, and
// those are enumerated exactly: the best segment starts at a mismatch (or up to |d| rows before one when the leading
// rows are the unpaired ones) and ends at a mismatch (or up to |d| rows after one).  More mismatches fall back to
// a shift histogram bound, and failing that to the DP.
// 4 or 5 diagonal mismatches: exact enumeration of the single-shift excursions (see above)
T1K_HDN T1K_NOINLINE inline bool diag_certified_45(const AlleleView &T, int tpos, const ReadView &Q, int ppos, int n, int mm) {
  int pos[5];
  {
    int c = 0;
    T1K_NOUNROLL
    for (int k = 0; k < n; k += 32) {
      u64 m = mm_chunk(T, tpos + k, Q, ppos + k, n - k);
      T1K_NOUNROLL
      while (m) { pos[c < 5 ? c : 4] = k + (ctz64(m) >> 1); ++c; m &= m - 1; }
    }
  }
  T1K_NOUNROLL
  for (int d = 1; d <= mm - 3; ++d) {
    T1K_NOUNROLL
    for (int i = 0; i + 2 + d < mm; ++i)            // at least 3 + d diagonal mismatches inside the segment
      T1K_NOUNROLL
      for (int j = i + 2 + d; j < mm; ++j)
        T1K_NOUNROLL
        for (int t = 0; t <= d; ++t) {
          // deletion first (allele ahead by d): rows [A, B - d] paired with columns r + d, last d rows unpaired
          {
            const int A = pos[i], B = pos[j] + t;
            if (B <= n - 1 && B - d >= A - 1) {
              int in = 0;
              T1K_NOUNROLL
              for (int q = 0; q < mm; ++q) in += pos[q] >= A && pos[q] <= B;
              if (in > 2 + d && in - shifted_mm(T, tpos, Q, ppos, n, d, A, B - d, in - 3 - d) > 2 + d) return false;
            }
          }
          // insertion first (allele behind by d): first d rows unpaired, rows [A + d, B] paired with columns r - d
          {
            const int A = pos[i] - t, B = pos[j];
            if (A >= 0 && B >= A + d - 1) {
              int in = 0;
              T1K_NOUNROLL
              for (int q = 0; q < mm; ++q) in += pos[q] >= A && pos[q] <= B;
              if (in > 2 + d && in - shifted_mm(T, tpos, Q, ppos, n, -d, A + d, B, in - 3 - d) > 2 + d) return false;
            }
          }
        }
  }
  return true;
}

// 6..24 diagonal mismatches: shift-histogram bound (F_d = diagonal mismatches that shift d turns into matches;
// sum_d max(0, F_d - 1) <= 1 leaves every excursion <= 0)
T1K_HDN T1K_NOINLINE inline bool diag_certified_hist(const AlleleView &T, int tpos, const ReadView &Q, int ppos, int n) {
  int F[2 * BAND + 1];
  T1K_NOUNROLL
  for (int d = 0; d <= 2 * BAND; ++d) F[d] = 0;
  T1K_NOUNROLL
  for (int k = 0; k < n; k += 32) {
    u64 m = mm_chunk(T, tpos + k, Q, ppos + k, n - k);
    T1K_NOUNROLL
    while (m) {
      int p = k + (ctz64(m) >> 1);
      m &= m - 1;
      T1K_NOUNROLL
      for (int d = -BAND; d <= BAND; ++d) {
        if (d == 0) continue;
        int q = p + d;
        if (q < 0 || q >= n) continue;
        if (base_eq(T, tpos + q, Q, ppos + p)) ++F[d + BAND];
      }
    }
  }
  int excess = 0;
  T1K_NOUNROLL
  for (int d = 0; d <= 2 * BAND; ++d) excess += F[d] > 1 ? F[d] - 1 : 0;
  return excess <= 1;
}

// Interval certificate (any number of diagonal mismatches <= 24).  A connected excursion from the diagonal over rows
// A..B with k segments at shifts d_1..d_k (k + 1 gap events, U unpaired rows) changes the score by
//     4 * (a_I - nu - |U| - (k + 1)),     a_I = diagonal mismatches in A..B, nu = mismatches of the shifted pairs,
// and a_I <= a_P + |U| (a_P = diagonal mismatches on paired rows), so a gain needs sum_s (a_{P_s} - nu_s) >= k + 2.
// Per segment a_{P_s} - nu_s <= G_d := the largest sum over a row interval of (+1: diagonal mismatch that shift d turns
// into a match, -1: diagonal match that shift d turns into a mismatch).  If G_d <= 1 for every shift of the band the sum
// is at most k: no excursion gains, the traceback stays on the diagonal.  G_d <= 1 iff between every two consecutive
// "+1" rows of shift d lies at least one "-1" row.
T1K_HDN T1K_NOINLINE inline bool diag_certified_interval_n(const AlleleView &T, int tpos, const ReadView &Q, int ppos, int n) {
  T1K_NOUNROLL
  for (int d = -BAND; d <= BAND; ++d) {
    if (d == 0) continue;
    const int lo = d < 0 ? -d : 0, hi = d > 0 ? n - 1 - d : n - 1;      // rows that have a column at this shift
    bool pending = false;                                               // a "+1" row with no "-1" row after it yet
    T1K_NOUNROLL
    for (int r = lo; r <= hi; r += 32) {
      const int cnt = hi - r + 1;
      const u64 a = mm_chunk(T, tpos + r, Q, ppos + r, cnt);            // diagonal mismatches of rows r .. r+31
      const u64 sft = mm_chunk(T, tpos + r + d, Q, ppos + r, cnt);      // mismatches of the same rows at shift d
      u64 plus = a & ~sft;
      const u64 minus = ~a & sft;
      u64 done = 0;                                                     // bits at and below the last "+1" row handled
      T1K_NOUNROLL
      while (plus) {
        const u64 b = plus & (~plus + 1);                               // lowest "+1" row
        const u64 below = b - 1;
        if (pending && (minus & below & ~done) == 0) return false;      // two "+1" rows with no "-1" row between: G_d >= 2
        pending = true;
        done = below | b;
        plus &= plus - 1;
      }
      if (minus & ~done) pending = false;
    }
  }
  return true;
}
// The same test for windows without N (the usual case), 32 rows at a time for all ten shifts at once: the allele bases
// tpos + r - 5 .. tpos + r + 36 are fetched once per 32 rows and every shift is a funnel shift of those two words.
T1K_HDN T1K_NOINLINE inline bool diag_certified_interval(const AlleleView &T, int tpos, const ReadView &Q, int ppos, int n) {
  if (T.useN) return diag_certified_interval_n(T, tpos, Q, ppos, n);
  u32 pendingBits = 0;                                                  // bit d + BAND: shift d has a "+1" row with no "-1" row after it yet
  T1K_NOUNROLL
  for (int r0 = 0; r0 < n; r0 += 32) {
    const int cnt = n - r0 < 32 ? n - r0 : 32;
    const u64 q = fetch32(Q.seq2, ppos + r0);
    const u64 tl = fetch32(T.seq, tpos + r0 - BAND), th = fetch32(T.seq, tpos + r0 - BAND + 32);
    const u64 rows = lowmask2(cnt) & M55;
    u64 a;
    { const u64 x = shr2w(tl, th, 2 * BAND) ^ q; a = (x | (x >> 1)) & rows; }
    T1K_NOUNROLL
    for (int d = -BAND; d <= BAND; ++d) {
      if (d == 0) continue;
      const u64 t = d == -BAND ? tl : shr2w(tl, th, 2 * (d + BAND));
      const u64 x = t ^ q;
      // rows of this chunk that have a column at shift d: max(0, -d) <= r <= min(n - 1, n - 1 - d)
      u64 vm = rows;
      if (r0 < BAND || r0 + 32 + BAND > n) {          // only the first and the last rows lack a column at some shift
        const int lo = (d < 0 ? -d : 0) - r0, hi = (d > 0 ? n - 1 - d : n - 1) - r0 + 1;
        vm = (hi <= 0 ? 0ull : hi >= 32 ? ~0ull : ((1ull << (2 * hi)) - 1)) & ~(lo <= 0 ? 0ull : lo >= 32 ? ~0ull : ((1ull << (2 * lo)) - 1)) & rows;
      }
      const u64 sft = (x | (x >> 1)) & vm;
      u64 plus = a & ~sft & vm;
      const u64 minus = ~a & sft;
      bool pending = (pendingBits >> (d + BAND)) & 1u;
      u64 done = 0;
      T1K_NOUNROLL
      while (plus) {
        const u64 b = plus & (~plus + 1);
        const u64 below = b - 1;
        if (pending && (minus & below & ~done) == 0) return false;
        pending = true;
        done = below | b;
        plus &= plus - 1;
      }
      if (minus & ~done) pending = false;
      pendingBits = (pendingBits & ~(1u << (d + BAND))) | ((u32)pending << (d + BAND));
    }
  }
  return true;
}

T1K_HD bool diag_certified(const AlleleView &T, int tpos, const ReadView &Q, int ppos, int n, int &mmOut) {
  int mm = 0;
  if (n <= 32) mm = popc64(mm_chunk(T, tpos, Q, ppos, n));     // the gap between two seed hits: one word
  else {
    T1K_NOUNROLL
    for (int k = 0; k < n; k += 32) mm += popc64(mm_chunk(T, tpos + k, Q, ppos + k, n - k));
  }
  mmOut = mm;
  if (mm <= 3) return true;
  // 4-5: the cheap sufficient tests first, the exact enumeration only when they cannot tell
  if (mm <= 5) return diag_certified_interval(T, tpos, Q, ppos, n) || diag_certified_45(T, tpos, Q, ppos, n, mm);
  if (mm > 24) return false;
  return diag_certified_interval(T, tpos, Q, ppos, n) || diag_certified_hist(T, tpos, Q, ppos, n);
}

// The equal-length case of dp_align below (99.9 % of the calls): band 5 on both sides, 13 window columns, the two rolling
// rows live in registers (fully unrolled column loop) and one 64-bit word per row holds the 13 direction nibbles.  Cell for
// cell the same arithmetic, sentinels and tie rules as dp_align; ~5x fewer instructions than the memory-resident rows.
T1K_HDN T1K_NOINLINE inline int dp_align_eq(const AlleleView &T, int tpos, const ReadView &Q, int ppos, int n, const LaneScratch &S, int &err) {
  u8 *ops = S.ops();
  constexpr int LB = BAND, WW = 2 * BAND + 3;       // columns i-LB-1 .. i+LB+1
  if (2 * n + 8 > S.opsCap || (n + 1) * 8 > S.dirBytes) { err |= ERR_BAND; return -1; }
  const int negInf = (n + 1) * (n + 1) * -4;
  const int stale = -4 + (n + 1) * -4;              // e[0][j], AlignAlgo.hpp:268 (Q5)
  u64 *dirRow = (u64 *)S.dir();                     // [n + 1] direction nibbles of row i, column window index jj at bits 4*jj
  int mP[WW + 1], eP[WW + 1];
#pragma unroll
  for (int jj = 0; jj < WW; ++jj) {
    const int j = jj - LB - 1;
    if (j < 0 || j > n) { mP[jj] = negInf; eP[jj] = negInf; }
    else if (j == 0) { mP[jj] = 0; eP[jj] = 0; }
    else { mP[jj] = -4 - 4 * j; eP[jj] = stale; }
  }
  mP[WW] = negInf; eP[WW] = negInf;
  T1K_NOUNROLL
  for (int i = 1; i <= n; ++i) {
    const int start = i - LB < 1 ? 1 : i - LB;
    const int pb = base2(Q.seq2, ppos + i - 1);
    const bool pn = T.useN && base2(Q.n2, ppos + i - 1);
    const int tb = tpos + start - 1;               // allele base of column `start`
    const u64 ts = fetch32(T.seq, tb);
    const u64 tn = T.useN ? fetch32(T.n2, tb) : 0;
    int mC[WW + 1], eC[WW + 1];
    u64 bitsRow = 0;
    int fPrev, mLeft;
    {                                               // jj = 0: column i-LB-1, never inside the band
      const int j = i - LB - 1;
      int mv = negInf, ev = negInf, fv = negInf;
      if (j == 0) { mv = -4 - 4 * i; ev = -4 - i; fv = -4 - 4 * i; }
      mC[0] = mv; eC[0] = ev; fPrev = fv; mLeft = mv;
    }
#pragma unroll
    for (int jj = 1; jj < WW - 1; ++jj) {
      const int j = i - LB - 1 + jj;
      int mv, ev, fv;
      if (j < 0 || j > n) { mv = ev = fv = negInf; }
      else if (j == 0) { mv = -4 - 4 * i; ev = -4 - i; fv = -4 - 4 * i; }
      else {
        const int mUp = mP[jj + 1], eUp = eP[jj + 1], mDiag = mP[jj];
        const int e2 = mUp - 5, f2 = mLeft - 5;
        ev = addmax(eUp, -1, e2);
        fv = addmax(fPrev, -1, f2);
        const int sh = (j - start) * 2;
        const bool eq = pn || ((tn >> sh) & 3) || (int)((ts >> sh) & 3) == pb;
        const int dv = mDiag + (eq ? 2 : -2);
        mv = max3(dv, ev, fv);
        const u64 bits = (u64)((dv == mv ? 1 : 0) | (fv >= ev ? 2 : 0) | (e2 == ev ? 4 : 0) | (f2 == fv ? 8 : 0));
        bitsRow |= bits << (4 * jj);
      }
      mC[jj] = mv; eC[jj] = ev;
      fPrev = fv; mLeft = mv;
    }
    mC[WW - 1] = negInf; eC[WW - 1] = negInf;       // jj = WW-1: column i+LB+1, outside the band (and never column 0)
    mC[WW] = negInf; eC[WW] = negInf;
    dirRow[i] = bitsRow;
#pragma unroll
    for (int jj = 0; jj <= WW; ++jj) { mP[jj] = mC[jj]; eP[jj] = eC[jj]; }
  }
  // traceback (AlignAlgo.hpp:323-408); boundary rows/columns by their closed forms
  int ti = n, tj = n, mat = 0, k = 0;
  T1K_NOUNROLL
  while (ti > 0 || tj > 0) {
    if (k >= S.opsCap - 2) { err |= ERR_BAND; return -1; }
    const int b = (ti > 0 && tj > 0) ? (int)((dirRow[ti] >> (4 * (tj - ti + LB + 1))) & 15) : 0;
    if (mat == 0) {
      int a;
      if (ti > 0 && tj > 0) {
        if (b & 1) a = base_eq(T, tpos + tj - 1, Q, ppos + ti - 1) ? 0 : 1;
        else a = (b & 2) ? 3 : 2;
      } else if (ti == 0) a = (-4 - tj >= stale) ? 3 : 2;
      else a = 2;                      // tj == 0, ti > 0: f = -4-4ti < e = -4-ti
      if (a <= 1) { ops[k++] = (u8)a; --ti; --tj; }
      else mat = a == 2 ? 1 : 2;
    } else if (mat == 1) {
      ops[k++] = 2;
      if (ti > 0) {
        const bool fromM = tj == 0 ? ti == 1 : (b & 4) != 0;
        --ti; mat = fromM ? 0 : 1;
      } else mat = 2;
    } else {
      ops[k++] = 3;
      if (tj > 0) {
        const bool fromM = ti == 0 ? tj == 1 : (b & 8) != 0;
        --tj; mat = fromM ? 0 : 2;
      } else mat = 1;
    }
  }
  T1K_NOUNROLL
  for (int a = 0, b = k - 1; a < b; ++a, --b) { u8 t = ops[a]; ops[a] = ops[b]; ops[b] = t; }
  return k;
}

// AlignAlgo::GlobalAlignment (AlignAlgo.hpp:215-421), one lane, band-only storage:
// two rolling rows of (m,e) and one direction nibble per band cell
//   bit0 diagonal predecessor reproduces m, bit1 f >= e, bit2 e opened from m, bit3 f opened from m.
// Writes the edit ops in forward order to S.ops() (0 M,1 X,2 I,3 D) and returns their count (<0: error).
T1K_HDN T1K_NOINLINE inline int dp_align(const AlleleView &T, int tpos, int lent, const ReadView &Q, int ppos, int lenp,
                            const LaneScratch &S, int &err) {
  u8 *ops = S.ops();
  if (lent == 0 || lenp == 0) return 0;
  if (lent == 1 && lenp == 1) { ops[0] = base_eq(T, tpos, Q, ppos) ? 0 : 1; return 1; }
#ifndef T1K_NO_DP_EQ
  if (lent == lenp) { T1K_COUNT(0, 1); T1K_COUNT(1, (long long)lenp * (2 * BAND + 3)); T1K_COUNT(2, 1); return dp_align_eq(T, tpos, Q, ppos, lent, S, err); }
#endif
  int lb = BAND, rb = BAND;
  if (lent > lenp) rb += lent - lenp; else if (lent < lenp) lb += lenp - lent;
  const int W = lb + rb + 3;          // columns i-lb-1 .. i+rb+1
  T1K_COUNT(0, 1); T1K_COUNT(1, (long long)lenp * W); T1K_COUNT(lent == lenp ? 2 : 3, 1);
  if (W > MAX_BAND_W || (lenp + 1) * W > S.dirBytes || lent + lenp + 8 > S.opsCap) { err |= ERR_BAND; return -1; }
  const int negInf = (lent + 1) * (lenp + 1) * -4;
  const int stale = -4 + (lenp + 1) * -4;   // e[0][j], AlignAlgo.hpp:268 (Q5)
  int *mP = S.rows(), *eP = mP + (MAX_BAND_W + 2), *mC = eP + (MAX_BAND_W + 2), *eC = mC + (MAX_BAND_W + 2);
  u8 *dir = S.dir();
  // row 0 window: columns -lb-1 .. rb+1
  T1K_NOUNROLL
  for (int jj = 0; jj < W; ++jj) {
    int j = jj - lb - 1;
    if (j < 0 || j > lent) { mP[jj] = negInf; eP[jj] = negInf; }
    else if (j == 0) { mP[jj] = 0; eP[jj] = 0; }
    else { mP[jj] = -4 - 4 * j; eP[jj] = stale; }
  }
  T1K_NOUNROLL
  for (int i = 1; i <= lenp; ++i) {
    int start = i - lb < 1 ? 1 : i - lb;
    int end = i + rb > lent ? lent : i + rb;
    int pb = base2(Q.seq2, ppos + i - 1), pn = T.useN ? base2(Q.n2, ppos + i - 1) : 0;
    int fPrev = negInf, mLeft = negInf;        // f and m of column j-1 in this row
    u8 *drow = dir + (size_t)i * W;
    // allele columns start..end of this row (at most W - 2 <= 62 of them) as two 32-base words: no loads per cell
    const int tb = tpos + start - 1;
    const u64 ts0 = fetch32(T.seq, tb), ts1 = fetch32(T.seq, tb + 32);
    u64 tn0 = 0, tn1 = 0;
    if (T.useN) { tn0 = fetch32(T.n2, tb); tn1 = fetch32(T.n2, tb + 32); }
    T1K_NOUNROLL
    for (int jj = 0; jj < W; ++jj) {
      int j = i - lb - 1 + jj;
      int mv, ev, fv;
      u8 bits = 0;
      if (j < 0 || j > lent) { mv = ev = fv = negInf; }
      else if (j == 0) { mv = -4 - 4 * i; ev = -4 - i; fv = -4 - 4 * i; }
      else if (j < start || j > end) { mv = ev = fv = negInf; }
      else {
        // previous row window starts one column earlier: column j is at jj+1, column j-1 at jj
        int mUp = mP[jj + 1], eUp = eP[jj + 1], mDiag = mP[jj];
        const int e2 = mUp - 5, f2 = mLeft - 5;
        ev = addmax(eUp, -1, e2);
        fv = addmax(fPrev, -1, f2);
        const int q = j - start, sh = (q & 31) * 2;
        const int tbase = (int)(((q < 32 ? ts0 : ts1) >> sh) & 3), tnn = (int)(((q < 32 ? tn0 : tn1) >> sh) & 3);
        bool eq = pn || tnn || tbase == pb;
        int dv = mDiag + (eq ? 2 : -2);
        mv = max3(dv, ev, fv);
        bits = (u8)((dv == mv ? 1 : 0) | (fv >= ev ? 2 : 0) | (e2 == ev ? 4 : 0) | (f2 == fv ? 8 : 0));
      }
      mC[jj] = mv; eC[jj] = ev;
      drow[jj] = bits;
      fPrev = fv; mLeft = mv;
    }
    // the window of row i must be readable at jj+1 by row i+1
    mC[W] = negInf; eC[W] = negInf;
    int *t = mP; mP = mC; mC = t;
    t = eP; eP = eC; eC = t;
  }
  // traceback (AlignAlgo.hpp:323-408); boundary rows/columns by their closed forms
  int ti = lenp, tj = lent, mat = 0, n = 0;
  T1K_NOUNROLL
  while (ti > 0 || tj > 0) {
    if (n >= S.opsCap - 2) { err |= ERR_BAND; return -1; }
    if (mat == 0) {
      int a;
      if (ti > 0 && tj > 0) {
        u8 b = dir[(size_t)ti * W + (tj - (ti - lb - 1))];
        if (b & 1) a = base_eq(T, tpos + tj - 1, Q, ppos + ti - 1) ? 0 : 1;
        else a = (b & 2) ? 3 : 2;
      } else if (ti == 0) a = (-4 - tj >= stale) ? 3 : 2;
      else a = 2;                      // tj == 0, ti > 0: f = -4-4ti < e = -4-ti
      if (a <= 1) { ops[n++] = (u8)a; --ti; --tj; }
      else mat = a == 2 ? 1 : 2;
    } else if (mat == 1) {
      ops[n++] = 2;
      if (ti > 0) {
        bool fromM;
        if (tj == 0) fromM = ti == 1;
        else fromM = (dir[(size_t)ti * W + (tj - (ti - lb - 1))] & 4) != 0;
        --ti; mat = fromM ? 0 : 1;
      } else mat = 2;
    } else {
      ops[n++] = 3;
      if (tj > 0) {
        bool fromM;
        if (ti == 0) fromM = tj == 1;
        else fromM = (dir[(size_t)ti * W + (tj - (ti - lb - 1))] & 8) != 0;
        --tj; mat = fromM ? 0 : 2;
      } else mat = 1;
    }
  }
  T1K_NOUNROLL
  for (int a = 0, b = n - 1; a < b; ++a, --b) { u8 t = ops[a]; ops[a] = ops[b]; ops[b] = t; }
  return n;
}

// number of EDIT_MATCH columns of GlobalAlignment(t, lent, p, lenp)  (GetAlignStats, SeqSet.hpp:438-455).
// Hot/cold split: nearly every call is an equal-length stretch of <= 32 columns with <= 3 mismatches (the gap between two
// seed hits, or a read overhang) and is answered by one XOR + popcount; everything else lives in the cold function so
// that the hot instruction footprint stays small (the kernel is instruction-cache sensitive).
T1K_HDN T1K_NOINLINE inline int align_matches_cold(const AlleleView &T, int tpos, int lent, const ReadView &Q, int ppos, int lenp,
                                      const LaneScratch &S, int &err) {
  if (lent == lenp) {
    int mm;
    if (diag_certified(T, tpos, Q, ppos, lent, mm)) return lent - mm;
  }
  T1K_COUNT(5, 1);
#if !defined(__CUDACC__) && defined(T1K_EMU_COUNTERS)
  if (lent == lenp) { int mmx; diag_certified(T, tpos, Q, ppos, lent, mmx); T1K_COUNT(24 + (mmx >= 11 ? 7 : mmx < 4 ? 0 : mmx - 4), 1); }
#endif
  int n = dp_align(T, tpos, lent, Q, ppos, lenp, S, err);
  int c = 0;
  const u8 *ops = S.ops();
  T1K_NOUNROLL
  for (int i = 0; i < n; ++i) c += ops[i] == 0;
  return c;
}
// the hot part alone: < 0 when the stretch needs the cold path
T1K_HDN T1K_NOINLINE inline int align_matches_hot(const AlleleView &T, int tpos, int lent, const ReadView &Q, int ppos, int lenp) {
  if (lent == 0 || lenp == 0) return 0;
  if (lent == lenp && lent <= 32) {
    const int mm = popc64(mm_chunk(T, tpos, Q, ppos, lent));
    if (mm <= 3) return lent - mm;
  }
  return -1;
}
T1K_HD int align_matches(const AlleleView &T, int tpos, int lent, const ReadView &Q, int ppos, int lenp, const LaneScratch &S, int &err) {
  T1K_COUNT(4, 1);
  const int m = align_matches_hot(T, tpos, lent, Q, ppos, lenp);
  return m >= 0 ? m : align_matches_cold(T, tpos, lent, Q, ppos, lenp, S, err);
}

// SeqSet::AddOverlapAlignmentInfo (SeqSet.hpp:2657-2680): the edit string of GlobalAlignment(allele[tpos, tpos + lent),
// read strand[ppos, ppos + lenp)) in S.ops() (0 M, 1 X, 2 I, 3 D), returns its length (< 0: error).  An equal-length pair whose
// diagonal is certified (diag_certified: the reference's traceback provably stays on the diagonal) is written straight from
// the mismatch plane, eight columns per store; everything else runs the band DP.  noDiag: always the DP (A/B switch).
T1K_HDN T1K_NOINLINE inline int align_info(const AlleleView &T, int tpos, int lent, const ReadView &Q, int ppos, int lenp, const LaneScratch &S,
                                           int &err, bool noDiag, bool &ranDp) {
  ranDp = false;
  if (lent == lenp && lent > 1 && !noDiag && lent + 16 <= S.opsCap) {
    int mm;
    if (diag_certified(T, tpos, Q, ppos, lent, mm)) {
      u64 *w = (u64 *)S.ops();
      T1K_NOUNROLL
      for (int k = 0; k < lent; k += 32) {
        const u64 d = mm_chunk(T, tpos + k, Q, ppos + k, lent - k);
        T1K_NOUNROLL
        for (int q = 0; q < 4 && k + 8 * q < lent; ++q) {
          const u32 x = (u32)(d >> (16 * q)) & 0xFFFFu;
          u64 v = 0;
#pragma unroll
          for (int c = 0; c < 8; ++c) v |= (u64)((x >> (2 * c)) & 1u) << (8 * c);
          w[(k >> 3) + q] = v;
        }
      }
      return lent;
    }
  }
  ranDp = lent > 0 && lenp > 0 && !(lent == 1 && lenp == 1);
  return dp_align(T, tpos, lent, Q, ppos, lenp, S, err);
}

// any N inside [s, e] of an N plane that starts at the allele's first word
T1K_HD bool n_in_range(const u64 *n2, int s, int e) {
  T1K_NOUNROLL
  for (int k = s; k <= e; k += 32)
    if (fetch32(n2, k) & lowmask2(e - k + 1)) return true;
  return false;
}
T1K_HD bool n_in_range(const RefView &R, u64 w0, int s, int e) { return n_in_range(R.n2 + w0, s, e); }
// IsSeparatorInRange (SeqSet.hpp:487-498): separators are -1, every N, and len
T1K_HD bool sep_in_range(const AlleleView &T, int s, int e) {
  if (s > e) return false;
  if (s <= -1 || e >= T.len) return true;
  return T.hasN && n_in_range(T.n2, s, e);
}

// IsOverlapLowComplex (SeqSet.hpp:458-485)
T1K_HDN T1K_NOINLINE inline bool low_complex(const ReadView &Q, int s, int e) {
  int cnt[4] = {0, 0, 0, 0};
  int n = e - s + 1;
  T1K_NOUNROLL
  for (int k = 0; k < n; k += 32) {
    u64 w = fetch32(Q.seq2, s + k), nm = Q.anyN ? fetch32(Q.n2, s + k) : 0;
    u64 keep = ~nm & M55 & lowmask2(n - k);
    u64 lo = w & M55, hi = (w >> 1) & M55;
    cnt[0] += popc64(~lo & ~hi & keep);
    cnt[1] += popc64(lo & ~hi & keep);
    cnt[2] += popc64(~lo & hi & keep);
    cnt[3] += popc64(lo & hi & keep);
  }
  int low = 0, lowTotal = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) if (cnt[i] <= 2) { ++low; lowTotal += cnt[i]; }
  if (lowTotal * 7 >= n) return false;
  return low >= 2;
}

// hit encoding: readOffset (10 bits) | seqOffset << 10 (alleles up to 4 Mbases)
constexpr int HIT_SHIFT = 10;
T1K_HD u32 hit_make(int a, u32 b) { return (u32)a | (b << HIT_SHIFT); }
T1K_HD int hit_a(u32 h) { return (int)(h & ((1u << HIT_SHIFT) - 1)); }
T1K_HD int hit_b(u32 h) { return (int)(h >> HIT_SHIFT); }
T1K_HD bool hit_diag_less(u32 x, u32 y) {   // CompSortHitCoordDiff (SeqSet.hpp:266-274): (a-b, b, a)
  int cx = hit_a(x) - hit_b(x), cy = hit_a(y) - hit_b(y);
  if (cx != cy) return cx < cy;
  return x < y;                              // numeric order of the encoding is (b, a)
}

// strand-selection key (maximise): the seed overlap that is smallest under _overlap::operator< with
// similarity 0 (SeqSet.hpp:1619-1627) decides the strand.
T1K_HD u64 strand_key(int matchCnt, int span, int seqIdx, int strand01) {
  return ((u64)matchCnt << 40) | ((u64)span << 32) | ((u64)(0xFFFFFFu - (u32)seqIdx) << 1) | (u64)(strand01 ? 0 : 1);
}

// list-order key, ascending = _overlap::operator< order (SeqSet.hpp:103-127): matchCnt desc, similarity desc (== denominator
// asc for equal matchCnt), read span desc.  The remaining fields (seqIdx, strand, readStart, readEnd, seqStart, seqEnd) are
// the order the candidates are EMITTED in — allele tiles ascending, one strand per range, and sort_emitted below inside an
// allele — so ties on the key fall to the candidate index.  Bit 0 stays free (Rec::key).
T1K_HD u64 order_key(int matchCnt, int denom, int span) {
  return ((u64)(16383 - matchCnt) << 50) | ((u64)denom << 35) | ((u64)(16383 - span) << 21);
}
T1K_HD int order_key_mc(u64 key) { return 16383 - (int)(key >> 50); }

// ---- chain consumer: GetOverlapsFromHits tail (SeqSet.hpp:1500-1550) + the matchCnt recomputation of
// GetOverlapsFromRead (SeqSet.hpp:1697-1845).  `C` yields the LIS chain as encoded hits; read and allele offsets
// both increase strictly along it.  OneDiag: every hit lies on one diagonal (the common case).
//   GetTotalHitLengthOnRead/Seq (SeqSet.hpp:1032-1069): runs of hits whose k-mers touch contribute last - first + k,
//   i.e. k for the first hit and min(step, k-or-step) for every later one -> one pass, one load per hit.
template <bool OneDiag, class Chain>
T1K_HDN T1K_NOINLINE inline void consume_chain(const RefView &R, const ReadView &Q, int strand01, int seqIdx, const Chain &C, int sz,
                                  const LaneScratch &S, int &nEmit, u64 &bestStrandKey, int &err) {
  if (sz * KMER < HIT_LEN_REQ) return;
  const u32 h0 = C(0);
  const int rs = hit_a(h0), ss = hit_b(h0);
  int pa = rs, pb = ss, hitLen = KMER, seqLenCov = KMER, nGap = 0;
  T1K_NOUNROLL
  for (int j = 1; j < sz; ++j) {
    const u32 h = C(j);
    const int a = hit_a(h), b = hit_b(h);
    const bool aOv = a <= pa + KMER - 1;
    hitLen += aOv ? a - pa : KMER;
    if (!OneDiag) { const bool bOv = b <= pb + KMER - 1; seqLenCov += bOv ? b - pb : KMER; nGap += !(aOv && bOv); }
    else nGap += !aOv;
    pa = a; pb = b;
  }
  if (OneDiag) seqLenCov = hitLen;
  if (hitLen < HIT_LEN_REQ || seqLenCov < HIT_LEN_REQ) return;
  const int re = pa + KMER - 1, se = pb + KMER - 1;
  u64 sk = strand_key(2 * hitLen, re - rs, seqIdx, strand01);
  if (sk > bestStrandKey) bestStrandKey = sk;
  int mc;
  if (OneDiag) {
    // every overlapping step adds 2*(a - pa), every gap 2k + 2*matches(gap): mc = 2*hitLen + 2*sum(matches).  The walk
    // stops at each gap, so that the lanes of a warp reach their (divergent, expensive) gap comparisons together.
    mc = 2 * hitLen;
    if (nGap > 0) {
      const AlleleView T = allele_view(R, seqIdx, Q);
      const int dg = ss - rs;
      int j = 1;
      pa = rs;
      T1K_NOUNROLL
      for (;;) {
        int a = 0;
        T1K_NOUNROLL
        while (j < sz) {
          a = hit_a(C(j));
          if (a > pa + KMER - 1) break;
          pa = a; ++j;
        }
        if (j >= sz) break;
        const int g = a - (pa + KMER);
        mc += 2 * align_matches(T, pa + dg + KMER, g, Q, pa + KMER, g, S, err);
        pa = a; ++j;
      }
    }
  } else {
    const AlleleView T = allele_view(R, seqIdx, Q);
    mc = 2 * KMER;
    pa = rs; pb = ss;
    T1K_NOUNROLL
    for (int j = 1; j < sz; ++j) {
      const u32 h = C(j);
      const int a = hit_a(h), b = hit_b(h);
      const bool aOv = pa + KMER - 1 >= a, bOv = pb + KMER - 1 >= b;
      if (pb - pa == b - a) {
        if (aOv) mc += 2 * (a - pa);
        else mc += 2 * KMER + 2 * align_matches(T, pb + KMER, b - (pb + KMER), Q, pa + KMER, a - (pa + KMER), S, err);
      } else if (aOv && !bOv) mc += 2 * (a - pa);
      else if (!aOv && bOv) mc += 2 * (b - pb);
      else if (aOv && bOv) mc += 2 * imin(a - pa, b - pb);
      else mc += 2 * KMER + 2 * align_matches(T, pb + KMER, b - (pb + KMER), Q, pa + KMER, a - (pa + KMER), S, err);
      pa = a; pb = b;
    }
  }
  bool below = sim_below(R, mc, se - ss + 1 + re - rs + 1, 0);
  if (!below && low_complex(Q, rs, re)) below = 0.0 < R.sim;        // similarity forced to 0 (SeqSet.hpp:1840-1842)
  if (below) return;
  if (nEmit >= MAX_EMIT) { err |= ERR_EMIT; return; }
  Cand &c = S.emit()[nEmit++];
  c.seqIdx = seqIdx; c.seqStart = ss; c.seqEnd = se;
  c.readStart = (u16)rs; c.readEnd = (u16)re; c.strand01 = (u8)strand01; c.flags = 0;
  c.matchCnt = (u16)mc; c.mmPos = 0;
}

// ---------------------------------------------------------------------------------------------------
// Seed table of one read strand, one u32 per read position a:
//   bits 0-7   number of seeds <= a            bits 8-15  number of seeds s <= a whose distance to the previous seed is > k-1
//   bits 16-23 first seed >= a (255: none)     bits 24-31 last seed <= a (255: none)
// "Seed" = a k-mer position GetHitsFromRead actually looks up with a non-empty posting list (SeqSet.hpp:1093-1153 after
// the skip rule).  Built once per strand (one lane); diag_fast below answers its range questions from it.
T1K_HDN inline void seed_table_build(const u16 *seedA, int nS, int len, u32 *stab) {
  int k = 0, cnt = 0, big = 0, last = 255;
  T1K_NOUNROLL
  for (int a = 0; a < len; ++a) {
    if (k < nS && seedA[k] == a) {
      if (last != 255 && a - last > KMER - 1) ++big;
      ++cnt; last = a; ++k;
    }
    stab[a] = (u32)cnt | ((u32)big << 8) | ((u32)last << 24);
  }
  int nxt = 255;
  T1K_NOUNROLL
  for (int a = len - 1; a >= 0; --a) {
    if ((stab[a] >> 24) == (u32)a) nxt = a;
    stab[a] |= (u32)nxt << 16;
  }
}
// a k-mer of one repeated base: the only k-mers the index build may drop although the allele holds them (KmerIndex.hpp:107-130, Q1)
T1K_HD bool kmer_homopolymer(u32 code) { return code == 0u || code == 0x155555u || code == 0x2AAAAAu || code == 0x3FFFFFu; }

constexpr int FAST_MAX_LEN = 160;    // read length the fast path handles (5 words of 32 bases)

// ---- The common case of GetOverlapsFromHits + GetOverlapsFromRead + ExtendOverlap + the full-read alignment in one
// pass over the read-vs-allele mismatch mask of ONE diagonal, without walking the hit list.
//
// Claim: let d = seqOffset - readOffset of the allele's first hit, [pLo, pHi) the read positions that fall inside the
// allele on that diagonal and suppose the window holds no N on either side.  A seed at a (pLo <= a, a + k <= pHi) whose k
// bases all match is a hit (a, a + d) of this allele (the index holds every non-homopolymer k-mer of every allele, and the
// strand is only eligible when no seed is a homopolymer).  If the NUMBER of such seeds equals the number n of postings the
// gather counted for this allele, the allele's hit list is exactly that set: one diagonal, one cluster, the LIS keeps
// everything (SeqSet.hpp:1303-1553) and the chain statistics follow from the mismatch positions alone:
//   * consecutive seeds inside a mismatch-free stretch are "touching" iff their distance is <= k-1 (checked through the
//     table's big-gap counter; a stretch with a wider step falls back), so the touching runs are the stretches that hold
//     at least one seed: hitLen = sum(last - first + k);
//   * the gap between two runs holds the mismatches in between; with <= 32 columns and <= 3 mismatches GlobalAlignment
//     of the gap is the pure diagonal (DESIGN.md "diagonal certificate"): matches = columns - mismatches.
//     A longer or dirtier gap runs the same align_matches_cold as the hit-list walk would.
// Everything else (count mismatch = hits on other diagonals, N in the window, a seed step > k-1 inside a stretch, reads
// > 160 bases) returns false with nothing written and the caller runs chain_allele on the gathered hit list.
// The same mismatch positions give ExtendOverlap (both overhangs lie on the diagonal) and the full-read alignment.
// lcMemo: per-lane memo of IsOverlapLowComplex for the last (readStart, readEnd) of this strand (0 = empty).
// n: hits of the allele; d: seqOffset - readOffset of its FIRST hit (smallest readOffset, then smallest seqOffset);
// onDiag / far: how many of the n hits lie on diagonal d / more than RADIUS diagonals away from it (the tile sweep counts
// them while it streams the index entries; only consulted when one or two postings are not on the diagonal).
// hot: the caller runs this for all 32 alleles of a tile at once; work that only a few alleles need and that would stall
// the warp — the GlobalAlignment of a long or dirty gap, a long or dirty overhang for ExtendOverlap — is not done: the
// function returns DF_DEFER (nothing written, the hit-count certificate already passed) and the caller queues the allele
// and runs it again with hot = false together with 31 other deferred alleles.
enum { DF_DECLINED = 0, DF_DONE = 1, DF_DEFER = 2 };
T1K_HDN T1K_NOINLINE inline int diag_fast(const RefView &R, const ReadView &Q, int strand01, int seqIdx, int n, int d, int onDiag, int far,
                              const u32 *stab, bool hot, Cand &out, bool &emitted, u64 &bestStrandKey, u32 &lcMemo, const LaneScratch &S, int &err) {
  emitted = false;
  bool needCold = false;
  u32 coldGap[4]; int nColdGap = 0;      // full mode: the first dirty gaps (read position | length << 16)
  const int len = Q.len;
#ifdef __CUDA_ARCH__
  const uint4 mt = *reinterpret_cast<const uint4 *>(R.meta + seqIdx);
  const u64 w0 = (u64)mt.x | ((u64)mt.y << 32);
  const int clen = (int)mt.z;
  const bool alleleHasN = mt.w != 0;
#else
  const u64 w0 = R.meta[seqIdx].wordOff;
  const int clen = R.meta[seqIdx].len;
  const bool alleleHasN = R.meta[seqIdx].hasN != 0;
#endif
  const int pLo = d < 0 ? -d : 0, pHi = imin(len, clen - d);
  const int W = pHi - pLo;
  if (W < KMER) return DF_DECLINED;
  // All allele words of the window in one round trip (the loads are independent; the window spans <= 5 chunks of 32
  // bases = <= 6 words, and the pad word after the allele makes word nW readable), then the mismatch masks of the chunks.
  const u64 *tp = R.seq2 + w0 + ((pLo + d) >> 5);
  const int sh = ((pLo + d) & 31) * 2;
  const int nW = (W + 31) >> 5;
  u64 tw[6];
#pragma unroll
  for (int j = 0; j < 6; ++j) tw[j] = j <= nW ? tp[j] : 0;
  if (alleleHasN && n_in_range(R.n2 + w0, pLo + d, pHi - 1 + d)) return DF_DECLINED;
  u64 t0 = tw[0], t1 = tw[1], t2 = tw[2], t3 = tw[3], t4 = tw[4], t5 = tw[5];     // word queue: later words move up
  int exMm = 0;
  // streaming state over the mismatches (ascending) and the closing sentinel pHi
  int prevMis = pLo - 1;
  int rs = -1, lastL = 0, hitLen = 0, cntSum = 0, gapMatches = 0, mmRun = 0, mmLeft = 0, mmTot = 0;
  u32 mmPos = 0;
  bool ok = true;
  T1K_NOUNROLL
  for (int c0 = 0; ok; c0 += 32) {
    const bool closing = c0 >= W;          // one extra round for the sentinel
    u64 m = 0;
    if (!closing) {
      const u64 t = (t0 >> sh) | ((t1 << 1) << (63 - sh));
      t0 = t1; t1 = t2; t2 = t3; t3 = t4; t4 = t5;
      const u64 x = t ^ fetch32(Q.seq2, pLo + c0);
      m = (x | (x >> 1)) & M55 & lowmask2(W - c0);
      if (m && R.relax) exMm += popc64(m & fetch32(R.ex2 + w0, pLo + d + c0));
    }
    T1K_NOUNROLL
    while (m || closing) {
      int p;
      if (m) { p = pLo + c0 + (ctz64(m) >> 1); m &= m - 1; }
      else p = pHi;
      // k-mer starts of the mismatch-free stretch (prevMis, p): [prevMis + 1, p - k]
      const int lo = prevMis + 1, hi = p - KMER;
      if (hi >= lo) {
        const int f = (int)((stab[lo] >> 16) & 255);
        if (f <= hi) {                      // (255 = none)
          const u32 tl = stab[hi], tf = stab[f];
          const int l = (int)(tl >> 24);
          if (((tl >> 8) & 255) != ((tf >> 8) & 255)) { ok = false; break; }     // a step > k-1 inside the stretch
          if (rs < 0) { rs = f; mmLeft = mmRun; }
          else {
            const int g = f - (lastL + KMER);
            if (g > 32 || mmRun > 3) {
              // a long or dirty gap: the same GlobalAlignment the hit-list walk would run (certificates, else the band
              // DP), the rest of the allele stays on this path.  No N in the window => the N planes are not consulted.
              if (hot) needCold = true;
              else if (nColdGap < 4) { coldGap[nColdGap++] = (u32)(lastL + KMER) | ((u32)g << 16); }       // aligned after the scan, see below
              else {
              AlleleView T;
              T.seq = R.seq2 + w0; T.n2 = R.n2 + w0; T.ex2 = R.ex2 + w0; T.len = clen; T.hasN = alleleHasN; T.useN = false;
              T1K_COUNT(32, 1); T1K_COUNT(33, g > 32); T1K_COUNT(40 + (mmRun > 15 ? 15 : mmRun), 1); T1K_COUNT(60 + (g / 8 > 15 ? 15 : g / 8), 1);
              gapMatches += align_matches_cold(T, lastL + KMER + d, g, Q, lastL + KMER, g, S, err);
              }
            } else gapMatches += g - mmRun;
          }
          hitLen += l - f + KMER;
          cntSum += (int)(tl & 255) - (int)(tf & 255) + 1;
          lastL = l; mmRun = 0;
        }
      }
      if (p == pHi) break;
      if (mmTot < 3) mmPos |= (u32)p << (8 * mmTot);
      ++mmTot; ++mmRun; prevMis = p;
    }
    if (closing) break;
  }
  T1K_COUNT(16, 1);
  if (!ok) T1K_COUNT(21, 1); else if (cntSum < n) T1K_COUNT(22, 1); else if (cntSum > n) T1K_COUNT(23, 1);
  if (!ok || cntSum > n) return DF_DECLINED;
  if (cntSum < n) {
    // One or two postings off the diagonal (a k-mer of the read that also occurs elsewhere in the allele).  If every one
    // of them lies more than RADIUS diagonals away, the diagonal sort puts a cluster break on both sides of the main
    // diagonal (SeqSet.hpp:1369-1374), so its cluster is exactly the seeds found above, and the strays form clusters of
    // fewer than three hits, which are dropped (SeqSet.hpp:1399-1404): the result is the single-diagonal one.
    const int extra = n - cntSum;
    if (extra > 2) return DF_DECLINED;
    if (onDiag != cntSum || far != extra) return DF_DECLINED;
  }
  T1K_COUNT(17, 1);
  // ---- from here on the result is the reference's: the tail of consume_chain<true>
  if (hitLen < HIT_LEN_REQ) return DF_DONE;
  if (needCold) return DF_DEFER;
  if (nColdGap > 0) {
    // the dirty gaps of the scan, aligned here so that the lanes of a warp (every one on an allele of its own in k_deferred)
    // run their alignments together instead of one after the other from inside the scan
    AlleleView T;
    T.seq = R.seq2 + w0; T.n2 = R.n2 + w0; T.ex2 = R.ex2 + w0; T.len = clen; T.hasN = alleleHasN; T.useN = false;
    T1K_NOUNROLL
    for (int i = 0; i < nColdGap; ++i) {
      const int gp = (int)(coldGap[i] & 0xffffu), g = (int)(coldGap[i] >> 16);
      gapMatches += align_matches_cold(T, gp + d, g, Q, gp, g, S, err);
    }
  }
  const int re = lastL + KMER - 1, mmRight = mmRun;
  const u64 sk = strand_key(2 * hitLen, re - rs, seqIdx, strand01);
  if (sk > bestStrandKey) bestStrandKey = sk;
  const int mc = 2 * hitLen + 2 * gapMatches;
  bool below = sim_below(R, mc, 2 * (re - rs + 1), 0);
  if (!below) {
    const u32 key = 0x10000u | (u32)rs | ((u32)re << 8);
    if ((lcMemo & 0x1FFFFu) != key) lcMemo = key | (low_complex(Q, rs, re) ? 0x20000u : 0u);
    if (lcMemo & 0x20000u) below = 0.0 < R.sim;
  }
  if (below) return DF_DONE;
  // ---- ExtendOverlap on the same diagonal (SeqSet.hpp:1994-2100): overhangs [pLo, rs) and (re, pHi)
  const int lo = rs - pLo, ro = pHi - 1 - re;
  const bool hotExt = (lo <= 32 && mmLeft <= 3) && (ro <= 32 && mmRight <= 3);
  if (hot && !hotExt) return DF_DEFER;
  Cand &c = out;
  c.seqIdx = seqIdx; c.seqStart = rs + d; c.seqEnd = re + d;
  c.readStart = (u16)rs; c.readEnd = (u16)re; c.strand01 = (u8)strand01; c.flags = 0;
  c.matchCnt = (u16)mc; c.mmPos = 0;
  emitted = true;
  if (hotExt) {
    const int mcE = mc + 2 * (lo - mmLeft + ro - mmRight);
    const int leftClip = pLo, rightClip = len - pHi;
    u8 flags = CF_PRE;
    if (d < 0 || d + len > clen) flags |= CF_NEEDCLIP;
    if (!sim_below(R, mcE, 2 * W, 0)) flags |= CF_RET;
    c.eReadStart = (u16)pLo; c.eReadEnd = (u16)(pHi - 1);
    c.eSeqStart = pLo + d; c.eSeqEnd = pHi - 1 + d;
    c.leftClip = (u16)leftClip; c.rightClip = (u16)rightClip;
    c.relaxed = mcE;
    c.eMatchCnt = mcE + 2 * leftClip + 2 * rightClip;
    // ---- the full-read alignment of [pLo, pHi) (SeqSet.hpp:2203-2274): <= 3 mismatches certify the diagonal
    T1K_COUNT(18, 1);
    if (mmTot <= 3) { T1K_COUNT(19, 1); flags |= CF_FA; c.mmPos = mmPos | ((u32)mmTot << 24) | ((u32)exMm << 26); }
    c.flags = flags;
  }
  return DF_DONE;
}

// ---- The same evaluation, bit-parallel: what the lanes of a tile run in hot mode (one allele per lane, no data-dependent
// loop in the common case, so the 32 alleles of a tile stay in lock-step).  Everything lives in 5 words of the 2-bit
// space indexed by READ position (bit 2p = position p; reads up to 160 bases):
//   M     mismatches of the read against the allele on diagonal d inside the window [pLo, pHi)
//   S     the strand's seeds (seed_bits2, built once per strand)
//   H     seeds whose k bases are all inside the window and match = S & ~(M or outside, OR-ed over the k positions)
//   U     positions covered by the k-mers of H (H dilated by k)
// The hit-count certificate is popcount(H) == n (or the one/two far strays rule) exactly as in diag_fast.  Given it, the
// allele's hits are the seeds of H on one diagonal, and (see consume_chain) hitLen = |U|; two consecutive hits either overlap
// or are separated by a gap that the reference aligns globally: with <= 3 mismatches the gap's alignment is the pure diagonal
// (DESIGN.md "diagonal certificate", any length), so matchCnt = 2 * (span - mismatches inside the span).  A gap with more
// mismatches, or an overhang with more than 3, needs the real alignment: DF_DEFER.  lcp: prefix base counts of the strand
// (lc_prefix_build) for IsOverlapLowComplex.
// bits 2p (p < b) of word j (positions 32j .. 32j+31)
T1K_HD u64 below_mask2(int j, int b) {
  const int hi = b - 32 * j;
  return hi <= 0 ? 0ull : hi >= 32 ? M55 : (((1ull << (2 * hi)) - 1) & M55);
}
// bits 2p (p in [a, b)) of word j (positions 32j .. 32j+31)
T1K_HD u64 range_mask2(int j, int a, int b) {
  int lo = a - 32 * j, hi = b - 32 * j;
  lo = lo < 0 ? 0 : lo > 32 ? 32 : lo; hi = hi < 0 ? 0 : hi > 32 ? 32 : hi;
  return hi > lo ? (lowmask2(hi) & ~lowmask2(lo) & M55) : 0ull;
}
// IsOverlapLowComplex (SeqSet.hpp:458-485) from prefix counts: lcp[p] = number of A | C << 8 | G << 16 | T << 24 among the
// bases [0, p) of an N-free strand
T1K_HD bool low_complex_lcp(const u32 *lcp, int s, int e) {
  const u32 a = lcp[e + 1], b = lcp[s];
  const int n = e - s + 1;
  int low = 0, lowTotal = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = (int)((a >> (8 * i)) & 255) - (int)((b >> (8 * i)) & 255);
    if (c <= 2) { ++low; lowTotal += c; }
  }
  if (lowTotal * 7 >= n) return false;
  return low >= 2;
}
T1K_HDN inline void lc_prefix_build(const ReadView &Q, u32 *lcp) {       // len + 1 entries (host / one lane)
  u32 c = 0;
  for (int p = 0; p <= Q.len; ++p) { lcp[p] = c; if (p < Q.len) c += 1u << (8 * base2(Q.seq2, p)); }
}
T1K_HDN inline void seed_bits2_build(const u16 *seedA, int nS, u32 *s2) {  // 10 u32 = 5 words of the 2-bit space
  for (int i = 0; i < 10; ++i) s2[i] = 0;
  for (int k = 0; k < nS; ++k) s2[seedA[k] >> 4] |= 1u << (2 * (seedA[k] & 15));
}

T1K_HDN T1K_NOINLINE inline int diag_hot(const RefView &R, const ReadView &Q, int strand01, int seqIdx, int n, int d, int onDiag, int far,
                             const u64 *S2, const u32 *lcp, Cand &out, bool &emitted, u64 &bestStrandKey) {
  emitted = false;
  const int len = Q.len;
#ifdef __CUDA_ARCH__
  const uint4 mt = *reinterpret_cast<const uint4 *>(R.meta + seqIdx);
  const u64 w0 = (u64)mt.x | ((u64)mt.y << 32);
  const int clen = (int)mt.z;
  const bool alleleHasN = mt.w != 0;
#else
  const u64 w0 = R.meta[seqIdx].wordOff;
  const int clen = R.meta[seqIdx].len;
  const bool alleleHasN = R.meta[seqIdx].hasN != 0;
#endif
  const int pLo = d < 0 ? -d : 0, pHi = imin(len, clen - d);
  const int W = pHi - pLo;
  if (W < KMER) return DF_DECLINED;
  if (alleleHasN && n_in_range(R.n2 + w0, pLo + d, pHi - 1 + d)) return DF_DECLINED;
  // allele bases at read positions 32j .. 32j+31 = allele positions 32j + d ..: six consecutive words (the planes are padded,
  // so words before the first / after the last allele are readable; what lies outside the window is masked)
  const u64 *tp = R.seq2 + (long long)w0 + (long long)(d >> 5);
  const int sh = (d & 31) * 2;
  u64 tw[6];
#pragma unroll
  for (int j = 0; j < 6; ++j) tw[j] = tp[j];
  u64 M[5], B[6];
  int mmTot = 0;
#pragma unroll
  for (int j = 0; j < 5; ++j) {
    const u64 t = sh ? shr2w(tw[j], tw[j + 1], sh) : tw[j];
    const u64 x = t ^ Q.seq2[j];
    const u64 in = below_mask2(j, pHi) & ~below_mask2(j, pLo);
    M[j] = (x | (x >> 1)) & in;
    B[j] = M[j] | (~in & M55);
    mmTot += popc64(M[j]);
  }
  B[5] = M55;
  // H: seeds whose k-mer [a, a+k) holds no B bit: OR of B over shifts 0 .. k-1 (towards lower positions), k = 11
  u64 H[5];
  {
    u64 w2[6], w4[6];
#pragma unroll
    for (int j = 0; j < 5; ++j) w2[j] = B[j] | shr2w(B[j], B[j + 1], 2);
    w2[5] = M55;
#pragma unroll
    for (int j = 0; j < 5; ++j) w4[j] = w2[j] | shr2w(w2[j], w2[j + 1], 4);
    w4[5] = M55;
    u64 w8[6];
#pragma unroll
    for (int j = 0; j < 5; ++j) w8[j] = w4[j] | shr2w(w4[j], w4[j + 1], 8);
    w8[5] = M55;
#pragma unroll
    for (int j = 0; j < 5; ++j) H[j] = S2[j] & ~(w8[j] | shr2w(w4[j], w4[j + 1], 14));
  }
  int cntSum = 0;
#pragma unroll
  for (int j = 0; j < 5; ++j) cntSum += popc64(H[j]);
  if (cntSum > n) return DF_DECLINED;
  if (cntSum < n) {
    // one or two postings off the diagonal, each more than RADIUS diagonals away (see diag_fast)
    const int extra = n - cntSum;
    if (extra > 2) return DF_DECLINED;
    if (onDiag != cntSum || far != extra) return DF_DECLINED;
  }
  if (cntSum == 0) return DF_DONE;
  // U: H dilated by k towards higher positions
  u64 U[5];
  {
    u64 u2[5], u4[5], u8[5];
#pragma unroll
    for (int j = 0; j < 5; ++j) u2[j] = H[j] | (j ? shl2w(H[j], H[j - 1], 2) : (H[j] << 2));
#pragma unroll
    for (int j = 0; j < 5; ++j) u4[j] = u2[j] | (j ? shl2w(u2[j], u2[j - 1], 4) : (u2[j] << 4));
#pragma unroll
    for (int j = 0; j < 5; ++j) u8[j] = u4[j] | (j ? shl2w(u4[j], u4[j - 1], 8) : (u4[j] << 8));
#pragma unroll
    for (int j = 0; j < 5; ++j) U[j] = u8[j] | (j ? shl2w(u4[j], u4[j - 1], 14) : (u4[j] << 14));
  }
  int hitLen = 0, rs = 0, lastL = 0;
#pragma unroll
  for (int j = 0; j < 5; ++j) hitLen += popc64(U[j]);
#pragma unroll
  for (int j = 4; j >= 0; --j) if (H[j]) rs = 32 * j + (ctz64(H[j]) >> 1);
#pragma unroll
  for (int j = 0; j < 5; ++j) if (H[j]) {
#ifdef __CUDA_ARCH__
    lastL = 32 * j + ((63 - __clzll((long long)H[j])) >> 1);
#else
    lastL = 32 * j + ((63 - __builtin_clzll(H[j])) >> 1);
#endif
  }
  if (hitLen < HIT_LEN_REQ) return DF_DONE;
  const int re = lastL + KMER - 1;
  int mmInside = 0, mmLeft = 0;
  u64 SP[5];                         // M restricted to the span [rs, re]
#pragma unroll
  for (int j = 0; j < 5; ++j) {
    const u64 bl = below_mask2(j, rs), bh = below_mask2(j, re + 1);
    SP[j] = M[j] & bh & ~bl;
    mmInside += popc64(SP[j]); mmLeft += popc64(M[j] & bl);
  }
  const int mmRight = mmTot - mmInside - mmLeft;
  if (mmInside > 3) {
    // does one gap (a maximal stretch of [rs, re] that U does not cover) hold more than 3 mismatches?
    int cnt = 0; bool pendingU = false, dirty = false;
#pragma unroll
    for (int j = 0; j < 5; ++j) {
      u64 g = SP[j], done = 0;
      T1K_NOUNROLL
      while (g) {
        const u64 b = g & (~g + 1), below = b - 1;
        if (pendingU || (U[j] & below & ~done)) cnt = 0;
        ++cnt; dirty |= cnt > 3; pendingU = false;
        done = below | b; g &= g - 1;
      }
      pendingU |= (U[j] & ~done) != 0;
    }
    if (dirty) return DF_DEFER;
  }
  const u64 sk = strand_key(2 * hitLen, re - rs, seqIdx, strand01);
  if (sk > bestStrandKey) bestStrandKey = sk;
  const int mc = 2 * (re - rs + 1 - mmInside);
  bool below = sim_below(R, mc, 2 * (re - rs + 1), 0);
  if (!below && low_complex_lcp(lcp, rs, re)) below = 0.0 < R.sim;
  if (below) return DF_DONE;
  // ---- ExtendOverlap on the same diagonal (SeqSet.hpp:1994-2100): overhangs [pLo, rs) and (re, pHi)
  if (mmLeft > 3 || mmRight > 3) return DF_DEFER;
  Cand &c = out;
  c.seqIdx = seqIdx; c.seqStart = rs + d; c.seqEnd = re + d;
  c.readStart = (u16)rs; c.readEnd = (u16)re; c.strand01 = (u8)strand01;
  c.matchCnt = (u16)mc; c.mmPos = 0;
  emitted = true;
  const int mcE = 2 * (W - mmTot);
  const int leftClip = pLo, rightClip = len - pHi;
  u8 flags = CF_PRE;
  if (d < 0 || d + len > clen) flags |= CF_NEEDCLIP;
  if (!sim_below(R, mcE, 2 * W, 0)) flags |= CF_RET;
  c.eReadStart = (u16)pLo; c.eReadEnd = (u16)(pHi - 1);
  c.eSeqStart = pLo + d; c.eSeqEnd = pHi - 1 + d;
  c.leftClip = (u16)leftClip; c.rightClip = (u16)rightClip;
  c.relaxed = mcE;
  c.eMatchCnt = mcE + 2 * leftClip + 2 * rightClip;
  // ---- the full-read alignment of [pLo, pHi) (SeqSet.hpp:2203-2274): <= 3 mismatches certify the diagonal
  if (mmTot <= 3) {
    u32 mmPos = 0; int k = 0, exMm = 0;
#pragma unroll
    for (int j = 0; j < 5; ++j) {
      u64 m = M[j];
      T1K_NOUNROLL
      while (m) { mmPos |= (u32)(32 * j + (ctz64(m) >> 1)) << (8 * k); ++k; m &= m - 1; }
    }
    if (R.relax && mmTot > 0) {
#pragma unroll
      for (int j = 0; j < 5; ++j) if (M[j]) exMm += popc64(M[j] & fetch32(R.ex2 + w0, 32 * j + d));
    }
    flags |= CF_FA;
    c.mmPos = mmPos | ((u32)mmTot << 24) | ((u32)exMm << 26);
  }
  c.flags = flags;
  return DF_DONE;
}

// full-read alignment of a CF_FA candidate: coverage and the exon-relaxed count from the stored mismatch positions
T1K_HD void full_align_known(const RefView &R, Cand &c, int weight) {
  const int mm = (int)((c.mmPos >> 24) & 3), exMm = (int)((c.mmPos >> 26) & 3);
  const int lent = c.eSeqEnd - c.eSeqStart + 1;
  if (weight > 0) {
    const size_t cb = (size_t)R.covOff[c.seqIdx];
    const int dd = c.eSeqStart - (int)c.eReadStart;
    cov_add(R.covDiff + cb + (size_t)COV_STRIDE * c.eSeqStart, weight); cov_add(R.covDiff + cb + (size_t)COV_STRIDE * (c.eSeqStart + lent), -weight);
    if (mm > 0) cov_add(R.covPoint + cb + (size_t)COV_STRIDE * (dd + (int)(c.mmPos & 255)), -weight);
    if (mm > 1) cov_add(R.covPoint + cb + (size_t)COV_STRIDE * (dd + (int)((c.mmPos >> 8) & 255)), -weight);
    if (mm > 2) cov_add(R.covPoint + cb + (size_t)COV_STRIDE * (dd + (int)((c.mmPos >> 16) & 255)), -weight);
  }
  c.relaxed = R.relax ? 2 * (lent - exMm) : c.eMatchCnt;
}

// The seed overlaps one (strand, allele) group emitted, into the tail order of _overlap::operator< (readStart, readEnd,
// seqStart, seqEnd; SeqSet.hpp:117-125): clusters come out by ascending diagonal, i.e. by DEscending seqStart for equal
// read coordinates (an allele that holds a segment twice), the reference lists them ascending.  n is almost always 1.
T1K_HD bool cand_tail_less(const Cand &a, const Cand &b) {
  if (a.readStart != b.readStart) return a.readStart < b.readStart;
  if (a.readEnd != b.readEnd) return a.readEnd < b.readEnd;
  if (a.seqStart != b.seqStart) return a.seqStart < b.seqStart;
  return a.seqEnd < b.seqEnd;
}
T1K_HDN T1K_NOINLINE inline void sort_emitted(Cand *e, int n) {
  T1K_NOUNROLL
  for (int i = 1; i < n; ++i) {
    if (!cand_tail_less(e[i], e[i - 1])) continue;
    const Cand v = e[i];
    int j = i - 1;
    T1K_NOUNROLL
    while (j >= 0 && cand_tail_less(v, e[j])) { e[j + 1] = e[j]; --j; }
    e[j + 1] = v;
  }
}

struct ChainDirect {   // contiguous run of a hit store
  const u32 *p; int stride;
  T1K_HD u32 operator()(int i) const { return p[(size_t)i * stride]; }
};

// the multi-diagonal cluster [s, e) of chain_allele: closest-to-dominant hit per read offset, LIS, chain (rare: kept out
// of line so that the common single-diagonal path stays small)
T1K_HDN T1K_NOINLINE inline void chain_cluster_general(const RefView &R, const ReadView &Q, int strand01, int seqIdx, u32 *h, int stride, int s, int e,
                                          int dom, const LaneScratch &S, int &nEmit, u64 &bestStrandKey, int &err) {
  const int m = e - s;
    // general path (SeqSet.hpp:1437-1456 + LIS :352-436)
    const size_t usedBytes = ((size_t)2 * (size_t)(Q.len + 1) + 15) & ~(size_t)15;
    if ((size_t)m * 12 + usedBytes > (size_t)S.dirBytes) { err |= ERR_SCRATCH; return; }
    u16 *used = (u16 *)S.dir();                 // min |diag - dom| per read offset
    u32 *conc = (u32 *)(S.dir() + usedBytes);
    u32 *chain = conc + m;
    u16 *top = (u16 *)(chain + m);
    u16 *link = top + m;
    T1K_NOUNROLL
    for (int k = s; k < e; ++k) used[hit_a(h[(size_t)k * stride])] = 0xFFFF;
    T1K_NOUNROLL
    for (int k = s; k < e; ++k) {
      u32 v = h[(size_t)k * stride];
      int d = iabs(hit_a(v) - hit_b(v) - dom);
      if (d > 0xFFFE) d = 0xFFFE;
      if (used[hit_a(v)] > d) used[hit_a(v)] = (u16)d;
    }
    int cn = 0;
    T1K_NOUNROLL
    for (int k = s; k < e; ++k) {
      u32 v = h[(size_t)k * stride];
      int d = iabs(hit_a(v) - hit_b(v) - dom);
      if (d > 0xFFFE) d = 0xFFFE;
      if (d == used[hit_a(v)]) {             // insertion into (b,a) order == numeric order
        int j = cn - 1;
        T1K_NOUNROLL
        while (j >= 0 && conc[j] > v) { conc[j + 1] = conc[j]; --j; }
        conc[j + 1] = v; ++cn;
      }
    }
    // LIS over read offsets (non-strict probe, strict extend; Q4)
    int ret = 1;
    top[0] = 0; link[0] = 0xFFFF;
    T1K_NOUNROLL
    for (int i = 1; i < cn; ++i) {
      int ai = hit_a(conc[i]);
      int tag;
      if (hit_a(conc[top[ret - 1]]) <= ai) tag = ret - 1;
      else {
        int l = 0, r = ret - 1; tag = -2;
        T1K_NOUNROLL
        while (l <= r) {
          int mid = (l + r) / 2, am = hit_a(conc[top[mid]]);
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
      int k = top[ret - 1];
      T1K_NOUNROLL
      for (int i = ret - 1; i >= 0; --i) { chain[i] = conc[k]; k = link[k]; }
    }
    int sz = 0;
    T1K_NOUNROLL
    for (int i = 0; i < ret; ++i)
      if (i == 0 || hit_b(chain[i]) != hit_b(chain[sz - 1])) chain[sz++] = chain[i];
    // the chain was built in S.dir(), which consume_chain's gap DPs overwrite: park it in its own region
    u32 *park = S.chain();
    if (sz > S.chainCap) { err |= ERR_SCRATCH; return; }
    T1K_NOUNROLL
    for (int i = 0; i < sz; ++i) park[i] = chain[i];
    ChainDirect cd; cd.p = park; cd.stride = 1;
    consume_chain<false>(R, Q, strand01, seqIdx, cd, sz, S, nEmit, bestStrandKey, err);
}

// ---- SeqSet::GetOverlapsFromHits for one (strand, allele) group (SeqSet.hpp:1303-1553; filter=0, isRef).
// hits: n encoded hits at h[i*stride]; on entry sorted by (readOffset, seqOffset); sorted in place by diagonal.
// Scratch use of the general (multi-diagonal) path: conc/chain/top/link live in S.dir().
T1K_HDN T1K_NOINLINE inline void chain_allele(const RefView &R, const ReadView &Q, int strand01, int seqIdx, u32 *h, int stride, int n,
                                 const LaneScratch &S, int &nEmit, u64 &bestStrandKey, int &err) {
  if (n < 3) return;
  // insertion sort by (diag, b, a); a single-diagonal group is already in order
  T1K_NOUNROLL
  for (int i = 1; i < n; ++i) {
    u32 v = h[(size_t)i * stride];
    if (!hit_diag_less(v, h[(size_t)(i - 1) * stride])) continue;
    int j = i - 1;
    T1K_NOUNROLL
    while (j >= 0 && hit_diag_less(v, h[(size_t)j * stride])) { h[(size_t)(j + 1) * stride] = h[(size_t)j * stride]; --j; }
    h[(size_t)(j + 1) * stride] = v;
  }
  int dom = 0;
  T1K_NOUNROLL
  for (int s = 0; s < n;) {
    int e, cur, curCnt = 1, domCnt = 0, prevC;
    { u32 v = h[(size_t)s * stride]; cur = hit_a(v) - hit_b(v); prevC = cur; }
    T1K_NOUNROLL
    for (e = s + 1; e < n; ++e) {
      u32 v = h[(size_t)e * stride];
      int c = hit_a(v) - hit_b(v);
      int diff = c - prevC;               // sorted ascending: diff >= 0
      if (diff > RADIUS) break;
      if (diff == 0) ++curCnt;
      else {
        if (curCnt > domCnt) { dom = cur; domCnt = curCnt; }
        cur = c; curCnt = 1;
      }
      prevC = c;
    }
    if (curCnt > domCnt) dom = cur;
    int m = e - s;
    if (m < 3 || m * KMER < HIT_LEN_REQ) { s = e; continue; }
    u32 first = h[(size_t)s * stride], last = h[(size_t)(e - 1) * stride];
    if (hit_a(first) - hit_b(first) == hit_a(last) - hit_b(last)) {
      // one diagonal: every read offset occurs once, (b,a) order == current order, LIS keeps everything
      ChainDirect cd; cd.p = h + (size_t)s * stride; cd.stride = stride;
      consume_chain<false>(R, Q, strand01, seqIdx, cd, m, S, nEmit, bestStrandKey, err);   // (the <true> variant is the same function specialised; one copy keeps the kernel's code small)
      s = e; continue;
    }
    chain_cluster_general(R, Q, strand01, seqIdx, h, stride, s, e, dom, S, nEmit, bestStrandKey, err);
    s = e;
  }
}

// ---- SeqSet::ExtendOverlap (SeqSet.hpp:1994-2100) + the separator tests of AssignRead (SeqSet.hpp:2163-2169)
// HOT: only the one-word comparisons are allowed; returns false (c untouched) when an overhang needs the cold path, so
// that the caller can run those candidates together afterwards instead of stalling the warp on a few lanes.
template <bool HOT>
T1K_HDN T1K_NOINLINE inline bool extend_cand(const RefView &R, const ReadView &Q, Cand &c, const LaneScratch &S, int &err) {
  const AlleleView T = allele_view(R, c.seqIdx, Q);
  const int clen = T.len, len = Q.len;
  int rs = c.readStart, re = c.readEnd, ss = c.seqStart, se = c.seqEnd;
  u8 flags = 0;
  if (sep_in_range(T, ss, se)) { c.flags = CF_SEP; return true; }
  if (sep_in_range(T, ss - rs, se + (len - re - 1))) flags |= CF_NEEDCLIP;
  int lo = imin(rs, ss), leftClip = 0, rightClip = 0;
  if (rs > ss) leftClip = rs - ss;
  if (T.hasN) {
    T1K_NOUNROLL
    for (int i = 0; i < lo; ++i)
      if (base2(T.n2, ss - i - 1)) { leftClip = lo - i; lo = i; break; }
  }
  int m = HOT ? align_matches_hot(T, ss - lo, lo, Q, rs - lo, lo) : align_matches(T, ss - lo, lo, Q, rs - lo, lo, S, err);
  if (HOT && m < 0) return false;
  int ro = imin(len - 1 - re, clen - 1 - se);
  if (len - 1 - re > clen - 1 - se) rightClip = len - 1 - re - (clen - 1 - se);
  if (T.hasN) {
    T1K_NOUNROLL
    for (int i = 0; i < ro; ++i)
      if (base2(T.n2, se + 1 + i)) { rightClip = ro - i; ro = i; break; }
  }
  {
    const int m2 = HOT ? align_matches_hot(T, se + 1, ro, Q, re + 1, ro) : align_matches(T, se + 1, ro, Q, re + 1, ro, S, err);
    if (HOT && m2 < 0) return false;
    m += m2;
  }
  c.eReadStart = (u16)(rs - lo); c.eReadEnd = (u16)(re + ro);
  c.eSeqStart = ss - lo; c.eSeqEnd = se + ro;
  int mc = 2 * m + c.matchCnt;
  if (!sim_below(R, mc, (re + ro) - (rs - lo) + 1 + (se + ro) - (ss - lo) + 1, 0)) flags |= CF_RET;
  c.leftClip = (u16)leftClip; c.rightClip = (u16)rightClip;
  c.relaxed = mc;                                  // SeqSet.hpp:2068 (before the clip bonus)
  c.eMatchCnt = mc + 2 * leftClip + 2 * rightClip;
  c.flags = flags;
  return true;
}

// ---- full-read alignment of an extended overlap (SeqSet.hpp:2203-2274): exon-relaxed match count and
// base coverage.  Coverage is kept as a range-add difference array plus point corrections, so a
// certified-diagonal record costs 2 + (#uncredited columns) atomics instead of one per base.
// Hot path (full_align): one pass over the window — mismatch count, exonic mismatch count and the first three mismatch
// positions; <= 3 mismatches certify the diagonal and, without N columns, those positions are the uncredited columns.
// Cold path (full_align_cold): 4+ mismatches (certificates, DP) and windows with N columns.
// allowDp = false: returns false (nothing written, no coverage added) when only the band DP can tell — the caller collects those
// and runs them together (k_align_dp), instead of one lane of a warp running a 150-row DP while 31 wait.
T1K_HDN T1K_NOINLINE inline bool full_align_cold(const RefView &R, const ReadView &Q, const AlleleView &T, Cand &c, int weight, int mm, int exMm,
                                    const LaneScratch &S, int &err, bool allowDp = true) {
  const int tpos = c.eSeqStart, ppos = c.eReadStart;
  const int lent = c.eSeqEnd - c.eSeqStart + 1, lenp = c.eReadEnd - c.eReadStart + 1;
  int32_t *covDiff = R.covDiff + (size_t)R.covOff[c.seqIdx], *covPoint = R.covPoint + (size_t)R.covOff[c.seqIdx];
  if (lent == lenp) {
    bool diag = mm <= 3;
    if (!diag) {
      if (mm <= 5) diag = diag_certified_interval(T, tpos, Q, ppos, lent) || diag_certified_45(T, tpos, Q, ppos, lent, mm);
      else if (mm <= 24) diag = diag_certified_interval(T, tpos, Q, ppos, lent) || diag_certified_hist(T, tpos, Q, ppos, lent);
    }
    if (diag) {
      if (weight > 0) {
        cov_add(covDiff + (size_t)COV_STRIDE * tpos, weight); cov_add(covDiff + (size_t)COV_STRIDE * (tpos + lent), -weight);
        T1K_NOUNROLL
        for (int k = 0; k < lent; k += 32) {
          u64 un = mm_chunk(T, tpos + k, Q, ppos + k, lent - k);
          if (T.useN) un |= (fetch32(T.n2, tpos + k) | fetch32(Q.n2, ppos + k)) & lowmask2(lent - k);
          T1K_NOUNROLL
          while (un) {
            const int p = k + (ctz64(un) >> 1);
            un &= un - 1;
            cov_add(covPoint + (size_t)COV_STRIDE * (tpos + p), -weight);
          }
        }
      }
      c.relaxed = R.relax ? 2 * (lent - exMm) : c.eMatchCnt;
      return true;
    }
    T1K_COUNT(8 + (mm >= 4 && mm <= 10 ? mm - 4 : 7), 1);
  }
  if (!allowDp) return false;
  T1K_COUNT(7, 1);
  int n = dp_align(T, tpos, lent, Q, ppos, lenp, S, err);
  if (n < 0) { c.relaxed = c.eMatchCnt; return true; }
  const u8 *ops = S.ops();
  int refPos = tpos, readPos = ppos, m = 0;
  T1K_NOUNROLL
  for (int k = 0; k < n; ++k) {
    int op = ops[k];
    if (R.relax) {
      if (base2(T.ex2, refPos)) { if (op == 0) ++m; } else ++m;
    }
    if (weight > 0 && op == 0 && readPos < Q.len && refPos < T.len && !base2(Q.n2, readPos) &&
        !base2(T.n2, refPos) && base2(T.seq, refPos) == base2(Q.seq2, readPos))
      cov_add(covPoint + (size_t)COV_STRIDE * refPos, weight);
    if (op != 2) ++refPos;
    if (op != 3) ++readPos;
  }
  c.relaxed = R.relax ? 2 * m : c.eMatchCnt;
  return true;
}

// HOT: returns false (nothing written, no coverage added) when the window needs full_align_cold.  allowDp: see full_align_cold.
template <bool HOT>
T1K_HDN T1K_NOINLINE inline bool full_align(const RefView &R, const ReadView &Q, Cand &c, int weight, const LaneScratch &S, int &err, bool allowDp = true) {
  const AlleleView T = allele_view(R, c.seqIdx, Q);
  const int tpos = c.eSeqStart, ppos = c.eReadStart;
  const int lent = c.eSeqEnd - c.eSeqStart + 1, lenp = c.eReadEnd - c.eReadStart + 1;
  T1K_COUNT(6, 1);
  int mm = 0, exMm = 0, p0 = 0, p1 = 0, p2 = 0;
  if (lent == lenp) {
    T1K_NOUNROLL
    for (int k = 0; k < lent; k += 32) {
      u64 d = mm_chunk(T, tpos + k, Q, ppos + k, lent - k);
      if (d) {
        if (R.relax) exMm += popc64(d & fetch32(T.ex2, tpos + k));
        T1K_NOUNROLL
        while (d && mm < 3) {
          const int p = k + (ctz64(d) >> 1);
          d &= d - 1;
          if (mm == 0) p0 = p; else if (mm == 1) p1 = p; else p2 = p;
          ++mm;
        }
        mm += popc64(d);
      }
    }
  }
  if (lent == lenp) {
    if (mm <= 3 && !T.useN) {
      if (weight > 0) {
        int32_t *covDiff = R.covDiff + (size_t)R.covOff[c.seqIdx], *covPoint = R.covPoint + (size_t)R.covOff[c.seqIdx];
        cov_add(covDiff + (size_t)COV_STRIDE * tpos, weight); cov_add(covDiff + (size_t)COV_STRIDE * (tpos + lent), -weight);
        if (mm > 0) cov_add(covPoint + (size_t)COV_STRIDE * (tpos + p0), -weight);
        if (mm > 1) cov_add(covPoint + (size_t)COV_STRIDE * (tpos + p1), -weight);
        if (mm > 2) cov_add(covPoint + (size_t)COV_STRIDE * (tpos + p2), -weight);
      }
      c.relaxed = R.relax ? 2 * (lent - exMm) : c.eMatchCnt;
      return true;
    }
  }
  if (HOT) return false;
  return full_align_cold(R, Q, T, c, weight, mm, exMm, S, err, allowDp);
}

// post-extension denominators / keys
T1K_HD int cand_denom_pre(const Cand &c) { return c.seqEnd - c.seqStart + 1 + c.readEnd - c.readStart + 1; }
T1K_HD int cand_denom_post(const Cand &c) {
  return c.eSeqEnd - c.eSeqStart + 1 + c.eReadEnd - c.eReadStart + 1 + 2 * c.leftClip + 2 * c.rightClip;
}
T1K_HD u64 cand_key_pre(const Cand &c) {
  return order_key(c.matchCnt, cand_denom_pre(c), c.readEnd - c.readStart);
}
T1K_HD u64 cand_key_post(const Cand &c) {
  return order_key(c.eMatchCnt, cand_denom_post(c), c.eReadEnd - c.eReadStart);
}

}  // namespace t1k
