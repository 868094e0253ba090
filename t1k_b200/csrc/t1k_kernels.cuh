// sm_100a kernels of the alignment path: read packing and SeqSet::AssignRead as rounds of kernels over a slice of the batch —
// k_seed (persistent, warp per read-end: seed selection, sweep over the allele tiles of the index, lane-per-allele mismatch-mask
// evaluation, candidates into the warp's arena), k_defer_group / k_deferred / k_defer_copy (alleles that need real gap or
// overhang alignments, distinct ones compacted), k_passes (the ordered scan, inclusion, > 1000 cut, records into the
// HBM-resident overlap store), k_align / k_align_dp (full-read alignments: certificates, band DP, coverage).
#pragma once
#include <cuda_runtime.h>

#include "t1k_core.cuh"

namespace t1k {

constexpr int WARPS_PER_BLOCK = 4;
constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
// 16-byte asynchronous copy global -> shared (LDGSTS): the next index entry of a seed lands in the warp's entry table
// while the lanes work on the current allele tile; no register staging, completion by cp.async.wait_all
__device__ __forceinline__ void cp_async16(void *smemDst, const void *gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smemDst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

struct ReadsDev {
  const u64 *planes;       // [(r*4 + plane) * rwords]; planes: fwd seq2, fwd n2, rc seq2, rc n2
  int rwords;              // words per plane = read_words(longest read of the batch)
  int maxLen;              // longest read of the batch (sizes the per-lane scratch)
  const u16 *len;
  const int32_t *weight;
  const u32 *workList;     // optional indirection (deferred re-runs); NULL = identity
  u32 nWork;
};

struct AssignOut {
  Rec *store;
  unsigned long long *storeCtr;
  u64 storeCap;
  u64 *readOff;            // per read-end: first record
  u32 *readCnt;
  u32 *readTop;            // per read-end: max over its records of matchCnt << 16 | (65535 - denominator) (pairing's "is anything better" test)
  u32 *maxCnt;             // longest record list of the batch (sizes the pairing kernel's per-warp scratch)
  int32_t *readRet;        // AssignRead's return value; -2 = deferred (store full)
  int *err;
  unsigned long long *stats;   // [0] postings visited, [1] candidates, [2] tiles, [3] alleles on the hit-list path (debug/roofline)
};

// ---- AssignRead runs as rounds of four kernels over a slice of the batch (the code a warp executes in steady state stays
// small enough for the instruction caches, and work that only a few lanes of a warp would do runs item-parallel instead):
//   k_seed      warp per read-end: seeds, sweep over the allele tiles of the index, mismatch-mask evaluation (diag_fast,
//               hot mode) -> candidates in the warp's arena of the candidate pool; alleles that need a dirty-gap alignment
//               or a dirty overhang -> DeferItem queue (their candidate slot reserved)
//   k_deferred  thread per DeferItem: the full evaluation, candidate written into its slot
//   k_passes    warp per read-end: AssignRead proper (ordered scan, inclusion, > 1000 cut, records into the store);
//               full-read alignments that are not known from the seeding stage -> AlignItem queue
//   k_align     thread per AlignItem: full-read alignment (certificates / band DP), coverage, relaxedMatchCnt patched
//               into the record
// State of one read-end between the kernels of a round:
struct alignas(16) ReadState {
  u64 candOff;             // first candidate in the pool
  u32 nFwd, nCand;         // candidates of the forward strand / of both strands
  u64 bestKey;             // strand-selection key (max; bit 0: the reverse strand won)
  // per strand (0 = forward): order word (order key | candidate index) of the first candidate in list order whose extension
  // fails / succeeds (goodMatchCnt of SeqSet.hpp:2156-2186 follows from those two: the list is matchCnt-descending)
  u64 fOrd[2], rOrd[2];
};
struct DeferItem { u32 read, seqIdx; int32_t d0; u32 n, onDiag, far, at, pass; };      // at: candidate index inside the read-end's list
struct AlignItem { u32 read, cand; u64 recSlot; };                                     // recSlot: store index of the record (~0: cut away)

struct AssignParams {
  RefView R;
  ReadsDev Q;
  AssignOut O;
  Cand *candPool;          // arenaCands candidates per warp of k_seed
  u64 arenaCands;
  u32 candCap;             // most candidates one read-end may produce (both strands)
  ReadState *state;        // [reads of the batch]
  u32 *stabBuf;            // [reads][2 strands][256]: seed tables of the strands that deferred something (k_deferred reads them)
  DeferItem *dq; unsigned int *dqCtr; u32 dqCap;
  u32 *lead; unsigned int *leadCtr;                     // queue indices of the distinct deferred items (NULL: no grouping); aliases aq, which is idle then
  AlignItem *aq; unsigned int *aqCtr; u32 aqCap;
  AlignItem *dpq; unsigned int *dpqCtr; u32 dpqCap;     // AlignItems whose alignment needs the band DP (k_align -> k_align_dp)
  u8 *laneScratch;         // per lane scr_bytes(Q.maxLen)
  u32 *hitBuf;             // per warp hitCap x 32: hit lists of the alleles of a tile that take the hit-list path (lane-interleaved)
  unsigned int *workCtr;   // next position of the work list (persists over the rounds of a batch)
  u32 workBegin, workEnd;  // k_passes: the positions this round's k_seed took
  int hitCap;              // hits per allele the hit-list path holds
  int seedCap;             // seeds per strand the shared-memory tables hold (>= longest read - k + 1, multiple of 32, >= 64)
  int noFast;              // 1: every allele goes through the hit-list path (A/B switch, T1K_NO_FAST)
  int noShare;             // 1: k_deferred / k_align evaluate every item on its own (A/B switch, T1K_NO_SHARE)
  u32 queueMargin;         // k_seed takes no new read-end once the DeferItem queue is this close to full
};

// warp reductions on the redux unit (one instruction per 32-bit reduction)
__device__ __forceinline__ u32 warp_min_u32(u32 v) { return __reduce_min_sync(FULL, v); }
__device__ __forceinline__ int warp_max_i32(int v) { return __reduce_max_sync(FULL, v); }
__device__ __forceinline__ int warp_sum_i32(int v) { return __reduce_add_sync(FULL, v); }
__device__ __forceinline__ u64 warp_max_u64(u64 v) {
  const u32 hi = __reduce_max_sync(FULL, (u32)(v >> 32));
  const u32 lo = __reduce_max_sync(FULL, (u32)(v >> 32) == hi ? (u32)v : 0u);
  return ((u64)hi << 32) | lo;
}
__device__ __forceinline__ u64 warp_min_u64(u64 v) {
  const u32 hi = __reduce_min_sync(FULL, (u32)(v >> 32));
  const u32 lo = __reduce_min_sync(FULL, (u32)(v >> 32) == hi ? (u32)v : 0xffffffffu);
  return ((u64)hi << 32) | lo;
}
// lexicographic min of (key, idx) over the warp
__device__ __forceinline__ void warp_min_pair(u64 &key, int &idx) {
  const u64 k = warp_min_u64(key);
  const int i = __reduce_min_sync(FULL, key == k ? idx : 0x7fffffff);
  key = k; idx = i;
}
__device__ __forceinline__ bool pair_less(u64 k, int i, u64 fk, int fi) { return k < fk || (k == fk && i < fi); }

// ---------------------------------------------------------------------------------------------------
// ASCII reads -> 2-bit planes of both strands (rc: SeqSet::ReverseComplement, SeqSet.hpp:2103-2114)
__global__ void k_pack_reads(const char *bases, const u64 *off, const u32 *len, u32 n, int RW, u64 *planes, u16 *lenOut, int *err) {
  u32 r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const char *s = bases + off[r];
  int L = (int)len[r];
  u64 *out = planes + (size_t)r * 4 * RW;
  if (L > (RW - 1) * 32 || L > MAX_READ_LEN) { atomicOr(err, ERR_READ_LEN); L = 0; }
  lenOut[r] = (u16)L;
  u64 fs = 0, fn = 0;
  for (int w = 0; w < RW; ++w) { out[w] = 0; out[RW + w] = 0; out[2 * RW + w] = 0; out[3 * RW + w] = 0; }
  for (int j = 0; j < L; ++j) {
    char c = s[j];
    int v = c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : c == 'T' ? 3 : c == 'N' ? 4 : 5;
    if (v == 5) { atomicOr(err, ERR_READ_CHAR); v = 4; }
    int sh = (j & 31) * 2;
    fs |= (u64)(v == 4 ? 3 : v) << sh;
    fn |= (u64)(v == 4) << sh;
    if ((j & 31) == 31 || j == L - 1) { out[j >> 5] = fs; out[RW + (j >> 5)] = fn; fs = fn = 0; }
  }
  u64 rs = 0, rn = 0;
  for (int j = 0; j < L; ++j) {
    char c = s[L - 1 - j];
    int v = c == 'A' ? 3 : c == 'C' ? 2 : c == 'G' ? 1 : c == 'T' ? 0 : 4;
    int sh = (j & 31) * 2;
    rs |= (u64)(v == 4 ? 3 : v) << sh;
    rn |= (u64)(v == 4) << sh;
    if ((j & 31) == 31 || j == L - 1) { out[2 * RW + (j >> 5)] = rs; out[3 * RW + (j >> 5)] = rn; rs = rn = 0; }
  }
}

// order word of a candidate: list-order key (bits 63..21) | index in the read-end's candidate list (< 2^20).  Ascending =
// _overlap::operator< order including its tie-breaks (see order_key), one 64-bit compare / atomicMin.
__device__ __forceinline__ u64 ord_word(u64 key, int idx) { return key | (u64)(u32)idx; }
constexpr u64 ORD_NONE = ~0ull;

// ---------------------------------------------------------------------------------------------------
// k_seed.  One warp = one read-end at a time (dynamic work queue).  Shared memory per warp (seedCap = S):
//   ent[S]    the CURRENT index entry {tile, off, mask, more} of every seed (cp.async destination); during seeding the
//             same storage holds {code, postings, first entry, end entry} of every k-mer window
//   seq/nn    the strand's 2-bit planes
//   cur[S], end[S]  entry cursor / end of every seed; stab[256] seed table of diag_fast; bits[16] its bit masks
//   seedA[S]  read offset of every seed; adv[S] entries the current tile consumed from the seed
//   q0[64], q1[64]  staging of deferred alleles (DF_DEFER): {allele, diagonal, hits, hits on the diagonal}, {hits far off
//             it, reserved candidate slot}; moved to the DeferItem queue 32 at a time
//   s2[5]     the strand's seeds in the 2-bit space, lcp[256] prefix base counts (diag_hot)
struct WarpSmem {
  uint4 *ent, *q0;
  u64 *seq, *nn, *s2;
  uint2 *q1;
  u32 *cur, *end, *stab, *bits, *lcp;
  u16 *seedA, *adv;
};
constexpr int DEFER_CAP = 64;
__host__ __device__ inline size_t warp_smem_bytes(int seedCap, int rwords) {
  return ((size_t)seedCap * (16 + 4 + 4 + 2 + 2) + DEFER_CAP * 24 + (2 * rwords + 6) * 8 + 2 * 256 * 4 + 16 * 4 + 15) & ~(size_t)15;
}

// returns whether the read holds an N
__device__ __forceinline__ bool load_planes(const AssignParams &P, u32 r, int strand01, const WarpSmem &W, int lane) {
  const int RW = P.Q.rwords;
  const u64 *src = P.Q.planes + ((size_t)r * 4 + (strand01 ? 0 : 2)) * RW;
  u64 nw = 0;
  T1K_NOUNROLL
  for (int w = lane; w < 2 * RW; w += 32) {              // n2 plane follows the seq plane
    const u64 x = src[w];
    if (w < RW) W.seq[w] = x; else { nw |= x; W.nn[w - RW] = x; }
  }
  __syncwarp();
  return __any_sync(FULL, nw != 0);
}

__device__ __forceinline__ Cand void_cand(u32 seqIdx, int strand01) {
  Cand v;                                         // a reserved slot that turned out to hold nothing: skipped like CF_SEP
  v.seqIdx = (int32_t)seqIdx; v.seqStart = v.seqEnd = 0; v.readStart = v.readEnd = 0; v.strand01 = (u8)strand01; v.flags = CF_SEP;
  v.matchCnt = 0; v.eSeqStart = v.eSeqEnd = 0; v.eReadStart = v.eReadEnd = v.leftClip = v.rightClip = 0;
  v.eMatchCnt = 0; v.relaxed = 0; v.mmPos = 0;
  return v;
}

// returns the number of candidates written to cands[]
__device__ u32 seed_one_read(const AssignParams &P, u32 r, const WarpSmem &W, Cand *cands, u64 candOff, u32 *hitTile, const LaneScratch &S, int lane) {
  const RefView &R = P.R;
  const int len = P.Q.len[r];
  const int CAP = P.hitCap;
  int err = 0;
  u32 nCand = 0, nFwd = 0;
  u64 bestKey = 0;
  unsigned long long stPost = 0, stTiles = 0, stSlow = 0;
  u64 fOrdF = ORD_NONE, rOrdF = ORD_NONE, fOrdR = ORD_NONE, rOrdR = ORD_NONE;
  ReadView Qv; Qv.seq2 = W.seq; Qv.n2 = W.nn; Qv.len = len; Qv.anyN = false;

  if (len >= KMER) {
    const int NP = len - KMER + 1;
    T1K_NOUNROLL
    for (int pass = 0; pass < 2; ++pass) {
      const int strand01 = pass == 0 ? 1 : 0;
      Qv.anyN = load_planes(P, r, strand01, W, lane);
      // ---- k-mer code, posting count and entry range of every window (GetHitsFromRead, SeqSet.hpp:1093-1153)
      T1K_NOUNROLL
      for (int a = lane; a < NP; a += 32) {
        const u32 code = (u32)(fetch32(W.seq, a) & 0x3FFFFFull);
        const bool valid = (fetch32(W.nn, a) & 0x155555ull) == 0;
        const uint2 k0 = *reinterpret_cast<const uint2 *>(R.kinfo + code), k1 = *reinterpret_cast<const uint2 *>(R.kinfo + code + 1);
        W.ent[a] = make_uint4(code, valid ? k1.y - k0.y : 0u, k0.x, valid ? k1.x : k0.x);
      }
      __syncwarp();
      // ---- the sequential skip rule (list >= 100, not first/last, <= K/2 in a row; Q2)
      // The strand is eligible for diag_fast when the read holds no N, fits 5 words and no seed is a homopolymer k-mer
      // (see the claim above diag_fast).
      int nS = 0;
      bool strandFast = !Qv.anyN && len <= FAST_MAX_LEN && !P.noFast;
      if (lane == 0) {
        u32 prev = 0; int skip = 0;
        T1K_NOUNROLL
        for (int a = 0; a < NP; ++a) {
          const uint4 w = W.ent[a];
          if (a == 0 || prev != w.x) {
            const int size = (int)w.y;
            if (size >= 100 && a != 0 && a != NP - 1 && skip < KMER / 2) { ++skip; continue; }
            skip = 0;
            if (size > 0) { W.seedA[nS] = (u16)a; W.cur[nS] = w.z; W.end[nS] = w.w; ++nS; stPost += w.y; if (kmer_homopolymer(w.x)) strandFast = false; }
          }
          prev = w.x;
        }
      }
      nS = __shfl_sync(FULL, nS, 0);
      strandFast = __shfl_sync(FULL, (int)strandFast, 0) != 0;
      u32 lcMemo = 0;
      __syncwarp();
      if (strandFast) {
        // seed table (see seed_table_build), warp-cooperative: seed / wide-step bit masks, then every lane derives the
        // entries of its read positions from the masks
        if (lane < 16) W.bits[lane] = 0;
        __syncwarp();
        T1K_NOUNROLL
        for (int k = lane; k < nS; k += 32) {
          const int a = W.seedA[k];
          atomicOr(&W.bits[a >> 5], 1u << (a & 31));
          if (k > 0 && a - (int)W.seedA[k - 1] > KMER - 1) atomicOr(&W.bits[8 + (a >> 5)], 1u << (a & 31));
        }
        __syncwarp();
        T1K_NOUNROLL
        for (int a = lane; a < len; a += 32) {
          const int wi = a >> 5, bit = a & 31;             // wi is warp-uniform
          u32 cnt = 0, big = 0;
          T1K_NOUNROLL
          for (int w = 0; w < wi; ++w) { cnt += __popc(W.bits[w]); big += __popc(W.bits[8 + w]); }
          const u32 upTo = 0xffffffffu >> (31 - bit);
          const u32 cw = W.bits[wi] & upTo;
          cnt += __popc(cw); big += __popc(W.bits[8 + wi] & upTo);
          u32 last = 255, nxt = 255;
          {
            int w = wi; u32 m = cw;
            T1K_NOUNROLL
            for (;;) { if (m) { last = (u32)(w * 32 + 31 - __clz(m)); break; } if (--w < 0) break; m = W.bits[w]; }
          }
          {
            int w = wi; u32 m = W.bits[wi] & (0xffffffffu << bit);
            T1K_NOUNROLL
            for (;;) { if (m) { nxt = (u32)(w * 32 + __ffs(m) - 1); break; } if (++w >= 8) break; m = W.bits[w]; }
          }
          W.stab[a] = cnt | (big << 8) | (nxt << 16) | (last << 24);
        }
      }
      if (strandFast) {
        // diag_hot's strand tables: seeds as bits of the 2-bit space, prefix base counts
        u32 *s2w = (u32 *)W.s2;
        if (lane < 10) s2w[lane] = 0;
        __syncwarp();
        T1K_NOUNROLL
        for (int k = lane; k < nS; k += 32) { const int a = W.seedA[k]; atomicOr(&s2w[a >> 4], 1u << (2 * (a & 15))); }
        T1K_NOUNROLL
        for (int a = lane; a <= len; a += 32) {
          const int wi = a >> 5, bit = a & 31;
          u32 c0 = 0, c1 = 0, c2 = 0, c3 = 0;
          T1K_NOUNROLL
          for (int w = 0; w <= wi; ++w) {
            const u64 x = W.seq[w], keep = w < wi ? M55 : (lowmask2(bit) & M55);
            const u64 lo = x & M55, hi = (x >> 1) & M55;
            c0 += __popcll(~lo & ~hi & keep); c1 += __popcll(lo & ~hi & keep); c2 += __popcll(~lo & hi & keep); c3 += __popcll(lo & hi & keep);
          }
          W.lcp[a] = c0 | (c1 << 8) | (c2 << 16) | (c3 << 24);
        }
      }
      // first entry of every seed (the window table is dead now)
      T1K_NOUNROLL
      for (int k = lane; k < nS; k += 32) cp_async16(&W.ent[k], R.entries + W.cur[k]);
      u64 laneKey = 0;
      u64 fOrd = ORD_NONE, rOrd = ORD_NONE;
      auto note = [&](const Cand &c, int i) {
        if (c.flags & CF_SEP) return;
        const u64 k = ord_word(cand_key_pre(c), i);
        if (c.flags & CF_RET) rOrd = min(rOrd, k); else fOrd = min(fOrd, k);
      };
      // deferred alleles leave the warp `take` at a time: into the DeferItem queue, or — queue full — evaluated right here
      int qn = 0;
      bool stabSaved = false;
      auto flush_deferred = [&](int take) {
        unsigned int base = 0;
        if (lane == 0) base = atomicAdd(P.dqCtr, (unsigned int)take);
        base = __shfl_sync(FULL, base, 0);
        if (base + (unsigned int)take <= P.dqCap) {
          if (!stabSaved) {          // k_deferred needs this strand's seed table
            u32 *dst = P.stabBuf + ((size_t)r * 2 + pass) * 256;
            T1K_NOUNROLL
            for (int a = lane; a < len; a += 32) dst[a] = W.stab[a];
            stabSaved = true;
          }
          if (lane < take) {
            const uint4 e0 = W.q0[qn - take + lane];
            const uint2 e1 = W.q1[qn - take + lane];
            DeferItem it;
            it.read = r; it.seqIdx = e0.x; it.d0 = (int32_t)e0.y; it.n = e0.z; it.onDiag = e0.w; it.far = e1.x; it.at = e1.y; it.pass = (u32)pass;
            P.dq[base + lane] = it;
          }
        } else if (lane < take) {
          // (the counter has moved on: the slots of this batch that still lie inside the queue become no-ops)
          if (base + (unsigned int)lane < P.dqCap) { DeferItem nop; nop.read = 0xffffffffu; nop.seqIdx = 0; nop.d0 = 0; nop.n = nop.onDiag = nop.far = nop.at = nop.pass = 0; P.dq[base + lane] = nop; }
          const uint4 e0 = W.q0[qn - take + lane];
          const uint2 e1 = W.q1[qn - take + lane];
          Cand c;
          bool em = false;
          const int df = diag_fast(R, Qv, strand01, (int)e0.x, (int)e0.z, (int)e0.y, (int)e0.w, (int)e1.x, W.stab, false, c, em, laneKey, lcMemo, S, err);
          if (df == DF_DECLINED) err |= ERR_SCRATCH;        // (cannot happen: the hit-count certificate passed in hot mode)
          if (df == DF_DONE && em) {
            if (!(c.flags & CF_PRE)) { extend_cand<false>(R, Qv, c, S, err); c.flags |= CF_PRE; }     // a long or dirty overhang
            cands[e1.y] = c;
            note(c, (int)e1.y);
          } else cands[e1.y] = void_cand(e0.x, strand01);
        }
        qn -= take;
        __syncwarp();
      };
      // ---- allele tiles (32 consecutive allele ids) in ascending order: the smallest tile any seed still has an entry for
      T1K_NOUNROLL
      for (;;) {
        cp_async_wait_all();
        __syncwarp();
        u32 mn = 0xffffffffu;
        T1K_NOUNROLL
        for (int k = lane; k < nS; k += 32) mn = min(mn, W.ent[k].x);
        const u32 T = warp_min_u32(mn);
        if (T == 0xffffffffu) break;
        ++stTiles;
        // sweep 1, lane = allele T*32 + lane: number of hits n, diagonal of the first hit, hits on / far off that diagonal.
        // The seeds are visited in read order and a seed's entries in offset order = the order of the allele's hit list.
        int n = 0, d0 = 0, onDiag = 0, far = 0;
        T1K_NOUNROLL
        for (int k0 = 0; k0 < nS; k0 += 32) {
          const int k = k0 + lane;
          uint4 me = make_uint4(0xffffffffu, 0, 0, 0);
          if (k < nS) me = W.ent[k];
          const bool act = me.x == T;
          unsigned bal = __ballot_sync(FULL, act);
          if (bal == 0) continue;
          // Usual case: every active seed of the chunk has ONE entry in this tile and all lie on one diagonal D (the alleles of
          // a tile are neighbours of one gene).  Then lane = seed holds a 32-allele mask, and the per-allele hit counts are the
          // column sums of that 32 x 32 bit matrix: transposed through five butterfly shuffles instead of walking the seeds.
          const int dgk = (int)me.y - (int)W.seedA[k < nS ? k : 0];
          const int D = __shfl_sync(FULL, dgk, __ffs(bal) - 1);
          if (!__any_sync(FULL, act && (me.w != 0 || dgk != D))) {
            u32 x = act ? me.z : 0u;
            u32 m = 0x0000ffffu;
#pragma unroll
            for (int j = 16; j; j >>= 1) {
              const u32 y = __shfl_xor_sync(FULL, x, j);
              x = (lane & j) ? ((x & ~m) | ((y >> j) & m)) : ((x & m) | ((y << j) & ~m));
              m ^= m << (j >> 1);
            }
            const int c = __popc(x);                    // hits of allele T*32 + lane from this chunk's seeds, all on diagonal D
            if (c) {
              if (n == 0) d0 = D;
              const int dd = D - d0;
              n += c; onDiag += dd == 0 ? c : 0; far += ((dd > RADIUS) | (dd < -RADIUS)) ? c : 0;
            }
            continue;
          }
          T1K_NOUNROLL
          while (bal) {
            const int kk = k0 + __ffs(bal) - 1;
            bal &= bal - 1;
            uint4 e = W.ent[kk];                        // broadcast
            const int a = W.seedA[kk];
            const u32 more = e.w;
            T1K_NOUNROLL
            for (u32 j = 0;; ++j) {
              if ((e.z >> lane) & 1u) {
                const int dg = (int)e.y - a;
                if (n == 0) d0 = dg;
                const int dd = dg - d0;
                ++n; onDiag += dd == 0; far += (dd > RADIUS) | (dd < -RADIUS);
              }
              if (j >= more) break;
              e = *reinterpret_cast<const uint4 *>(R.entries + W.cur[kk] + j + 1);     // same k-mer at another offset inside this tile
            }
          }
        }
        __syncwarp();
        // the consumed seeds move on; their next entries arrive while the lanes work on this tile
        T1K_NOUNROLL
        for (int k = lane; k < nS; k += 32) {
          const uint4 e = W.ent[k];
          u32 adv = 0;
          if (e.x == T) {
            adv = 1 + e.w;
            const u32 c = W.cur[k] + adv;
            W.cur[k] = c;
            if (c < W.end[k]) cp_async16(&W.ent[k], R.entries + c); else W.ent[k].x = 0xffffffffu;
          }
          W.adv[k] = (u16)adv;
        }
        // ---- lane per allele: the mismatch-mask path first (uniform work); the few alleles it declines get their hit list
        int nEmit = 0;
        Cand fc;
        bool fastEmit = false, handled = n < 3, defer = false;
        if (strandFast && n >= 3) {
          const int df = diag_hot(R, Qv, strand01, (int)(T * 32 + lane), n, d0, onDiag, far, W.s2, W.lcp, fc, fastEmit, laneKey);
          handled = df != DF_DECLINED;
          defer = df == DF_DEFER;          // needs work the other lanes do not: queued, its candidate slot reserved
          if (fastEmit || defer) nEmit = 1;
        }
        if (__any_sync(FULL, !handled)) {
          // sweep 2: hit lists (readOffset | seqOffset << 8, in (readOffset, seqOffset) order) of the declined alleles into the
          // warp's lane-interleaved tile in HBM scratch, from the entries the tile just consumed
          __syncwarp();
          int cnt = 0;
          T1K_NOUNROLL
          for (int k0 = 0; k0 < nS; k0 += 32) {
            const int k = k0 + lane;
            unsigned bal = __ballot_sync(FULL, k < nS && W.adv[k] != 0);
            T1K_NOUNROLL
            while (bal) {
              const int kk = k0 + __ffs(bal) - 1;
              bal &= bal - 1;
              const u32 a = W.seedA[kk], nAdv = W.adv[kk], c1 = W.cur[kk];
              T1K_NOUNROLL
              for (u32 j = 0; j < nAdv; ++j) {
                const uint4 e = *reinterpret_cast<const uint4 *>(R.entries + (c1 - nAdv + j));
                if (!handled && ((e.z >> lane) & 1u)) {
                  if (cnt < CAP) hitTile[(size_t)cnt * 32 + lane] = hit_make((int)a, e.y);
                  ++cnt;
                }
              }
            }
          }
          if (!handled) {
            ++stSlow;
            if (cnt > CAP) err |= ERR_HITS;
            else {
              chain_allele(R, Qv, strand01, (int)(T * 32 + lane), hitTile + lane, 32, n, S, nEmit, laneKey, err);
              if (nEmit > 1) sort_emitted(S.emit(), nEmit);       // several clusters on one allele: the tail order of _overlap::operator<
              Cand *em = S.emit();
              T1K_NOUNROLL
              for (int j = 0; j < nEmit; ++j) { Cand c = em[j]; c.mmPos = 0; extend_cand<false>(R, Qv, c, S, err); c.flags |= CF_PRE; em[j] = c; }
            }
          }
        }
        // ---- ordered emission (allele order == lane order)
        int incl = nEmit;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(FULL, incl, o); if (lane >= o) incl += t; }
        const int tot = __shfl_sync(FULL, incl, 31);
        if (tot > 0) {
          if (nCand + tot > P.candCap) err |= ERR_CAND;
          else {
            const int at = (int)nCand + incl - nEmit;
            const unsigned balD = __ballot_sync(FULL, defer);
            if (defer) {
              const int q = qn + __popc(balD & ((1u << lane) - 1));
              W.q0[q] = make_uint4(T * 32 + lane, (u32)d0, (u32)n, (u32)onDiag);
              W.q1[q] = make_uint2((u32)far, (u32)at);
            } else if (fastEmit) { cands[at] = fc; note(fc, at); }
            else {
              const Cand *em = S.emit();
              T1K_NOUNROLL
              for (int j = 0; j < nEmit; ++j) { const Cand c = em[j]; cands[at + j] = c; note(c, at + j); }
            }
            nCand += tot;
            qn += __popc(balD);
          }
        }
        if (qn >= 32) { __syncwarp(); flush_deferred(32); }
      }
      T1K_NOUNROLL
      while (qn > 0) { __syncwarp(); flush_deferred(min(qn, 32)); }
      bestKey = max(bestKey, warp_max_u64(laneKey));
      fOrd = warp_min_u64(fOrd);
      rOrd = warp_min_u64(rOrd);
      if (pass == 0) { fOrdF = fOrd; rOrdF = rOrd; nFwd = nCand; }
      else { fOrdR = fOrd; rOrdR = rOrd; }
    }
  }
  if (lane == 0) {
    ReadState st;
    st.candOff = candOff; st.nFwd = nFwd; st.nCand = nCand; st.bestKey = bestKey;
    st.fOrd[0] = fOrdF; st.fOrd[1] = fOrdR; st.rOrd[0] = rOrdF; st.rOrd[1] = rOrdR;
    P.state[r] = st;
  }
  err = __reduce_or_sync(FULL, (unsigned)err);
  if (lane == 0 && err) atomicOr(P.O.err, err);
  if (P.O.stats) { stSlow = __reduce_add_sync(FULL, (u32)stSlow); }
  if (lane == 0 && P.O.stats) {
    atomicAdd(P.O.stats + 0, stPost);
    atomicAdd(P.O.stats + 1, (unsigned long long)nCand);
    atomicAdd(P.O.stats + 2, stTiles);
    atomicAdd(P.O.stats + 3, stSlow);
  }
  return nCand;
}

extern __shared__ u64 t1k_smem[];

// MINB = resident blocks per SM the register budget is compiled for
template <int MINB>
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32, MINB) k_seed(AssignParams P) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t gwarp = (size_t)blockIdx.x * WARPS_PER_BLOCK + warp;
  const int SC = P.seedCap;
  const int RW = P.Q.rwords;
  u8 *sm = (u8 *)t1k_smem + (size_t)warp * warp_smem_bytes(SC, RW);
  WarpSmem W;
  W.ent = (uint4 *)sm; W.q0 = W.ent + SC;
  W.seq = (u64 *)(W.q0 + DEFER_CAP); W.nn = W.seq + RW; W.s2 = W.nn + RW;
  W.q1 = (uint2 *)(W.s2 + 6);
  W.cur = (u32 *)(W.q1 + DEFER_CAP); W.end = W.cur + SC; W.stab = W.end + SC; W.lcp = W.stab + 256; W.bits = W.lcp + 256;
  W.seedA = (u16 *)(W.bits + 16); W.adv = W.seedA + SC;
  const LaneScratch S = lane_scratch(P.laneScratch + (gwarp * 32 + lane) * scr_bytes(P.Q.maxLen), P.Q.maxLen);
  u32 *hitTile = P.hitBuf + gwarp * (size_t)P.hitCap * 32;
  const u64 arena0 = gwarp * P.arenaCands;
  u64 used = 0;
  // a warp takes read-ends while its arena of the candidate pool can hold the most a read-end may produce
  while (used + P.candCap <= P.arenaCands) {
    // (soft limit: when the item queues are nearly full the round ends early rather than falling back to in-place evaluation)
    if (*(volatile unsigned int *)P.dqCtr + P.queueMargin > P.dqCap) break;
    u32 w = 0;
    if (lane == 0) w = atomicAdd(P.workCtr, 1u);
    w = __shfl_sync(FULL, w, 0);
    if (w >= P.Q.nWork) break;
    const u32 r = P.Q.workList ? P.Q.workList[w] : w;
    used += seed_one_read(P, r, W, P.candPool + arena0 + used, arena0 + used, hitTile, S, lane);
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------------------
// The deferred alleles in three steps.  The items that are queued together are mostly consecutive alleles of ONE read-end (a
// tile's deferred lanes are queued in lane order), and the alleles of a gene are nearly identical: inside the window the read
// covers, most of them hold the very same bases (KIR-DNA-like set: 5.4 items per distinct one within a warp).  Everything
// diag_fast + extend_cand read of an allele lies inside that window (plus its length), so items whose (read-end, strand,
// diagonal, hit counts, allele length, window bases[, exon mask]) are EQUAL — compared word for word, the hash only forms the
// groups — get the same candidate up to seqIdx.  Identical lanes of a warp cost what one lane costs, so the saving comes from
// COMPACTING the distinct items:
//   k_defer_group  warp per 32 consecutive items: groups equal items, the lowest lane of a group becomes its leader; leaders
//                  are appended to a dense list, a member keeps the queue index of its leader (n) and is flagged (pass bit 31)
//   k_deferred     thread per leader: the full evaluation, candidate into its slot; status and strand key left in the item
//   k_defer_copy   thread per member: the leader's candidate with its own seqIdx
// T1K_NO_SHARE=1 (P.noShare): k_deferred alone over every item (A/B switch).
__device__ __forceinline__ u64 mix64(u64 h, u64 v) { h ^= v + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2); return h * 0xff51afd7ed558ccdull; }
constexpr u32 DEFER_MEMBER = 0x80000000u;

__global__ void __launch_bounds__(128) k_defer_group(AssignParams P) {
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nThreads = (size_t)gridDim.x * blockDim.x;
  const int lane = threadIdx.x & 31;
  const u32 nItems = min(*P.dqCtr, P.dqCap);
  const RefView &R = P.R;
  for (size_t i0 = tid - lane; i0 < nItems; i0 += nThreads) {        // warp-uniform: 32 consecutive items per round
    const size_t i = i0 + lane;
    DeferItem it;
    bool valid = i < nItems;
    if (valid) { it = P.dq[i]; valid = it.read != 0xffffffffu; }
    if (!valid) { it.read = 0xffffffffu; it.seqIdx = 0; it.d0 = 0; it.n = it.onDiag = it.far = it.at = it.pass = 0; }
    // ---- signature of what the evaluation reads
    u64 w[5] = {0, 0, 0, 0, 0}, x[5] = {0, 0, 0, 0, 0};
    int clen = 0;
    bool shareable = false;
    u64 h = ~(u64)lane;
    if (valid) {
      const uint4 mt = *reinterpret_cast<const uint4 *>(R.meta + it.seqIdx);
      const u64 w0 = (u64)mt.x | ((u64)mt.y << 32);
      clen = (int)mt.z;
      const int len = P.Q.len[it.read];
      const int pLo = it.d0 < 0 ? -it.d0 : 0, pHi = imin(len, clen - it.d0), Wn = pHi - pLo;
      if (Wn >= KMER && Wn <= 160) {
        shareable = true;
#pragma unroll
        for (int j = 0; j < 5; ++j)
          if (32 * j < Wn) {
            w[j] = fetch32(R.seq2 + w0, pLo + it.d0 + 32 * j) & lowmask2(Wn - 32 * j);
            if (R.relax) x[j] = fetch32(R.ex2 + w0, pLo + it.d0 + 32 * j) & lowmask2(Wn - 32 * j);
          }
        h = mix64(mix64((u64)it.read << 1 | it.pass, (u64)(u32)it.d0 | ((u64)it.n << 32)), (u64)it.onDiag | ((u64)it.far << 20) | ((u64)(u32)clen << 40));
#pragma unroll
        for (int j = 0; j < 5; ++j) h = mix64(h, w[j] ^ (x[j] * 0x9e3779b97f4a7c15ull));
        h &= ~(1ull << 63);              // (never equal to an unshareable lane's ~lane)
      }
    }
    const unsigned grp = __match_any_sync(FULL, h);
    int leader = __ffs(grp) - 1;
    {
      // word-for-word check against the group's lowest lane; a lane that differs (hash collision) evaluates for itself
      bool same = shareable;
      same &= __shfl_sync(FULL, it.read, leader) == it.read && __shfl_sync(FULL, it.pass, leader) == it.pass && __shfl_sync(FULL, it.d0, leader) == it.d0;
      same &= __shfl_sync(FULL, it.n, leader) == it.n && __shfl_sync(FULL, it.onDiag, leader) == it.onDiag && __shfl_sync(FULL, it.far, leader) == it.far;
      same &= __shfl_sync(FULL, clen, leader) == clen;
#pragma unroll
      for (int j = 0; j < 5; ++j) { same &= __shfl_sync(FULL, w[j], leader) == w[j]; same &= __shfl_sync(FULL, x[j], leader) == x[j]; }
      if (!same) leader = lane;
    }
    const bool isLeader = valid && leader == lane;
    const unsigned bal = __ballot_sync(FULL, isLeader);
    unsigned int base = 0;
    if (lane == 0 && bal) base = atomicAdd(P.leadCtr, (unsigned int)__popc(bal));
    base = __shfl_sync(FULL, base, 0);
    if (isLeader) P.lead[base + __popc(bal & ((1u << lane) - 1))] = (u32)i;
    else if (valid) { P.dq[i].n = (u32)(i0 + leader); P.dq[i].pass = it.pass | DEFER_MEMBER; }
  }
  if (tid == 0) atomicAdd(P.O.stats + 4, (unsigned long long)nItems);
}

// thread per leader (P.lead != NULL) or per item
__global__ void __launch_bounds__(128) k_deferred(AssignParams P) {
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nThreads = (size_t)gridDim.x * blockDim.x;
  const u32 nWork = P.lead ? *P.leadCtr : min(*P.dqCtr, P.dqCap);
  const LaneScratch S = lane_scratch(P.laneScratch + tid * scr_bytes(P.Q.maxLen), P.Q.maxLen);
  const RefView &R = P.R;
  int err = 0;
  for (size_t k = tid; k < nWork; k += nThreads) {
    const size_t i = P.lead ? P.lead[k] : k;
    const DeferItem it = P.dq[i];
    if (it.read == 0xffffffffu) continue;
    const int strand01 = it.pass == 0 ? 1 : 0;
    const u64 *pl = P.Q.planes + ((size_t)it.read * 4 + (strand01 ? 0 : 2)) * P.Q.rwords;
    ReadView Q; Q.seq2 = pl; Q.n2 = pl + P.Q.rwords; Q.len = P.Q.len[it.read]; Q.anyN = false;     // (deferring strands hold no N)
    ReadState *st = P.state + it.read;
    Cand *slot = P.candPool + st->candOff + it.at;
    Cand c;
    bool em = false;
    u64 laneKey = 0; u32 lcMemo = 0;
    const int df = diag_fast(R, Q, strand01, (int)it.seqIdx, (int)it.n, it.d0, (int)it.onDiag, (int)it.far,
                             P.stabBuf + ((size_t)it.read * 2 + it.pass) * 256, false, c, em, laneKey, lcMemo, S, err);
    if (df == DF_DECLINED) err |= ERR_SCRATCH;        // (cannot happen: the hit-count certificate passed in hot mode)
    const bool done = df == DF_DONE && em;
    if (done) {
      if (!(c.flags & CF_PRE)) { extend_cand<false>(R, Q, c, S, err); c.flags |= CF_PRE; }     // a long or dirty overhang
      *slot = c;
      if (!(c.flags & CF_SEP)) {
        const u64 kk = ord_word(cand_key_pre(c), (int)it.at);
        atomicMin((unsigned long long *)((c.flags & CF_RET) ? &st->rOrd[it.pass] : &st->fOrd[it.pass]), (unsigned long long)kk);
      }
    } else *slot = void_cand(it.seqIdx, strand01);
    if (laneKey) atomicMax((unsigned long long *)&st->bestKey, (unsigned long long)laneKey);
    if (P.lead) {                                     // what the members of this leader's group need besides the candidate
      DeferItem *o = P.dq + i;
      o->n = done ? 1u : 0u; o->onDiag = (u32)laneKey; o->far = (u32)(laneKey >> 32);
    }
  }
  if (err) atomicOr(P.O.err, err);
  if (tid == 0) atomicAdd(P.O.stats + 5, (unsigned long long)nWork);
}

__global__ void __launch_bounds__(256) k_defer_copy(AssignParams P) {
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nThreads = (size_t)gridDim.x * blockDim.x;
  const u32 nItems = min(*P.dqCtr, P.dqCap);
  for (size_t i = tid; i < nItems; i += nThreads) {
    const DeferItem it = P.dq[i];
    if (it.read == 0xffffffffu || !(it.pass & DEFER_MEMBER)) continue;
    const u32 pass = it.pass & ~DEFER_MEMBER;
    const DeferItem L = P.dq[it.n];                   // same read-end, strand and diagonal
    ReadState *st = P.state + it.read;
    const Cand *cands = P.candPool + st->candOff;
    Cand *slot = P.candPool + st->candOff + it.at;
    if (L.n) {
      Cand c = cands[L.at];
      c.seqIdx = (int32_t)it.seqIdx;
      *slot = c;
      if (!(c.flags & CF_SEP)) {
        const u64 kk = ord_word(cand_key_pre(c), (int)it.at);
        atomicMin((unsigned long long *)((c.flags & CF_RET) ? &st->rOrd[pass] : &st->fOrd[pass]), (unsigned long long)kk);
      }
    } else *slot = void_cand(it.seqIdx, pass == 0 ? 1 : 0);
    const u64 keyL = (u64)L.onDiag | ((u64)L.far << 32);
    if (keyL) atomicMax((unsigned long long *)&st->bestKey, (unsigned long long)((keyL & ~(0xFFFFFFull << 1)) | ((u64)(0xFFFFFFu - it.seqIdx) << 1)));
  }
}

// ---------------------------------------------------------------------------------------------------
// k_passes: AssignRead proper (SeqSet.hpp:2132-2300) on the best strand's candidates, warp per read-end
__device__ void passes_one_read(const AssignParams &P, u32 r, u32 *qCand, u64 *qSlot, const LaneScratch &S, int lane) {
  const RefView &R = P.R;
  const ReadState st = P.state[r];
  const int weight = P.Q.weight[r];
  const Cand *cands = P.candPool + st.candOff;
  int err = 0;
  const int best01 = (st.bestKey & 1) ? 0 : 1;
  const int c0 = best01 ? 0 : (int)st.nFwd, c1 = best01 ? (int)st.nFwd : (int)st.nCand;
  const u64 fOrd = best01 ? st.fOrd[0] : st.fOrd[1], rOrd = best01 ? st.rOrd[0] : st.rOrd[1];
  int ret = -1, nFinal = 0;
  unsigned long long pos = 0;
  bool deferred = false;
  if (c1 - c0 > 0) {
    const int RW = P.Q.rwords;
    const u64 *pl = P.Q.planes + ((size_t)r * 4 + (best01 ? 0 : 2)) * RW;
    ReadView Qv; Qv.seq2 = pl; Qv.n2 = pl + RW; Qv.len = P.Q.len[r];
    { u64 nw = 0; for (int w = lane; w < RW; w += 32) nw |= pl[RW + w]; Qv.anyN = __any_sync(FULL, nw != 0); }
    // goodMatchCnt (SeqSet.hpp:2156-2186) = the largest matchCnt among the returned candidates that precede the first
    // failing one = the first returned candidate if it precedes the failure, and nothing otherwise.
    const int good = rOrd < fOrd ? order_key_mc(rOrd) : -1;
    auto included = [&](const Cand &c, int i) {          // the ordered scan's verdict on one candidate (SeqSet.hpp:2163-2186)
      if ((c.flags & CF_SEP) || !(c.flags & CF_RET)) return false;
      if (ord_word(cand_key_pre(c), i) < fOrd) return true;
      return !((int)c.matchCnt < good && (!(c.flags & CF_NEEDCLIP) || sim_below(R, c.matchCnt, cand_denom_pre(c), 1)));
    };
    // pass A (read-only): what is kept, the best matchCnt, the head of the post-extension order and the smallest
    // similarity (as an exact fraction), which tells whether the > 1000 cut (SeqSet.hpp:2290-2298) removes anything
    int bestMc = -1, nInc = 0;
    u64 bOrd = ORD_NONE;
    int minNum = 1, minDen = 0;                          // minDen == 0: none yet
    T1K_NOUNROLL
    for (int i = c0 + lane; i < c1; i += 32) {
      if (i + 64 < c1) prefetch_l2(&cands[i + 64]);
      const Cand c = cands[i];
      if (!included(c, i)) continue;
      bestMc = max(bestMc, c.eMatchCnt);
      ++nInc;
      bOrd = min(bOrd, ord_word(cand_key_post(c), i));
      const int den = cand_denom_post(c);
      if (minDen == 0 || (long long)c.eMatchCnt * minDen < (long long)minNum * den) { minNum = c.eMatchCnt; minDen = den; }
    }
    bestMc = warp_max_i32(bestMc);
    nInc = warp_sum_i32(nInc);
    bOrd = warp_min_u64(bOrd);
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      const int on = __shfl_xor_sync(FULL, minNum, o), od = __shfl_xor_sync(FULL, minDen, o);
      if (od != 0 && (minDen == 0 || (long long)on * minDen < (long long)minNum * od)) { minNum = on; minDen = od; }
    }
    // reserve the store before touching coverage, so that a full store can be retried without double counting
    if (lane == 0 && nInc > 0) pos = atomicAdd(P.O.storeCtr, (unsigned long long)nInc);
    pos = __shfl_sync(FULL, pos, 0);
    if (nInc > 0 && pos + nInc > P.O.storeCap) deferred = true;
    if (!deferred && nInc > 0) {
      const bool usePost = nInc > 1000;      // SeqSet.hpp:2290-2298
      u64 cOrd = ORD_NONE;                   // first candidate (post order) the cut removes
      if (usePost) {
        const int bIdx = (int)(bOrd & 0xFFFFFu);
        const Cand cb = cands[bIdx];
        const double cutSim = (double)cb.eMatchCnt / (double)cand_denom_post(cb) - 0.1;
        if ((double)minNum / (double)minDen < cutSim) {
          T1K_NOUNROLL
          for (int i = c0 + lane; i < c1; i += 32) {
            const Cand c = cands[i];
            if (!included(c, i) || i == bIdx) continue;
            if ((double)c.eMatchCnt / (double)cand_denom_post(c) < cutSim) cOrd = min(cOrd, ord_word(cand_key_post(c), i));
          }
          cOrd = warp_min_u64(cOrd);
        }
      }
      // pass B: ordered compaction into the store (allele order is kept: pairing searches it), minus the cut, and the
      // full-read alignment + coverage of everything within 10 of the best (Q8).  Alignments the seeding stage already knows
      // (CF_FA) cost nothing; the others leave as AlignItems (k_align patches relaxedMatchCnt into the record), or —
      // queue full — are aligned here 32 at a time.
      const bool doAlign = weight >= 0;
      int running = 0, qn = 0;
      u32 top = 0;
      auto emit_rec = [&](const Cand &c, u64 slot) {
        Rec o;
        o.seqIdx = c.seqIdx; o.seqStart = c.eSeqStart; o.seqEnd = c.eSeqEnd;
        o.readStart = c.eReadStart; o.readEnd = c.eReadEnd; o.leftClip = c.leftClip; o.rightClip = c.rightClip;
        o.mcx = rec_mcx(c.eMatchCnt, c.relaxed, c.strand01);
        o.key = usePost ? (cand_key_post(c) | 1ull) : cand_key_pre(c);
        P.O.store[slot] = o;
      };
      auto flush_cold = [&](int take) {
        unsigned int base = 0;
        if (lane == 0) base = atomicAdd(P.aqCtr, (unsigned int)take);
        base = __shfl_sync(FULL, base, 0);
        if (lane < take) {
          const u32 j = qCand[qn - take + lane];
          const u64 sl = qSlot[qn - take + lane];
          if (base + (unsigned int)take <= P.aqCap) {
            AlignItem it; it.read = r; it.cand = j; it.recSlot = sl;
            P.aq[base + lane] = it;
          } else {
            if (base + (unsigned int)lane < P.aqCap) { AlignItem nop; nop.read = 0xffffffffu; nop.cand = 0; nop.recSlot = ~0ull; P.aq[base + lane] = nop; }
            Cand cc = cands[j];
            full_align<false>(R, Qv, cc, weight, S, err);
            if (sl != ~0ull) emit_rec(cc, sl);
          }
        }
        qn -= take;
        __syncwarp();
      };
      T1K_NOUNROLL
      for (int b = c0; b < c1; b += 32) {
        const int i = b + lane;
        bool inc = false, cold = false, wr = false;
        Cand c;
        if (i + 64 < c1) prefetch_l2(&cands[i + 64]);
        if (i < c1) {
          c = cands[i];
          inc = included(c, i);
          if (inc) {
            wr = !usePost || ord_word(cand_key_post(c), i) < cOrd;
            if (doAlign) {
              if (c.eMatchCnt < bestMc - 10) c.relaxed = 0;
              else if (c.flags & CF_FA) full_align_known(R, c, weight);
              else cold = true;
            }
          }
        }
        const unsigned balW = __ballot_sync(FULL, wr), balC = __ballot_sync(FULL, cold);
        const u64 slot = pos + (u64)running + __popc(balW & ((1u << lane) - 1));
        if (wr) {
          top = max(top, ((u32)c.eMatchCnt << 16) | (u32)(65535 - cand_denom_post(c)));
          emit_rec(c, slot);                         // (a cold candidate's relaxedMatchCnt is patched in by k_align)
        }
        if (cold) {
          const int q = qn + __popc(balC & ((1u << lane) - 1));
          qCand[q] = (u32)i; qSlot[q] = wr ? slot : ~0ull;
        }
        running += __popc(balW);
        qn += __popc(balC);
        __syncwarp();
        if (qn >= 32) flush_cold(32);
      }
      T1K_NOUNROLL
      while (qn > 0) flush_cold(min(qn, 32));
      top = __reduce_max_sync(FULL, top);
      if (lane == 0) P.O.readTop[r] = top;
      nFinal = running;
      ret = nFinal;
    } else if (!deferred) { ret = 0; if (lane == 0) P.O.readTop[r] = 0; }
  }
  if (lane == 0) {
    P.O.readOff[r] = pos;
    P.O.readCnt[r] = deferred ? 0 : (u32)nFinal;
    P.O.readRet[r] = deferred ? -2 : ret;
    if (nFinal > 0 && !deferred) atomicMax(P.O.maxCnt, (u32)nFinal);
    if (deferred) atomicOr(P.O.err, ERR_STORE);
  }
  err = __reduce_or_sync(FULL, (unsigned)err);
  if (lane == 0 && err) atomicOr(P.O.err, err);
}

__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32) k_passes(AssignParams P) {
  __shared__ u32 qCandS[WARPS_PER_BLOCK][DEFER_CAP];
  __shared__ u64 qSlotS[WARPS_PER_BLOCK][DEFER_CAP];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t gwarp = (size_t)blockIdx.x * WARPS_PER_BLOCK + warp, nWarps = (size_t)gridDim.x * WARPS_PER_BLOCK;
  const LaneScratch S = lane_scratch(P.laneScratch + (gwarp * 32 + lane) * scr_bytes(P.Q.maxLen), P.Q.maxLen);
  for (size_t w = P.workBegin + gwarp; w < P.workEnd; w += nWarps) {
    const u32 r = P.Q.workList ? P.Q.workList[w] : (u32)w;
    passes_one_read(P, r, qCandS[warp], qSlotS[warp], S, lane);
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------------------
// k_align: thread per AlignItem — mismatch count, certificates, coverage.  The few items only the band DP can decide are
// compacted into a second queue and run by k_align_dp with every lane on a DP of its own.
__device__ __forceinline__ void align_item(const AssignParams &P, const AlignItem &it, const LaneScratch &S, int &err, bool allowDp, bool &done) {
  const RefView &R = P.R;
  const ReadState *st = P.state + it.read;
  const int best01 = (st->bestKey & 1) ? 0 : 1;
  const int RW = P.Q.rwords;
  const u64 *pl = P.Q.planes + ((size_t)it.read * 4 + (best01 ? 0 : 2)) * RW;
  ReadView Q; Q.seq2 = pl; Q.n2 = pl + RW; Q.len = P.Q.len[it.read];
  u64 nw = 0;
  T1K_NOUNROLL
  for (int k = 0; k < RW; ++k) nw |= pl[RW + k];
  Q.anyN = nw != 0;
  Cand c = P.candPool[st->candOff + it.cand];
  done = full_align<false>(R, Q, c, P.Q.weight[it.read], S, err, allowDp);
  if (done && it.recSlot != ~0ull) P.O.store[it.recSlot].mcx = rec_mcx(c.eMatchCnt, c.relaxed, c.strand01);
}

__global__ void __launch_bounds__(128) k_align(AssignParams P) {
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nThreads = (size_t)gridDim.x * blockDim.x;
  const u32 nItems = min(*P.aqCtr, P.aqCap);
  const LaneScratch S = lane_scratch(P.laneScratch + tid * scr_bytes(P.Q.maxLen), P.Q.maxLen);
  const int lane = threadIdx.x & 31;
  int err = 0;
  const size_t nRound = ((size_t)nItems + nThreads - 1) / nThreads;
  for (size_t rnd = 0; rnd < nRound; ++rnd) {
    const size_t i = rnd * nThreads + tid;
    bool needDp = false;
    AlignItem it;
    if (i < nItems) {
      it = P.aq[i];
      if (it.read != 0xffffffffu) { bool done = true; align_item(P, it, S, err, false, done); needDp = !done; }
    }
    const unsigned bal = __ballot_sync(FULL, needDp);
    if (bal) {
      unsigned int base = 0;
      if (lane == 0) base = atomicAdd(P.dpqCtr, (unsigned int)__popc(bal));
      base = __shfl_sync(FULL, base, 0);
      if (needDp) {
        const unsigned int at = base + __popc(bal & ((1u << lane) - 1));
        if (at < P.dpqCap) P.dpq[at] = it;
        else { bool done; align_item(P, it, S, err, true, done); }          // queue full: right here
      }
    }
  }
  if (err) atomicOr(P.O.err, err);
}

__global__ void __launch_bounds__(128) k_align_dp(AssignParams P) {
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nThreads = (size_t)gridDim.x * blockDim.x;
  const u32 nItems = min(*P.dpqCtr, P.dpqCap);
  const LaneScratch S = lane_scratch(P.laneScratch + tid * scr_bytes(P.Q.maxLen), P.Q.maxLen);
  int err = 0;
  for (size_t i = tid; i < nItems; i += nThreads) { bool done; align_item(P, P.dpq[i], S, err, true, done); }
  if (err) atomicOr(P.O.err, err);
}

}  // namespace t1k
