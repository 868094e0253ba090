// sm_100a kernels of the alignment path: read packing, the persistent warp-per-read-end assignment kernel
// (seed selection -> allele-tile gather -> lane-per-allele chaining/rescoring -> extension -> full-read
// alignment + coverage -> ordered compaction into the HBM-resident overlap store) and coverage finalisation.
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
  const u64 *planes;       // [(r*4 + plane) * RWORDS]; planes: fwd seq2, fwd n2, rc seq2, rc n2
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

struct AssignParams {
  RefView R;
  ReadsDev Q;
  AssignOut O;
  Cand *candBuf;           // per warp
  u32 candCap;
  u8 *laneScratch;         // per lane SCR_BYTES
  u32 *hitBuf;             // per warp hitCap x 32: hit lists of the alleles of a tile that take the hit-list path (lane-interleaved)
  unsigned int *workCtr;
  int hitCap;              // hits per allele the hit-list path holds
  int seedCap;             // seeds per strand the shared-memory tables hold (>= longest read - k + 1, multiple of 32, >= 64)
  int noFast;              // 1: every allele goes through the hit-list path (A/B switch, T1K_NO_FAST)
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
__global__ void k_pack_reads(const char *bases, const u64 *off, const u32 *len, u32 n, u64 *planes, u16 *lenOut, int *err) {
  u32 r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const char *s = bases + off[r];
  int L = (int)len[r];
  u64 *out = planes + (size_t)r * 4 * RWORDS;
  if (L > 255) { atomicOr(err, ERR_READ_LEN); L = 0; }
  lenOut[r] = (u16)L;
  u64 fs = 0, fn = 0;
  for (int w = 0; w < RWORDS; ++w) { out[w] = 0; out[RWORDS + w] = 0; out[2 * RWORDS + w] = 0; out[3 * RWORDS + w] = 0; }
  for (int j = 0; j < L; ++j) {
    char c = s[j];
    int v = c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : c == 'T' ? 3 : c == 'N' ? 4 : 5;
    if (v == 5) { atomicOr(err, ERR_READ_CHAR); v = 4; }
    int sh = (j & 31) * 2;
    fs |= (u64)(v == 4 ? 3 : v) << sh;
    fn |= (u64)(v == 4) << sh;
    if ((j & 31) == 31 || j == L - 1) { out[j >> 5] = fs; out[RWORDS + (j >> 5)] = fn; fs = fn = 0; }
  }
  u64 rs = 0, rn = 0;
  for (int j = 0; j < L; ++j) {
    char c = s[L - 1 - j];
    int v = c == 'A' ? 3 : c == 'C' ? 2 : c == 'G' ? 1 : c == 'T' ? 0 : 4;
    int sh = (j & 31) * 2;
    rs |= (u64)(v == 4 ? 3 : v) << sh;
    rn |= (u64)(v == 4) << sh;
    if ((j & 31) == 31 || j == L - 1) { out[2 * RWORDS + (j >> 5)] = rs; out[3 * RWORDS + (j >> 5)] = rn; rs = rn = 0; }
  }
}

// ---------------------------------------------------------------------------------------------------
// One warp = one read-end at a time (dynamic work queue).  Shared memory per warp (seedCap = S):
//   ent[S]    the CURRENT index entry {tile, off, mask, more} of every seed (cp.async destination); during seeding the
//             same storage holds {code, postings, first entry, end entry} of every k-mer window
//   seq/nn    the strand's 2-bit planes
//   cur[S], end[S]  entry cursor / end of every seed; stab[256] seed table of diag_fast; bits[16] its bit masks
//   seedA[S]  read offset of every seed; adv[S] entries the current tile consumed from the seed
//   q0[64], q1[64]  alleles whose mismatch-mask evaluation was deferred (DF_DEFER): {allele, diagonal, hits, hits on the
//             diagonal}, {hits far off it, reserved candidate slot}; run 32 at a time
struct WarpSmem {
  uint4 *ent, *q0;
  u64 *seq, *nn;
  uint2 *q1;
  u32 *cur, *end, *stab, *bits;
  u16 *seedA, *adv;
};
constexpr int DEFER_CAP = 64;
__host__ __device__ inline size_t warp_smem_bytes(int seedCap) {
  return ((size_t)seedCap * (16 + 4 + 4 + 2 + 2) + DEFER_CAP * 24 + 2 * RWORDS * 8 + 256 * 4 + 16 * 4 + 15) & ~(size_t)15;
}

// returns whether the read holds an N
__device__ __forceinline__ bool load_planes(const AssignParams &P, u32 r, int strand01, const WarpSmem &W, int lane) {
  const u64 *src = P.Q.planes + ((size_t)r * 4 + (strand01 ? 0 : 2)) * RWORDS;
  u64 nw = 0;
  if (lane < RWORDS) W.seq[lane] = src[lane];
  else if (lane < 2 * RWORDS) { nw = src[lane]; W.nn[lane - RWORDS] = nw; }   // n2 plane follows the seq plane
  __syncwarp();
  return __any_sync(FULL, nw != 0);
}

__device__ void assign_one_read(const AssignParams &P, u32 r, const WarpSmem &W, Cand *cands, u32 *hitTile, const LaneScratch &S, int lane) {
  const RefView &R = P.R;
  const int len = P.Q.len[r];
  const int weight = P.Q.weight[r];
  const int CAP = P.hitCap;
  int err = 0;
  u32 nCand = 0, nFwd = 0;
  u64 bestKey = 0;
  unsigned long long stPost = 0, stTiles = 0, stSlow = 0;
  // per strand, in list order (_overlap::operator<): the first candidate whose extension fails and the first one whose
  // extension succeeds (goodMatchCnt of SeqSet.hpp:2156-2186 follows from those two: the list is matchCnt-descending)
  u64 fKeyF = ~0ull, rKeyF = ~0ull, fKeyR = ~0ull, rKeyR = ~0ull;
  int fIdxF = 0x7fffffff, rIdxF = 0x7fffffff, fIdxR = 0x7fffffff, rIdxR = 0x7fffffff;
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
      // first entry of every seed (the window table is dead now)
      T1K_NOUNROLL
      for (int k = lane; k < nS; k += 32) cp_async16(&W.ent[k], R.entries + W.cur[k]);
      u64 laneKey = 0;
      u64 fKey = ~0ull, rKey = ~0ull; int fIdx = 0x7fffffff, rIdx = 0x7fffffff;
      auto note = [&](const Cand &c, int i) {
        if (c.flags & CF_SEP) return;
        const u64 k = cand_key_pre(c);
        if (c.flags & CF_RET) { if (pair_less(k, i, rKey, rIdx)) { rKey = k; rIdx = i; } }
        else if (pair_less(k, i, fKey, fIdx)) { fKey = k; fIdx = i; }
      };
      // the deferred alleles, `take` at a time with all lanes busy: the whole evaluation incl. dirty gaps and overhangs
      int qn = 0;
      auto run_deferred = [&](int take) {
        if (lane < take) {
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
          } else {
            Cand v;                                         // nothing to emit: the reserved slot stays void
            v.seqIdx = (int32_t)e0.x; v.seqStart = v.seqEnd = 0; v.readStart = v.readEnd = 0; v.strand01 = (u8)strand01; v.flags = CF_SEP;
            v.matchCnt = 0; v.pad = 0; v.eSeqStart = v.eSeqEnd = 0; v.eReadStart = v.eReadEnd = v.leftClip = v.rightClip = 0;
            v.eMatchCnt = 0; v.relaxed = 0; v.mmPos = 0;
            cands[e1.y] = v;
          }
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
          unsigned bal = __ballot_sync(FULL, k < nS && W.ent[k].x == T);
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
          const int df = diag_fast(R, Qv, strand01, (int)(T * 32 + lane), n, d0, onDiag, far, W.stab, true, fc, fastEmit, laneKey, lcMemo, S, err);
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
                  if (cnt < CAP) hitTile[(size_t)cnt * 32 + lane] = a | (e.y << 8);
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
        if (qn >= 32) { __syncwarp(); run_deferred(32); }
      }
      T1K_NOUNROLL
      while (qn > 0) { __syncwarp(); run_deferred(min(qn, 32)); }
      bestKey = max(bestKey, warp_max_u64(laneKey));
      warp_min_pair(fKey, fIdx);
      warp_min_pair(rKey, rIdx);
      if (pass == 0) { fKeyF = fKey; fIdxF = fIdx; rKeyF = rKey; rIdxF = rIdx; nFwd = nCand; }
      else { fKeyR = fKey; fIdxR = fIdx; rKeyR = rKey; rIdxR = rIdx; }
    }
  }
  // ---- AssignRead proper (SeqSet.hpp:2132-2300) on the best strand's candidates
  const int best01 = (bestKey & 1) ? 0 : 1;
  const int c0 = best01 ? 0 : (int)nFwd, c1 = best01 ? (int)nFwd : (int)nCand;
  const u64 fKey = best01 ? fKeyF : fKeyR, rKey = best01 ? rKeyF : rKeyR;
  const int fIdx = best01 ? fIdxF : fIdxR, rIdx = best01 ? rIdxF : rIdxR;
  int ret = -1, nFinal = 0;
  unsigned long long pos = 0;
  bool deferred = false;
  if (c1 - c0 > 0) {
    if (best01 == 1) Qv.anyN = load_planes(P, r, 1, W, lane);
    __threadfence_block();
    __syncwarp();
    // goodMatchCnt (SeqSet.hpp:2156-2186) = the largest matchCnt among the returned candidates that precede the first
    // failing one = the first returned candidate if it precedes the failure, and nothing otherwise.
    const int good = pair_less(rKey, rIdx, fKey, fIdx) ? order_key_mc(rKey) : -1;
    auto included = [&](const Cand &c, int i) {          // the ordered scan's verdict on one candidate (SeqSet.hpp:2163-2186)
      if ((c.flags & CF_SEP) || !(c.flags & CF_RET)) return false;
      if (pair_less(cand_key_pre(c), i, fKey, fIdx)) return true;
      return !((int)c.matchCnt < good && (!(c.flags & CF_NEEDCLIP) || sim_below(R, c.matchCnt, cand_denom_pre(c), 1)));
    };
    // pass A (read-only): what is kept, the best matchCnt, the head of the post-extension order and the smallest
    // similarity (as an exact fraction), which tells whether the > 1000 cut (SeqSet.hpp:2290-2298) removes anything
    int bestMc = -1, nInc = 0;
    u64 bKey = ~0ull; int bIdx = 0x7fffffff;
    int minNum = 1, minDen = 0;                          // minDen == 0: none yet
    T1K_NOUNROLL
    for (int i = c0 + lane; i < c1; i += 32) {
      if (i + 64 < c1) prefetch_l2(&cands[i + 64]);
      const Cand c = cands[i];
      if (!included(c, i)) continue;
      bestMc = max(bestMc, c.eMatchCnt);
      ++nInc;
      const u64 k = cand_key_post(c);
      if (pair_less(k, i, bKey, bIdx)) { bKey = k; bIdx = i; }
      const int den = cand_denom_post(c);
      if (minDen == 0 || (long long)c.eMatchCnt * minDen < (long long)minNum * den) { minNum = c.eMatchCnt; minDen = den; }
    }
    bestMc = warp_max_i32(bestMc);
    nInc = warp_sum_i32(nInc);
    warp_min_pair(bKey, bIdx);
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      const int on = __shfl_xor_sync(FULL, minNum, o), od = __shfl_xor_sync(FULL, minDen, o);
      if (od != 0 && (minDen == 0 || (long long)on * minDen < (long long)minNum * od)) { minNum = on; minDen = od; }
    }
    __syncwarp();
    // reserve the store before touching coverage, so that a full store can be retried without double counting
    if (lane == 0 && nInc > 0) pos = atomicAdd(P.O.storeCtr, (unsigned long long)nInc);
    pos = __shfl_sync(FULL, pos, 0);
    if (nInc > 0 && pos + nInc > P.O.storeCap) deferred = true;
    if (!deferred && nInc > 0) {
      const bool usePost = nInc > 1000;      // SeqSet.hpp:2290-2298
      u64 cKey = ~0ull; int cIdx = 0x7fffffff;       // first candidate (post order) the cut removes
      if (usePost) {
        const Cand cb = cands[bIdx];
        const double cutSim = (double)cb.eMatchCnt / (double)cand_denom_post(cb) - 0.1;
        if ((double)minNum / (double)minDen < cutSim) {
          T1K_NOUNROLL
          for (int i = c0 + lane; i < c1; i += 32) {
            const Cand c = cands[i];
            if (!included(c, i) || i == bIdx) continue;
            if ((double)c.eMatchCnt / (double)cand_denom_post(c) < cutSim) {
              const u64 k = cand_key_post(c);
              if (pair_less(k, i, cKey, cIdx)) { cKey = k; cIdx = i; }
            }
          }
          warp_min_pair(cKey, cIdx);
        }
      }
      // pass B: full-read alignment + coverage of everything within 10 of the best (Q8) and ordered compaction into the
      // store (allele order is kept: pairing searches it), minus the cut.  Two-speed: candidates whose full-read alignment
      // the seeding stage already knows (CF_FA) cost nothing; the others are queued (W.cur: candidate, W.end: store slot)
      // and aligned 32 at a time, all lanes busy.
      const bool doAlign = weight >= 0;
      int running = 0, qn = 0;
      u32 top = 0;
      auto emit_rec = [&](const Cand &c, u32 slot) {
        Rec o;
        o.seqIdx = c.seqIdx; o.seqStart = c.eSeqStart; o.seqEnd = c.eSeqEnd;
        o.readStart = c.eReadStart; o.readEnd = c.eReadEnd; o.leftClip = c.leftClip; o.rightClip = c.rightClip;
        o.mcx = rec_mcx(c.eMatchCnt, c.relaxed, c.strand01);
        o.key = usePost ? (cand_key_post(c) | 1ull) : cand_key_pre(c);
        P.O.store[pos + slot] = o;
      };
      T1K_NOUNROLL
      for (int b = c0; b < c1 || qn > 0; b += 32) {
        const int i = b + lane;
        bool inc = false, cold = false, wr = false;
        Cand c;
        if (i + 64 < c1) prefetch_l2(&cands[i + 64]);
        if (i < c1) {
          c = cands[i];
          inc = included(c, i);
          if (inc) {
            wr = !usePost || pair_less(cand_key_post(c), i, cKey, cIdx);
            if (doAlign) {
              if (c.eMatchCnt < bestMc - 10) c.relaxed = 0;
              else if (c.flags & CF_FA) full_align_known(R, c, weight);
              else cold = true;
            }
          }
        }
        const unsigned balW = __ballot_sync(FULL, wr), balC = __ballot_sync(FULL, cold);
        const u32 slot = (u32)running + __popc(balW & ((1u << lane) - 1));
        if (wr) top = max(top, ((u32)c.eMatchCnt << 16) | (u32)(65535 - cand_denom_post(c)));
        if (cold) {
          const int q = qn + __popc(balC & ((1u << lane) - 1));
          W.cur[q] = (u32)i; W.end[q] = wr ? slot : 0xffffffffu;
        } else if (wr) emit_rec(c, slot);
        running += __popc(balW);
        qn += __popc(balC);
        __syncwarp();
        if (qn >= 32 || (b + 32 >= c1 && qn > 0)) {          // flush a full batch, or the remainder at the end
          const int take = min(qn, 32);
          if (lane < take) {
            const int j = (int)W.cur[qn - take + lane];
            const u32 sl = W.end[qn - take + lane];
            Cand cc = cands[j];
            full_align<false>(R, Qv, cc, weight, S, err);
            if (sl != 0xffffffffu) emit_rec(cc, sl);
          }
          qn -= take;
          __syncwarp();
        }
      }
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
  if (P.O.stats) { stSlow = __reduce_add_sync(FULL, (u32)stSlow); }
  if (lane == 0 && P.O.stats) {
    atomicAdd(P.O.stats + 0, stPost);
    atomicAdd(P.O.stats + 1, (unsigned long long)nCand);
    atomicAdd(P.O.stats + 2, stTiles);
    atomicAdd(P.O.stats + 3, stSlow);
  }
}

extern __shared__ u64 t1k_smem[];

// MINB = resident blocks per SM the register budget is compiled for
template <int MINB>
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32, MINB) k_assign(AssignParams P) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t gwarp = (size_t)blockIdx.x * WARPS_PER_BLOCK + warp;
  const int SC = P.seedCap;
  u8 *sm = (u8 *)t1k_smem + (size_t)warp * warp_smem_bytes(SC);
  WarpSmem W;
  W.ent = (uint4 *)sm; W.q0 = W.ent + SC;
  W.seq = (u64 *)(W.q0 + DEFER_CAP); W.nn = W.seq + RWORDS;
  W.q1 = (uint2 *)(W.nn + RWORDS);
  W.cur = (u32 *)(W.q1 + DEFER_CAP); W.end = W.cur + SC; W.stab = W.end + SC; W.bits = W.stab + 256;
  W.seedA = (u16 *)(W.bits + 16); W.adv = W.seedA + SC;
  LaneScratch S; S.base = P.laneScratch + (gwarp * 32 + lane) * (size_t)SCR_BYTES;
  Cand *cands = P.candBuf + gwarp * (size_t)P.candCap;
  u32 *hitTile = P.hitBuf + gwarp * (size_t)P.hitCap * 32;
  for (;;) {
    u32 w = 0;
    if (lane == 0) w = atomicAdd(P.workCtr, 1u);
    w = __shfl_sync(FULL, w, 0);
    if (w >= P.Q.nWork) break;
    const u32 r = P.Q.workList ? P.Q.workList[w] : w;
    assign_one_read(P, r, W, cands, hitTile, S, lane);
    __syncwarp();
  }
}

}  // namespace t1k
