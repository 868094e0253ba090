// sm_100a kernels of the alignment path: read packing, the persistent warp-per-read-end assignment kernel
// (seed selection -> allele-tile gather -> lane-per-allele chaining/rescoring -> extension -> full-read
// alignment + coverage -> ordered compaction into the HBM-resident overlap store) and coverage finalisation.
#pragma once
#include <cuda_runtime.h>

#include "t1k_core.cuh"

namespace t1k {

constexpr int WARPS_PER_BLOCK = 4;
constexpr unsigned FULL = 0xffffffffu;
constexpr int GATHER_DEPTH = 8;      // stages of the posting-block ring of the tile gather (power of two)

__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
// ---- TMA (bulk async copy engine), 1-D form: one lane moves a whole posting block global -> shared and the block's
// mbarrier flips when the bytes have landed (cp.async.bulk + mbarrier complete_tx; source, destination and size are
// multiples of 16 bytes)
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(u64 *bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(u64 *bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void *smemDst, const void *gsrc, unsigned bytes, u64 *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(smemDst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(u64 *bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "T1K_MBAR_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra T1K_MBAR_DONE;\n"
      "bra T1K_MBAR_WAIT;\n"
      "T1K_MBAR_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
constexpr int RING_POSTINGS = 34;      // 32 postings of the block + the one that follows + 1 for the 16-byte alignment of the source
constexpr int RING_BYTES = RING_POSTINGS * 8;

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
  int32_t *readRet;        // AssignRead's return value; -2 = deferred (store full), -3 = deferred (hit tile too small)
  int *err;
  unsigned long long *stats;   // [0] postings visited, [1] candidates, [2] tiles, [3] dp calls (debug/roofline)
};

struct AssignParams {
  RefView R;
  ReadsDev Q;
  AssignOut O;
  Cand *candBuf;           // per warp
  u32 candCap;
  u8 *laneScratch;         // per lane SCR_BYTES
  unsigned int *workCtr;
  int hitCap;              // hits per allele kept in shared memory
  int noFast;              // 1: every allele goes through the hit-list path (A/B switch, T1K_NO_FAST)
};

// warp reductions on the redux unit (one instruction per 32-bit reduction; the kernel is instruction-footprint bound)
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
// One warp = one read-end at a time (dynamic work queue).  Shared memory per warp:
//   read planes (2 x RWORDS words), posting ring (GATHER_DEPTH x 272 B, TMA destination) + one mbarrier per stage,
//   H[hitCap][32]  encoded hits of the current allele tile, lane-interleaved (bank = lane)
//   cnt[32], cur[256], end[256], nxt[256], stab[256] (seed table of diag_fast), seedA[256], act[256]
struct WarpSmem {
  u32 *H, *cnt, *cur, *end, *nxt, *stab;
  u8 *seedA, *act;
  u64 *seq, *nn;
  Posting *ring;           // GATHER_DEPTH x RING_POSTINGS postings (TMA destination, 16-byte aligned stages)
  u64 *bars;               // GATHER_DEPTH mbarriers, one per ring stage
};
__host__ __device__ inline size_t warp_smem_bytes(int hitCap) {
  return (size_t)hitCap * 32 * 4 + 32 * 4 + 4 * 256 * 4 + 2 * 256 + 2 * RWORDS * 8 + GATHER_DEPTH * RING_BYTES + GATHER_DEPTH * 8;
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

// stage i of the tile gather: one TMA copy brings the 32-posting block of active seed i (from the even posting at or
// below its cursor, so that the source is 16-byte aligned) plus the posting that follows it into ring stage i mod DEPTH
__device__ __forceinline__ void gather_issue(const RefView &R, const WarpSmem &W, int i, int nAct, int lane) {
  if (i < nAct && lane == 0) {
    const int k = W.act[i];
    const u32 c0 = W.cur[k] & ~1u;
    const int st = i & (GATHER_DEPTH - 1);
    mbar_expect_tx(W.bars + st, RING_BYTES);
    tma_load_1d(W.ring + st * RING_POSTINGS, R.post + c0, RING_BYTES, W.bars + st);    // (the posting array is padded: reading past a list's end is harmless)
  }
}

// ringPhase: bit s = phase parity the warp waits for next on ring stage s (the mbarriers live as long as the warp)
__device__ void assign_one_read(const AssignParams &P, u32 r, const WarpSmem &W, Cand *cands, const LaneScratch &S, int lane, u32 &ringPhase) {
  const RefView &R = P.R;
  const int len = P.Q.len[r];
  const int weight = P.Q.weight[r];
  const int CAP = P.hitCap;
  int err = 0;
  u32 nCand = 0, nFwd = 0;
  u64 bestKey = 0;
  unsigned long long stPost = 0, stTiles = 0;
  bool overflow = false;   // some allele has more hits than the shared-memory tile holds: re-run with the big tile
  ReadView Qv; Qv.seq2 = W.seq; Qv.n2 = W.nn; Qv.len = len; Qv.anyN = false;

  if (len >= KMER) {
    const int NP = len - KMER + 1;
    T1K_NOUNROLL
    for (int pass = 0; pass < 2 && !overflow; ++pass) {
      const int strand01 = pass == 0 ? 1 : 0;
      Qv.anyN = load_planes(P, r, strand01, W, lane);
      // ---- k-mer codes and posting ranges of every window (GetHitsFromRead, SeqSet.hpp:1093-1153)
      T1K_NOUNROLL
      for (int a = lane; a < NP; a += 32) {
        u32 code = (u32)(fetch32(W.seq, a) & 0x3FFFFFull);
        bool valid = (fetch32(W.nn, a) & 0x155555ull) == 0;
        u32 lo = R.kstart[code], hi = R.kstart[code + 1];
        W.nxt[a] = code; W.cur[a] = lo; W.end[a] = valid ? hi : lo;
      }
      __syncwarp();
      // ---- the sequential skip rule (list >= 100, not first/last, <= K/2 in a row; Q2)
      // The strand is eligible for diag_fast when the read holds no N, fits 5 words and no seed is a homopolymer k-mer
      // (see the claim above diag_fast); lane 0 then also builds the seed table.
      int nS = 0;
      bool strandFast = !Qv.anyN && len <= FAST_MAX_LEN && !P.noFast;
      if (lane == 0) {
        u32 prev = 0; int skip = 0;
        T1K_NOUNROLL
        for (int a = 0; a < NP; ++a) {
          u32 code = W.nxt[a];
          if (a == 0 || prev != code) {
            u32 lo = W.cur[a], hi = W.end[a];
            int size = (int)(hi - lo);
            if (size >= 100 && a != 0 && a != NP - 1 && skip < KMER / 2) { ++skip; continue; }
            skip = 0;
            if (size > 0) { W.seedA[nS] = (u8)a; W.cur[nS] = lo; W.end[nS] = hi; ++nS; if (kmer_homopolymer(code)) strandFast = false; }
          }
          prev = code;
        }
      }
      nS = __shfl_sync(FULL, nS, 0);
      strandFast = __shfl_sync(FULL, (int)strandFast, 0) != 0;
      u32 lcMemo = 0;
      __syncwarp();
      if (strandFast) {
        // seed table (see seed_table_build), warp-cooperative: seed / wide-step bit masks in W.cnt (free until the first
        // tile), then every lane derives the entries of its read positions from the masks
        if (lane < 16) W.cnt[lane] = 0;
        __syncwarp();
        T1K_NOUNROLL
        for (int k = lane; k < nS; k += 32) {
          const int a = W.seedA[k];
          atomicOr(&W.cnt[a >> 5], 1u << (a & 31));
          if (k > 0 && a - (int)W.seedA[k - 1] > KMER - 1) atomicOr(&W.cnt[8 + (a >> 5)], 1u << (a & 31));
        }
        __syncwarp();
        T1K_NOUNROLL
        for (int a = lane; a < len; a += 32) {
          const int wi = a >> 5, bit = a & 31;             // wi is warp-uniform
          u32 cnt = 0, big = 0;
          T1K_NOUNROLL
          for (int w = 0; w < wi; ++w) { cnt += __popc(W.cnt[w]); big += __popc(W.cnt[8 + w]); }
          const u32 upTo = 0xffffffffu >> (31 - bit);
          const u32 cw = W.cnt[wi] & upTo;
          cnt += __popc(cw); big += __popc(W.cnt[8 + wi] & upTo);
          u32 last = 255, nxt = 255;
          {
            int w = wi; u32 m = cw;
            T1K_NOUNROLL
            for (;;) { if (m) { last = (u32)(w * 32 + 31 - __clz(m)); break; } if (--w < 0) break; m = W.cnt[w]; }
          }
          {
            int w = wi; u32 m = W.cnt[wi] & (0xffffffffu << bit);
            T1K_NOUNROLL
            for (;;) { if (m) { nxt = (u32)(w * 32 + __ffs(m) - 1); break; } if (++w >= 8) break; m = W.cnt[w]; }
          }
          W.stab[a] = cnt | (big << 8) | (nxt << 16) | (last << 24);
        }
        __syncwarp();
      }
      T1K_NOUNROLL
      for (int k = lane; k < nS; k += 32) { W.nxt[k] = R.post[W.cur[k]].idx; stPost += W.end[k] - W.cur[k]; }
      __syncwarp();
      u64 laneKey = 0;
      // ---- allele tiles: 32 consecutive allele ids starting at the smallest pending one
      T1K_NOUNROLL
      for (;;) {
        u32 mn = 0xffffffffu;
        T1K_NOUNROLL
        for (int k = lane; k < nS; k += 32) mn = min(mn, W.nxt[k]);
        mn = warp_min_u32(mn);
        if (mn == 0xffffffffu) break;
        const u32 base = mn;
        ++stTiles;
        W.cnt[lane] = 0;
        __syncwarp();
        // seeds with postings inside this tile, in read-offset order (=> per-allele hits sorted by (a, b))
        int nAct = 0;
        T1K_NOUNROLL
        for (int k0 = 0; k0 < nS; k0 += 32) {
          const int k = k0 + lane;
          const bool act = k < nS && W.nxt[k] < base + 32;
          const unsigned bal = __ballot_sync(FULL, act);
          if (act) W.act[nAct + __popc(bal & ((1u << lane) - 1))] = (u8)k;
          nAct += __popc(bal);
        }
        __syncwarp();
        // The 256-byte posting blocks of the active seeds stream through a GATHER_DEPTH-stage cp.async ring in shared
        // memory (each lane copies and later reads its own 8 bytes), so GATHER_DEPTH - 1 blocks are in flight while one
        // is scattered into the tile; the loop body exists once (its instruction footprint matters).
        T1K_NOUNROLL
        for (int s0 = 0; s0 < GATHER_DEPTH - 1; ++s0) gather_issue(R, W, s0, nAct, lane);
        T1K_NOUNROLL
        for (int i = 0; i < nAct; ++i) {
          gather_issue(R, W, i + GATHER_DEPTH - 1, nAct, lane);
          const int st = i & (GATHER_DEPTH - 1);
          mbar_wait(W.bars + st, (ringPhase >> st) & 1u);
          ringPhase ^= 1u << st;
          const int k = W.act[i];
          const u32 a = W.seedA[k], e = W.end[k];
          u32 c = W.cur[k];
          const Posting *blk = W.ring + st * RING_POSTINGS + (c & 1u);     // blk[j] = posting c + j
          Posting p; p.idx = 0xffffffffu; p.off = 0;
          if (c + lane < e) p = blk[lane];
          const u32 follow = c + 32 < e ? blk[32].idx : 0xffffffffu;        // allele id of the posting after the block (broadcast read)
          bool first = true;                 // p came through the ring (`follow` knows what comes after it)
          T1K_NOUNROLL
          for (;;) {
            const bool in = p.idx < base + 32;
            const unsigned bal = __ballot_sync(FULL, in);
            const int consumed = __popc(bal);
            const u32 prevIdx = __shfl_up_sync(FULL, p.idx, 1);
            const bool dup = lane > 0 && p.idx == prevIdx;
            const u32 local = p.idx - base;
            if (!__any_sync(FULL, dup && in)) {
              // every allele of the tile occurs at most once in this block (the usual case): plain scatter
              if (in) {
                const u32 slot = W.cnt[local];
                if ((int)slot < CAP) W.H[slot * 32 + local] = a | (p.off << 8);
                W.cnt[local] = slot + 1;
              }
            } else {
              // a k-mer repeated inside an allele: rank the postings of each allele run
              const unsigned sm = __ballot_sync(FULL, !dup);
              const int runStart = 31 - __clz(sm & (0xffffffffu >> (31 - lane)));
              const int rank = lane - runStart;
              const bool lastOfRun = lane == 31 || ((sm >> (lane + 1)) & 1u);
              u32 slot = 0;
              if (in) {
                slot = W.cnt[local] + rank;
                if ((int)slot < CAP) W.H[slot * 32 + local] = a | (p.off << 8);
              }
              __syncwarp();
              if (in && lastOfRun) W.cnt[local] = slot + 1;
            }
            __syncwarp();
            const u32 nc = c + consumed;
            u32 nextIdx;                             // allele id of the first posting left in the list (all branches warp-uniform)
            if (consumed < 32) nextIdx = __shfl_sync(FULL, p.idx, consumed);
            else if (nc >= e) nextIdx = 0xffffffffu;
            else if (first) nextIdx = follow;
            else nextIdx = 0;                        // unknown: look at the next block
            if (consumed == 32 && nc < e && nextIdx < base + 32) {   // the list continues inside this tile (repeated k-mer)
              c = nc; first = false;
              p.idx = 0xffffffffu; p.off = 0;
              if (c + lane < e) p = R.post[c + lane];
              continue;
            }
            if (lane == 0) { W.cur[k] = nc; W.nxt[k] = nextIdx; }
            break;
          }
        }
        __syncwarp();
        // ---- lane-per-allele chaining + rescoring
        const int n = (int)W.cnt[lane];
        int nEmit = 0;
        if (__any_sync(FULL, n > CAP)) { overflow = true; break; }
        // all lanes first try the mismatch-mask path (uniform work); the few it declines walk their hit lists
        Cand fc;
        bool fastEmit = false, handled = n < 3;
        if (strandFast && n >= 3) {
          handled = diag_fast(R, Qv, strand01, (int)(base + lane), n, W.H + lane, 32, W.stab, fc, fastEmit, laneKey, lcMemo, S, err);
          if (fastEmit) nEmit = 1;
        }
        if (!handled) {
          chain_allele(R, Qv, strand01, (int)(base + lane), W.H + lane, 32, n, S, nEmit, laneKey, err);
          if (nEmit > 1) sort_emitted(S.emit(), nEmit);       // several clusters on one allele: the tail order of _overlap::operator<
        }
        // ---- ordered emission (allele order == lane order)
        int incl = nEmit;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(FULL, incl, o); if (lane >= o) incl += t; }
        const int tot = __shfl_sync(FULL, incl, 31);
        if (tot > 0) {
          if (nCand + tot > P.candCap) err |= ERR_CAND;
          else {
            if (fastEmit) cands[nCand + incl - 1] = fc;
            else {
              const Cand *em = S.emit();
              T1K_NOUNROLL
              for (int j = 0; j < nEmit; ++j) cands[nCand + incl - nEmit + j] = em[j];
            }
            nCand += tot;
          }
        }
        __syncwarp();
      }
      bestKey = max(bestKey, warp_max_u64(laneKey));
      if (pass == 0) nFwd = nCand;
    }
  }
  // ---- AssignRead proper (SeqSet.hpp:2132-2300) on the best strand's candidates
  const int best01 = (bestKey & 1) ? 0 : 1;
  const int c0 = best01 ? 0 : (int)nFwd, c1 = best01 ? (int)nFwd : (int)nCand;
  int ret = -1, nFinal = 0;
  unsigned long long pos = 0;
  bool deferred = false;
  if (c1 - c0 > 0 && !overflow) {
    if (best01 == 1) Qv.anyN = load_planes(P, r, 1, W, lane);
    __threadfence_block();
    __syncwarp();
    // pass 1: extension; the first candidate (list order) whose extension fails, and the first one that succeeds.
    // Two-speed loop: candidates the seeding stage already extended (CF_PRE) cost nothing; the others are queued
    // (W.cur is free after the gather) and extended 32 at a time, all lanes busy.
    u64 fKey = ~0ull; int fIdx = 0x7fffffff;        // first failing (not CF_RET)
    u64 rKey = ~0ull; int rIdx = 0x7fffffff;        // first returned (CF_RET)
    auto note = [&](const Cand &c, int i) {
      if (c.flags & CF_SEP) return;
      const u64 k = cand_key_pre(c);
      if (c.flags & CF_RET) { if (pair_less(k, i, rKey, rIdx)) { rKey = k; rIdx = i; } }
      else if (pair_less(k, i, fKey, fIdx)) { fKey = k; fIdx = i; }
    };
    int qn = 0;
    T1K_NOUNROLL
    for (int b = c0; b < c1 || qn > 0; b += 32) {
      int i = b + lane;
      bool have = i < c1, cold = false, pre = false;
      Cand c;
      if (i + 64 < c1) prefetch_l2(&cands[i + 64]);       // the candidate buffers live in HBM: next rounds' records into L2
      if (have) {
        c = cands[i];
        pre = (c.flags & CF_PRE) != 0;              // extension already known from the seeding stage
        cold = !pre;                                // the others are queued and extended 32 at a time
      }
      const unsigned bal = __ballot_sync(FULL, cold);
      if (cold) { W.cur[qn + __popc(bal & ((1u << lane) - 1))] = (u32)i; have = false; }
      qn += __popc(bal);
      __syncwarp();
      if (qn >= 32 || (b + 32 >= c1 && qn > 0)) {          // flush a full batch, or the remainder at the end
        const int take = min(qn, 32);
        if (lane < take) {
          if (have) {                                       // this lane's hot result first
            if (!pre) cands[i] = c;
            note(c, i);
          }
          i = (int)W.cur[qn - take + lane];
          c = cands[i];
          extend_cand<false>(R, Qv, c, S, err);
          have = true; pre = false;
        }
        qn -= take;
        __syncwarp();
      }
      if (have) {
        if (!pre) cands[i] = c;
        note(c, i);
      }
    }
    warp_min_pair(fKey, fIdx);
    warp_min_pair(rKey, rIdx);
    __syncwarp();
    // goodMatchCnt (SeqSet.hpp:2156-2186) = the largest matchCnt among the returned candidates that precede the first
    // failing one.  The list order is matchCnt-descending, so that is the first returned candidate if it precedes the
    // failure, and nothing otherwise: no pass of its own.
    const int good = pair_less(rKey, rIdx, fKey, fIdx) ? order_key_mc(rKey) : -1;
    // pass 2: inclusion; the list head under the post-extension order (for the > 1000 cut, SeqSet.hpp:2290-2298)
    int bestMc = -1, nInc = 0;
    u64 bKey = ~0ull; int bIdx = 0x7fffffff;
    T1K_NOUNROLL
    for (int i = c0 + lane; i < c1; i += 32) {
      if (i + 64 < c1) prefetch_l2(&cands[i + 64]);
      Cand c = cands[i];
      if ((c.flags & CF_SEP) || !(c.flags & CF_RET)) continue;
      bool before = pair_less(cand_key_pre(c), i, fKey, fIdx);
      if (!before && (int)c.matchCnt < good && (!(c.flags & CF_NEEDCLIP) || sim_below(R, c.matchCnt, cand_denom_pre(c), 1))) continue;
      cands[i].flags = c.flags | CF_INCLUDE;
      bestMc = max(bestMc, c.eMatchCnt);
      ++nInc;
      const u64 k = cand_key_post(c);
      if (pair_less(k, i, bKey, bIdx)) { bKey = k; bIdx = i; }
    }
    bestMc = warp_max_i32(bestMc);
    nInc = warp_sum_i32(nInc);
    warp_min_pair(bKey, bIdx);
    __syncwarp();
    // reserve the store before touching coverage, so that a full store can be retried without double counting
    if (lane == 0 && nInc > 0) pos = atomicAdd(P.O.storeCtr, (unsigned long long)nInc);
    pos = __shfl_sync(FULL, pos, 0);
    if (nInc > 0 && pos + nInc > P.O.storeCap) deferred = true;
    if (!deferred) {
      const bool usePost = nInc > 1000;      // SeqSet.hpp:2290-2298
      double cutSim = 0;
      if (usePost) { const Cand cb = cands[bIdx]; cutSim = (double)cb.eMatchCnt / (double)cand_denom_post(cb) - 0.1; }
      u64 cKey = ~0ull; int cIdx = 0x7fffffff;       // first candidate (post order) the cut removes
      // pass 3: full-read alignment of everything within 10 of the best (Q8); where the cut starts
      {
        const bool doAlign = weight >= 0;
        int qn = 0;                                          // same two-speed structure as the extension pass
        T1K_NOUNROLL
        for (int b = c0; b < c1 || qn > 0; b += 32) {
          const int i = b + lane;
          bool cold = false;
          if (i + 64 < c1) prefetch_l2(&cands[i + 64]);
          if (i < c1) {
            Cand c = cands[i];
            if (c.flags & CF_INCLUDE) {
              if (usePost && i != bIdx && (double)c.eMatchCnt / (double)cand_denom_post(c) < cutSim) {
                const u64 k = cand_key_post(c);
                if (pair_less(k, i, cKey, cIdx)) { cKey = k; cIdx = i; }
              }
              if (doAlign) {
                if (c.eMatchCnt < bestMc - 10) c.relaxed = 0;
                else if (c.flags & CF_FA) full_align_known(R, c, weight);
                else cold = true;
                if (!cold) cands[i].relaxed = c.relaxed;
              }
            }
          }
          const unsigned bal = __ballot_sync(FULL, cold);
          if (cold) W.cur[qn + __popc(bal & ((1u << lane) - 1))] = (u32)i;
          qn += __popc(bal);
          __syncwarp();
          if (qn >= 32 || (b + 32 >= c1 && qn > 0)) {
            const int take = min(qn, 32);
            if (lane < take) {
              const int j = (int)W.cur[qn - take + lane];
              Cand c = cands[j];
              full_align<false>(R, Qv, c, weight, S, err);
              cands[j].relaxed = c.relaxed;
            }
            qn -= take;
            __syncwarp();
          }
        }
      }
      warp_min_pair(cKey, cIdx);
      __syncwarp();
      // pass 4: ordered compaction into the store (allele order is kept: pairing searches it), minus the cut
      int running = 0;
      u32 top = 0;
      T1K_NOUNROLL
      for (int b = c0; b < c1; b += 32) {
        const int i = b + lane;
        bool inc = false;
        Cand c;
        u64 kPost = 0;
        if (i + 64 < c1) prefetch_l2(&cands[i + 64]);
        if (i < c1) {
          c = cands[i];
          inc = (c.flags & CF_INCLUDE) != 0;
          if (inc && usePost) { kPost = cand_key_post(c); inc = pair_less(kPost, i, cKey, cIdx); }
        }
        const unsigned bal = __ballot_sync(FULL, inc);
        if (inc) {
          Rec o;
          o.seqIdx = c.seqIdx; o.seqStart = c.eSeqStart; o.seqEnd = c.eSeqEnd;
          o.readStart = c.eReadStart; o.readEnd = c.eReadEnd; o.leftClip = c.leftClip; o.rightClip = c.rightClip;
          o.mcx = rec_mcx(c.eMatchCnt, c.relaxed, c.strand01);
          o.key = usePost ? (kPost | 1ull) : cand_key_pre(c);
          top = max(top, ((u32)c.eMatchCnt << 16) | (u32)(65535 - cand_denom_post(c)));
          P.O.store[pos + running + __popc(bal & ((1u << lane) - 1))] = o;
        }
        running += __popc(bal);
      }
      top = __reduce_max_sync(FULL, top);
      if (lane == 0) P.O.readTop[r] = top;
      nFinal = running;
      ret = nFinal;
    }
  }
  if (lane == 0) {
    P.O.readOff[r] = pos;
    P.O.readCnt[r] = (deferred || overflow) ? 0 : (u32)nFinal;
    P.O.readRet[r] = overflow ? -3 : deferred ? -2 : ret;
    if (nFinal > 0 && !deferred && !overflow) atomicMax(P.O.maxCnt, (u32)nFinal);
    if (deferred) atomicOr(P.O.err, ERR_STORE);
    if (overflow) atomicOr(P.O.err, ERR_HITS);
  }
  err = __reduce_or_sync(FULL, (unsigned)err);
  if (lane == 0 && err) atomicOr(P.O.err, err);
  if (P.O.stats) stPost = __reduce_add_sync(FULL, (u32)stPost);     // < 2^32 postings per read-end
  if (lane == 0 && P.O.stats) {
    atomicAdd(P.O.stats + 0, stPost);
    atomicAdd(P.O.stats + 1, (unsigned long long)nCand);
    atomicAdd(P.O.stats + 2, stTiles);
  }
}

extern __shared__ u64 t1k_smem[];

// MINB = resident blocks per SM the register budget is compiled for (4: 128 registers, 5: 96, 6: 80)
template <int MINB>
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32, MINB) k_assign(AssignParams P) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t gwarp = (size_t)blockIdx.x * WARPS_PER_BLOCK + warp;
  u8 *sm = (u8 *)t1k_smem + (size_t)warp * ((warp_smem_bytes(P.hitCap) + 15) & ~(size_t)15);
  WarpSmem W;
  W.seq = (u64 *)sm; W.nn = W.seq + RWORDS;
  W.ring = (Posting *)(W.nn + RWORDS);
  W.bars = (u64 *)(W.ring + GATHER_DEPTH * RING_POSTINGS);
  W.H = (u32 *)(W.bars + GATHER_DEPTH);
  if (lane < GATHER_DEPTH) mbar_init(W.bars + lane, 1);
  mbar_fence_init();
  __syncwarp();
  W.cnt = W.H + (size_t)P.hitCap * 32;
  W.cur = W.cnt + 32; W.end = W.cur + 256; W.nxt = W.end + 256; W.stab = W.nxt + 256;
  W.seedA = (u8 *)(W.stab + 256); W.act = W.seedA + 256;
  LaneScratch S; S.base = P.laneScratch + (gwarp * 32 + lane) * (size_t)SCR_BYTES;
  Cand *cands = P.candBuf + gwarp * (size_t)P.candCap;
  u32 ringPhase = 0;
  for (;;) {
    u32 w = 0;
    if (lane == 0) w = atomicAdd(P.workCtr, 1u);
    w = __shfl_sync(FULL, w, 0);
    if (w >= P.Q.nWork) break;
    const u32 r = P.Q.workList ? P.Q.workList[w] : w;
    assign_one_read(P, r, W, cands, S, lane, ringPhase);
    __syncwarp();
  }
}

}  // namespace t1k
