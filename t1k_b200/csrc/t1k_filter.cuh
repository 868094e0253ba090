// the candidate filter of fastq-extractor on the device.  One warp = one read: IsGoodCandidate
// (FastqExtractor.cpp:113-118) = !IsLowComplexity (:89-112) && SeqSet::HasHitInSet (SeqSet.hpp:1915-1990).  On whole-genome data
// nearly every read has no or a few stray hits and leaves after the seed look-ups, so the kernel is a stream over the reads
// with two random 8-byte look-ups per k-mer window into the k-mer table (4^k entries, up to 2 GB at k = 14): HBM bound.
//
//   1. low complexity, length < k                                                     -> not a candidate
//   2. seeds of both strands with the skip rule (the previous code carries over from the forward into the reverse pass)
//   3. hits per (strand, sequence) = column sums over the tile index, reverse strand first, sequences ascending: the first
//      largest bucket (SeqSet.hpp:1929-1957)
//   4. k * (largest count) < hitLenRequired                                           -> not a candidate
//   5. the hits of that bucket, chained by the lane code of t1k_filter_lane.cuh (one lane: only candidate reads get here)
#pragma once
#include "t1k_filter_lane.cuh"
#include "t1k_kernels.cuh"

namespace t1k {

struct FilterParams {
  const KmerInfo *kinfo;
  const KmerEntry *entries;
  const u32 *present;      // one bit per k-mer code: the code has postings (32 MB at k = 14: L2-resident, so the windows of a read
                           // that is not from the reference's genes never reach the 2 GB table in HBM)
  int k, hitLenReq;
  double sim;
  const u64 *planes;       // [(r*4 + plane) * rwords]
  int rwords;
  const u16 *len;
  u32 nReads;
  u8 *good;                // out: IsGoodCandidate per read
  u32 *hitBuf;             // per warp: hitCap hits + scratch for the chaining
  int hitCap;
  int seedCap;
  unsigned int *workCtr;
  int *err;
  unsigned long long *stats;   // [0] k-mer windows looked up, [1] entries swept, [2] reads that reached the chaining
};

// shared memory per warp: planes (4 x rwords), window table / current entries ent[S], per strand seedA[S], beg[S], end[S]; cur[S]
__host__ __device__ inline size_t filter_smem_bytes(int seedCap, int rwords) {
  return ((size_t)seedCap * (16 + 4 + 2 * (2 + 4 + 4)) + 4 * (size_t)rwords * 8 + 15) & ~(size_t)15;
}

__device__ void filter_one_read(const FilterParams &P, u32 r, u8 *sm, u32 *hits, int lane) {
  const int SC = P.seedCap, RW = P.rwords, k = P.k;
  uint4 *ent = (uint4 *)sm;
  u64 *pl = (u64 *)(ent + SC);                       // fwd seq2, fwd n2, rc seq2, rc n2
  u32 *cur = (u32 *)(pl + 4 * RW);
  u32 *beg = cur + SC, *end = beg + 2 * SC;          // [2][SC]
  u16 *seedA = (u16 *)(end + 2 * SC);                // [2][SC]
  const int len = P.len[r];
  const u64 *src = P.planes + (size_t)r * 4 * RW;
  for (int w = lane; w < 4 * RW; w += 32) pl[w] = src[w];
  __syncwarp();
  int verdict = 0;
  unsigned long long stWin = 0, stEnt = 0, stChain = 0;
  // ---- 1. IsLowComplexity on the forward strand
  int c0 = 0, c1 = 0, c2 = 0, c3 = 0, cn = 0;
  for (int w = lane; w * 32 < len; w += 32) {
    const u64 x = pl[w], nm = pl[RW + w];
    const u64 inl = lowmask2(len - 32 * w) & M55, keep = ~nm & inl;
    const u64 lo = x & M55, hi = (x >> 1) & M55;
    c0 += __popcll(~lo & ~hi & keep); c1 += __popcll(lo & ~hi & keep); c2 += __popcll(~lo & hi & keep); c3 += __popcll(lo & hi & keep);
    cn += __popcll(nm & inl);
  }
  c0 = warp_sum_i32(c0); c1 = warp_sum_i32(c1); c2 = warp_sum_i32(c2); c3 = warp_sum_i32(c3); cn = warp_sum_i32(cn);
  const bool anyN = cn > 0;                          // (a read without N needs no validity test per window)
  bool lowc = c0 >= len / 2 || c1 >= len / 2 || c2 >= len / 2 || c3 >= len / 2 || cn >= len / 10;
  if (!lowc) lowc = (c0 <= 2) + (c1 <= 2) + (c2 <= 2) + (c3 <= 2) >= 2;
  if (!lowc && len >= k) {
    const int NP = len - k + 1;
    const u64 codeMask = (1ull << (2 * k)) - 1, nMask = codeMask & M55;
    // ---- 2. seeds of both strands
    int nS[2] = {0, 0};
    u32 prev = 0;
    for (int pass = 0; pass < 2; ++pass) {
      const u64 *seq = pl + (pass ? 2 * RW : 0), *nn = seq + RW;
      bool big = false;
      for (int a = lane; a < NP; a += 32) {
        const u32 code = (u32)(fetch32(seq, a) & codeMask);
        const bool valid = !anyN || (fetch32(nn, a) & nMask) == 0;
        uint2 k0 = make_uint2(0, 0), k1 = make_uint2(0, 0);
        if (valid && ((P.present[code >> 5] >> (code & 31u)) & 1u)) {
          k0 = *reinterpret_cast<const uint2 *>(P.kinfo + code); k1 = *reinterpret_cast<const uint2 *>(P.kinfo + code + 1);
        }
        ent[a] = make_uint4(code, k1.y - k0.y, k0.x, k1.x);
        big |= k1.y - k0.y >= 100u;
      }
      stWin += NP;
      __syncwarp();
      int n = 0;
      if (!__any_sync(FULL, big)) {
        // no list reaches the skip rule's threshold (every read that is not from the reference's genes): a window is a seed iff
        // its list is not empty and its code differs from the previous window's -> ordered compaction, no sequential pass
        for (int a0 = 0; a0 < NP; a0 += 32) {
          const int a = a0 + lane;
          bool seed = false;
          uint4 w = make_uint4(0, 0, 0, 0);
          if (a < NP) {
            w = ent[a];
            const u32 before = a == 0 ? prev : ent[a - 1].x;
            seed = w.y > 0 && (a == 0 || before != w.x);
          }
          const unsigned bal = __ballot_sync(FULL, seed);
          if (seed) {
            const int at = pass * SC + n + __popc(bal & ((1u << lane) - 1));
            seedA[at] = (u16)a; beg[at] = w.z; end[at] = w.w;
          }
          n += __popc(bal);
        }
        if (lane == 0) prev = ent[NP - 1].x;
      } else if (lane == 0) {
        int skip = 0;
        for (int a = 0; a < NP; ++a) {
          const uint4 w = ent[a];
          if (a == 0 || prev != w.x) {
            const int size = (int)w.y;
            if (size >= 100 && a != 0 && a != NP - 1 && skip < k / 2) { ++skip; continue; }
            skip = 0;
            if (size > 0) { seedA[pass * SC + n] = (u16)a; beg[pass * SC + n] = w.z; end[pass * SC + n] = w.w; ++n; }
          }
          prev = w.x;
        }
      }
      n = __shfl_sync(FULL, n, 0);
      prev = __shfl_sync(FULL, prev, 0);
      nS[pass] = n;
      __syncwarp();
    }
    // ---- 3. the first largest (strand, sequence) bucket: reverse strand first, sequences ascending
    int bestCnt = 0, bestPass = 0;
    u32 bestSeq = 0;
    for (int q = 0; q < 2; ++q) {
      const int pass = 1 - q;
      const int n = nS[pass];
      const u32 *B = beg + pass * SC, *E = end + pass * SC;
      for (int s = lane; s < n; s += 32) { cur[s] = B[s]; ent[s] = *reinterpret_cast<const uint4 *>(P.entries + B[s]); }
      __syncwarp();
      for (;;) {
        u32 mn = 0xffffffffu;
        for (int s = lane; s < n; s += 32) mn = min(mn, ent[s].x);
        const u32 T = warp_min_u32(mn);
        if (T == 0xffffffffu) break;
        int cnt = 0;
        for (int s0 = 0; s0 < n; s0 += 32) {
          const int s = s0 + lane;
          uint4 me = make_uint4(0xffffffffu, 0, 0, 0);
          if (s < n) me = ent[s];
          const bool act = me.x == T;
          unsigned bal = __ballot_sync(FULL, act);
          if (bal == 0) continue;
          if (!__any_sync(FULL, act && me.w != 0)) {
            u32 x = act ? me.z : 0u, m = 0x0000ffffu;
#pragma unroll
            for (int j = 16; j; j >>= 1) {
              const u32 y = __shfl_xor_sync(FULL, x, j);
              x = (lane & j) ? ((x & ~m) | ((y >> j) & m)) : ((x & m) | ((y << j) & ~m));
              m ^= m << (j >> 1);
            }
            cnt += __popc(x);
            stEnt += __popc(bal);
            continue;
          }
          while (bal) {
            const int ss = s0 + __ffs(bal) - 1;
            bal &= bal - 1;
            uint4 e = ent[ss];
            const u32 more = e.w;
            for (u32 j = 0;; ++j) {
              cnt += (e.z >> lane) & 1u;
              ++stEnt;
              if (j >= more) break;
              e = *reinterpret_cast<const uint4 *>(P.entries + cur[ss] + j + 1);
            }
          }
        }
        const int m = warp_max_i32(cnt);
        if (m > bestCnt) {
          bestCnt = m; bestPass = pass;
          bestSeq = T * 32 + (u32)(__ffs(__ballot_sync(FULL, cnt == m)) - 1);
        }
        __syncwarp();
        for (int s = lane; s < n; s += 32) {
          const uint4 e = ent[s];
          if (e.x == T) {
            const u32 c = cur[s] + 1 + e.w;
            cur[s] = c;
            if (c < E[s]) ent[s] = *reinterpret_cast<const uint4 *>(P.entries + c); else ent[s].x = 0xffffffffu;
          }
        }
        __syncwarp();
      }
    }
    // ---- 4. / 5.
    if (bestCnt > 0 && k * bestCnt >= P.hitLenReq) {
      ++stChain;
      // first entry of every seed at or after the bucket's tile (the lists are sorted by tile): binary search, lane per seed
      const int n = nS[bestPass];
      const u32 *B = beg + bestPass * SC, *E = end + bestPass * SC;
      const u32 T = bestSeq >> 5, bit = bestSeq & 31u;
      for (int s = lane; s < n; s += 32) {
        u32 lo = B[s], hi = E[s];
        while (lo < hi) { const u32 mid = (lo + hi) >> 1; if (P.entries[mid].tile < T) lo = mid + 1; else hi = mid; }
        cur[s] = lo;
      }
      __syncwarp();
      if (lane == 0) {
        int nh = 0;
        bool over = false;
        for (int s = 0; s < n; ++s)
          for (u32 c = cur[s]; c < E[s]; ++c) {
            const KmerEntry e = P.entries[c];
            if (e.tile != T) break;
            if ((e.mask >> bit) & 1u) { if (nh < P.hitCap) hits[nh] = hit_make((int)seedA[bestPass * SC + s], e.off); else over = true; ++nh; }
          }
        if (over || nh != bestCnt) atomicOr(P.err, over ? ERR_HITS : ERR_SCRATCH);
        else {
          const int bestLen = filt::bucket_best_hit_len(hits, nh, k, P.hitLenReq, (u8 *)(hits + P.hitCap));
          verdict = filt::hit_length_passes(len, bestLen, k, P.sim) ? 1 : 0;
        }
      }
    }
  }
  if (lane == 0) {
    P.good[r] = (u8)verdict;
    if (P.stats) { atomicAdd(P.stats + 0, stWin); atomicAdd(P.stats + 2, stChain); }
  }
  if (P.stats) {
    const unsigned long long e = __reduce_add_sync(FULL, (u32)stEnt);
    if (lane == 0) atomicAdd(P.stats + 1, e / 32 + (e % 32 ? 1 : 0));
  }
}

__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32) k_filter(FilterParams P) {
  extern __shared__ u64 t1k_filter_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t gwarp = (size_t)blockIdx.x * WARPS_PER_BLOCK + warp;
  u8 *sm = (u8 *)t1k_filter_smem + (size_t)warp * filter_smem_bytes(P.seedCap, P.rwords);
  u32 *hits = P.hitBuf + gwarp * ((size_t)4 * P.hitCap + filt::FILTER_USED_BYTES / 4);
  for (;;) {
    u32 w = 0;
    if (lane == 0) w = atomicAdd(P.workCtr, 1u);
    w = __shfl_sync(FULL, w, 0);
    if (w >= P.nReads) break;
    filter_one_read(P, w, sm, hits, lane);
    __syncwarp();
  }
}

}  // namespace t1k
