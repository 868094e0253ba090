// the analyzer's second pass over the alignments — SeqSet::AddOverlapAlignmentInfo (SeqSet.hpp:2657-2680),
// called through AddFragmentAlignmentInfo (SeqSet.hpp:2757-2778) for every (fragment, allele) the analyzer kept
// (Analyzer.cpp:624-669): the edit string of GlobalAlignment(allele[seqStart..seqEnd], read strand[readStart..readEnd]).
//
// k_align_info: one lane per (read, overlap) item, the lanes of a warp on 32 consecutive items.  Each lane leaves its edit
// string in its scratch row (certified diagonal: written from the mismatch plane; otherwise the band DP with DPX min/max and
// the traceback), then the warp copies the 32 strings to their output slots 16 bytes per lane (slots are 16-byte aligned and
// consecutive for consecutive items, so the stores coalesce).  A string ends with -1 as the reference's does
// (AlignAlgo.hpp:409).  Also k_dpx_peak: the DPX issue-rate micro-benchmark the band DP's cells/s are reported against.
#pragma once
#include "t1k_core.cuh"
#include "t1k_kernels.cuh"

namespace t1k {

struct OvIn { int32_t seqIdx, readStart, readEnd, seqStart, seqEnd, strand, matchCnt, relaxedMatchCnt, leftClip, rightClip; };   // == T1KOverlap

struct AlnInfoParams {
  RefView R;
  const u64 *planes; int rwords, maxLen; const u16 *len;     // reads packed by k_pack_reads
  const OvIn *ov; const u32 *readIdx; u32 nItems;
  const u64 *slot;         // byte offset of every item's output slot (multiple of 16; ~0: no alignment, seqIdx == -1)
  u8 *out;
  u8 *laneScratch;         // per lane scr_bytes(maxLen)
  int *err;
  unsigned long long *stats;   // [0] certified diagonals, [1] band DPs, [2] band cells of those DPs
  int noDiag;
};

__global__ void __launch_bounds__(128) k_align_info(AlnInfoParams P) {
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nThreads = (size_t)gridDim.x * blockDim.x;
  const int lane = threadIdx.x & 31;
  const LaneScratch S = lane_scratch(P.laneScratch + tid * scr_bytes(P.maxLen), P.maxLen);
  const int RW = P.rwords;
  int err = 0;
  unsigned long long nDiag = 0, nDp = 0, cells = 0;
  const size_t nRound = ((size_t)P.nItems + nThreads - 1) / nThreads;
  for (size_t rnd = 0; rnd < nRound; ++rnd) {
    const size_t i = rnd * nThreads + tid;
    int n = -1;
    u64 slot = ~0ull;
    if (i < P.nItems) {
      slot = P.slot[i];
      if (slot != ~0ull) {
        const OvIn o = P.ov[i];
        const u32 r = P.readIdx[i];
        const u64 *pl = P.planes + ((size_t)r * 4 + (o.strand == 1 ? 0 : 2)) * RW;
        ReadView Q; Q.seq2 = pl; Q.n2 = pl + RW; Q.len = P.len[r];
        u64 nw = 0;
        T1K_NOUNROLL
        for (int k = 0; k < RW; ++k) nw |= pl[RW + k];
        Q.anyN = nw != 0;
        const AlleleView T = allele_view(P.R, o.seqIdx, Q);
        const int lent = o.seqEnd - o.seqStart + 1, lenp = o.readEnd - o.readStart + 1;
        bool ranDp;
        n = align_info(T, o.seqStart, lent, Q, o.readStart, lenp, S, err, P.noDiag != 0, ranDp);
        if (n >= 0) {
          S.ops()[n] = 0xFF;
          if (ranDp) { ++nDp; cells += (unsigned long long)lenp * (unsigned)(2 * BAND + 1 + iabs(lent - lenp)); } else ++nDiag;
        }
      }
    }
    __syncwarp();
    T1K_NOUNROLL
    for (int j = 0; j < 32; ++j) {
      const int nj = __shfl_sync(FULL, n, j);
      if (nj < 0) continue;
      const u64 sj = __shfl_sync(FULL, slot, j);
      const uint4 *src = (const uint4 *)(P.laneScratch + (tid - lane + j) * scr_bytes(P.maxLen));
      uint4 *dst = (uint4 *)(P.out + sj);
      for (int c = lane; c < ((nj + 1 + 15) >> 4); c += 32) dst[c] = src[c];
    }
    __syncwarp();
  }
  if (err) atomicOr(P.err, err);
  if (nDiag) atomicAdd(P.stats, nDiag);
  if (nDp) { atomicAdd(P.stats + 1, nDp); atomicAdd(P.stats + 2, cells); }
}

// DPX issue rate: every thread runs `iters` rounds of 8 independent max(a + b, c) chains (VIADDMNMX); result written so that
// nothing is optimised away.  ops = threads * iters * 8.
__global__ void __launch_bounds__(256) k_dpx_peak(int iters, int seed, int *sink) {
  int a0 = seed + threadIdx.x, a1 = a0 ^ 0x55, a2 = a0 + 7, a3 = a0 - 9, a4 = a0 * 3, a5 = a0 ^ 0x1234, a6 = a0 + 77, a7 = a0 - 123;
  const int b = (int)blockIdx.x - 3, c = seed - 1000;
#pragma unroll 8
  for (int i = 0; i < iters; ++i) {
    a0 = __viaddmax_s32(a0, b, c); a1 = __viaddmax_s32(a1, b, c); a2 = __viaddmax_s32(a2, b, c); a3 = __viaddmax_s32(a3, b, c);
    a4 = __viaddmax_s32(a4, b, c); a5 = __viaddmax_s32(a5, b, c); a6 = __viaddmax_s32(a6, b, c); a7 = __viaddmax_s32(a7, b, c);
  }
  const int v = a0 ^ a1 ^ a2 ^ a3 ^ a4 ^ a5 ^ a6 ^ a7;
  if (v == 0x7fffffff) *sink = v;
}

}  // namespace t1k
