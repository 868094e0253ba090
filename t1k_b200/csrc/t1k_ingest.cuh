// read ingest on the device — packing and de-duplication of the read-ends of a chunk of fragments
// (Genotyper.cpp:363-454: reads into memory, `allReads` sorted by sequence so that equal read-ends are aligned once with
// weight = number of copies; only the grouping matters, not the order).
//
// The chunk's raw reads (fixed stride, NUL padded, as t1k_genotype receives them) are copied to HBM as they are and
//   k_ingest_pack     thread per read-end: length, 2-bit planes of both strands (as k_pack_reads), N flag, 64-bit hash
//   k_dedup_insert    open-addressing table of read-end indices: a read-end claims a free slot (atomicCAS) or, where the
//                     slot's occupant has the SAME bases (planes compared word for word), lowers it to the smaller index
//                     (atomicMin): after the kernel a slot holds the first read-end of its class
//   k_dedup_resolve   every read-end finds its class again: representative, duplicate counts
//   k_scan_*          exclusive scan of the representative flags (unique ids in first-appearance order)
//   k_dedup_emit      representatives: planes, length, weight into the compact arrays the AssignRead kernels read
//   k_dedup_map       fragments: (end1, end2) as unique ids, N flag of the fragment
// Everything is deterministic (the smallest index represents a class; ids follow first appearance), so a chunk always gives
// the same batch.  Host work per chunk: one H2D copy.
#pragma once
#include "t1k_kernels.cuh"

namespace t1k {

struct IngestParams {
  const char *raw1, *raw2;   // [m * stride] each; raw2 NULL: single-end
  u32 stride, m, nEnds;      // nEnds = m * mates; read-end k = fragment k / mates, mate k % mates
  int mates, RW;
  u64 *planesAll;            // [nEnds][4][RW]
  u16 *lenAll;
  u8 *endHasN;
  u64 *hash;
  u32 *table; u32 tabMask;   // slots: read-end index or 0xffffffff
  u32 *repOf;                // representative of every read-end
  u32 *cnt;                  // per representative: copies
  u32 *uid;                  // exclusive scan of the representative flags
  u32 *blockSum;
  // outputs
  u64 *planes; u16 *len16; int32_t *w;       // compact, per unique read-end
  u32 *e1, *e2; u8 *fragHasN;                // per fragment
  u32 *nUnique;              // [0] number of unique read-ends, [1] longest read
  int *err;
};

__device__ __forceinline__ const char *ingest_end(const IngestParams &P, u32 k) {
  const u32 f = P.mates == 2 ? k >> 1 : k;
  return ((P.mates == 2 && (k & 1)) ? P.raw2 : P.raw1) + (size_t)f * P.stride;
}

__global__ void k_ingest_pack(IngestParams P) {
  const u32 k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= P.nEnds) return;
  const char *s = ingest_end(P, k);
  const int RW = P.RW;
  int L = 0;
  while (L < (int)P.stride && s[L]) ++L;
  if (L > (RW - 1) * 32 || L > MAX_READ_LEN) { atomicOr(P.err, ERR_READ_LEN); L = 0; }
  u64 *out = P.planesAll + (size_t)k * 4 * RW;
  for (int w = 0; w < RW; ++w) { out[w] = 0; out[RW + w] = 0; out[2 * RW + w] = 0; out[3 * RW + w] = 0; }
  u64 fs = 0, fn = 0, h = 0x9e3779b97f4a7c15ull ^ (u64)L;
  bool anyN = false;
  for (int j = 0; j < L; ++j) {
    const char c = s[j];
    int v = c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : c == 'T' ? 3 : c == 'N' ? 4 : 5;
    if (v == 5) { atomicOr(P.err, ERR_READ_CHAR); v = 4; }
    anyN |= v == 4;
    const int sh = (j & 31) * 2;
    fs |= (u64)(v == 4 ? 3 : v) << sh;
    fn |= (u64)(v == 4) << sh;
    if ((j & 31) == 31 || j == L - 1) {
      out[j >> 5] = fs; out[RW + (j >> 5)] = fn;
      h = (h ^ fs) * 0xff51afd7ed558ccdull; h ^= h >> 32; h = (h ^ fn) * 0xc4ceb9fe1a85ec53ull; h ^= h >> 29;
      fs = fn = 0;
    }
  }
  u64 rs = 0, rn = 0;
  for (int j = 0; j < L; ++j) {
    const char c = s[L - 1 - j];
    const int v = c == 'A' ? 3 : c == 'C' ? 2 : c == 'G' ? 1 : c == 'T' ? 0 : 4;
    const int sh = (j & 31) * 2;
    rs |= (u64)(v == 4 ? 3 : v) << sh;
    rn |= (u64)(v == 4) << sh;
    if ((j & 31) == 31 || j == L - 1) { out[2 * RW + (j >> 5)] = rs; out[3 * RW + (j >> 5)] = rn; rs = rn = 0; }
  }
  P.lenAll[k] = (u16)L;
  P.endHasN[k] = anyN ? 1 : 0;
  P.hash[k] = h;
  P.cnt[k] = 0;
  atomicMax(P.nUnique + 1, (u32)L);
}

// same bases: equal length and equal forward planes (the reverse planes follow from them)
__device__ __forceinline__ bool ingest_same(const IngestParams &P, u32 a, u32 b) {
  if (a == b) return true;
  const int L = P.lenAll[a];
  if (L != P.lenAll[b]) return false;
  const int nw = (L + 31) >> 5, RW = P.RW;
  const u64 *pa = P.planesAll + (size_t)a * 4 * RW, *pb = P.planesAll + (size_t)b * 4 * RW;
  for (int w = 0; w < nw; ++w) if (pa[w] != pb[w] || pa[RW + w] != pb[RW + w]) return false;
  return true;
}

__global__ void k_dedup_insert(IngestParams P) {
  const u32 k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= P.nEnds) return;
  u32 slot = (u32)(P.hash[k] >> 20) & P.tabMask;
  for (;;) {
    u32 cur = P.table[slot];
    if (cur == 0xffffffffu) {
      cur = atomicCAS(P.table + slot, 0xffffffffu, k);
      if (cur == 0xffffffffu) return;               // claimed
    }
    // the occupant's class never changes (only its index, to another member of the same class), so one compare decides
    if (ingest_same(P, cur, k)) { atomicMin(P.table + slot, k); return; }
    slot = (slot + 1) & P.tabMask;
  }
}

__global__ void k_dedup_resolve(IngestParams P) {
  const u32 k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= P.nEnds) return;
  u32 slot = (u32)(P.hash[k] >> 20) & P.tabMask;
  for (;;) {
    const u32 cur = P.table[slot];                  // (never empty on this probe path: k was inserted along it)
    if (ingest_same(P, cur, k)) { P.repOf[k] = cur; atomicAdd(P.cnt + cur, 1u); return; }
    slot = (slot + 1) & P.tabMask;
  }
}

// exclusive scan of flag[k] = (repOf[k] == k), 1024 read-ends per block
__global__ void __launch_bounds__(1024) k_scan_block(IngestParams P) {
  __shared__ u32 warpSum[32];
  const u32 k = blockIdx.x * 1024 + threadIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const u32 flag = (k < P.nEnds && P.repOf[k] == k) ? 1u : 0u;
  u32 incl = flag;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const u32 t = __shfl_up_sync(FULL, incl, o); if (lane >= o) incl += t; }
  if (lane == 31) warpSum[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    u32 v = warpSum[lane], iv = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const u32 t = __shfl_up_sync(FULL, iv, o); if (lane >= o) iv += t; }
    warpSum[lane] = iv - v;
    if (lane == 31) P.blockSum[blockIdx.x] = iv;
  }
  __syncthreads();
  if (k < P.nEnds) P.uid[k] = warpSum[warp] + incl - flag;
}
// one block: blockSum -> exclusive offsets, total into nUnique[0]
__global__ void __launch_bounds__(1024) k_scan_sums(IngestParams P, u32 nBlocks) {
  __shared__ u32 warpSum[32];
  __shared__ u32 carry;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (u32 b0 = 0; b0 < nBlocks; b0 += 1024) {
    const u32 b = b0 + threadIdx.x;
    const u32 v = b < nBlocks ? P.blockSum[b] : 0u;
    u32 incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const u32 t = __shfl_up_sync(FULL, incl, o); if (lane >= o) incl += t; }
    if (lane == 31) warpSum[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      u32 w = warpSum[lane], iw = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const u32 t = __shfl_up_sync(FULL, iw, o); if (lane >= o) iw += t; }
      warpSum[lane] = iw - w;
    }
    __syncthreads();
    const u32 excl = carry + warpSum[warp] + incl - v;
    if (b < nBlocks) P.blockSum[b] = excl;
    __syncthreads();
    if (threadIdx.x == 1023) carry = excl + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) P.nUnique[0] = carry;
}

__global__ void k_dedup_emit(IngestParams P) {
  const u32 k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= P.nEnds) return;
  const u32 id = P.uid[k] + P.blockSum[k >> 10];
  P.uid[k] = id;
  if (P.repOf[k] != k) return;
  const int n = 4 * P.RW;
  const u64 *src = P.planesAll + (size_t)k * n;
  u64 *dst = P.planes + (size_t)id * n;
  for (int w = 0; w < n; ++w) dst[w] = src[w];
  P.len16[id] = P.lenAll[k];
  P.w[id] = (int32_t)P.cnt[k];
}

__global__ void k_dedup_map(IngestParams P) {
  const u32 f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= P.m) return;
  if (P.mates == 2) {
    P.e1[f] = P.uid[P.repOf[2 * f]]; P.e2[f] = P.uid[P.repOf[2 * f + 1]];
    P.fragHasN[f] = P.endHasN[2 * f] | P.endHasN[2 * f + 1];
  } else {
    P.e1[f] = P.uid[P.repOf[f]];
    P.fragHasN[f] = P.endHasN[f];
  }
}

// records of both mates' lists summed over the fragments (k_pair's algorithmic reads; statistics only)
__global__ void k_pair_records(const u32 *e1, const u32 *e2, u32 nFrag, const u32 *readCnt, unsigned long long *out) {
  const u32 f = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long v = 0;
  if (f < nFrag) { v = readCnt[e1[f]]; if (e2) v += readCnt[e2[f]]; }
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  if ((threadIdx.x & 31) == 0 && v) atomicAdd(out, v);
}

}  // namespace t1k
