// The global tail of the hot path on the device: Genotyper::FinalizeReadAssignments' inverse lists + BuildAlleleEquivalentClass
// (Genotyper.hpp:912-939, 1072-1139) and the assembly of the EM problem (Genotyper.hpp:1155-1232) from the coalesced read groups.
// On the host this is three passes over the 10^7..10^8 (group, allele) entries (transposition, fingerprints, distinct classes
// per group) that every rank of a read-sharded run repeats; here the incidence lives as an ALLELE x GROUP BIT MATRIX in HBM:
//   k_tail_bits     warp per group: bit (allele, group) for every entry
//   k_tail_fp       thread per allele: its ascending group list = the set bits of its row -> the reference's polynomial
//                   fingerprint (Genotyper.hpp:1089, same 32-bit arithmetic), a 64-bit row hash, the list length
//   (host: alleles sorted by (fingerprint desc, allele asc) as the reference sorts them; alleles of one (fingerprint, hash,
//    length) bucket are candidates for one class)
//   k_tail_cmp      warp per (first of bucket, other) pair: rows compared word for word — equal lists, not equal hashes,
//                   make a class; a bucket that fails the comparison sends the whole call to the host path
//   k_tail_rowlen / k_tail_rowfill   warp per group: the group's classes in first-appearance order = the classes of the entries
//                   whose allele is its class's first member (members of a class sit in the same groups, rows are in ascending
//                   allele order) -> CSR of the EM's matrix, ordered compaction by ballot
//   k_tail_collen / k_tail_colfill   warp per class: the set bits of its first member's row inside the rank's row range -> CSC
//                   (ascending group order: the EM's fixed summation order)
//   k_scan_i64      one block: exclusive scan of a length array (<= a few 10^5 entries)
// The EM kernels then run on these arrays where they lie: nothing of the matrix crosses PCIe.
#pragma once
#include "t1k_kernels.cuh"

namespace t1k {

struct TailParams {
  int32_t nGroups, nAlleles, nEc;
  const int64_t *gPtr;          // [nGroups + 1]
  const int32_t *gAllele;       // allele id of every entry (rows in ascending allele order)
  u64 *bits; int32_t wordsPerRow;      // [nAlleles][wordsPerRow]
  int32_t *fp; u64 *rowHash; int32_t *listLen;      // per allele
  const int32_t *alleleEc;      // per allele: class (-1: none)
  const u8 *isRep;              // per allele: first member of its class
  const int32_t *ecRep;         // per class: its first member
  int64_t *rowLen, *rowPtr;     // per group (+1)
  int32_t *col;
  int32_t g0, g1;               // row range of this rank
  int64_t *colLen, *colBeg;     // per class (+1)
  int32_t *rowIdx;
};

__global__ void k_tail_bits(TailParams P) {
  const int g = (int)(((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (g >= P.nGroups) return;
  const u64 bit = 1ull << (g & 63);
  const int w = g >> 6;
  for (int64_t k = P.gPtr[g] + lane; k < P.gPtr[g + 1]; k += 32)
    atomicOr((unsigned long long *)(P.bits + (size_t)P.gAllele[k] * P.wordsPerRow + w), (unsigned long long)bit);
}

__global__ void k_tail_fp(TailParams P) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= P.nAlleles) return;
  const u64 *row = P.bits + (size_t)a * P.wordsPerRow;
  const u32 readCnt = (u32)P.nGroups;
  u32 b = 0; int n = 0;
  u64 h = 0x9e3779b97f4a7c15ull;
  for (int w = 0; w < P.wordsPerRow; ++w) {
    u64 x = row[w];
    if (!x) continue;
    h = (h ^ x ^ ((u64)w << 48)) * 0xff51afd7ed558ccdull; h ^= h >> 31;
    while (x) {
      const u32 g = (u32)(w * 64 + __ffsll((long long)x) - 1);
      x &= x - 1;
      b = (b * readCnt + g) % 1000003u;               // Genotyper.hpp:1089, unsigned 32-bit wrap-around included
      ++n;
    }
  }
  P.fp[a] = n ? (int32_t)b : -1;
  P.rowHash[a] = h;
  P.listLen[a] = n;
}

// pairs[2i], pairs[2i+1]: alleles whose rows must be equal; differ[0] != 0 afterwards: some pair is not
__global__ void k_tail_cmp(TailParams P, const int32_t *pairs, int nPairs, int *differ) {
  const int i = (int)(((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (i >= nPairs) return;
  const u64 *ra = P.bits + (size_t)pairs[2 * i] * P.wordsPerRow, *rb = P.bits + (size_t)pairs[2 * i + 1] * P.wordsPerRow;
  bool d = false;
  for (int w = lane; w < P.wordsPerRow; w += 32) d |= ra[w] != rb[w];
  if (__any_sync(FULL, d) && lane == 0) atomicOr(differ, 1);
}

__global__ void k_tail_rowlen(TailParams P) {
  const int g = (int)(((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (g >= P.nGroups) return;
  int n = 0;
  for (int64_t k = P.gPtr[g] + lane; k < P.gPtr[g + 1]; k += 32) n += P.isRep[P.gAllele[k]];
  n = __reduce_add_sync(FULL, n);
  if (lane == 0) P.rowLen[g] = n;
}
__global__ void k_tail_rowfill(TailParams P) {
  const int g = (int)(((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (g >= P.nGroups) return;
  int64_t out = P.rowPtr[g];
  const int64_t k1 = P.gPtr[g + 1];
  for (int64_t k0 = P.gPtr[g]; k0 < k1; k0 += 32) {
    const int64_t k = k0 + lane;
    int a = 0; bool keep = false;
    if (k < k1) { a = P.gAllele[k]; keep = P.isRep[a] != 0; }
    const unsigned bal = __ballot_sync(FULL, keep);
    if (keep) P.col[out + __popc(bal & ((1u << lane) - 1))] = P.alleleEc[a];
    out += __popc(bal);
  }
}

__device__ __forceinline__ u64 tail_range_mask(int w, int g0, int g1) {      // bits of word w that lie in [g0, g1)
  const int lo = w * 64, hi = lo + 64;
  if (hi <= g0 || lo >= g1) return 0;
  u64 m = ~0ull;
  if (g0 > lo) m &= ~0ull << (g0 - lo);
  if (g1 < hi) m &= ~0ull >> (hi - g1);
  return m;
}
__global__ void k_tail_collen(TailParams P) {
  const int e = (int)(((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (e >= P.nEc) return;
  const u64 *row = P.bits + (size_t)P.ecRep[e] * P.wordsPerRow;
  int n = 0;
  for (int w = (P.g0 >> 6) + lane; w <= ((P.g1 - 1) >> 6) && P.g1 > P.g0; w += 32) n += __popcll(row[w] & tail_range_mask(w, P.g0, P.g1));
  n = __reduce_add_sync(FULL, n);
  if (lane == 0) P.colLen[e] = n;
}
__global__ void k_tail_colfill(TailParams P) {
  const int e = (int)(((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (e >= P.nEc || P.g1 <= P.g0) return;
  const u64 *row = P.bits + (size_t)P.ecRep[e] * P.wordsPerRow;
  int64_t out = P.colBeg[e];
  const int wEnd = (P.g1 - 1) >> 6;
  for (int w0 = P.g0 >> 6; w0 <= wEnd; w0 += 32) {
    const int w = w0 + lane;
    u64 x = w <= wEnd ? (row[w] & tail_range_mask(w, P.g0, P.g1)) : 0;
    const int c = __popcll(x);
    int incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(FULL, incl, o); if (lane >= o) incl += t; }
    int64_t at = out + incl - c;
    while (x) { P.rowIdx[at++] = w * 64 + __ffsll((long long)x) - 1; x &= x - 1; }
    out += __shfl_sync(FULL, incl, 31);
  }
}

// allele runs of the merged partitions (as the all-gather left them in HBM) into the global first-appearance order:
// group k's run starts at byte srcOff[k] of `blobs` and goes to dst[dstPtr[k] .. dstPtr[k + 1])
__global__ void k_tail_gather(const uint8_t *blobs, const int64_t *srcOff, const int64_t *dstPtr, int nGroups, int32_t *dst) {
  const int g = (int)(((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (g >= nGroups) return;
  const int32_t *src = reinterpret_cast<const int32_t *>(blobs + srcOff[g]);
  const int64_t b = dstPtr[g], n = dstPtr[g + 1] - b;
  for (int64_t k = lane; k < n; k += 32) dst[b + k] = src[k];
}

// out[0..n] = exclusive scan of len[0..n) (one block)
__global__ void __launch_bounds__(1024) k_scan_i64(const int64_t *len, int n, int64_t *out) {
  __shared__ int64_t warpSum[32];
  __shared__ int64_t carry;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int b0 = 0; b0 < n; b0 += 1024) {
    const int i = b0 + threadIdx.x;
    const int64_t v = i < n ? len[i] : 0;
    int64_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int64_t t = __shfl_up_sync(FULL, incl, o); if (lane >= o) incl += t; }
    if (lane == 31) warpSum[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      const int64_t w = warpSum[lane]; int64_t iw = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int64_t t = __shfl_up_sync(FULL, iw, o); if (lane >= o) iw += t; }
      warpSum[lane] = iw - w;
    }
    __syncthreads();
    const int64_t excl = carry + warpSum[warp] + incl - v;
    if (i < n) out[i] = excl;
    __syncthreads();
    if (threadIdx.x == 1023) carry = excl + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) out[n] = carry;
}

}  // namespace t1k
