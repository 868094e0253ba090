// EM over read groups x allele equivalence classes on the device: Genotyper::EMupdate (Genotyper.hpp:372-421),
// SQUAREMalpha (:424-437) and the vector steps of QuantifyAlleleEquivalentClass (:1234-1290).  All fp64.
//
// The incidence matrix is binary and sparse (<1 % dense at HLA scale), so the E-step is two
// segmented reductions with a fixed summation order (bit-reproducible run to run, and independent of the grid):
//   k_em_rowsum  warp per read group:  psum[g] = sum_{e in row g} x[e]            (CSR, 4 B/nnz)
//   k_em_colsum  warp per EC:          rc[e]   = sum_{g in col e} count[g]*(x[e]/psum[g])   (CSC, 4 B/nnz)
// followed by single-block vector kernels (E <= #alleles) for the M-step and the SQUAREM extrapolation.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace t1k {

__device__ __forceinline__ double warp_sum_f64(double v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// deterministic block sum (blockDim.x = 1024); result valid in every thread
__device__ __forceinline__ double block_sum_f64(double v, double *sh) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  v = warp_sum_f64(v);
  __syncthreads();
  if (lane == 0) sh[w] = v;
  __syncthreads();
  double t = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.0;
  if (w == 0) t = warp_sum_f64(t);
  if (threadIdx.x == 0) sh[32] = t;
  __syncthreads();
  return sh[32];
}

__global__ void k_em_rowsum(int nGroups, const int64_t *__restrict__ rowPtr, const int32_t *__restrict__ col,
                            const double *__restrict__ x, double *__restrict__ psum) {
  const int g = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  if (g >= nGroups) return;
  const int64_t b = rowPtr[g], e = rowPtr[g + 1];
  double s = 0;
  for (int64_t k = b + lane; k < e; k += 32) s += x[col[k]];
  s = warp_sum_f64(s);
  if (lane == 0) psum[g] = s == 0 ? 1.0 : s;      // Genotyper.hpp:393-394
}

__global__ void k_em_colsum(int nEc, const int64_t *__restrict__ colBeg, const int64_t *__restrict__ colEnd, const int32_t *__restrict__ rowIdx,
                            const double *__restrict__ count, const double *__restrict__ psum,
                            const double *__restrict__ x, double *__restrict__ rc) {
  const int e = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  if (e >= nEc) return;
  const int64_t b = colBeg[e], en = colEnd[e];
  const double xe = x[e];
  double s = 0;
  for (int64_t k = b + lane; k < en; k += 32) { const int g = rowIdx[k]; s += count[g] * (xe / psum[g]); }
  s = warp_sum_f64(s);
  if (lane == 0) rc[e] = s;
}

// M-step (Genotyper.hpp:406-419): xNext[e] = rc[e]/len[e] / sum(rc/len).  One block of 1024 threads.
__global__ void __launch_bounds__(1024) k_em_mstep(int nEc, const double *__restrict__ rc, const int32_t *__restrict__ len,
                                                    double *__restrict__ xNext) {
  __shared__ double sh[33];
  double s = 0;
  for (int e = threadIdx.x; e < nEc; e += blockDim.x) s += rc[e] / len[e];
  const double norm = block_sum_f64(s, sh);
  for (int e = threadIdx.x; e < nEc; e += blockDim.x) xNext[e] = rc[e] / len[e] / norm;
}

// SQUAREM extrapolation (Genotyper.hpp:1242-1261): alpha from (x0,x1,x2), x3 = x0 - 2a(x1-x0) + a^2(x2-2x1+x0)
__global__ void __launch_bounds__(1024) k_em_squarem(int nEc, const double *__restrict__ x0, const double *__restrict__ x1,
                                                      const double *__restrict__ x2, double minAlpha, double *__restrict__ x3) {
  __shared__ double sh[33];
  double sr = 0, sv = 0;
  for (int e = threadIdx.x; e < nEc; e += blockDim.x) {
    const double r = x1[e] - x0[e], v = x2[e] - 2 * x1[e] + x0[e];
    sr += r * r; sv += v * v;
  }
  sr = block_sum_f64(sr, sh);
  sv = block_sum_f64(sv, sh);
  double alpha = sv == 0 ? -1.0 : -sqrt(sr) / sqrt(sv);
  if (minAlpha < 0 && alpha < minAlpha) alpha = minAlpha;
  for (int e = threadIdx.x; e < nEc; e += blockDim.x)
    x3[e] = x0[e] - 2 * alpha * (x1[e] - x0[e]) + alpha * alpha * (x2[e] - 2 * x1[e] + x0[e]);
}

// diffSum = sum |x1 - x0| ; x0 = x1   (Genotyper.hpp:1279-1287)
__global__ void __launch_bounds__(1024) k_em_advance(int nEc, double *__restrict__ x0, const double *__restrict__ x1,
                                                      double *__restrict__ diffOut) {
  __shared__ double sh[33];
  double d = 0;
  for (int e = threadIdx.x; e < nEc; e += blockDim.x) { d += fabs(x1[e] - x0[e]); x0[e] = x1[e]; }
  d = block_sum_f64(d, sh);
  if (threadIdx.x == 0) *diffOut = d;
}

// ---- reference-order variants: every sum runs in the reference's own (sequential) order with separately rounded
// multiplies and adds (no FMA contraction), so the abundances equal the reference's x86 results bit for bit and the
// convergence test `diffSum < 1e-5` (Genotyper.hpp:1289) sees the very same number.  Elementwise work stays parallel;
// only the dependent add chains are serial (one thread per row / per column / per vector sum).
__global__ void k_em_rowsum_seq(int nGroups, const int64_t *__restrict__ rowPtr, const int32_t *__restrict__ col,
                                const double *__restrict__ x, double *__restrict__ psum) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= nGroups) return;
  double s = 0;
  for (int64_t k = rowPtr[g]; k < rowPtr[g + 1]; ++k) s = __dadd_rn(s, x[col[k]]);
  psum[g] = s == 0 ? 1.0 : s;
}

// One warp per EC column: the lanes compute 32 terms count[g] * (x[e] / psum[g]) at once (coalesced index loads, the
// divisions in parallel), then every lane adds them in ascending group order through shuffles — the same roundings in
// the same order as the reference's serial loop (Genotyper.hpp:391-404), without one thread issuing 60 instructions per
// entry of a 50 k-entry column.
__global__ void k_em_colsum_seq(int nEc, const int64_t *__restrict__ colBeg, const int64_t *__restrict__ colEnd, const int32_t *__restrict__ rowIdx,
                                const double *__restrict__ count, const double *__restrict__ psum,
                                const double *__restrict__ x, double *__restrict__ rc) {
  const int e = (int)(((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (e >= nEc) return;
  const double xe = x[e];
  double s = 0;
  const int64_t k1 = colEnd[e];
  for (int64_t b = colBeg[e]; b < k1; b += 32) {
    const int64_t k = b + lane;
    double t = 0;
    if (k < k1) { const int g = rowIdx[k]; t = __dmul_rn(count[g], __ddiv_rn(xe, psum[g])); }
    const int m = k1 - b < 32 ? (int)(k1 - b) : 32;
    if (m == 32) {
#pragma unroll
      for (int j = 0; j < 32; ++j) s = __dadd_rn(s, __shfl_sync(0xffffffffu, t, j));
    } else {
      for (int j = 0; j < m; ++j) s = __dadd_rn(s, __shfl_sync(0xffffffffu, t, j));
    }
  }
  if (lane == 0) rc[e] = s;
}

// tmp[e] = f(e) elementwise, then tmp[0..n) added in index order.  Called by a whole warp: the lanes load 32 consecutive terms
// at once (the next 32 are already in flight) and every lane runs the same chain of separately rounded adds over them through
// shuffles — the reference's order and roundings, without one thread waiting for 26 k dependent global loads.
__device__ __forceinline__ double seq_sum(const double *tmp, int n) {
  const int lane = threadIdx.x & 31;
  double s = 0;
  double t = lane < n ? tmp[lane] : 0.0;
  for (int b = 0; b < n; b += 32) {
    const int nb = b + 32 + lane;
    const double tn = nb < n ? tmp[nb] : 0.0;
    const int m = n - b < 32 ? n - b : 32;
    if (m == 32) {
#pragma unroll
      for (int j = 0; j < 32; ++j) s = __dadd_rn(s, __shfl_sync(0xffffffffu, t, j));
    } else {
      for (int j = 0; j < m; ++j) s = __dadd_rn(s, __shfl_sync(0xffffffffu, t, j));
    }
    t = tn;
  }
  return s;
}

__global__ void __launch_bounds__(1024) k_em_mstep_seq(int nEc, const double *__restrict__ rc, const int32_t *__restrict__ len,
                                                        double *__restrict__ tmp, double *__restrict__ xNext) {
  __shared__ double norm;
  for (int e = threadIdx.x; e < nEc; e += blockDim.x) tmp[e] = __ddiv_rn(rc[e], (double)len[e]);
  __syncthreads();
  if (threadIdx.x < 32) { const double v = seq_sum(tmp, nEc); if (threadIdx.x == 0) norm = v; }
  __syncthreads();
  const double nrm = norm;
  for (int e = threadIdx.x; e < nEc; e += blockDim.x) xNext[e] = __ddiv_rn(tmp[e], nrm);
}

__global__ void __launch_bounds__(1024) k_em_squarem_seq(int nEc, const double *__restrict__ x0, const double *__restrict__ x1,
                                                          const double *__restrict__ x2, double minAlpha, double *__restrict__ tmpR,
                                                          double *__restrict__ tmpV, double *__restrict__ x3) {
  __shared__ double sAlpha;
  for (int e = threadIdx.x; e < nEc; e += blockDim.x) {
    const double r = __dsub_rn(x1[e], x0[e]);
    const double v = __dadd_rn(__dsub_rn(x2[e], __dmul_rn(2.0, x1[e])), x0[e]);
    tmpR[e] = __dmul_rn(r, r); tmpV[e] = __dmul_rn(v, v);
  }
  __syncthreads();
  __shared__ double sSv;
  if (threadIdx.x < 32) { const double sr = seq_sum(tmpR, nEc); if (threadIdx.x == 0) sAlpha = sr; }
  else if (threadIdx.x < 64) { const double sv = seq_sum(tmpV, nEc); if (threadIdx.x == 32) sSv = sv; }
  __syncthreads();
  if (threadIdx.x == 0) {
    const double sr = sAlpha, sv = sSv;
    double alpha = sv == 0 ? -1.0 : __ddiv_rn(-sqrt(sr), sqrt(sv));
    if (minAlpha < 0 && alpha < minAlpha) alpha = minAlpha;
    sAlpha = alpha;
  }
  __syncthreads();
  const double alpha = sAlpha;
  const double a2 = __dmul_rn(2.0, alpha), aa = __dmul_rn(alpha, alpha);
  for (int e = threadIdx.x; e < nEc; e += blockDim.x) {
    const double r = __dsub_rn(x1[e], x0[e]);
    const double v = __dadd_rn(__dsub_rn(x2[e], __dmul_rn(2.0, x1[e])), x0[e]);
    x3[e] = __dadd_rn(__dsub_rn(x0[e], __dmul_rn(a2, r)), __dmul_rn(aa, v));
  }
}

__global__ void __launch_bounds__(1024) k_em_advance_seq(int nEc, double *__restrict__ x0, const double *__restrict__ x1,
                                                          double *__restrict__ tmp, double *__restrict__ diffOut) {
  for (int e = threadIdx.x; e < nEc; e += blockDim.x) { tmp[e] = fabs(__dsub_rn(x1[e], x0[e])); x0[e] = x1[e]; }
  __syncthreads();
  if (threadIdx.x < 32) { const double v = seq_sum(tmp, nEc); if (threadIdx.x == 0) *diffOut = v; }
}

}  // namespace t1k
