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

__global__ void k_em_colsum(int nEc, const int64_t *__restrict__ colPtr, const int32_t *__restrict__ rowIdx,
                            const double *__restrict__ count, const double *__restrict__ psum,
                            const double *__restrict__ x, double *__restrict__ rc) {
  const int e = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  if (e >= nEc) return;
  const int64_t b = colPtr[e], en = colPtr[e + 1];
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

}  // namespace t1k
