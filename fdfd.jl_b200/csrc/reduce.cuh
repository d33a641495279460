// reduce.cuh -- deterministic two-stage reductions: warp shuffles -> one partial per CTA in HBM ->
// a fixed-order final sum done by the (tiny) scalar kernels of the Krylov loop.  No atomics, so a solve
// is bit-reproducible run to run.
#pragma once
#include <cuda_runtime.h>

template <typename T> __device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// every thread passes K accumulators; thread 0 stores the K block sums to out[0..K)
template <int THREADS, int K> __device__ __forceinline__ void block_reduce_store(const double* acc, double* out) {
  constexpr int NW = THREADS / 32;
  __shared__ double sm[K][NW];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    double v = warp_sum(acc[k]);
    if (lane == 0) sm[k][w] = v;
  }
  __syncthreads();
  if (threadIdx.x < K) {
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < NW; ++i) s += sm[threadIdx.x][i];
    out[threadIdx.x] = s;
  }
}

// final stage: one CTA sums `count` partial K-tuples laid out [block][K]; result valid in every thread
template <int THREADS, int K> __device__ __forceinline__ void final_reduce(const double* partials, int count, double* res) {
  constexpr int NW = THREADS / 32;
  __shared__ double sm2[K][NW];
  __shared__ double tot[K];
  double acc[K];
#pragma unroll
  for (int k = 0; k < K; ++k) acc[k] = 0.0;
  for (int i = threadIdx.x; i < count; i += THREADS) {
#pragma unroll
    for (int k = 0; k < K; ++k) acc[k] += partials[(size_t)i * K + k];
  }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    double v = warp_sum(acc[k]);
    if (lane == 0) sm2[k][w] = v;
  }
  __syncthreads();
  if (threadIdx.x < K) {
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < NW; ++i) s += sm2[threadIdx.x][i];
    tot[threadIdx.x] = s;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < K; ++k) res[k] = tot[k];
  __syncthreads();
}
