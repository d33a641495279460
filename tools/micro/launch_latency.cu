// Microbenchmark (round 2): how fast does a B200 retire chains of small dependent kernels, from 1 and from 4 streams?
// The solver's coarse multigrid levels are such chains (~280 k launches per 4096^2 solve).  Variants: same / mixed shared-memory
// footprints (a different carve-out between consecutive kernels forces an SM reconfiguration), eager launches vs CUDA graphs,
// programmatic dependent launch.  nvcc -O3 -gencode arch=compute_100a,code=sm_100a launch_latency.cu -o launch_latency
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>
#include <thread>
#include <chrono>

template <int SMEM> __global__ void k_small(float* p, int n, int pdl) {
  extern __shared__ float sm[];
  if (pdl) asm volatile("griddepcontrol.launch_dependents;");
  if (pdl) asm volatile("griddepcontrol.wait;" ::: "memory");
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (SMEM > 0) { sm[threadIdx.x] = i < n ? p[i] : 0.f; __syncthreads(); if (i < n) p[i] = sm[threadIdx.x ^ 1] * 0.999f + 1e-3f; }
  else if (i < n) p[i] = p[i] * 0.999f + 1e-3f;
}

static void launch(int variant, int k, float* p, int n, cudaStream_t st, bool pdl) {
  const int threads = 128, blocks = (n + threads - 1) / threads;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(blocks); cfg.blockDim = dim3(threads); cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
  const int mixed = variant == 1 ? (k % 3) : 0;
  const int ip = pdl ? 1 : 0;
  if (variant == 2) { cfg.dynamicSmemBytes = 0; cudaLaunchKernelEx(&cfg, k_small<0>, p, n, ip); return; }
  if (mixed == 0) { cfg.dynamicSmemBytes = 512; cudaLaunchKernelEx(&cfg, k_small<1>, p, n, ip); }
  else if (mixed == 1) { cfg.dynamicSmemBytes = 20 * 1024; cudaLaunchKernelEx(&cfg, k_small<1>, p, n, ip); }
  else { cfg.dynamicSmemBytes = 0; cudaLaunchKernelEx(&cfg, k_small<0>, p, n, ip); }
}

int main() {
  const int NS = 4, CHAIN = 2000, REPS = 5;
  cudaFuncSetAttribute(k_small<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  std::vector<cudaStream_t> st(NS);
  std::vector<float*> buf(NS);
  for (int s = 0; s < NS; ++s) { cudaStreamCreateWithFlags(&st[s], cudaStreamNonBlocking); }
  for (int npts : {4096, 65536, 1 << 20}) {
    for (int s = 0; s < NS; ++s) { cudaMalloc(&buf[s], npts * sizeof(float)); cudaMemset(buf[s], 0, npts * sizeof(float)); }
    for (int variant = 0; variant < 3; ++variant) {           // 0: same smem, 1: mixed smem footprints, 2: no smem
      for (int pdl = 0; pdl < 2; ++pdl) {
        // graphs
        std::vector<cudaGraphExec_t> ge(NS);
        for (int s = 0; s < NS; ++s) {
          cudaGraph_t g;
          cudaStreamBeginCapture(st[s], cudaStreamCaptureModeRelaxed);
          for (int k = 0; k < CHAIN; ++k) launch(variant, k, buf[s], npts, st[s], pdl);
          cudaStreamEndCapture(st[s], &g);
          cudaGraphInstantiate(&ge[s], g, 0);
          cudaGraphDestroy(g);
        }
        for (int nstreams : {1, 2, 4}) {
          for (int s = 0; s < nstreams; ++s) cudaGraphLaunch(ge[s], st[s]);
          cudaDeviceSynchronize();
          auto t0 = std::chrono::steady_clock::now();
          for (int r = 0; r < REPS; ++r) for (int s = 0; s < nstreams; ++s) cudaGraphLaunch(ge[s], st[s]);
          cudaDeviceSynchronize();
          const double us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count();
          printf("npts=%7d variant=%d pdl=%d graph streams=%d : %.2f us per kernel per stream, %.2f us per kernel aggregate  (%s)\n", npts, variant, pdl, nstreams,
                 us / (REPS * CHAIN), us / (REPS * CHAIN * nstreams), cudaGetErrorString(cudaGetLastError()));
        }
        for (int s = 0; s < NS; ++s) cudaGraphExecDestroy(ge[s]);
      }
    }
    // eager launches from 4 host threads
    for (int nstreams : {1, 4}) {
      auto t0 = std::chrono::steady_clock::now();
      std::vector<std::thread> th;
      for (int s = 0; s < nstreams; ++s) th.emplace_back([&, s]() { for (int k = 0; k < CHAIN * REPS; ++k) launch(0, k, buf[s], npts, st[s], false); cudaStreamSynchronize(st[s]); });
      for (auto& t : th) t.join();
      const double us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count();
      printf("npts=%7d eager threads=%d : %.2f us per kernel per stream, %.2f aggregate\n", npts, nstreams, us / (REPS * CHAIN), us / (REPS * CHAIN * nstreams));
    }
    for (int s = 0; s < NS; ++s) cudaFree(buf[s]);
  }
  return 0;
}
