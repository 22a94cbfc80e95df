// developer probe (GPU box, >= 2 GPUs): what a copy-engine pull of remote slices costs.
//   nvcc -O2 -arch=sm_100a -o p2p_probe p2p_probe.cu && ./p2p_probe
// Pulls `count` chunks of `bytes` each from GPU 1 into GPU 0 with cudaMemcpyAsync on 1/2/4/8
// streams, idle and while a DMMA-bound kernel occupies every SM of GPU 0; prints GB/s and the
// per-copy time.  Sizes follow the engine's slices: 143 KB (c2 B slice), 1.6 MB (merged B range),
// 5.7 MB (c2 A slice), 80 MB (c4 A slice).
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#define OK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__global__ void busy(double *out, int iters) {
  double acc[8][2];
  for (int i = 0; i < 8; i++) acc[i][0] = acc[i][1] = 0;
  double a = threadIdx.x * 1e-3, b = threadIdx.x * 2e-3;
  for (int it = 0; it < iters; it++)
#pragma unroll
    for (int i = 0; i < 8; i++)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(acc[i][0]), "+d"(acc[i][1]) : "d"(a), "d"(b));
  double s = 0;
  for (int i = 0; i < 8; i++) s += acc[i][0] + acc[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  int n = 0;
  OK(cudaGetDeviceCount(&n));
  if (n < 2) { printf("needs 2 GPUs\n"); return 0; }
  const size_t pool = 1ull << 30;
  char *src, *dst;
  OK(cudaSetDevice(1));
  OK(cudaMalloc(&src, pool));
  OK(cudaMemset(src, 1, pool));
  OK(cudaDeviceSynchronize());
  OK(cudaSetDevice(0));
  int can = 0;
  OK(cudaDeviceCanAccessPeer(&can, 0, 1));
  printf("peer access 0<-1: %d\n", can);
  if (can) OK(cudaDeviceEnablePeerAccess(1, 0));
  OK(cudaMalloc(&dst, pool));
  double *scratch;
  OK(cudaMalloc(&scratch, 148 * 256 * 8 * 4));
  const int NS = 8;
  cudaStream_t st[NS], ks;
  for (auto &s : st) OK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
  OK(cudaStreamCreateWithFlags(&ks, cudaStreamNonBlocking));
  cudaEvent_t e0, e1, done[NS];
  OK(cudaEventCreate(&e0));
  OK(cudaEventCreate(&e1));
  for (auto &d : done) OK(cudaEventCreateWithFlags(&d, cudaEventDisableTiming));
  const size_t sizes[] = {143360, 1600000, 5734400, 80000000};
  for (int load = 0; load < 2; load++)
    for (size_t bytes : sizes)
      for (int ns : {1, 2, 4, 8}) {
        const int count = (int)std::min<size_t>(256, (pool / 2) / bytes);
        if (load) busy<<<148 * 2, 256, 0, ks>>>(scratch, 3000000);  // ~100+ ms of DMMA on every SM
        OK(cudaEventRecord(e0, st[0]));
        for (int s = 1; s < ns; s++) OK(cudaStreamWaitEvent(st[s], e0, 0));
        for (int i = 0; i < count; i++)
          OK(cudaMemcpyAsync(dst + (size_t)i * bytes, src + (size_t)((i * 7) % count) * bytes, bytes,
                             cudaMemcpyDeviceToDevice, st[i % ns]));
        for (int s = 1; s < ns; s++) {
          OK(cudaEventRecord(done[s], st[s]));
          OK(cudaStreamWaitEvent(st[0], done[s], 0));
        }
        OK(cudaEventRecord(e1, st[0]));
        OK(cudaEventSynchronize(e1));
        float ms = 0;
        OK(cudaEventElapsedTime(&ms, e0, e1));
        OK(cudaDeviceSynchronize());
        printf("%s bytes %9zu x %3d on %d streams: %8.3f ms  %7.1f GB/s  %7.1f us/copy\n", load ? "busy" : "idle", bytes,
               count, ns, ms, bytes * (double)count / ms / 1e6, ms * 1e3 / count);
      }
  return 0;
}
