// Microbenchmark: FP64 ceilings of this GPU (no reference counterpart; BASELINE.json asks for
// "fraction of FP64 tensor peak", MEASURED_PEAKS.json has no FP64 figure).
//   dmma  : register-resident mma.sync.m8n8k4.f64 loop   (SASS DMMA.8x8x4)
//   dmma16: register-resident mma.sync.m16n8k16.f64 loop
//   dfma  : register-resident fma.rn.f64 loop
//   lds   : DMMA fed by LDS.64 fragments from a conflict-free smem tile (no global traffic)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak fp64_peak.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

template <int NACC>
__global__ void k_dmma(double *out, int iters) {
  double acc[NACC][2];
  for (int i = 0; i < NACC; i++) acc[i][0] = acc[i][1] = 0.0;
  double a = threadIdx.x * 1e-3, b = threadIdx.x * 2e-3;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++) dmma884(acc[i][0], acc[i][1], a, b);
  }
  double s = 0;
  for (int i = 0; i < NACC; i++) s += acc[i][0] + acc[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_dmma16(double *out, int iters) {
  double acc[8][4];
  for (int i = 0; i < 8; i++) for (int j = 0; j < 4; j++) acc[i][j] = 0.0;
  double a[8], b[4];
  for (int i = 0; i < 8; i++) a[i] = threadIdx.x * 1e-3 + i;
  for (int i = 0; i < 4; i++) b[i] = threadIdx.x * 2e-3 + i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
                   : "+d"(acc[i][0]), "+d"(acc[i][1]), "+d"(acc[i][2]), "+d"(acc[i][3])
                   : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                     "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
  }
  double s = 0;
  for (int i = 0; i < 8; i++) for (int j = 0; j < 4; j++) s += acc[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void k_dfma(double *out, int iters) {
  double acc[NACC];
  for (int i = 0; i < NACC; i++) acc[i] = i;
  double a = 1.0 + threadIdx.x * 1e-9, b = threadIdx.x * 2e-9;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++) acc[i] = fma(acc[i], a, b);
  }
  double s = 0;
  for (int i = 0; i < NACC; i++) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// warp tile (8*MI) x (8*NI), K chunk of 16 doubles per row in smem (128 B rows, XOR-swizzled
// 16-B chunks like TMA SWIZZLE_128B); fragments via LDS.64.
template <int MI, int NI>
__global__ void k_lds(double *out, int iters) {
  extern __shared__ __align__(1024) unsigned char smem[];
  double *sm = (double *)smem;
  const int nrows = 8 * (MI + NI) * (blockDim.x / 32);
  for (int i = threadIdx.x; i < nrows * 16; i += blockDim.x) sm[i] = (i % 7) * 1e-3;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int rowin8 = 2 * (g & 3) + (g >> 2);
  double acc[MI][NI][2];
  for (int i = 0; i < MI; i++) for (int j = 0; j < NI; j++) acc[i][j][0] = acc[i][j][1] = 0;
  const unsigned char *wbase = smem + (size_t)warp * 8 * (MI + NI) * 128;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int s = 0; s < 4; s++) {
      double a[MI], b[NI];
      const int chunk = ((2 * s + (t >> 1)) ^ rowin8) * 16 + 8 * (t & 1);
#pragma unroll
      for (int i = 0; i < MI; i++) a[i] = *(const volatile double *)(wbase + (i * 8 + rowin8) * 128 + chunk);
#pragma unroll
      for (int j = 0; j < NI; j++) b[j] = *(const volatile double *)(wbase + ((MI + j) * 8 + rowin8) * 128 + chunk);
#pragma unroll
      for (int i = 0; i < MI; i++)
#pragma unroll
        for (int j = 0; j < NI; j++) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
  }
  double s = 0;
  for (int i = 0; i < MI; i++) for (int j = 0; j < NI; j++) s += acc[i][j][0] + acc[i][j][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
static double timeit(F launch, int reps = 5) {
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  launch(); launch(); CK(cudaDeviceSynchronize());
  double best = 1e30;
  for (int r = 0; r < reps; r++) {
    CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
  }
  CK(cudaGetLastError());
  return best * 1e-3;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int nsm = p.multiProcessorCount;
  printf("device %s sms %d\n", p.name, nsm);
  double *out; CK(cudaMalloc(&out, sizeof(double) * nsm * 16 * 1024));
  const int iters = 20000;
  for (int warps : {4, 8, 16}) {
    for (int bps : {1, 2}) {
      int grid = nsm * bps, thr = warps * 32;
      double s = timeit([&] { k_dmma<8><<<grid, thr>>>(out, iters); });
      double fl = (double)grid * warps * iters * 8 * 512.0;
      printf("dmma884  acc8  warps/cta %2d cta/sm %d : %7.2f TFLOP/s\n", warps, bps, fl / s / 1e12);
      s = timeit([&] { k_dmma<16><<<grid, thr>>>(out, iters); });
      fl = (double)grid * warps * iters * 16 * 512.0;
      printf("dmma884  acc16 warps/cta %2d cta/sm %d : %7.2f TFLOP/s\n", warps, bps, fl / s / 1e12);
      s = timeit([&] { k_dmma16<<<grid, thr>>>(out, iters / 4); });
      fl = (double)grid * warps * (iters / 4) * 8 * (2.0 * 16 * 8 * 16);
      printf("dmma16816      warps/cta %2d cta/sm %d : %7.2f TFLOP/s\n", warps, bps, fl / s / 1e12);
      s = timeit([&] { k_dfma<16><<<grid, thr>>>(out, iters); });
      fl = (double)grid * thr * iters * 16 * 2.0;
      printf("dfma     acc16 warps/cta %2d cta/sm %d : %7.2f TFLOP/s\n", warps, bps, fl / s / 1e12);
    }
  }
  CK(cudaFuncSetAttribute(k_lds<4, 5>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  CK(cudaFuncSetAttribute(k_lds<4, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  CK(cudaFuncSetAttribute(k_lds<2, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  CK(cudaFuncSetAttribute(k_lds<2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  for (int warps : {4, 8, 16}) {
    int thr = warps * 32;
    double s = timeit([&] { k_lds<4, 5><<<nsm, thr, warps * 72 * 128>>>(out, iters / 8); });
    printf("lds-fed 4x5 warps %2d : %7.2f TFLOP/s\n", warps, (double)nsm * warps * (iters / 8) * 4 * 20 * 512.0 / s / 1e12);
    s = timeit([&] { k_lds<4, 4><<<nsm, thr, warps * 64 * 128>>>(out, iters / 8); });
    printf("lds-fed 4x4 warps %2d : %7.2f TFLOP/s\n", warps, (double)nsm * warps * (iters / 8) * 4 * 16 * 512.0 / s / 1e12);
    s = timeit([&] { k_lds<2, 4><<<nsm, thr, warps * 48 * 128>>>(out, iters / 8); });
    printf("lds-fed 2x4 warps %2d : %7.2f TFLOP/s\n", warps, (double)nsm * warps * (iters / 8) * 4 * 8 * 512.0 / s / 1e12);
    s = timeit([&] { k_lds<2, 2><<<nsm, thr, warps * 32 * 128>>>(out, iters / 8); });
    printf("lds-fed 2x2 warps %2d : %7.2f TFLOP/s\n", warps, (double)nsm * warps * (iters / 8) * 4 * 4 * 512.0 / s / 1e12);
  }
  // sustained: 3 s of the best DMMA config, report clocks via nvidia-smi separately
  {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0));
    int n = 0; double fl = 0;
    for (; n < 60; n++) { k_dmma<16><<<nsm * 2, 256>>>(out, iters * 4); fl += (double)nsm * 2 * 8 * iters * 4 * 16 * 512.0; }
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("dmma884 sustained %.2f s : %7.2f TFLOP/s\n", ms * 1e-3, fl / (ms * 1e-3) / 1e12);
  }
  return 0;
}
