// Micro-benchmarks of the FP64 tensor pipe on B200 (sm_100a), to locate the practical ceiling of the
// Cholesky k-loop:  (1) register-resident DMMA rate vs resident warps per SM sub-partition;
// (2) the k-loop's instruction mix (10 LDS.64 + 16 DMMA per k-step, 16 rows x 64 columns per warp)
// fed from shared memory that is never refilled (no global traffic, no barriers).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/micro/dmma_micro scripts/micro/dmma_micro.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double &d0, double &d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

template <int NACC>
__global__ void reg_kernel(int iters, double *sink) {
  double acc[NACC][2];
#pragma unroll
  for (int i = 0; i < NACC; ++i) acc[i][0] = acc[i][1] = 0.0;
  double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) dmma(acc[i][0], acc[i][1], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += acc[i][0] + acc[i][1];
  if (s == 123.456) sink[0] = s;
}

// k-loop mix: TMR rows of A (16 per warp) and 64 rows of B, KC = 16 columns, 128-byte swizzle
template <int NWARPS>
__global__ void __launch_bounds__(NWARPS * 32) smem_kernel(int chunks, double *sink) {
  extern __shared__ double sm[];
  double *As = sm;                       // [NWARPS*16][16]
  double *Bs = sm + NWARPS * 16 * 16;    // [64][16]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, tg = lane & 3;
  for (int i = tid; i < (NWARPS * 16 + 64) * 16; i += NWARPS * 32) sm[i] = 1.0 + 1e-9 * i;
  __syncthreads();
  double acc[2][8][2];
#pragma unroll
  for (int m = 0; m < 2; ++m)
#pragma unroll
    for (int n = 0; n < 8; ++n) acc[m][n][0] = acc[m][n][1] = 0.0;
  int koff[4];
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    const int k = kk * 4 + tg;
    koff[kk] = (((k >> 1) ^ (g & 7)) << 1) | (k & 1);
  }
  const double *Aw = As + (warp * 16 + g) * 16;
  const double *Bw = Bs + g * 16;
  for (int ch = 0; ch < chunks; ++ch) {
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      double a[2], b[8];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) a[mt] = -Aw[mt * 8 * 16 + koff[kk]];
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) b[nt] = Bw[nt * 8 * 16 + koff[kk]];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) dmma(acc[mt][nt][0], acc[mt][nt][1], a[mt], b[nt]);
    }
    asm volatile("" ::: "memory");
  }
  double s = 0.0;
#pragma unroll
  for (int m = 0; m < 2; ++m)
#pragma unroll
    for (int n = 0; n < 8; ++n) s += acc[m][n][0] + acc[m][n][1];
  if (s == 123.456) sink[0] = s;
}

template <class F>
float time_ms(F f) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  f();
  cudaEventRecord(e0);
  f();
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  return ms;
}

int main() {
  cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
  const int nsm = prop.multiProcessorCount;
  double *sink; cudaMalloc(&sink, 8);
  printf("%s, %d SMs\n", prop.name, nsm);
  printf("== register-resident DMMA: warps per CTA x CTAs per SM -> TFLOP/s\n");
  const int iters = 20000;
  for (int wpc : {4, 8, 12, 16})
    for (int cps : {1, 2}) {
      if (wpc * cps > 32) continue;
      float ms = time_ms([&] { reg_kernel<16><<<nsm * cps, wpc * 32>>>(iters, sink); });
      double fl = (double)nsm * cps * wpc * 16.0 * iters * 512.0;
      printf("  %2d warps/CTA x %d CTA/SM (%4.1f warps/SMSP), 16 acc: %6.2f TF/s\n", wpc, cps, wpc * cps / 4.0, fl / ms / 1e9);
    }
  for (int wpc : {4, 8}) {
    float ms = time_ms([&] { reg_kernel<4><<<nsm, wpc * 32>>>(iters, sink); });
    double fl = (double)nsm * wpc * 4.0 * iters * 512.0;
    printf("  %2d warps/CTA x 1 CTA/SM, 4 acc (dependent every 4): %6.2f TF/s\n", wpc, fl / ms / 1e9);
  }
  printf("== k-loop mix from shared memory (10 LDS.64 + 16 DMMA per k-step)\n");
  const int chunks = 20000;
  cudaFuncSetAttribute(smem_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  cudaFuncSetAttribute(smem_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  for (int cps : {1, 2}) {
    float ms = time_ms([&] { smem_kernel<8><<<nsm * cps, 256, (8 * 16 + 64) * 16 * 8>>>(chunks, sink); });
    double fl = (double)nsm * cps * 8 * 64.0 * chunks * 512.0;
    printf("  8 warps x %d CTA/SM: %6.2f TF/s\n", cps, fl / ms / 1e9);
  }
  for (int cps : {1, 2, 3, 4}) {
    float ms = time_ms([&] { smem_kernel<4><<<nsm * cps, 128, (4 * 16 + 64) * 16 * 8>>>(chunks, sink); });
    double fl = (double)nsm * cps * 4 * 64.0 * chunks * 512.0;
    printf("  4 warps x %d CTA/SM: %6.2f TF/s\n", cps, fl / ms / 1e9);
  }
  return 0;
}
