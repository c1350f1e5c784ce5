// Micro-benchmark of the Cholesky k-loop WITH its operand ring (TMA -> 3-slot smem ring -> mbarrier
// hand-off -> 10 LDS.64 + 16 DMMA per k-step), on an L2-resident source, no other phases.  Variants
// isolate what the hand-off costs relative to the smem-only loop (scripts/micro/dmma_micro.cu: 98.5 %).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I starry_process_b200/csrc -o scripts/micro/ring_micro scripts/micro/ring_micro.cu -lcuda
#include <cstdio>
#include <cstring>
#include <cuda.h>
#include <cuda_runtime.h>
#include "common.cuh"

void spb_set_error(const std::string &) {}

constexpr int NB = 64, KC = 16;

template <int TM, int STAGES>
struct Sm {
  double As[STAGES][TM][KC];
  double Bs[STAGES][NB][KC];
  uint64_t full[STAGES], empty[STAGES];
};

__device__ __forceinline__ int swz(int row, int k) { return (((k >> 1) ^ (row & 7)) << 1) | (k & 1); }
__device__ __forceinline__ double negate(double x) {
  return __hiloint2double(__double2hiint(x) ^ (int)0x80000000, __double2loint(x));
}

// MODE 0: as the product kernel (issue ch+LA at the top of chunk ch, non-blocking then blocking)
// MODE 1: dedicated producer warp (extra warp: NW compute warps + 1 producer)
// MODE 2: as 0, plus fragments of the next chunk's first k-step loaded before the chunk boundary
template <int TM, int STAGES, int MODE>
__global__ void __launch_bounds__(TM * 2 + (MODE == 1 ? 32 : 0))
    ring_kernel(const __grid_constant__ CUtensorMap tm, int nchunks_per_tile, int ntiles, int nrows_mat, double *sink, int nmat) {
  extern __shared__ __align__(1024) unsigned char raw[];
  using S = Sm<TM, STAGES>;
  S &sm = *reinterpret_cast<S *>(raw);
  constexpr int NCW = TM / 16;             // compute warps
  constexpr int NTH = NCW * 32;
  constexpr int LA = STAGES - 1;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, tg = lane & 3;
  const int item = blockIdx.x % nmat;       // nmat = 8 source matrices of 1024 x 1024: L2 resident
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&sm.full[s], (MODE == 1 || MODE == 3) ? 1 : NTH);
      mbar_init(&sm.empty[s], NCW);
    }
  }
  __syncthreads();
  double acc[2][8][2];
#pragma unroll
  for (int m = 0; m < 2; ++m)
#pragma unroll
    for (int n = 0; n < 8; ++n) acc[m][n][0] = acc[m][n][1] = 0.0;
  int koff[4];
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) koff[kk] = swz(g, kk * 4 + tg);
  const int total = nchunks_per_tile * ntiles;

  auto tma_issue = [&](unsigned x) {
    const unsigned st = x % STAGES;
    const int tile = x / nchunks_per_tile, ch = x % nchunks_per_tile;
    const int row0 = (tile * TM) % (nrows_mat - TM);
    mbar_arrive_expect_tx(&sm.full[st], (TM + NB) * KC * 8u);
#pragma unroll
    for (int bx = 0; bx < TM / 64; ++bx)
      tma_load_3d(&sm.As[st][64 * bx][0], &tm, &sm.full[st], ch * KC, row0 + 64 * bx, item);
    tma_load_3d(&sm.Bs[st][0][0], &tm, &sm.full[st], ch * KC, (row0 + 192) % (nrows_mat - 64), item);
  };

  if (MODE == 1) {
    if (warp == NCW) {   // producer warp
      if (lane == 0) {
        for (unsigned x = 0; x < (unsigned)total; ++x) {
          const unsigned st = x % STAGES;
          if (x >= STAGES) mbar_wait(&sm.empty[st], ((x / STAGES) + 1u) & 1u);
          tma_issue(x);
        }
      }
      return;
    }
    for (unsigned x = 0; x < (unsigned)total; ++x) {
      const unsigned st = x % STAGES;
      mbar_wait(&sm.full[st], (x / STAGES) & 1u);
      const double *Aw = &sm.As[st][warp * 16 + g][0];
      const double *Bw = &sm.Bs[st][g][0];
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        double a[2], b[8];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) a[mt] = negate(Aw[mt * 8 * KC + koff[kk]]);
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) b[nt] = Bw[nt * 8 * KC + koff[kk]];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
          for (int nt = 0; nt < 8; ++nt) dmma_m8n8k4(acc[mt][nt][0], acc[mt][nt][1], a[mt], b[nt]);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm.empty[st]);
    }
  } else {
    auto issue = [&](unsigned x, bool blocking) -> bool {
      const unsigned st = x % STAGES;
      if (MODE == 3) {   // only thread 0 touches the barriers on the issue side
        if (tid != 0) return true;
        if (blocking) mbar_wait(&sm.empty[st], ((x / STAGES) + 1u) & 1u);
        else if (!mbar_test(&sm.empty[st], ((x / STAGES) + 1u) & 1u)) return false;
        tma_issue(x);
        return true;
      }
      if (blocking) mbar_wait(&sm.empty[st], ((x / STAGES) + 1u) & 1u);
      else if (!mbar_test(&sm.empty[st], ((x / STAGES) + 1u) & 1u)) return false;
      if (tid == 0) tma_issue(x);
      else mbar_arrive(&sm.full[st]);
      return true;
    };
    for (int j = 0; j < LA; ++j) issue(j, true);
    double an[2], bn[8];
    bool have_next = false;
    for (unsigned x = 0; x < (unsigned)total; ++x) {
      const unsigned st = x % STAGES;
      bool early = false;
      if (x + LA < (unsigned)total) early = issue(x + LA, false);
      if (!(MODE == 2 && have_next)) mbar_wait(&sm.full[st], (x / STAGES) & 1u);
      const double *Aw = &sm.As[st][warp * 16 + g][0];
      const double *Bw = &sm.Bs[st][g][0];
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        double a[2], b[8];
        if (MODE == 2 && kk == 0 && have_next) {
#pragma unroll
          for (int mt = 0; mt < 2; ++mt) a[mt] = an[mt];
#pragma unroll
          for (int nt = 0; nt < 8; ++nt) b[nt] = bn[nt];
        } else {
#pragma unroll
          for (int mt = 0; mt < 2; ++mt) a[mt] = negate(Aw[mt * 8 * KC + koff[kk]]);
#pragma unroll
          for (int nt = 0; nt < 8; ++nt) b[nt] = Bw[nt * 8 * KC + koff[kk]];
        }
        if (MODE == 2 && kk == 3) {
          // next chunk's first fragments, if it has already landed
          have_next = false;
          if (x + 1 < (unsigned)total) {
            const unsigned st2 = (x + 1) % STAGES;
            if (mbar_test(&sm.full[st2], ((x + 1) / STAGES) & 1u)) {
              have_next = true;
              const double *A2 = &sm.As[st2][warp * 16 + g][0];
              const double *B2 = &sm.Bs[st2][g][0];
#pragma unroll
              for (int mt = 0; mt < 2; ++mt) an[mt] = negate(A2[mt * 8 * KC + koff[0]]);
#pragma unroll
              for (int nt = 0; nt < 8; ++nt) bn[nt] = B2[nt * 8 * KC + koff[0]];
            }
          }
        }
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
          for (int nt = 0; nt < 8; ++nt) dmma_m8n8k4(acc[mt][nt][0], acc[mt][nt][1], a[mt], b[nt]);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm.empty[st]);
      if (!early && x + LA < (unsigned)total) issue(x + LA, true);
    }
  }
  double s = 0.0;
#pragma unroll
  for (int m = 0; m < 2; ++m)
#pragma unroll
    for (int n = 0; n < 8; ++n) s += acc[m][n][0] + acc[m][n][1];
  if (s == 123.456) sink[0] = s;
}

typedef CUresult (*encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                              const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                              CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int TM, int STAGES, int MODE>
void go(const CUtensorMap &tm, int nsm, int cps, double *sink, const char *name, int nmat = 8) {
  auto k = ring_kernel<TM, STAGES, MODE>;
  const int smem = sizeof(Sm<TM, STAGES>) + 1024;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  const int nchunks = 32, ntiles = 40;   // 32 chunks of 16 columns per tile (c0 = 512), 40 tiles
  const int nth = TM * 2 + (MODE == 1 ? 32 : 0);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<<<nsm * cps, nth, smem>>>(tm, nchunks, ntiles, 1024, sink, nmat);
  cudaEventRecord(e0);
  k<<<nsm * cps, nth, smem>>>(tm, nchunks, ntiles, 1024, sink, nmat);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  cudaError_t err = cudaGetLastError();
  double fl = (double)nsm * cps * (double)nchunks * ntiles * TM * NB * KC * 2.0;
  printf("  %-44s %d CTA/SM: %7.3f ms %6.2f TF/s %s\n", name, cps, ms, fl / ms / 1e9, err == cudaSuccess ? "" : cudaGetErrorString(err));
}

int main() {
  cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
  const int nsm = prop.multiProcessorCount;
  double *sink; cudaMalloc(&sink, 8);
  double *src; const size_t n = 1024, B = 600;
  cudaMalloc(&src, B * n * n * 8); cudaMemset(src, 0, B * n * n * 8);
  void *ptr = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q);
  encode_fn enc = (encode_fn)ptr;
  CUtensorMap tm; memset(&tm, 0, sizeof(tm));
  cuuint64_t dims[3] = {n, n, B}; cuuint64_t strides[2] = {n * 8, n * n * 8};
  cuuint32_t box[3] = {KC, 64, 1}; cuuint32_t es[3] = {1, 1, 1};
  CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, src, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("%s, %d SMs, tensor map %d\n", prop.name, nsm, (int)r);
  printf("== k-loop + TMA ring from an L2-resident source (64 MB)\n");
  go<128, 3, 0>(tm, nsm, 1, sink, "TM128 S3 product-style");
  go<128, 3, 0>(tm, nsm, 2, sink, "TM128 S3 product-style");
  go<128, 3, 2>(tm, nsm, 1, sink, "TM128 S3 + next-chunk fragment prefetch");
  go<128, 3, 2>(tm, nsm, 2, sink, "TM128 S3 + next-chunk fragment prefetch");
  go<128, 3, 1>(tm, nsm, 1, sink, "TM128 S3 producer warp");
  go<128, 3, 1>(tm, nsm, 2, sink, "TM128 S3 producer warp");
  go<128, 3, 3>(tm, nsm, 1, sink, "TM128 S3 thread-0-only issue");
  go<128, 3, 3>(tm, nsm, 2, sink, "TM128 S3 thread-0-only issue");
  go<64, 3, 3>(tm, nsm, 3, sink, "TM64 S3 thread-0-only issue");
  go<128, 6, 0>(tm, nsm, 1, sink, "TM128 S6 product-style");
  go<128, 6, 1>(tm, nsm, 1, sink, "TM128 S6 producer warp");
  go<64, 3, 0>(tm, nsm, 3, sink, "TM64 S3 product-style");
  go<64, 3, 2>(tm, nsm, 3, sink, "TM64 S3 + next-chunk fragment prefetch");
  go<64, 3, 1>(tm, nsm, 3, sink, "TM64 S3 producer warp");
  go<64, 4, 1>(tm, nsm, 2, sink, "TM64 S4 producer warp");
  printf("== same, every CTA streams its own 8 MB matrix (600 matrices, 4.9 GB: DRAM)\n");
  go<128, 3, 0>(tm, nsm, 1, sink, "TM128 S3 product-style", 600);
  go<128, 3, 0>(tm, nsm, 2, sink, "TM128 S3 product-style", 600);
  go<128, 3, 1>(tm, nsm, 1, sink, "TM128 S3 producer warp", 600);
  go<128, 3, 1>(tm, nsm, 2, sink, "TM128 S3 producer warp", 600);
  go<128, 6, 1>(tm, nsm, 1, sink, "TM128 S6 producer warp", 600);
  go<64, 3, 0>(tm, nsm, 3, sink, "TM64 S3 product-style", 600);
  go<64, 3, 1>(tm, nsm, 3, sink, "TM64 S3 producer warp", 600);
  go<64, 4, 1>(tm, nsm, 2, sink, "TM64 S4 producer warp", 600);
  return 0;
}
