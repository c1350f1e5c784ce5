// Cholesky factor of the Ylm covariance and prior draws (sp.py:265-271, 489-509).
//   L_ylm = cho_factor(cov_ylm)            -> the batched DMMA Cholesky kernel with nt = 256
//   y     = (mean_ylm[:, None] + L u)^T    -> one tensor-core NT GEMM  y[s][i] = sum_k u[s][k] L[i][k]
#include "gemm_nt.cuh"

namespace {
__global__ void zero_upper_kernel(int B, double *L) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)B * 65536) return;
  const int ij = (int)(idx & 65535);
  if ((ij & 255) > (ij >> 8)) L[idx] = 0.0;
}
// generalised: zero the strict upper triangle of B (n x n) matrices with leading dimension ld
__global__ void tril_kernel(int B, int n, int ld, long long stride, double *L) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t per = (size_t)n * n;
  if (idx >= (size_t)B * per) return;
  const size_t b = idx / per;
  const int ij = (int)(idx - b * per);
  const int i = ij / n, j = ij - i * n;
  if (j > i) L[b * (size_t)stride + (size_t)i * ld + j] = 0.0;
}
}  // namespace

// C = alpha A Bm^T + beta C on the FP64 tensor pipe (the contractions of predict / conditional
// sampling, sp.py:767-1002: K_ts_t = (A_ts Sigma) A_t^T, mu = mean + V w, K = K_ts_ts - V V^T, ...)
extern "C" int spb_gemm_nt(spb_context *ctx, int batch, int M, int N, int K, double alpha,
                           const double *A, int lda, long long strideA, const double *Bm, int ldb,
                           long long strideB, double beta, double *C, int ldc, long long strideC,
                           void *stream) {
  SPB_REQUIRE(ctx != nullptr && A != nullptr && Bm != nullptr && C != nullptr, "gemm_nt: null argument");
  SPB_CHECK_CUDA(cudaSetDevice(ctx->device));
  gnt::Desc d = {};
  d.A = A;
  d.strideA = strideA;
  d.lda = lda;
  d.Bm = Bm;
  d.strideB = strideB;
  d.ldb = ldb;
  d.C = C;
  d.strideC = strideC;
  d.ldc = ldc;
  d.M = M;
  d.N = N;
  d.K = K;
  d.batch = batch;
  d.ksplit = 1;
  d.alpha = alpha;
  d.beta = beta;
  return gnt::launch<gnt::EPI_AXPBY>(ctx, d, (cudaStream_t)stream);
}

extern "C" int spb_tril(spb_context *ctx, int B, int n, double *L, int ld, long long stride,
                        void *stream) {
  SPB_REQUIRE(ctx != nullptr && L != nullptr && B > 0 && n > 0 && ld >= n, "tril: bad arguments");
  SPB_CHECK_CUDA(cudaSetDevice(ctx->device));
  const size_t total = (size_t)B * n * n;
  tril_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(B, n, ld, stride, L);
  SPB_LAUNCH_CHECK(ctx);
  return 0;
}

extern "C" int spb_cho_cov_ylm(spb_context *ctx, int B, const double *cov_ylm, double *L_ylm,
                               int32_t *info, void *stream_) {
  SPB_REQUIRE(ctx != nullptr && B > 0, "cho_cov_ylm: bad arguments");
  cudaStream_t stream = (cudaStream_t)stream_;
  SPB_CHECK_CUDA(cudaSetDevice(ctx->device));
  if (L_ylm != cov_ylm)
    SPB_CHECK_CUDA(cudaMemcpyAsync(L_ylm, cov_ylm, (size_t)B * 65536 * sizeof(double),
                                   cudaMemcpyDeviceToDevice, stream));
  int st = spb_cholesky_lnlike(ctx, B, 256, L_ylm, 256, 65536, 0, nullptr, 256, 0, nullptr, nullptr,
                               nullptr, info, stream_);
  if (st) return st;
  zero_upper_kernel<<<(unsigned)(((size_t)B * 65536 + 255) / 256), 256, 0, stream>>>(B, L_ylm);
  SPB_LAUNCH_CHECK(ctx);
  return 0;
}

extern "C" int spb_sample_ylm(spb_context *ctx, int B, int nsamples, const double *mean_ylm,
                              const double *L_ylm, const double *unit_normals, double *y,
                              void *stream) {
  SPB_REQUIRE(ctx != nullptr && B > 0 && nsamples > 0, "sample_ylm: bad arguments");
  SPB_CHECK_CUDA(cudaSetDevice(ctx->device));
  gnt::Desc d = {};
  d.A = unit_normals;
  d.strideA = (long long)nsamples * 256;
  d.lda = 256;
  d.Bm = L_ylm;
  d.strideB = 65536;
  d.ldb = 256;
  d.C = y;
  d.strideC = (long long)nsamples * 256;
  d.ldc = 256;
  d.M = nsamples;
  d.N = 256;
  d.K = 256;
  d.batch = B;
  d.ksplit = 1;
  d.vec = mean_ylm;
  d.strideVec = 256;
  d.alpha = 1.0;
  return gnt::launch<gnt::EPI_ADD_ROWVEC>(ctx, d, (cudaStream_t)stream);
}
