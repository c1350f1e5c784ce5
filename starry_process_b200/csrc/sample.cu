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
}  // namespace

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
