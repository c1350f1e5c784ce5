// Flux-space GP moments.
//
//   spb_flux_operator     rTA1 / rTA1L(u)      ops/flux/rTA1.cc:7-30, rTA1L.cc:22-60, flux.h:302-523
//   spb_flux_marginal     inclination-marginalised mean / variance / kernel on the covpts grid
//                         flux.py:55-62, 181-231, 295-333; ops/wigner/special_tensordotRz.cc:10-65
//   spb_flux_conditional  K = A cov_ylm A^T, mean = (A mean_ylm)[0]      flux.py:335-343
//
// B200-first re-organisation of the marginal kernel (DESIGN.md "marginal kernel").  The reference
// rotates Sigma + mu mu^T into the polar frame (two block-diagonal 256x256 products per sample)
// and then evaluates, for each of the 304 grid lags, two (256 x 256) contractions.  Because the
// z-rotation only multiplies row i by cos/sin(m_i theta), the lag dependence collapses to
//      f(theta_k) = sum_m a_m cos(m theta_k),   a_m = sum_{i: |m_i| = m} sum_j W_ij Ez_ij
// plus the sine lane b_m sin(m theta_k), which vanishes for an exactly symmetric Ez but is carried
// at its rounding-level value (~1e-9 of a_0) so that the kernel equals the reference's.  With Ez = Rx^T (Sigma + mu mu^T) Rx the Sigma part of a_m is a fixed
// quadratic form <Omega_m(u), Sigma>; the 16 + 15 forms are pre-folded on the host (tables) and
// evaluated for the whole batch as ONE tensor-core GEMM (B x 65536) . (65536 x 16) that streams
// each cov_ylm exactly once from HBM; the rank-one mu mu^T part is evaluated directly from
// ez = mu . Rx(pi/2).
#include "gemm_nt.cuh"
#include "spb_tables.h"

namespace {

__host__ __device__ __forceinline__ int nwig(int l) { return ((l + 1) * (2 * l + 1) * (2 * l + 3)) / 3; }

__device__ __forceinline__ void lm_of(int n, int &l, int &m) {
  l = (int)floor(sqrt((double)n) + 1e-9);
  m = n - l * l - l;
}

// rTA1L[l(l+1)] = 2 sqrt(2l+1) int_0^1 P_l(x) x I(x) dx / (1 - u1/3 - u2/6),
// I(x) = 1 - u1 (1-x) - u2 (1-x)^2; all m != 0 entries vanish by azimuthal symmetry.
// This is the closed form of flux.h:501-523 (the reference reaches the same numbers through a
// sparse change of basis with ~1e-12 cancellation noise); 32-point Gauss-Legendre is exact here.
__global__ void flux_operator_kernel(int nu, const double *u, const double *tab, double *out) {
  const int iu = blockIdx.x, n = threadIdx.x;
  if (iu >= nu) return;
  int l, m;
  lm_of(n, l, m);
  double val = 0.0;
  if (m == 0) {
    const double u1 = u ? u[2 * iu] : 0.0, u2 = u ? u[2 * iu + 1] : 0.0;
    const double norm = 1.0 - u1 / 3.0 - u2 / 6.0;
    double acc = 0.0;
    for (int gi = 0; gi < 32; ++gi) {
      const double x = tab[SPB_TAB_GL_X + gi], wq = tab[SPB_TAB_GL_W + gi];
      double p0 = 1.0, p1 = x;
      double pl = (l == 0) ? p0 : p1;
      for (int k = 2; k <= l; ++k) {
        pl = ((2 * k - 1) * x * p1 - (k - 1) * p0) / k;
        p0 = p1;
        p1 = pl;
      }
      const double om = 1.0 - x;
      acc += wq * pl * x * (1.0 - u1 * om - u2 * om * om);
    }
    val = 2.0 * sqrt(2.0 * l + 1.0) * acc / norm;
  }
  out[(size_t)iu * 256 + n] = val;
}

// ---- marginal, stage 1: per sample  ez, flux mean, and the mu-part of a_m -------------------
struct MargParams {
  int B;
  const double *mean_ylm;  // (B,256)
  const double *rTA1;      // (256)
  const double *tab;
  double *gp_mean;         // (B)
  double *amu;             // (B,32): a_0..a_15, b_1..b_15 (mu mu^T part)
};

__global__ void __launch_bounds__(256) marginal_mu_kernel(MargParams p) {
  __shared__ double mu[256], ez[256], zez[256], arow[256], red[16];
  const int b = blockIdx.x, n = threadIdx.x, warp = n >> 5, lane = n & 31;
  int l, m;
  lm_of(n, l, m);
  const int w = 2 * l + 1, j = m + l;
  mu[n] = p.mean_ylm[(size_t)b * 256 + n];
  if (n < 16) red[n] = 0.0;
  __syncthreads();
  // ez = mu . Rx(pi/2)  (flux.py:55-57)
  const double zl = p.rTA1[l * l + l];
  {
    const double *rx = p.tab + SPB_TAB_RX90 + nwig(l - 1) + j;
    double acc = 0.0;
    for (int mp = 0; mp < w; ++mp) acc = fma(mu[l * l + mp], rx[mp * w], acc);
    ez[n] = acc;
    zez[n] = zl * acc;
  }
  __syncthreads();
  // flux mean: sum_l w_l . ez_l with w_l = rTA1_l . wnp[l]  (flux.py:188-191, 298-303); only the
  // m = 0 entry of rTA1_l is non-zero
  {
    const double wn = zl * p.tab[SPB_TAB_FLUX_WNP + nwig(l - 1) + l * w + j];
    double v = warp_sum(wn * ez[n]);
    if (lane == 0) arow[warp] = v;
  }
  __syncthreads();
  if (n == 0) {
    double s = 0.0;
    for (int k = 0; k < 8; ++k) s += arow[k];
    p.gp_mean[b] = s;
  }
  __syncthreads();
  // A_i = ez_i z_{l_i} sum_j Wnp_ij z_{l_j} ez_j   (W = Wnp o Z blocks, flux.py:194-209)
  for (int i = warp; i < 256; i += 8) {
    const double *wr = p.tab + SPB_TAB_FLUX_W + (size_t)i * 256;
    double acc = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) acc = fma(wr[lane + 32 * k], zez[lane + 32 * k], acc);
    acc = warp_sum(acc);
    if (lane == 0) arow[i] = acc;
  }
  __syncthreads();
  ez[n] = zez[n] * arow[n];  // A_i (ez is no longer needed)
  mu[n] = zez[l * l + l - m];  // zez at the (l, -m) partner (mu is no longer needed)
  __syncthreads();
  if (n < 16) {
    // a_m (mu part) = sum over rows with |m_i| = m, fixed order => deterministic
    double s = 0.0;
    for (int ll = n; ll <= SPB_LMAX; ++ll) {
      s += ez[ll * ll + ll + n];
      if (n > 0) s += ez[ll * ll + ll - n];
    }
    p.amu[(size_t)b * 32 + n] = s;
  }
  __syncthreads();
  // sine lane: B_i = zez_i sum_j Wnp_ij zez_jbar
  for (int i = warp; i < 256; i += 8) {
    const double *wr = p.tab + SPB_TAB_FLUX_W + (size_t)i * 256;
    double acc = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) acc = fma(wr[lane + 32 * k], mu[lane + 32 * k], acc);
    acc = warp_sum(acc);
    if (lane == 0) arow[i] = acc;
  }
  __syncthreads();
  ez[n] = zez[n] * arow[n];
  __syncthreads();
  if (n >= 1 && n < 16) {
    double s = 0.0;
    for (int ll = n; ll <= SPB_LMAX; ++ll) s += ez[ll * ll + ll + n] - ez[ll * ll + ll - n];
    p.amu[(size_t)b * 32 + 15 + n] = s;
  }
  (void)red;
}

// ---- Omega_m(u)[p][q] = Omega_m[p][q] z_{l_p} z_{l_q} ---------------------------------------
__global__ void omega_scale_kernel(const double *tab, const double *rTA1, double *out) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)31 * 65536) return;
  const int pq = (int)(idx & 65535);
  const int pr = pq >> 8, q = pq & 255;
  int lp, mp, lq, mq;
  lm_of(pr, lp, mp);
  lm_of(q, lq, mq);
  out[idx] = tab[SPB_TAB_FLUX_OMEGA + idx] * rTA1[lp * lp + lp] * rTA1[lq * lq + lq];
}

// ---- marginal, stage 3: a_m -> kernel grid -> cubic coefficients ---------------------------
struct CoefParams {
  int B, covpts, ksplit;
  const double *acov_part;  // (ksplit, B, 64) partial GEMM results, columns 0..15 used
  long long strideSplit;
  const double *amu, *gp_mean;
  double *var, *coef;       // coef: (B, 4, covpts+1)
};

__global__ void __launch_bounds__(128) marginal_coef_kernel(CoefParams p) {
  extern __shared__ double yp[];  // covpts + 4
  __shared__ double am[32];  // a_0..a_15, b_1..b_15
  const int b = blockIdx.x, tid = threadIdx.x;
  if (tid < 31) {
    double s = p.amu[(size_t)b * 32 + tid];
    for (int k = 0; k < p.ksplit; ++k) s += p.acov_part[(size_t)k * p.strideSplit + (size_t)b * 64 + tid];
    am[tid] = s;
  }
  __syncthreads();
  const double mean = p.gp_mean[b];
  const int npts = p.covpts + 4;
  const double dx = 2.0 * 3.14159265358979323846 / p.covpts;
  for (int k = tid; k < npts; k += 128) {
    const double x = -dx + k * dx;  // flux.py:311-314 (numpy arange: start + k*step)
    const double c1 = cos(x), s1 = sin(x);
    double cm1 = 1.0, cm = c1, sm1 = 0.0, sm = s1;
    double s = am[0] + am[1] * c1 + am[16] * s1;
    for (int mm = 2; mm < 16; ++mm) {
      const double cn = 2.0 * cm * c1 - cm1;  // wigner.h:311-316
      const double sn = 2.0 * sm * c1 - sm1;
      cm1 = cm;
      cm = cn;
      sm1 = sm;
      sm = sn;
      s += am[mm] * cn + am[15 + mm] * sn;
    }
    yp[k] = s - mean * mean;  // flux.py:320
  }
  if (tid == 0) {
    double s = 0.0;
    for (int mm = 0; mm < 16; ++mm) s += am[mm];
    p.var[b] = s - mean * mean;  // flux.py:306-308
  }
  __syncthreads();
  const int nc = p.covpts + 1;
  double *cf = p.coef + (size_t)b * 4 * nc;
  for (int k = tid; k < nc; k += 128) {
    const double y0 = yp[k], y1 = yp[k + 1], y2 = yp[k + 2], y3 = yp[k + 3];
    cf[k] = y1;                                                   // flux.py:327-330
    cf[nc + k] = -y0 / 3.0 - 0.5 * y1 + y2 - y3 / 6.0;
    cf[2 * nc + k] = 0.5 * (y0 + y2) - y1;
    cf[3 * nc + k] = 0.5 * ((y1 - y2) + (y3 - y0) / 3.0);
  }
}

__global__ void cond_mean_kernel(int B, const double *A, long long A_stride, const double *mean_ylm,
                                 double *gp_mean) {
  // mean = (A . mean_ylm)[0]  (flux.py:340)
  const int b = blockIdx.x, n = threadIdx.x;
  __shared__ double red[8];
  double v = warp_sum(A[(size_t)b * A_stride + n] * mean_ylm[(size_t)b * 256 + n]);
  if ((n & 31) == 0) red[n >> 5] = v;
  __syncthreads();
  if (n == 0) {
    double s = 0.0;
    for (int k = 0; k < 8; ++k) s += red[k];
    gp_mean[b] = s;
  }
}

// Split-K factor of the marginal-kernel GEMM (B x 65536) . (65536 x 31).  FIXED (independent of the
// batch size, so that a sample's result does not depend on how the batch was chunked or sharded):
// 32 slices give 32 * ceil(B / 64) CTAs -- 256 at the 512 samples one GPU of an 8-way split sees
// (8 slices left 64 CTAs on 148 SMs: 0.72 ms against 0.22 ms for its share of a 4096 batch).
constexpr int MARG_KSPLIT = 32;
constexpr int COND_CHUNK = 512;
constexpr int COND_I8_MIN_NT = 1024;   // lower-only K = T A^T on the INT8 tensor cores from this size

}  // namespace

extern "C" int spb_flux_operator(spb_context *ctx, int nu, const double *u, double *rTA1,
                                 void *stream) {
  SPB_REQUIRE(ctx != nullptr && nu > 0, "flux_operator: bad arguments");
  SPB_REQUIRE(ctx->tables_count == SPB_TAB_TOTAL, "flux_operator: context has no constant tables");
  SPB_CHECK_CUDA(cudaSetDevice(ctx->device));
  flux_operator_kernel<<<nu, 256, 0, (cudaStream_t)stream>>>(nu, u, ctx->d_tables, rTA1);
  SPB_LAUNCH_CHECK(ctx);
  return 0;
}

extern "C" size_t spb_flux_marginal_workspace_bytes(const spb_context *ctx, int B) {
  (void)ctx;
  size_t bytes = (size_t)31 * 65536 * 8;                 // Omega(u)
  bytes += (size_t)MARG_KSPLIT * B * 64 * 8;             // split-K partials
  bytes += (size_t)B * 32 * 8;                           // mu part
  return bytes + 1024;
}

extern "C" int spb_flux_marginal(spb_context *ctx, int B, const double *mean_ylm,
                                 const double *cov_ylm, const double *rTA1, int covpts,
                                 double *gp_mean, double *var, double *coef, void *workspace,
                                 size_t workspace_bytes, void *stream_) {
  SPB_REQUIRE(ctx != nullptr && B > 0 && covpts >= 4, "flux_marginal: bad arguments");
  SPB_REQUIRE(ctx->tables_count == SPB_TAB_TOTAL, "flux_marginal: context has no constant tables");
  SPB_REQUIRE(workspace != nullptr && workspace_bytes >= spb_flux_marginal_workspace_bytes(ctx, B),
              "flux_marginal: workspace too small");
  SPB_REQUIRE(((uintptr_t)workspace % 256) == 0, "flux_marginal: workspace must be 256-byte aligned");
  cudaStream_t stream = (cudaStream_t)stream_;
  SPB_CHECK_CUDA(cudaSetDevice(ctx->device));
  double *omega_u = reinterpret_cast<double *>(workspace);
  double *part = omega_u + (size_t)31 * 65536;
  double *amu = part + (size_t)MARG_KSPLIT * B * 64;

  MargParams mp;
  mp.B = B;
  mp.mean_ylm = mean_ylm;
  mp.rTA1 = rTA1;
  mp.tab = ctx->d_tables;
  mp.gp_mean = gp_mean;
  mp.amu = amu;
  marginal_mu_kernel<<<B, 256, 0, stream>>>(mp);
  SPB_LAUNCH_CHECK(ctx);
  omega_scale_kernel<<<(31 * 65536 + 255) / 256, 256, 0, stream>>>(ctx->d_tables, rTA1, omega_u);
  SPB_LAUNCH_CHECK(ctx);

  // acov[b][m] = sum_k cov[b][k] Omega_u[m][k], K = 65536, split-K partials
  for (int b0 = 0; b0 < B; b0 += 32768) {
    const int Bc = (B - b0 < 32768) ? B - b0 : 32768;
    gnt::Desc d = {};
    d.A = cov_ylm + (size_t)b0 * 65536;
    d.strideA = 0;
    d.lda = 65536;
    d.Bm = omega_u;
    d.strideB = 0;
    d.ldb = 65536;
    d.C = part + (size_t)b0 * 64;
    d.strideC = 0;
    d.ldc = 64;
    d.M = Bc;
    d.N = 31;
    d.K = 65536;
    d.batch = 1;
    d.ksplit = MARG_KSPLIT;
    d.strideSplit = (long long)B * 64;
    d.lower_only = 0;
    d.alpha = 1.0;
    int st = gnt::launch<gnt::EPI_STORE>(ctx, d, stream);
    if (st) return st;
  }
  CoefParams cp;
  cp.B = B;
  cp.covpts = covpts;
  cp.ksplit = MARG_KSPLIT;
  cp.acov_part = part;
  cp.strideSplit = (long long)B * 64;
  cp.amu = amu;
  cp.gp_mean = gp_mean;
  cp.var = var;
  cp.coef = coef;
  marginal_coef_kernel<<<B, 128, (covpts + 4) * sizeof(double), stream>>>(cp);
  SPB_LAUNCH_CHECK(ctx);
  return 0;
}

extern "C" size_t spb_flux_conditional_workspace_bytes(const spb_context *ctx, int B, int nt) {
  (void)ctx;
  const int Bc = B < COND_CHUNK ? B : COND_CHUNK;
  // T = A Sigma of one chunk, plus (long light curves) the digit planes of the INT8 product T A^T
  return (size_t)Bc * nt * 256 * 8 + 1024 + (nt >= COND_I8_MIN_NT ? spb_gemm_i8_lower_workspace_bytes(Bc, nt) : 0);
}

static int flux_conditional_impl(spb_context *ctx, int B, int nt, const double *A, long long A_stride,
                                 const double *mean_ylm, const double *cov_ylm, double *gp_mean, double *K,
                                 int ldk, void *workspace, size_t workspace_bytes, void *stream_,
                                 bool lower_only) {
  SPB_REQUIRE(ctx != nullptr && B > 0 && nt > 0, "flux_conditional: bad arguments");
  SPB_REQUIRE(workspace != nullptr &&
                  workspace_bytes >= spb_flux_conditional_workspace_bytes(ctx, B, nt),
              "flux_conditional: workspace too small");
  SPB_REQUIRE(ldk >= nt && (ldk % 2) == 0, "flux_conditional: ldk must be even and >= nt");
  cudaStream_t stream = (cudaStream_t)stream_;
  SPB_CHECK_CUDA(cudaSetDevice(ctx->device));
  double *T = reinterpret_cast<double *>(workspace);
  cond_mean_kernel<<<B, 256, 0, stream>>>(B, A, A_stride, mean_ylm, gp_mean);
  SPB_LAUNCH_CHECK(ctx);
  for (int b0 = 0; b0 < B; b0 += COND_CHUNK) {
    const int Bc = (B - b0 < COND_CHUNK) ? B - b0 : COND_CHUNK;
    // T[b] = A cov[b]   (cov symmetric: T[i][j] = sum_k A[i][k] cov[j][k])
    gnt::Desc d1 = {};
    d1.A = A + (size_t)b0 * A_stride;
    d1.strideA = A_stride;
    d1.lda = 256;
    d1.Bm = cov_ylm + (size_t)b0 * 65536;
    d1.strideB = 65536;
    d1.ldb = 256;
    d1.C = T;
    d1.strideC = (long long)nt * 256;
    d1.ldc = 256;
    d1.M = nt;
    d1.N = 256;
    d1.K = 256;
    d1.batch = Bc;
    d1.ksplit = 1;
    d1.alpha = 1.0;
    int st = gnt::launch<gnt::EPI_STORE>(ctx, d1, stream);
    if (st) return st;
    // K[b] = T[b] A^T: lower tiles, mirrored into the upper triangle unless the caller only needs the
    // lower one (the log-likelihood path: the Cholesky kernels never read above the diagonal, and the
    // mirror is 134 MB of strided 8-byte stores per nt = 4096 sample)
    gnt::Desc d2 = {};
    d2.A = T;
    d2.strideA = (long long)nt * 256;
    d2.lda = 256;
    d2.Bm = A + (size_t)b0 * A_stride;
    d2.strideB = A_stride;
    d2.ldb = 256;
    d2.C = K + (size_t)b0 * nt * ldk;
    d2.strideC = (long long)nt * ldk;
    d2.ldc = ldk;
    d2.M = nt;
    d2.N = nt;
    d2.K = 256;
    d2.batch = Bc;
    d2.ksplit = 1;
    d2.lower_only = 1;
    d2.alpha = 1.0;
    if (lower_only && nt >= COND_I8_MIN_NT && ctx->opt_syrk_i8) {
      // long light curves: the 4.3 GFLOP (nt = 4096) product on the INT8 tensor cores (syrk_i8.cu)
      void *wi8 = reinterpret_cast<unsigned char *>(workspace) +
                  (((size_t)(B < COND_CHUNK ? B : COND_CHUNK) * nt * 256 * 8 + 1023) & ~(size_t)1023);
      st = spb_gemm_i8_lower(ctx, Bc, nt, T, A + (size_t)b0 * A_stride, A_stride, d2.C, ldk, wi8, stream);
    } else {
      st = lower_only ? gnt::launch<gnt::EPI_STORE>(ctx, d2, stream)
                      : gnt::launch<gnt::EPI_MIRROR>(ctx, d2, stream);
    }
    if (st) return st;
  }
  return 0;
}

extern "C" int spb_flux_conditional(spb_context *ctx, int B, int nt, const double *A,
                                    long long A_stride, const double *mean_ylm,
                                    const double *cov_ylm, double *gp_mean, double *K, int ldk,
                                    void *workspace, size_t workspace_bytes, void *stream_) {
  return flux_conditional_impl(ctx, B, nt, A, A_stride, mean_ylm, cov_ylm, gp_mean, K, ldk, workspace,
                               workspace_bytes, stream_, false);
}

extern "C" int spb_flux_conditional_lower(spb_context *ctx, int B, int nt, const double *A,
                                          long long A_stride, const double *mean_ylm,
                                          const double *cov_ylm, double *gp_mean, double *K, int ldk,
                                          void *workspace, size_t workspace_bytes, void *stream_) {
  return flux_conditional_impl(ctx, B, nt, A, A_stride, mean_ylm, cov_ylm, gp_mean, K, ldk, workspace,
                               workspace_bytes, stream_, true);
}
