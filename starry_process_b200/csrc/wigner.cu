// Real Wigner rotations and the flux design matrix.
//
//   spb_Rx            Rx(theta), all l <= 15, packed (5456)      ops/wigner/Rx.cc:10-49 ->
//                     rotar / dlmn                               ops/include/wigner.h:37-284
//   spb_tensordotRz   f = M . Rz(theta_k) row by row             ops/wigner/tensordotRz.cc:10-56,
//                                                                wigner.h:290-339
//   spb_design_matrix A(t; i, p, u) = rTA1 . Rx(-i) . Rz(theta) . Rx(pi/2)
//                                                                flux.py:88-105, 278-281
//
// Built with -fmad=false: the Alvarez-Collado recurrences are evaluated with the same unfused
// multiply/add sequence as the reference's `g++ -O2` build, but PARALLEL over the (m', m) entries
// of each degree (the reference walks them serially): one CTA per angle, the complex matrices
// D[l] for all l kept in shared memory (43.6 KB).
#include "common.cuh"
#include "spb_tables.h"

namespace {

__host__ __device__ __forceinline__ int nwig(int l) { return ((l + 1) * (2 * l + 1) * (2 * l + 3)) / 3; }

// integer cos/sin of k*pi/2 style phase tables used by rotar (wigner.h:232-270)
__device__ __forceinline__ void phase_al(int mp, int &c, int &s) {
  // (cosmal, sinmal) starts (0,-1) at mp = 1 and maps (c, s) -> (s, -c)
  const int k = (mp - 1) & 3;
  c = (k == 0) ? 0 : (k == 1) ? -1 : (k == 2) ? 0 : 1;
  s = (k == 0) ? -1 : (k == 1) ? 0 : (k == 2) ? 1 : 0;
}
__device__ __forceinline__ void phase_ga(int m, int &c, int &s) {
  // (cosmga, sinmga) starts (0, 1) at m = 1 and maps (c, s) -> (-s, c)
  const int k = (m - 1) & 3;
  c = (k == 0) ? 0 : (k == 1) ? -1 : (k == 2) ? 0 : 1;
  s = (k == 0) ? 1 : (k == 1) ? 0 : (k == 2) ? -1 : 0;
}

__global__ void __launch_bounds__(256) rx_kernel(int nang, const double *theta, double sign_in,
                                                 double *Rout) {
  __shared__ double D[SPB_NWIG];
  const int tid = threadIdx.x;
  const int ia = blockIdx.x;
  if (ia >= nang) return;
  double *R = Rout + (size_t)ia * SPB_NWIG;
  const double th = sign_in * theta[ia];
  const double root_two = sqrt(2.0);
  const double c2 = cos(th), s2 = sin(th);

  if (tid == 0) {
    // wigner.h:164-206
    D[0] = 1.0;
    D[9] = 0.5 * (1.0 + c2);
    D[8] = -s2 / root_two;
    D[7] = 0.5 * (1.0 - c2);
    D[6] = -D[8];
    D[5] = D[9] - D[7];
    D[4] = D[8];
    D[3] = D[7];
    D[2] = D[6];
    D[1] = D[9];
    R[0] = 1.0;
    R[1] = D[9] - D[7];
    R[2] = -root_two * D[6];
    R[3] = 0;
    R[4] = -root_two * D[8];
    R[5] = D[5];
    R[6] = 0;
    R[7] = 0;
    R[8] = 0;
    R[9] = D[9] + D[7];
  }
  __syncthreads();
  double tgbet2;
  if (fabs(s2) < 1.0e-14) tgbet2 = s2;  // SP_WIGNER_TOL, constants.h:71-73
  else tgbet2 = (1.0 - c2) / s2;

  for (int l = 2; l <= SPB_LMAX; ++l) {
    const int w = 2 * l + 1, w1 = w - 2, w2 = w - 4;
    double *Dl = D + nwig(l - 1);
    const double *Dm1 = D + nwig(l - 2);
    const double *Dm2 = D + nwig(l - 3);
    double *Rl = R + nwig(l - 1);
#define DL(r, c) Dl[(r)*w + (c)]
#define DM1(r, c) Dm1[(r)*w1 + (c)]
#define DM2(r, c) Dm2[(r)*w2 + (c)]
    // (a) top row by recurrence, wigner.h:57-73 (serial in m)
    if (tid == 0) {
      const int isup = l - 1, iinf = 1 - l;
      DL(2 * l, 2 * l) = 0.5 * DM1(isup + l - 1, isup + l - 1) * (1.0 + c2);
      DL(2 * l, 0) = 0.5 * DM1(isup + l - 1, -isup + l - 1) * (1.0 - c2);
      for (int m = isup; m > iinf - 1; --m)
        DL(2 * l, m + l) = -tgbet2 * sqrt((double)(l + m + 1) / (l - m)) * DL(2 * l, m + 1 + l);
    }
    // (b) upper quarter triangle, wigner.h:77-110: rows mp in [0, l-1], columns m in [-mp, mp]
    {
      const int al = l, al1 = l - 1, tal1 = al + al1;
      const double ali = 1.0 / al1;
      const double cosaux = c2 * al * al1;
      for (int idx = tid; idx < l * w; idx += 256) {
        const int mp = idx / w, m = idx % w - l;
        if (m < -mp || m > mp) continue;
        const int laux = l + mp, lbux = l - mp;
        const double aux = ali / sqrt((double)(laux * lbux));
        const double cux = sqrt((double)((laux - 1) * (lbux - 1))) * al;
        const int lauz = l + m, lbuz = l - m;
        const double auz = 1.0 / sqrt((double)(lauz * lbuz));
        const double fact = aux * auz;
        double term = tal1 * (cosaux - (double)(m * mp)) * DM1(mp + l - 1, m + l - 1);
        if ((lbuz != 1) && (lbux != 1)) {
          const double cuz = sqrt((double)((lauz - 1) * (lbuz - 1)));
          term = term - DM2(mp + l - 2, m + l - 2) * cux * cuz;
        }
        DL(mp + l, m + l) = fact * term;
      }
    }
    __syncthreads();
    // (c) reflection, wigner.h:117-129: D(mp, m) = (-1)^(mp+m) D(m, mp), m in [1, l], mp in [-m, m-1]
    for (int idx = tid; idx < l * w; idx += 256) {
      const int m = idx / w + 1, mp = idx % w - l;
      if (mp < -m || mp > m - 1) continue;
      const double sg = ((mp + m) & 1) ? -1.0 : 1.0;
      DL(mp + l, m + l) = sg * DL(m + l, mp + l);
    }
    __syncthreads();
    // (d) inversion, wigner.h:131-142: D(mp, m) = (-1)^(mp+m) D(-mp, -m), m in [-l, l-1], mp in [-l, -m-1]
    for (int idx = tid; idx < w * w; idx += 256) {
      const int m = idx / w - l, mp = idx % w - l;
      if (m > l - 1 || mp > -m - 1) continue;
      const double sg = ((mp + m) & 1) ? -1.0 : 1.0;
      DL(mp + l, m + l) = sg * DL(-mp + l, -m + l);
    }
    __syncthreads();
    // (e) real matrices from the complex ones, wigner.h:226-270
    if (tid == 0) Rl[l * w + l] = DL(l, l);
    for (int idx = tid; idx < l; idx += 256) {
      const int mp = idx + 1;
      int cal, sal;
      phase_al(mp, cal, sal);
      Rl[(mp + l) * w + l] = root_two * DL(l, mp + l) * cal;
      Rl[(-mp + l) * w + l] = root_two * DL(l, mp + l) * sal;
      // the axis entries R(l, +-m) are rewritten identically for every mp in the reference
      const int m = mp;
      int cga, sga;
      phase_ga(m, cga, sga);
      Rl[l * w + (m + l)] = root_two * DL(m + l, l) * cga;
      Rl[l * w + (-m + l)] = -root_two * DL(m + l, l) * sga;
    }
    for (int idx = tid; idx < l * l; idx += 256) {
      const int mp = idx / l + 1, m = idx % l + 1;
      int cal, sal, cga, sga;
      phase_al(mp, cal, sal);
      phase_ga(m, cga, sga);
      const int sgn = (mp & 1) ? -1 : 1;
      const double d1 = DL(-mp + l, -m + l);
      const double d2 = sgn * DL(mp + l, -m + l);
      const int cosag = cal * cga - sal * sga;
      const int cosagm = cal * cga + sal * sga;
      const int sinag = sal * cga + cal * sga;
      const int sinagm = sal * cga - cal * sga;
      Rl[(mp + l) * w + (m + l)] = d1 * cosag + d2 * cosagm;
      Rl[(mp + l) * w + (-m + l)] = -d1 * sinag + d2 * sinagm;
      Rl[(-mp + l) * w + (m + l)] = d1 * sinag + d2 * sinagm;
      Rl[(-mp + l) * w + (-m + l)] = d1 * cosag - d2 * cosagm;
    }
    __syncthreads();
#undef DL
#undef DM1
#undef DM2
  }
}

// cos(n theta), sin(n theta) by the recurrences of wigner.h:307-316
__device__ __forceinline__ void cheb_cs(double theta, double *cs, double *sn) {
  cs[0] = 1.0;
  sn[0] = 0.0;
  cs[1] = cos(theta);
  sn[1] = sin(theta);
  for (int n = 2; n <= SPB_LMAX; ++n) {
    cs[n] = 2.0 * cs[n - 1] * cs[1] - cs[n - 2];
    sn[n] = 2.0 * sn[n - 1] * cs[1] - sn[n - 2];
  }
}

__global__ void __launch_bounds__(256) tensordotRz_kernel(int K, const double *M, const double *theta,
                                                          double *f) {
  __shared__ double cs[16], sn[16];
  const int k = blockIdx.x, n = threadIdx.x;
  if (k >= K) return;
  if (n == 0) cheb_cs(theta[k], cs, sn);
  __syncthreads();
  const int l = (int)floor(sqrt((double)n) + 1e-9);
  const int j = n - l * l, m = j - l;
  const double *Mk = M + (size_t)k * 256;
  const double cm = cs[m < 0 ? -m : m];
  const double sm = m < 0 ? -sn[-m] : sn[m];
  f[(size_t)k * 256 + n] = Mk[n] * cm + Mk[l * l + 2 * l - j] * sm;
}

// ------------------------------------------------------------------------------------------
// Design matrix.  grid = (row tiles of 32 timestamps, inclinations)
//   v        = rTA1 . Rx(-i)                      (hoisted: identical for every timestamp)
//   f[t]     = v . Rz(theta_t),  theta_t = 2 pi mod(t/p, 1)
//   A[t, :]  = f[t] . Rx(pi/2)                    (block diagonal, 5456 MAC per row)
// thread n owns output column n for the 32 rows of the tile; rows are written as full 2 KB lines.
// ------------------------------------------------------------------------------------------
constexpr int DM_ROWS = 32;

struct DesignParams {
  int I, nt;
  const double *t, *inc, *period, *rTA1;
  int rTA1_stride;
  const double *RxInc;  // (I, 5456)  Rx(-inc)
  const double *Rx90;   // (5456)
  double *A;
};

__global__ void __launch_bounds__(256) design_kernel(DesignParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double *Rsh = reinterpret_cast<double *>(smem_raw);  // 5456
  double *v = Rsh + SPB_NWIG;                          // 256
  double *fsh = v + 256;                               // DM_ROWS x 256
  double *csn = fsh + DM_ROWS * 256;                   // DM_ROWS x 32 (cos | sin)
  const int n = threadIdx.x;
  const int ii = blockIdx.y;
  const int t0 = blockIdx.x * DM_ROWS;
  const int l = (int)floor(sqrt((double)n) + 1e-9);
  const int j = n - l * l, m = j - l, w = 2 * l + 1;

  for (int idx = n; idx < SPB_NWIG; idx += 256) Rsh[idx] = p.Rx90[idx];
  // v[(l, m)] = sum_m' rTA1[(l, m')] Rx(-i)_l[m'][m]   (flux.py:95-96, 74-86)
  {
    const double *rt = p.rTA1 + (size_t)ii * p.rTA1_stride + l * l;
    const double *rx = p.RxInc + (size_t)ii * SPB_NWIG + nwig(l - 1);
    double acc = 0.0;
    for (int mp = 0; mp < w; ++mp) acc += rt[mp] * rx[mp * w + j];
    v[n] = acc;
  }
  if (n < DM_ROWS) {
    const int t = t0 + n;
    if (t < p.nt) {
      const double per = p.period ? p.period[ii] : 1.0;
      const double x = p.t[t] / per;
      const double theta = 2.0 * 3.14159265358979323846 * (x - floor(x));  // tt.mod(t/p, 1)
      cheb_cs(theta, csn + n * 32, csn + n * 32 + 16);
    }
  }
  __syncthreads();
  const int nrows = min(DM_ROWS, p.nt - t0);
  const double vn = v[n], vb = v[l * l + 2 * l - j];
  const int am = m < 0 ? -m : m;
  for (int r = 0; r < nrows; ++r) {
    const double cm = csn[r * 32 + am];
    const double sm = (m < 0) ? -csn[r * 32 + 16 + am] : csn[r * 32 + 16 + am];
    fsh[r * 256 + n] = vn * cm + vb * sm;  // wigner.h:331-337
  }
  __syncthreads();
  double acc[DM_ROWS];
#pragma unroll
  for (int r = 0; r < DM_ROWS; ++r) acc[r] = 0.0;
  const double *rx = Rsh + nwig(l - 1) + j;
  const double *fl = fsh + l * l;
  for (int mp = 0; mp < w; ++mp) {
    const double rv = rx[mp * w];
#pragma unroll
    for (int r = 0; r < DM_ROWS; ++r) acc[r] = fma(fl[r * 256 + mp], rv, acc[r]);
  }
  double *Ab = p.A + ((size_t)ii * p.nt + t0) * 256 + n;
#pragma unroll
  for (int r = 0; r < DM_ROWS; ++r)
    if (r < nrows) Ab[(size_t)r * 256] = acc[r];
}

}  // namespace

extern "C" int spb_Rx(spb_context *ctx, int nang, const double *theta, double *Rx, void *stream) {
  SPB_REQUIRE(ctx != nullptr && nang > 0, "Rx: bad arguments");
  SPB_CHECK_CUDA(cudaSetDevice(ctx->device));
  rx_kernel<<<nang, 256, 0, (cudaStream_t)stream>>>(nang, theta, 1.0, Rx);
  SPB_LAUNCH_CHECK(ctx);
  return 0;
}

extern "C" int spb_tensordotRz(spb_context *ctx, int K, const double *M, const double *theta,
                               double *f, void *stream) {
  SPB_REQUIRE(ctx != nullptr && K > 0, "tensordotRz: bad arguments");
  SPB_CHECK_CUDA(cudaSetDevice(ctx->device));
  tensordotRz_kernel<<<K, 256, 0, (cudaStream_t)stream>>>(K, M, theta, f);
  SPB_LAUNCH_CHECK(ctx);
  return 0;
}

extern "C" size_t spb_design_matrix_workspace_bytes(const spb_context *ctx, int I, int nt) {
  (void)ctx;
  (void)nt;
  return (size_t)I * SPB_NWIG * sizeof(double);
}

extern "C" int spb_design_matrix(spb_context *ctx, int I, int nt, const double *t,
                                 const double *inc_rad, const double *period, const double *rTA1,
                                 int rTA1_stride, double *A, void *workspace,
                                 size_t workspace_bytes, void *stream_) {
  SPB_REQUIRE(ctx != nullptr && I > 0 && nt > 0, "design_matrix: bad arguments");
  SPB_REQUIRE(ctx->tables_count == SPB_TAB_TOTAL, "design_matrix: context has no constant tables");
  SPB_REQUIRE(workspace != nullptr && workspace_bytes >= (size_t)I * SPB_NWIG * sizeof(double),
              "design_matrix: workspace too small");
  cudaStream_t stream = (cudaStream_t)stream_;
  SPB_CHECK_CUDA(cudaSetDevice(ctx->device));
  double *RxInc = reinterpret_cast<double *>(workspace);
  rx_kernel<<<I, 256, 0, stream>>>(I, inc_rad, -1.0, RxInc);  // Rx(-i), flux.py:96
  SPB_LAUNCH_CHECK(ctx);
  DesignParams p;
  p.I = I;
  p.nt = nt;
  p.t = t;
  p.inc = inc_rad;
  p.period = period;
  p.rTA1 = rTA1;
  p.rTA1_stride = rTA1_stride;
  p.RxInc = RxInc;
  p.Rx90 = ctx->d_tables + SPB_TAB_RX90;
  p.A = A;
  const size_t smem = (SPB_NWIG + 256 + DM_ROWS * 256 + DM_ROWS * 32) * sizeof(double);
  static bool attr = false;
  if (!attr) {
    SPB_CHECK_CUDA(cudaFuncSetAttribute(design_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)smem));
    attr = true;
  }
  dim3 grid((nt + DM_ROWS - 1) / DM_ROWS, I);
  SPB_REQUIRE(I <= 65535, "design_matrix: too many inclinations for one launch");
  design_kernel<<<grid, 256, smem, stream>>>(p);
  SPB_LAUNCH_CHECK(ctx);
  return 0;
}
