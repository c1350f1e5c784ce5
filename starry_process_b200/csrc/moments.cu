// Ylm moment integrals: (r, a, b, c, n) -> mean_ylm (256), cov_ylm (256 x 256), batched.
//
// Replaces the reference chain  SizeIntegral -> LatitudeIntegral -> LongitudeIntegral ->
// ContrastIntegral  (size.py:93-115; latitude.py:171-212 + ops/include/latitude.h:22-173;
// integrals.py:116-151; math.py:121-139 + ops/eigh/eigh.py:11-20; longitude.py:9-49;
// contrast.py:9-33).  Built with -fmad=false: the Beta-moment recurrences and the binomial double
// sums of latitude.h are reproduced with the reference's unfused operation order; bulk
// contractions use explicit fma().
//
// B200-first re-organisation (DESIGN.md "moments"):
//   K1  one CTA per hyperparameter sample:
//         spot profile -> q_size (16 values)                         size.py:45-53
//         Beta moments B(k), even/even binomial table term(2a,2b)    latitude.h:48-60,112-143
//           (the odd-index "F"/2F1 lane of latitude.h:63-109 only ever feeds entries that the
//            scatter of latitude.h:146-172 never reads -- l+m and l-m have equal parity -- so it
//            is not evaluated)
//         the 256x256 eigenproblem of math.py:121-139 is solved in the 31-dimensional range of
//         Q (constant orthonormal basis Z): S = Z^T Q Z, parallel cyclic Jacobi in shared
//         memory, eigenvalues <= 1e-15 clipped exactly as matrix_sqrt does
//         -> sqrtC_lat (256 x r), first moments through the folded tensors R0, t_lon
//   K2  sqrtC_lon = T_lon . sqrtC_lat  (integrals.py:133-138), T_lon slices staged in shared
//       memory and re-used across 32 samples
//   K3  cov = (pi c)^2 n (sqrtC_lon sqrtC_lon^T - mom1 mom1^T) + diag(lambda) on the FP64 tensor
//       pipe (gemm_nt.cuh, lower tiles mirrored); the longitude re-factorisation of
//       integrals.py:144-150 only clips <=1e-15 modes of an explicitly PSD product and is skipped.
#include "gemm_nt.cuh"
#include "spb_tables.h"

namespace {

constexpr int NT1 = 256;
constexpr int JS = 33;  // Jacobi smem stride

struct K1Params {
  const double *r_deg, *a, *b, *c, *n;
  int B;
  const double *tab;
  double *mom1;      // (B,256)   first moment before the contrast scaling
  double *mean_ylm;  // (B,256)
  double *S_lat;     // (B,256,32) sqrtC_lat, first rkeep columns non-zero
  double *scale;     // (B)       (pi c)^2 n
  int *rkeep;        // (B)
  int32_t *info;     // (B)
};

struct K1Smem {
  double bprof[1000];
  double qs[16];
  double Bk[64];
  double tt[31][31];      // term(2a, 2b)
  double Y[256][33];      // Q Z, later re-used
  double A[32][JS];       // Jacobi iterate
  double V[32][JS];       // eigenvectors
  double Xs[32][JS];      // V diag(sqrt w), kept modes compacted
  double cs[16][2];
  double m1lat[256];
  int pp[16], qq[16];
  int rotated;
  int order[32];
  int rk;
};

__device__ __forceinline__ void lm_of(int n, int &l, int &m) {
  l = (int)floor(sqrt((double)n));
  while (l * l > n) --l;
  while ((l + 1) * (l + 1) <= n) ++l;
  m = n - l * l - l;
}

__global__ void __launch_bounds__(NT1, 2) moments_k1(K1Params p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  K1Smem &sm = *reinterpret_cast<K1Smem *>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.x;
  const double *tab = p.tab;

  const double ang = 3.14159265358979323846 / 180.0;
  const double r = p.r_deg[b] * ang;
  double aa = p.a[b], bb = p.b[b];
  const double cc = p.c[b], nn = p.n[b];
  // CheckBoundsOp semantics (ops/exceptions.py:30-48; size.py:103, latitude.py:179-181,
  // contrast.py:15): the batched path flags the element instead of raising.
  const double tol = 1e-6;
  bool bad = !(r >= -tol && r <= 0.5 * 3.14159265358979323846 + tol) || !(aa >= -tol && aa <= 1 + tol) ||
             !(bb >= -tol && bb <= 1 + tol) || !(nn >= -tol);
  if (aa < 1e-12) aa = 1e-12;  // latitude.py:180-182 (abmin)
  if (bb < 1e-12) bb = 1e-12;

  // ---- spot profile, size.py:45-53 (sfac = 300)
  for (int s = tid; s < 1000; s += NT1) {
    const double z = 300.0 * (tab[SPB_TAB_THETA + s] - r);
    sm.bprof[s] = 1.0 / (1.0 + exp(-z)) - 1.0;
  }
  __syncthreads();
  for (int row = warp; row < 16; row += NT1 / 32) {
    const double *bp = tab + SPB_TAB_BP + (size_t)row * 1000;
    double acc = 0.0;
    for (int s = lane; s < 1000; s += 32) acc = fma(bp[s], sm.bprof[s], acc);
    acc = warp_sum(acc);
    if (lane == 0) sm.qs[row] = acc;
  }

  // ---- Beta moments, latitude.py:197-200 and latitude.h:48-60
  if (tid == 0) {
    const double alpha0 = exp(aa * 10.0);
    const double beta0 = exp(log(0.5) + bb * (10.0 - log(0.5)));
    const double alpha = alpha0 > 0.0 ? alpha0 : 0.0;  // ops/latitude/latitude.cc:47-48
    const double beta = beta0 > 0.0 ? beta0 : 0.0;
    sm.Bk[0] = 1.0;
    for (int k = 1; k < 61; ++k) {
      const double c1 = 1.0 / (alpha + beta + k - 1.0);
      const double c2 = (alpha + k - 1.0) * c1;
      sm.Bk[k] = c2 * sm.Bk[k - 1];
    }
  }
  __syncthreads();

  // ---- term(2a, 2b) = sum_k1 sum_k2 C(a,k1) (-1)^k2 C(b,k2) B(k1+k2), latitude.h:112-143
  // The signed binomial products come from the LAT_FAC table (built on the host with the
  // reference's ratio recurrences, bit-identical to evaluating them here); the accumulation order
  // is the reference's.  The heavy (a, b) pairs are dealt out first so the tail is short.
  {
    const double *fac = p.tab + SPB_TAB_LAT_FAC;
    const double *facoff = p.tab + SPB_TAB_LAT_FACOFF;
    for (int idx = tid; idx < 31 * 31; idx += NT1) {
      const int a2 = idx / 31, b2 = idx % 31;
      double acc = 0.0;
      if (a2 + b2 <= 30) {
        const double *f = fac + (int)facoff[idx];
        for (int k1 = 0; k1 < a2 + 1; ++k1) {
          const double *bk = sm.Bk + k1;
#pragma unroll 4
          for (int k2 = 0; k2 < b2 + 1; ++k2) acc += f[k2] * bk[k2];
          f += b2 + 1;
        }
      }
      sm.tt[a2][b2] = acc;
    }
  }
  __syncthreads();

  // ---- Y = Q Z with Q(n1,n2) = term(j1+j2, i1+i2) 2^-(l1+l2)   (latitude.h:146-172)
  int l1, m1;
  lm_of(tid, l1, m1);
  const int j1 = m1 + l1, i1 = l1 - m1;
  {
    double y[32];   // column 31 of the padded basis is zero
#pragma unroll
    for (int a = 0; a < 32; ++a) y[a] = 0.0;
    const double *Z = tab + SPB_TAB_LAT_Z;
    int l2 = 0, m2 = 0;
    for (int n2 = 0; n2 < 256; ++n2) {
      const int J = j1 + m2 + l2, I = i1 + l2 - m2;
      // Q(n1, n2) vanishes unless l+m has the same parity for both indices (latitude.h:146-172):
      // the warp-uniform row of Z is only fetched (as 16-byte loads) when this thread needs it
      if (!(I & 1)) {
        const double qv = ldexp(sm.tt[J >> 1][I >> 1], -(l1 + l2));
        const double2 *zr = reinterpret_cast<const double2 *>(Z + n2 * 32);
#pragma unroll
        for (int a = 0; a < 16; ++a) {
          const double2 z2 = __ldg(zr + a);
          y[2 * a] = fma(qv, z2.x, y[2 * a]);
          y[2 * a + 1] = fma(qv, z2.y, y[2 * a + 1]);
        }
      }
      if (++m2 > l2) {
        ++l2;
        m2 = -l2;
      }
    }
#pragma unroll
    for (int a = 0; a < 31; ++a) sm.Y[tid][a] = y[a];
  }
  __syncthreads();
  // ---- S = Z^T Y (31 x 31), symmetrised, padded to 32
  {
    const double *Z = tab + SPB_TAB_LAT_Z;
    for (int idx = tid; idx < 32 * 32; idx += NT1) {
      const int a = idx >> 5, c2 = idx & 31;
      double s = 0.0;
      if (a < 31 && c2 < 31) {
        for (int n1 = 0; n1 < 256; ++n1) s = fma(Z[n1 * 32 + a], sm.Y[n1][c2], s);
      }
      sm.V[a][c2] = s;  // staging
    }
  }
  __syncthreads();
  for (int idx = tid; idx < 32 * 32; idx += NT1) {
    const int a = idx >> 5, c2 = idx & 31;
    sm.A[a][c2] = 0.5 * (sm.V[a][c2] + sm.V[c2][a]);
  }
  __syncthreads();
  for (int idx = tid; idx < 32 * 32; idx += NT1) {
    const int a = idx >> 5, c2 = idx & 31;
    sm.V[a][c2] = (a == c2) ? 1.0 : 0.0;
  }
  __syncthreads();

  // ---- parallel cyclic Jacobi (round-robin ordering, 16 disjoint rotations per round)
  double amax = 0.0;
  for (int a = 0; a < 31; ++a) amax = fmax(amax, fabs(sm.A[a][a]));
  const double rot_tol = 1e-20 * amax;
  bool converged = false;
  for (int sweep = 0; sweep < 16 && !converged; ++sweep) {
    if (tid == 0) sm.rotated = 0;
    __syncthreads();
    for (int rnd = 0; rnd < 31; ++rnd) {
      if (tid < 16) {
        int pi, qi;
        if (tid == 0) {
          pi = rnd;
          qi = 31;
        } else {
          pi = (rnd + tid) % 31;
          qi = (rnd - tid + 31) % 31;
        }
        if (pi > qi) {
          const int t = pi;
          pi = qi;
          qi = t;
        }
        const double apq = sm.A[pi][qi];
        double cth = 1.0, sth = 0.0;
        // rotate unless a_pq is negligible against sqrt(a_pp a_qq) (relative criterion for PSD
        // matrices) or against the absolute floor 1e-20 max|a_ii|
        const double thr = fmax(rot_tol, 4.0e-15 * sqrt(fabs(sm.A[pi][pi] * sm.A[qi][qi])));
        if (fabs(apq) > thr && qi < 31) {
          const double tau = (sm.A[qi][qi] - sm.A[pi][pi]) / (2.0 * apq);
          const double t = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
          cth = 1.0 / sqrt(1.0 + t * t);
          sth = t * cth;
          sm.rotated = 1;
        }
        sm.pp[tid] = pi;
        sm.qq[tid] = qi;
        sm.cs[tid][0] = cth;
        sm.cs[tid][1] = sth;
      }
      __syncthreads();
      // rows: A <- J^T A
      for (int idx = tid; idx < 16 * 32; idx += NT1) {
        const int k = idx >> 5, j = idx & 31;
        const int pi = sm.pp[k], qi = sm.qq[k];
        const double cth = sm.cs[k][0], sth = sm.cs[k][1];
        const double ap = sm.A[pi][j], aq = sm.A[qi][j];
        sm.A[pi][j] = cth * ap - sth * aq;
        sm.A[qi][j] = sth * ap + cth * aq;
      }
      __syncthreads();
      // columns: A <- A J, V <- V J
      for (int idx = tid; idx < 16 * 32; idx += NT1) {
        const int k = idx >> 5, i = idx & 31;
        const int pi = sm.pp[k], qi = sm.qq[k];
        const double cth = sm.cs[k][0], sth = sm.cs[k][1];
        const double ap = sm.A[i][pi], aq = sm.A[i][qi];
        sm.A[i][pi] = cth * ap - sth * aq;
        sm.A[i][qi] = sth * ap + cth * aq;
        const double vp = sm.V[i][pi], vq = sm.V[i][qi];
        sm.V[i][pi] = cth * vp - sth * vq;
        sm.V[i][qi] = sth * vp + cth * vq;
      }
      __syncthreads();
    }
    converged = (sm.rotated == 0);
    __syncthreads();
  }
  if (!converged) {
    // rotations at the rounding floor can keep a sweep "busy"; the solve has failed only if a
    // significant off-diagonal element survives
    double offmax = 0.0;
    for (int a = 0; a < 31; ++a)
      for (int c2 = a + 1; c2 < 31; ++c2) offmax = fmax(offmax, fabs(sm.A[a][c2]));
    converged = offmax <= 1e-13 * amax;
  }

  // ---- matrix_sqrt clip (math.py:133-136): keep w > 1e-15, compact kept modes to the front
  if (tid == 0) {
    int rk = 0;
    for (int e = 0; e < 31; ++e)
      if (sm.A[e][e] > 1e-15) sm.order[rk++] = e;
    sm.rk = rk;
  }
  __syncthreads();
  const int rk = sm.rk;
  for (int idx = tid; idx < 32 * 32; idx += NT1) {
    const int a = idx >> 5, e = idx & 31;
    double v = 0.0;
    if (e < rk && a < 31) {
      const int src = sm.order[e];
      v = sm.V[a][src] * sqrt(sm.A[src][src]);
    }
    sm.Xs[a][e] = v;
  }
  __syncthreads();

  // ---- sqrtC_lat row (integrals.py:133-138 with eigE = q_size column, T = R_lat U)
  const double qsl = sm.qs[l1];
  {
    const double *H = tab + SPB_TAB_LAT_H + (size_t)tid * 32;
    double out[32];
#pragma unroll
    for (int e = 0; e < 32; ++e) out[e] = 0.0;
    for (int a = 0; a < 31; ++a) {
      const double h = H[a];
#pragma unroll
      for (int e = 0; e < 32; ++e) out[e] = fma(h, sm.Xs[a][e], out[e]);
    }
    double *dst = p.S_lat + ((size_t)b * 256 + tid) * 32;
#pragma unroll
    for (int e = 0; e < 32; ++e) dst[e] = bad ? NAN : qsl * out[e];
  }

  // ---- first moments (integrals.py:126-131): latitude then longitude
  {
    const int w = 2 * l1 + 1;
    const double *R0 = tab + SPB_TAB_LAT_R0 + (size_t)(l1 * (2 * l1 - 1) * (2 * l1 + 1)) / 3 +
                       (size_t)(m1 + l1) * w;
    double acc = 0.0;
    for (int k = 0; k < w; ++k) {
      const int I = 2 * l1 - k;
      double ql = 0.0;
      if (!(k & 1)) ql = ldexp(sm.tt[k >> 1][I >> 1], -l1);
      acc = fma(R0[k], ql, acc);
    }
    sm.m1lat[tid] = qsl * acc;
  }
  __syncthreads();
  {
    const int w = 2 * l1 + 1;
    const double *T1 = tab + SPB_TAB_LON_T1 + (size_t)(l1 * (2 * l1 - 1) * (2 * l1 + 1)) / 3 +
                       (size_t)(m1 + l1) * w;
    double acc = 0.0;
    for (int m = 0; m < w; ++m) acc = fma(T1[m], sm.m1lat[l1 * l1 + m], acc);
    const double pi = 3.14159265358979323846;
    p.mom1[(size_t)b * 256 + tid] = bad ? NAN : acc;
    p.mean_ylm[(size_t)b * 256 + tid] = bad ? NAN : (pi * cc * nn) * acc;  // contrast.py:22
  }
  if (tid == 0) {
    const double pi = 3.14159265358979323846;
    p.scale[b] = (pi * cc) * (pi * cc) * nn;  // contrast.py:23-25
    p.rkeep[b] = rk;
    int flag = 0;
    if (bad) flag |= SPB_INFO_BOUNDS;
    if (!converged) flag |= SPB_INFO_EIG_NOCONV;
    p.info[b] = flag;
  }
}

// ------------------------------------------------------------------------------------------
// K2: X[b][(l,m')][e2][e] = sum_m T_lon[l][m'][e2][m] * S_lat[b][(l,m)][e]
// grid.x = (l, 256-row chunk) work list, grid.y = groups of 32 samples
// ------------------------------------------------------------------------------------------
struct K2Item {
  int l, row0, nrows, toff;
};
__constant__ K2Item k2_items[64];

struct K2Params {
  const double *tab;
  const double *S_lat;  // (B,256,32)
  const int *rkeep;
  double *X;  // (B,256,31*32)
  int B;
};

constexpr int K2_GROUP = 32;
constexpr int K2_NRP = 256;  // row pitch of the transposed T tile

// Thread tile: 4 rows x 4 kept eigen-columns (16 accumulators).  Per m the thread reads 4 T values
// (one 32-byte run of the TRANSPOSED tile) and 4 S values and issues 16 FMAs, i.e. one shared-memory
// wavefront per ~4 FMAs instead of one per FMA (the row-per-thread form was LSU-bound).  Only
// the first rk4 = 4 ceil(rkeep / 4) columns are computed and written; the SYRK never reads the rest.
__global__ void __launch_bounds__(256, 2) moments_k2(K2Params p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double *Tt = reinterpret_cast<double *>(smem_raw);    // [w][K2_NRP]  T_lon tile, transposed
  const K2Item it = k2_items[blockIdx.x];
  const int w = 2 * it.l + 1;
  double *Ssh = Tt + 31 * K2_NRP;                        // [2][31][32]
  const int tid = threadIdx.x;
  const double *Tl = p.tab + SPB_TAB_LON_T + it.toff + (size_t)it.row0 * w;
  const int nq = (it.nrows + 3) >> 2;
  for (int idx = tid; idx < w * K2_NRP; idx += 256) {
    const int m = idx >> 8, r = idx & (K2_NRP - 1);
    Tt[idx] = (r < it.nrows) ? Tl[(size_t)r * w + m] : 0.0;
  }
  const int b0 = blockIdx.y * K2_GROUP;
  const int b1 = min(p.B, b0 + K2_GROUP);
  for (int b = b0; b < b1; ++b) {
    double *Sb = Ssh + ((b - b0) & 1) * (31 * 32);
    const double *src = p.S_lat + ((size_t)b * 256 + it.l * it.l) * 32;
    for (int idx = tid; idx < w * 32; idx += 256) Sb[idx] = src[idx];
    __syncthreads();
    const int ng = (p.rkeep[b] + 3) >> 2;   // groups of 4 kept columns
    double *Xb = p.X + (size_t)b * (256 * 992) + ((size_t)it.l * it.l * 31 + it.row0) * 32;
    for (int tile = tid; tile < nq * ng; tile += 256) {
      const int rq = tile / ng, eg = tile - rq * ng;
      double acc[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[i][e] = 0.0;
      const double *tp = Tt + 4 * rq;
      const double *sp = Sb + 4 * eg;
      for (int m = 0; m < w; ++m) {
        const double2 t01 = *reinterpret_cast<const double2 *>(tp + m * K2_NRP);
        const double2 t23 = *reinterpret_cast<const double2 *>(tp + m * K2_NRP + 2);
        const double2 s01 = *reinterpret_cast<const double2 *>(sp + m * 32);
        const double2 s23 = *reinterpret_cast<const double2 *>(sp + m * 32 + 2);
        const double tv[4] = {t01.x, t01.y, t23.x, t23.y};
        const double sv[4] = {s01.x, s01.y, s23.x, s23.y};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int e = 0; e < 4; ++e) acc[i][e] = fma(tv[i], sv[e], acc[i][e]);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = 4 * rq + i;
        if (r < it.nrows) {
          double2 *dst = reinterpret_cast<double2 *>(Xb + (size_t)r * 32 + 4 * eg);
          dst[0] = make_double2(acc[i][0], acc[i][1]);
          dst[1] = make_double2(acc[i][2], acc[i][3]);
        }
      }
    }
    // the double-buffered Ssh makes one barrier per sample sufficient
  }
}

int k2_upload_items(int *nitems_out) {
  static bool done = false;
  static int nitems = 0;
  if (!done) {
    K2Item items[64];
    int toff = 0;
    for (int l = 0; l <= SPB_LMAX; ++l) {
      const int w = 2 * l + 1, rows = w * 31;
      for (int r0 = 0; r0 < rows; r0 += 256) {
        items[nitems].l = l;
        items[nitems].row0 = r0;
        items[nitems].nrows = (rows - r0 < 256) ? rows - r0 : 256;
        items[nitems].toff = toff;
        ++nitems;
      }
      toff += w * 31 * w;
    }
    if (cudaMemcpyToSymbol(k2_items, items, sizeof(K2Item) * nitems) != cudaSuccess) return 1;
    done = true;
  }
  *nitems_out = nitems;
  return 0;
}

constexpr int MOM_CHUNK = 1024;  // samples per pass (bounds the sqrtC_lon workspace to 2 GB)

struct MomWs {
  double *mom1, *S_lat, *scale, *X;
  int *rkeep;
};

size_t mom_ws_layout(int B, unsigned char *base, MomWs *ws) {
  const int Bc = B < MOM_CHUNK ? B : MOM_CHUNK;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off += (bytes + 255) & ~(size_t)255;
    return o;
  };
  size_t o_mom1 = take((size_t)B * 256 * 8);
  size_t o_S = take((size_t)B * 256 * 32 * 8);
  size_t o_scale = take((size_t)B * 8);
  size_t o_rk = take((size_t)B * 4);
  size_t o_X = take((size_t)Bc * 256 * 992 * 8);
  if (ws) {
    ws->mom1 = reinterpret_cast<double *>(base + o_mom1);
    ws->S_lat = reinterpret_cast<double *>(base + o_S);
    ws->scale = reinterpret_cast<double *>(base + o_scale);
    ws->rkeep = reinterpret_cast<int *>(base + o_rk);
    ws->X = reinterpret_cast<double *>(base + o_X);
  }
  return off;
}

// gauss2beta, latitude.py:62-77
__global__ void gauss2beta_kernel(int B, const double *mu, const double *sigma, double *a,
                                  double *b) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B) return;
  const double ang = 3.14159265358979323846 / 180.0;
  const double m = mu[i] * ang;
  const double v = (sigma[i] * ang) * (sigma[i] * ang);
  const double c1 = cos(m), c2 = cos(2 * m), c3 = cos(3 * m);
  const double ch = cos(0.5 * m);
  const double term = 1.0 / (16 * v * (ch * ch * ch * ch));
  const double alpha = (2 + 4 * v + (3 + 8 * v) * c1 + 2 * c2 + c3) * term;
  const double beta = (c1 + 2 * v * (3 + c2) - c3) * term;
  a[i] = log(alpha) / 10.0;
  b[i] = fmax(0.0, (log(beta) - log(0.5)) / (10.0 - log(0.5)));
}

// log |d(a, b) / d(mu, sigma)|, latitude.py:221-241 (mode and width of the latitude pdf from the
// Beta shape parameters) and 281-316; -inf when sigma > sigma_max.
__global__ void log_jac_kernel(int B, const double *a, const double *b, double sigma_max_rad,
                               double *out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B) return;
  double aa = a[i], bb = b[i];
  if (aa < 1e-12) aa = 1e-12;  // latitude.py:178-182 (abmin)
  if (bb < 1e-12) bb = 1e-12;
  const double al = exp(aa * 10.0);
  const double be = exp(log(0.5) + bb * (10.0 - log(0.5)));
  double term = 4 * al * al - 8 * al - 6 * be + 4 * al * be + be * be + 5;
  const double mu = 2 * atan(sqrt(2 * al + be - 2 - sqrt(term)));
  const double cm = cos(mu), sn = sin(mu);
  term = 1 - al + be + (be - 1) * cm + (al - 1) / (cm * cm);
  const double sigma = sqrt(sn * sn / term);
  const double c1 = 1 + cm, s2 = sin(2 * mu);
  const double num = al * be * (c1 * c1 * c1) * (s2 * s2 * s2);
  const double f1 = -3 + 2 * al + be + (-1 + 2 * al + be) * cm;
  const double f2 = 2 * (-1 + al + be) + 3 * (-1 + be) * cm - 2 * (-1 + al - be) * cos(2 * mu) +
                    (-1 + be) * cos(3 * mu);
  const double lj = log(fabs(num / (sigma * f1 * (f2 * f2))));
  out[i] = (sigma > sigma_max_rad) ? -INFINITY : lj;
}

}  // namespace

extern "C" int spb_log_jac(spb_context *ctx, int B, const double *a, const double *b,
                           double sigma_max_deg, double *log_jac, void *stream) {
  SPB_REQUIRE(ctx != nullptr && B > 0, "log_jac: bad arguments");
  SPB_CHECK_CUDA(cudaSetDevice(ctx->device));
  log_jac_kernel<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
      B, a, b, sigma_max_deg * 3.14159265358979323846 / 180.0, log_jac);
  SPB_LAUNCH_CHECK(ctx);
  return 0;
}

extern "C" int spb_gauss2beta(spb_context *ctx, int B, const double *mu_deg,
                              const double *sigma_deg, double *a, double *b, void *stream) {
  SPB_REQUIRE(ctx != nullptr && B > 0, "gauss2beta: bad arguments");
  SPB_CHECK_CUDA(cudaSetDevice(ctx->device));
  gauss2beta_kernel<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(B, mu_deg, sigma_deg, a, b);
  SPB_LAUNCH_CHECK(ctx);
  return 0;
}

extern "C" size_t spb_ylm_moments_workspace_bytes(const spb_context *ctx, int B) {
  (void)ctx;
  return mom_ws_layout(B, nullptr, nullptr);
}

extern "C" int spb_ylm_moments(spb_context *ctx, int B, const double *r_deg, const double *a,
                               const double *b, const double *c, const double *n, double *mean_ylm,
                               double *cov_ylm, int32_t *info, void *workspace,
                               size_t workspace_bytes, void *stream_) {
  SPB_REQUIRE(ctx != nullptr && B > 0, "ylm_moments: bad arguments");
  SPB_REQUIRE(ctx->tables_count == SPB_TAB_TOTAL, "ylm_moments: context has no constant tables");
  SPB_REQUIRE(workspace_bytes >= mom_ws_layout(B, nullptr, nullptr) && workspace != nullptr,
              "ylm_moments: workspace too small");
  SPB_REQUIRE(((uintptr_t)workspace % 256) == 0, "ylm_moments: workspace must be 256-byte aligned");
  cudaStream_t stream = (cudaStream_t)stream_;
  SPB_CHECK_CUDA(cudaSetDevice(ctx->device));
  MomWs ws;
  mom_ws_layout(B, reinterpret_cast<unsigned char *>(workspace), &ws);

  static bool attr1 = false;
  if (!attr1) {
    SPB_CHECK_CUDA(cudaFuncSetAttribute(moments_k1, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)sizeof(K1Smem)));
    SPB_CHECK_CUDA(cudaFuncSetAttribute(moments_k2, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)((K2_NRP * 31 + 2 * 31 * 32) * sizeof(double))));
    attr1 = true;
  }
  K1Params p1;
  p1.r_deg = r_deg;
  p1.a = a;
  p1.b = b;
  p1.c = c;
  p1.n = n;
  p1.B = B;
  p1.tab = ctx->d_tables;
  p1.mom1 = ws.mom1;
  p1.mean_ylm = mean_ylm;
  p1.S_lat = ws.S_lat;
  p1.scale = ws.scale;
  p1.rkeep = ws.rkeep;
  p1.info = info;
  moments_k1<<<B, NT1, sizeof(K1Smem), stream>>>(p1);
  SPB_LAUNCH_CHECK(ctx);

  int nitems = 0;
  SPB_REQUIRE(k2_upload_items(&nitems) == 0, "ylm_moments: constant upload failed");
  for (int b0 = 0; b0 < B; b0 += MOM_CHUNK) {
    const int Bc = (B - b0 < MOM_CHUNK) ? B - b0 : MOM_CHUNK;
    K2Params p2;
    p2.tab = ctx->d_tables;
    p2.S_lat = ws.S_lat + (size_t)b0 * 256 * 32;
    p2.rkeep = ws.rkeep + b0;
    p2.X = ws.X;
    p2.B = Bc;
    dim3 grid2(nitems, (Bc + K2_GROUP - 1) / K2_GROUP);
    moments_k2<<<grid2, 256, (K2_NRP * 31 + 2 * 31 * 32) * sizeof(double), stream>>>(p2);
    SPB_LAUNCH_CHECK(ctx);

    gnt::Desc d = {};
    d.A = ws.X;
    d.strideA = 256 * 992;
    d.lda = 992;
    d.Bm = ws.X;
    d.strideB = 256 * 992;
    d.ldb = 992;
    d.C = cov_ylm + (size_t)b0 * 65536;
    d.strideC = 65536;
    d.ldc = 256;
    d.M = 256;
    d.N = 256;
    d.K = 992;
    d.batch = Bc;
    d.ksplit = 1;
    d.strideSplit = 0;
    d.lower_only = 1;
    d.scale = ws.scale + b0;
    d.vec = ws.mom1 + (size_t)b0 * 256;
    d.strideVec = 256;
    d.diag = ctx->d_tables + SPB_TAB_LAMBDA;
    d.rkeep = ws.rkeep + b0;
    d.alpha = 1.0;
    int st = gnt::launch<gnt::EPI_SYRK_COV>(ctx, d, stream);
    if (st) return st;
  }
  return 0;
}
