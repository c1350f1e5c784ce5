// Batched Cholesky -> lnlike with the trailing updates on the INT8 tensor cores (tcgen05 + TMEM).
// Included by potrf.cu inside its anonymous namespace (shares RowMap, init_acc, potf2_regs, trsm_warp).
//
// FP64 has no tcgen05 form, but the panel update  P = K - L L^T  is a sum of products of numbers whose
// magnitude is known in advance (|L_ik| <= sqrt(K_ii)), so it can be evaluated EXACTLY in integer
// arithmetic (Ozaki-style error-free slicing; accuracy study: oracle/study/ozaki_cholesky.py,
// micro-benchmark: scripts/micro/i8_emul_micro.cu):
//
//   * every stored row of L (and of the appended right-hand sides y = L^-1 r) is scaled by a power of
//     two 2^-e_i fixed before the factorisation starts, rounded to S x RB bits and cut into S signed
//     RB-bit digits  L_ik 2^-e_i = sum_s d_s(i,k) 2^{-RB (s+1)};  the digits are stored as S int8 PLANES
//     next to K.  Default: S = 7 planes of balanced RB = 8-bit digits (55 bits per row, 28 plane
//     products); S = 8 x RB = 7 (56 bits, 36 products) and S = 7 x RB = 7 (49 bits) are kept;
//   * for a 128 x 64 tile of the panel, one thread issues  tcgen05.mma.kind::i8  products of digit plane
//     A_s (128 x 32 bytes of k) with the concatenated planes B_0 .. B_{D-s} (N up to 256) into TMEM:
//     pairs with s + t = d accumulate EXACTLY (int32) in the 64-column block d, all 512 columns of an SM
//     hold the D + 1 = S blocks of the tile; operands arrive by TMA (32-byte swizzle) in a shared
//     memory ring; the tensor core runs from smem descriptors, so no warp touches the operands;
//   * the 8 compute warps read the S blocks back with fragment-shaped tcgen05.ld (16x256b: the same
//     (row g, columns 2 tg) layout as the DMMA accumulators), combine them in fp64
//     (Horner in 2^-RB), apply the row scales and subtract from the K values of the tile (copied to
//     shared memory by cp.async during the previous tile when nt <= 1536, loaded directly otherwise);
//     from there on the tile is the round-1 kernel's: in-register potf2 of the diagonal block, in-register
//     TRSM on DMMA, |y|^2 and log-determinant reductions;
//   * the new rows of L are cut into digit planes straight from the accumulator registers and stored;
//     the producer thread waits for exactly the tiles a k-chunk needs (a monotone counter of stored
//     tiles), so the MMA stream of the next tile -- and of the next PANEL, up to its last two k-chunks --
//     runs while the warps are still in the TRSM / stores of the current one (look-ahead for free).
//
// With 55 / 56 bits relative to the row maximum the lnlike differences to the DMMA kernel are at the level of
// fp64 rounding noise (<= 3.8e-12 / 1.2e-12 relative on the bench covariances, cond 1e3 .. 2e7).  K itself is
// only READ.
//
// Safety net: a digit that does not fit (a row bound violated: cannot happen for rows of L, and for
// the right-hand-side rows only if K' - diag were indefinite by more than 15/16 of diag) sets
// SPB_INFO_I8_RANGE for the matrix; the launcher then re-runs exactly those matrices through the DMMA
// kernel (K is untouched, so nothing is lost).

#ifdef SPB_POTRF_PROF
#define I8_PROF_DECL(row_) unsigned long long _it = clock64(); const int _ir = (row_)
#define I8_PROF(sm, k) do { if (_ir >= 0) { unsigned long long _n = clock64(); (sm).prof[_ir][k] += _n - _it; _it = _n; } } while (0)
#else
#define I8_PROF_DECL(row_)
#define I8_PROF(sm, k)
#endif

constexpr int I8_KCH = 32;           // bytes of k per ring stage = K of one tcgen05.mma.kind::i8
// digit width RB (template parameter): 7 -> digits in [-64, 63] except the leading one, which uses the whole
// int8 range (|x| <= 0.99); 8 -> balanced base-256 digits in [-128, 127], all planes use the whole range, the
// rows are scaled to |x| <= 1/2: 7 planes then carry 55 bits -- one bit less than 8 planes of 7 bits with 28
// instead of 36 plane products and 7/8 of the operand traffic
constexpr int I8_NCT = 256;          // compute threads (warps 0-7)
#ifndef I8_ROLE_WARP0
#define I8_ROLE_WARP0 8
#endif
constexpr int I8_PRODUCER_WARP = I8_ROLE_WARP0;        // TMA producer (one thread)
constexpr int I8_ISSUER_WARP = I8_ROLE_WARP0 + 1;      // MMA issuer (one thread)
constexpr int I8_NTHREADS = 32 * (I8_ROLE_WARP0 + 2);
constexpr int I8_MAXP = 256;         // panels per matrix (nt <= 16384)
constexpr int I8_TM = 128;

struct I8Params {
  uint8_t *Q;            // (B, S, NR, LDQ) digit planes
  double *E;             // (B, NR) row scales 2^e
  long long strideQ;     // bytes per matrix
  int NR, LDQ;
  int store_factor;      // also write L (fp64) over K like the DMMA kernel
  double lambda_min;     // > 0: lower bound of the smallest eigenvalue of K' (else min(aff.diag))
  int gate;              // pause the MMA stream while warps are in the shuffle / shared-memory bound phases
};

// Virtual rows of a panel for this kernel.  A tile's A operand must be ONE box of 128 consecutive plane
// rows, i.e. virtual row v of the panel at column c0 has to be plane row c0 + v; the last panel of a
// matrix whose size is not a multiple of 64 (identity-padded diagonal block, right-hand sides at virtual
// row 64 + m) fixes the plane rows of the right-hand sides at n64 + m, n64 = round_up(nt, 64).  Two layouts:
//   * "gap" (M <= n64 - nt, the bench case): the right-hand-side planes are kept TWICE, at rows nt + m
//     (inside the unused gap nt .. n64 - 1, read by the full panels, where they follow the last matrix row
//     exactly as in the FP64 kernel: no extra tile) and at rows n64 + m (read by the last panel);
//   * otherwise: once, at n64 + m, and every panel carries the n64 - nt unused rows as KIND_NONE rows
//     between the matrix rows and the right-hand sides (nbel counts them).
struct RowMapI8 {
  double *Kb, *Rb;
  int n, M, ld, ldr, c0, nbel;   // nbel: virtual rows between the diagonal block and the right-hand sides
  __device__ __forceinline__ double *row(int v, int &kind) const {
    if (v < NB) {
      const int r = c0 + v;
      if (r < n) {
        kind = KIND_DIAG;
        return Kb + (size_t)r * ld;
      }
      kind = KIND_PAD;
      return nullptr;
    }
    int w = v - NB;
    if (w < nbel) {
      const int r = c0 + NB + w;
      if (r < n) {
        kind = KIND_BELOW;
        return Kb + (size_t)r * ld;
      }
      kind = KIND_NONE;
      return nullptr;
    }
    w -= nbel;
    if (w < M) {
      kind = KIND_RHS;
      return Rb + (size_t)w * ldr;
    }
    kind = KIND_NONE;
    return nullptr;
  }
};

template <int S_, int STAGES_, bool KPRE_ = false>
struct SmemI8 {
  using G = Geo<128>;
  static constexpr int S = S_;
  static constexpr int STAGES = STAGES_;
  static constexpr int NTHREADS = I8_NCT;
  uint8_t A[STAGES_][S_][I8_TM][I8_KCH];   // 4 KB per plane
  uint8_t B[STAGES_][S_][NB][I8_KCH];      // 2 KB per plane; planes t0..t1 form one N = 64 (t1 - t0 + 1) operand
  // KPRE: the 128 x 64 K tile of the NEXT tile, copied by cp.async while the current one is processed; chunk c
  // (of 16) of thread t at [(c * 256 + t)]: every thread reads back what it copied, conflict-free both ways
  double2 Kpre[KPRE_ ? 16 * I8_NCT : 1];
  LdStore<false> Ld;
  double Dv[NB][DS];
  double red[I8_NCT / 32];
  double dpiv[NB];
  double af[4];
  double dgmin;
  uint64_t full[STAGES_], empty[STAGES_], tmem_full, tmem_empty;
  unsigned first[I8_MAXP + 1];   // sequence number of the first tile of every panel of this matrix
  uint32_t tmem_base;
  volatile unsigned stored;      // tiles whose planes are in global memory (monotone over the launch)
  int quiet;                     // number of warps inside potf2 / TRSM (the issuer holds back while > 0)
  int bad;
  int bad_range;
  int next_item;
#ifdef SPB_POTRF_PROF
  unsigned long long prof[2][16];
#endif
};

__device__ __forceinline__ void tma_load_4d(void *sdst, const void *tmap, uint64_t *bar, int c0, int c1,
                                            int c2, int c3) {
  unsigned s = (unsigned)__cvta_generic_to_shared(sdst);
  unsigned b = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];\n" ::"r"(s),
      "l"(tmap), "r"(b), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ uint64_t i8_smem_desc(const void *p) {
  // K-major operand, 32-byte swizzle: rows of 32 B, 8-row groups 256 B apart, descriptor version 1
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
  uint64_t d = (uint64_t)((a >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(256 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)6 << 61;
  return d;
}
__device__ __forceinline__ uint32_t i8_idesc(int n) {
  // D = s32, A = B = signed 8 bit, both K-major, M = 128
  return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(I8_TM >> 4) << 24);
}
__device__ __forceinline__ void i8_mma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc,
                                       uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void i8_commit(uint64_t *bar) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(bar);
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(a)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
}
// 16 lanes x 256 bit, x2 = 16 columns: for each 8-column block b, v[4 b], v[4 b + 1] = (row T/4,
// columns 2 (T%4), +1) and v[4 b + 2], v[4 b + 3] = (row T/4 + 8, same columns): the DMMA C layout
__device__ __forceinline__ void tmem_ld_frag16(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }
// Waits of the two single-thread roles: poll with a sleep in between.  A tight try_wait loop keeps the
// warp permanently eligible and its barrier probes share the LSU path with the SHFL / LDS of the compute
// warp on the same scheduler (measured: the shuffle-bound 8x8 pivot chain ran 3.5x slower next to one).
__device__ __forceinline__ void mbar_wait_sleep(uint64_t *bar, unsigned parity, unsigned ns) {
  while (!mbar_test(bar, parity)) __nanosleep(ns);
}
// The tensor core reads its operands through the same shared-memory datapath the warps' SHFL / LDS use:
// while the MMA stream runs, the shuffle-bound pivot chain of potf2 and the in-register TRSM run 2.5x
// slower (measured with phase timers).  Warps announce those phases; the issuer holds back meanwhile.
__device__ __forceinline__ void quiet_enter(int *q, int lane, bool on) {
  if (on && lane == 0) atomicAdd(q, 1);
}
__device__ __forceinline__ void quiet_leave(int *q, int lane, bool on) {
  if (on && lane == 0) atomicSub(q, 1);
}
__device__ __forceinline__ void cbar256() { asm volatile("bar.sync 2, 256;\n" ::: "memory"); }

// smallest power of two 2^e with 1.0101 x < 2^e  (x / 2^e <= 0.99: the leading digit fits int8)
__device__ __forceinline__ double pow2_above(double x) {
  if (!(x > 0.0) || !(x < 1e300)) return 1.0;
  int ex;
  (void)frexp(x * 1.0101, &ex);
  return ldexp(1.0, ex);
}

// The fused affine map (spb_affine) of one accumulator row that holds RAW values of a full panel (the
// arithmetic half of init_acc; used when the values come from the cp.async prefetch buffer).
template <class SM>
__device__ __forceinline__ void i8_apply_aff(const SM &sm, const RowMapI8 &rm, const AffRow &af, int v, int c0,
                                             int tg, double (&accrow)[8][2]) {
  int kind;
  (void)rm.row(v, kind);
  if (kind != KIND_DIAG && kind != KIND_BELOW) return;
  const int gi = c0 + v;
  double Ci = sm.af[3], Di = 0.0;
  if (af.norm) {
    const double qi = af.q[gi];
    const double a = sm.af[1] * (1.0 - qi);
    Ci += a;
    Di = a + sm.af[2] * qi;
  }
  aff_full_row(sm, af, kind == KIND_DIAG, gi, v, c0, tg, Ci, Di, accrow);
}

template <class SM>
__device__ __forceinline__ double block_sum_i8(SM &sm, double v) {
  const int tid = threadIdx.x, pw = tid >> 5, lane = tid & 31;
  v = warp_sum(v);
  cbar256();
  if (lane == 0) sm.red[pw] = v;
  cbar256();
  double t = 0.0;
#pragma unroll
  for (int w = 0; w < I8_NCT / 32; ++w) t += sm.red[w];
  cbar256();
  return t;
}
template <class SM>
__device__ __forceinline__ double block_min_i8(SM &sm, double v) {
  const int tid = threadIdx.x, pw = tid >> 5, lane = tid & 31;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
  cbar256();
  if (lane == 0) sm.red[pw] = v;
  cbar256();
  double t = sm.red[0];
#pragma unroll
  for (int w = 1; w < I8_NCT / 32; ++w) t = fmin(t, sm.red[w]);
  cbar256();
  return t;
}

// sum_{j < S} 64 128^j: makes every base-128 digit of a signed S-digit number non-negative
__host__ __device__ constexpr long long i8_bias(int S, int RB) {
  long long b = 0;
  for (int j = 0; j < S; ++j) b = (b << RB) + (1ll << (RB - 1));
  return b;
}

// number of 128-row tiles of the panel starting at column c0 (NR = nt + M virtual rows in total)
__device__ __forceinline__ int i8_nbel(int n, int n64, bool gap, int c0) {
  return gap ? ((c0 + NB <= n) ? n - c0 - NB : 0) : n64 - c0 - NB;
}
__device__ __forceinline__ int i8_nvirt(int n, int n64, int M, bool gap, int c0) {
  return NB + i8_nbel(n, n64, gap, c0) + M;
}
__device__ __forceinline__ int i8_ntiles(int nvirt) { return (nvirt + I8_TM - 1) / I8_TM; }

template <int S, int STAGES, int RB, bool KPRE>
__global__ void __launch_bounds__(I8_NTHREADS, 1)
    potrf_i8_kernel(PotrfParams p, I8Params ip, const __grid_constant__ CUtensorMap tmA,
                    const __grid_constant__ CUtensorMap tmB) {
  constexpr int D = S - 1;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  using SM = SmemI8<S, STAGES, KPRE>;
  SM &sm = *reinterpret_cast<SM *>(smem_raw);
  const int tid = threadIdx.x, pw = tid >> 5, lane = tid & 31, g = lane >> 2, tg = lane & 3;
  // TMEM lanes a warp may read: 32 (pw % 4) .. + 31; the fragment-shaped loads take 16 of them, so compute
  // warp pw owns tile rows 32 (pw % 4) + 16 (pw / 4) .. + 15, i.e. it plays LOGICAL warp 2 (pw % 4) + pw / 4
  // of the 8-warp geometry (rows 16 warp .. 16 warp + 15)
  const int warp = 2 * (pw & 3) + ((pw >> 2) & 1);
  const int NR = ip.NR;

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&sm.full[s], 1);
      mbar_init(&sm.empty[s], 1);
    }
    mbar_init(&sm.tmem_full, 1);
    mbar_init(&sm.tmem_empty, I8_NCT / 32);
    sm.stored = 0;
    sm.quiet = 0;
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
#ifdef SPB_POTRF_PROF
  if (tid < 32) sm.prof[tid / 16][tid % 16] = 0;
#endif
  if (pw == I8_ISSUER_WARP) {
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(&sm.tmem_base);
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(a), "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sm.tmem_base;

  unsigned xchunk = 0;     // producer / issuer: running k-chunk count (ring slot, barrier phase)
  unsigned tq = 0;         // issuer / compute: running count of tiles that went through TMEM
  unsigned seq_base = 0;   // sequence number of the first tile of the current matrix

  for (int item = blockIdx.x; item < p.B;) {
    // ---------------------------------------------------------------- per-matrix set-up
    RowMapI8 rm;
    rm.n = p.n;
    rm.ld = p.ld;
    rm.ldr = p.ldr;
    rm.Kb = p.K + (size_t)item * p.strideK;
    rm.Rb = p.R ? p.R + (size_t)item * p.strideR : nullptr;
    rm.M = p.R ? p.M : 0;
    const int n64 = (p.n + NB - 1) & ~(NB - 1);
    const int M_ = p.R ? p.M : 0;
    const bool gap = (M_ <= n64 - p.n);   // also true when M = 0; false when nt is a multiple of 64 and M > 0
    double *quad_out = p.quad ? p.quad + (size_t)item * p.M : nullptr;
    AffRow af;
    af.q = nullptr;
    af.dg = nullptr;
    af.dg_vec = 0;
    af.norm = false;
    if (p.aff_on) {
      af.norm = (p.aff.scal != nullptr) && (p.aff.q != nullptr);
      if (af.norm) af.q = p.aff.q + (size_t)item * p.n;
      if (p.aff.diag) {
        af.dg = p.aff.diag + (size_t)item * p.aff.diag_stride;
        af.dg_vec = (p.aff.diag_kind == 1);
      }
    }
    double *Eb = ip.E + (size_t)item * NR;
    uint8_t *Qb = ip.Q + (size_t)item * ip.strideQ;
    if (tid == 0) {
      sm.af[0] = af.norm ? p.aff.scal[4 * (size_t)item + 0] : 1.0;
      sm.af[1] = af.norm ? p.aff.scal[4 * (size_t)item + 1] : 0.0;
      sm.af[2] = af.norm ? p.aff.scal[4 * (size_t)item + 2] : 0.0;
      sm.af[3] = (p.aff_on && p.aff.offset) ? p.aff.offset[(size_t)item * p.aff.offset_stride] : 0.0;
      sm.bad = 0;
      sm.bad_range = 0;
      unsigned f = seq_base;
      int j = 0;
      for (int c0 = 0; c0 < p.n; c0 += NB, ++j) {
        sm.first[j] = f;
        f += (unsigned)i8_ntiles(i8_nvirt(p.n, n64, M_, gap, c0));
      }
      sm.first[j] = f;
    }
    __syncthreads();
    const int npanels = (p.n + NB - 1) / NB;
    const unsigned seq_end = sm.first[npanels];

    if (pw < 8) {
      // ============================================================== COMPUTE WARPS
      I8_PROF_DECL(tid == 0 ? 0 : -1);
      if (quad_out)
        for (int m = tid; m < rm.M; m += I8_NCT) quad_out[m] = 0.0;
      // row scales: 2^e_i > 1.0101 sqrt(K'_ii) for the rows of L;  for the right-hand-side rows
      // |y_k| <= |y| <= |r| / sqrt(lambda_min(K')) <= sqrt(nt) max|r| / sqrt(min diag) (x 4 of margin)
      double dgm = 1e300;
      if (ip.lambda_min > 0.0) {
        dgm = ip.lambda_min;
      } else if (af.dg) {
        if (af.dg_vec) {
          for (int r = tid; r < p.n; r += I8_NCT) dgm = fmin(dgm, af.dg[r]);
        } else {
          dgm = af.dg[0];
        }
      }
      dgm = block_min_i8(sm, dgm);
      for (int r = tid; r < p.n; r += I8_NCT) {
        double kd = rm.Kb[(size_t)r * p.ld + r];
        if (p.aff_on) {
          double Ci = sm.af[3], Di = 0.0;
          if (af.norm) {
            const double qi = af.q[r];
            const double a = sm.af[1] * (1.0 - qi);
            Ci += a;
            Di = a + sm.af[2] * qi;
          }
          kd = aff_apply(sm, af, kd, Ci, Di, true, r, r);
        }
        Eb[r] = pow2_above(sqrt(kd)) * (RB == 8 ? 2.0 : 1.0);
      }
      for (int m = pw; m < rm.M; m += I8_NCT / 32) {
        const double *rr = rm.Rb + (size_t)m * p.ldr;
        double mx = 0.0;
        for (int k = lane; k < p.n; k += 32) mx = fmax(mx, fabs(rr[k]));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        if (lane == 0)
          Eb[n64 + m] = pow2_above(4.0 * sqrt((double)p.n) * mx * rsqrt(dgm)) * (RB == 8 ? 2.0 : 1.0);
      }
      for (int r = p.n + tid; r < n64; r += I8_NCT) Eb[r] = 1.0;
      __threadfence_block();
      cbar256();
      I8_PROF(sm, 8);

      double logdet_part = 0.0, quad_part = 0.0;
      double acc[2][8][2];
      bool bad_range = false;
      bool have_pre = false;   // KPRE: the prefetch buffer holds this tile's raw values (block-uniform)
      unsigned seq = seq_base;
      for (int c0 = 0; c0 < p.n; c0 += NB) {
        rm.c0 = c0;
        rm.nbel = i8_nbel(p.n, n64, gap, c0);
        const int nvirt = NB + rm.nbel + rm.M;
        const bool full_panel = (c0 + NB <= p.n);
        const bool need_planes = (c0 + NB < p.n);   // a later panel will read these columns
        const int nchunks = c0 / I8_KCH;
        for (int v0 = 0; v0 < nvirt; v0 += I8_TM, ++seq) {
          const bool diag_tile = (v0 == 0);
          const int vr = v0 + warp * 16;            // first tile row of this warp
          if (KPRE && have_pre) {
            // the raw K (or right-hand-side) values of this tile were copied to shared memory during the
            // previous tile: pick them up and apply the affine map
            cp_async_wait<0>();
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
              for (int nt = 0; nt < 8; ++nt) {
                const double2 kv = sm.Kpre[(mt * 8 + nt) * I8_NCT + tid];
                acc[mt][nt][0] = kv.x;
                acc[mt][nt][1] = kv.y;
              }
              if (p.aff_on) i8_apply_aff(sm, rm, af, vr + mt * 8 + g, c0, tg, acc[mt]);
            }
          } else {
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
              init_acc(sm, rm, af, p.aff_on != 0, vr + mt * 8 + g, c0, tg, full_panel, acc[mt]);
          }
          {
            // the K rows of the NEXT tile are read exactly once, from HBM, and their latency (3 us per tile
            // when loaded at the top of the tile) is the largest single item of the per-tile chain: with
            // KPRE they are copied now by cp.async (per thread: its own 16 x 16 bytes) and waited for at the
            // top of the next tile; without, at least requested into L2
            RowMapI8 rn = rm;
            int v0n = v0 + I8_TM;
            if (v0n >= nvirt) {
              v0n = 0;
              rn.c0 = c0 + NB;
              rn.nbel = i8_nbel(p.n, n64, gap, rn.c0);
            }
            have_pre = false;
            if (rn.c0 < p.n) {
              if (KPRE) {
                if (rn.c0 + NB <= p.n) {      // full panels only (block-uniform)
#pragma unroll
                  for (int mt = 0; mt < 2; ++mt) {
                    int kind;
                    const double *pn = rn.row(v0n + warp * 16 + mt * 8 + g, kind);
                    const int bytes = pn ? 16 : 0;
                    const double *src = (pn ? pn : rn.Kb) + rn.c0 + 2 * tg;
#pragma unroll
                    for (int nt = 0; nt < 8; ++nt)
                      cp_async16(&sm.Kpre[(mt * 8 + nt) * I8_NCT + tid], src + nt * 8, bytes);
                  }
                  cp_async_commit();
                  have_pre = true;
                }
              } else {
#pragma unroll
                for (int mt = 0; mt < 2; ++mt) {
                  int kind;
                  const double *pn = rn.row(v0n + warp * 16 + mt * 8 + g, kind);
                  if (pn != nullptr && rn.c0 + 16 * tg < p.n)
                    asm volatile("prefetch.global.L2 [%0];\n" ::"l"(pn + rn.c0 + 16 * tg));
                }
              }
            }
          }
#ifdef SPB_POTRF_PROF
          if (tid == 0 && c0 == 0) sm.prof[0][9] += clock64() - _it;   // K loads of panel 0 (no MMA / TMA in flight)
#endif
          I8_PROF(sm, 0);
          if (nchunks > 0) {
            // scales of this thread's rows (plane row = c0 + virtual row for matrix and RHS rows alike)
            double si[2];
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
              const int v = vr + mt * 8 + g;
              int kind;
              (void)rm.row(v, kind);
              const int er = (kind == KIND_RHS) ? n64 + (v - NB - rm.nbel) : c0 + v;   // row of the scale table
              const bool valid = (kind == KIND_DIAG || kind == KIND_BELOW || kind == KIND_RHS);
              si[mt] = valid ? -ldexp(Eb[er], -2 * RB) : 0.0;
            }
            mbar_wait_sleep(&sm.tmem_full, tq & 1u, 50);
            tc_fence_after();
            I8_PROF(sm, 1);
            const uint32_t tw = tmem + ((uint32_t)(32 * (pw & 3) + 16 * ((pw >> 2) & 1)) << 16);
#pragma unroll
            for (int cq = 0; cq < 4; ++cq) {
              uint32_t v[S][8];
#pragma unroll
              for (int d = 0; d < S; ++d) tmem_ld_frag16(tw + (uint32_t)(NB * d + 16 * cq), v[d]);
              tmem_ld_wait();
#pragma unroll
              for (int b = 0; b < 2; ++b) {
                const int nt = 2 * cq + b;
                const int gc = c0 + nt * 8 + 2 * tg;
                const double sj0 = (gc < p.n) ? Eb[gc] : 0.0;
                const double sj1 = (gc + 1 < p.n) ? Eb[gc + 1] : 0.0;
#pragma unroll
                for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
                  for (int e = 0; e < 2; ++e) {
                    double t = (double)(int)v[D][4 * b + 2 * mt + e];
#pragma unroll
                    for (int d = D - 1; d >= 0; --d)
                      t = fma(t, RB == 8 ? 0.00390625 : 0.0078125, (double)(int)v[d][4 * b + 2 * mt + e]);
                    acc[mt][nt][e] = fma(si[mt] * (e ? sj1 : sj0), t, acc[mt][nt][e]);
                  }
                }
              }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&sm.tmem_empty);
            ++tq;
            I8_PROF(sm, 2);
          }
          const int lw = warp;
          if (diag_tile) {
            if (lw < 4) {
              quiet_enter(&sm.quiet, lane, (ip.gate & 1) != 0);
#ifdef SPB_POTRF_PROF
              const unsigned long long t8_before = sm.prof[0][11];
#endif
              potf2_regs(sm, acc, lw, lane);
#ifdef SPB_POTRF_PROF
              if (tid == 0 && c0 == 0) sm.prof[0][15] += sm.prof[0][11] - t8_before;   // pivot chains of panel 0
#endif
              quiet_leave(&sm.quiet, lane, (ip.gate & 1) != 0);
            }
            cbar256();   // L_jj and the inverses of its diagonal tiles are in shared memory
            if (tid < min(NB, p.n - c0)) logdet_part += 0.5 * log(sm.dpiv[tid]);
            if (ip.store_factor) {
              for (int idx = tid; idx < NB * NB; idx += I8_NCT) {
                const int i = idx >> 6, j = idx & 63;
                if (j <= i && c0 + i < p.n) rm.Kb[(size_t)(c0 + i) * p.ld + c0 + j] = sm.Ld.at(i, j);
              }
            }
          }
          I8_PROF(sm, 3);
          // a matrix already found not positive definite turns into NaNs from here on: its digits
          // overflow, but there is nothing to rescue by re-running it on the FP64 kernel
          const bool already_bad = (*(volatile int *)&sm.bad) != 0;
          const int vr2 = v0 + lw * 16;             // first tile row of this warp from here on
          if (!(diag_tile && lw < 4) && vr2 < nvirt) {
            quiet_enter(&sm.quiet, lane, (ip.gate & 1) != 0);
            trsm_warp(sm, acc, lane);
            quiet_leave(&sm.quiet, lane, (ip.gate & 1) != 0);
            I8_PROF(sm, 4);
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
              const int v = vr2 + mt * 8 + g;
              int kind;
              double *prow = rm.row(v, kind);
              const bool live = (kind == KIND_BELOW || kind == KIND_RHS);
              // ---- fp64: y rows always (output), rows of L only on request
              double q = 0.0;
              if (live && (kind == KIND_RHS || ip.store_factor)) {
#pragma unroll
                for (int nt = 0; nt < 8; ++nt) {
                  const int gc = c0 + nt * 8 + 2 * tg;
                  if (gc + 1 < rm.n) {
                    *reinterpret_cast<double2 *>(prow + gc) = make_double2(acc[mt][nt][0], acc[mt][nt][1]);
                  } else if (gc < rm.n) {
                    prow[gc] = acc[mt][nt][0];
                  }
                }
              }
              if (kind == KIND_RHS) {
#pragma unroll
                for (int nt = 0; nt < 8; ++nt) {
                  const int gc = c0 + nt * 8 + 2 * tg;
                  if (gc < rm.n) q += acc[mt][nt][0] * acc[mt][nt][0];
                  if (gc + 1 < rm.n) q += acc[mt][nt][1] * acc[mt][nt][1];
                }
              }
              quad_part += q;
              if (quad_out) {
                q += __shfl_xor_sync(0xffffffffu, q, 1);
                q += __shfl_xor_sync(0xffffffffu, q, 2);
                if (kind == KIND_RHS && tg == 0) atomicAdd(quad_out + (v - NB - rm.nbel), q);
              }
              // ---- digit planes of the new rows (read by the MMA stream of later panels).
              // Inside a 64-column panel block the bytes are stored in the order 16 tg + 2 nt + e
              // (column 8 nt + 2 tg + e): a dot product over k does not care about the order as long
              // as both operands use the same one, and this one makes the 16 entries a thread holds of
              // a row contiguous -- one 16-byte store per (row, plane), no shuffles.
              // Digits: X = rint(x 2^(7S - e)), Xb = X + sum_j 64 128^j has base-128 digits u_j in
              // [0, 127] (plain bit fields, no carry chain), d_j = u_j - 64; the top digit keeps the sign.
              if (need_planes && live) {
                const int mrhs = v - NB - rm.nbel;                      // right-hand-side index (KIND_RHS)
                const int pr = (kind == KIND_RHS) ? n64 + mrhs : c0 + v;   // plane row == row of the scale table
                const double sinv = ldexp(1.0 / Eb[pr], RB * S);   // 2^(RB S - e): exact
                uint8_t *qrow = Qb + (size_t)pr * ip.LDQ + c0 + 16 * tg;
                const size_t pstride = (size_t)NR * ip.LDQ;
                constexpr long long BIAS = i8_bias(S, RB);
                constexpr unsigned FMASK = (1u << RB) - 1u;
                constexpr unsigned HALF4 = (RB == 8) ? 0x80808080u : 0x40404040u;
                uint32_t W[S][4];
#pragma unroll
                for (int s = 0; s < S; ++s)
#pragma unroll
                  for (int w = 0; w < 4; ++w) W[s][w] = 0u;
#pragma unroll
                for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
                  for (int e = 0; e < 2; ++e) {
                    const long long Xb = __double2ll_rn(acc[mt][nt][e] * sinv) + BIAS;
                    const unsigned mul = 1u << (8 * ((nt & 1) * 2 + e));   // byte position in the word
#pragma unroll
                    for (int jd = 0; jd < S - 1; ++jd)
                      W[S - 1 - jd][nt >> 1] += ((uint32_t)(Xb >> (RB * jd)) & FMASK) * mul;
                    const int top = (int)(Xb >> (RB * (S - 1))) - (1 << (RB - 1));
                    bad_range = bad_range || ((top < -128 || top > 127) && !already_bad);
                    W[0][nt >> 1] += ((uint32_t)top & 255u) * mul;
                  }
                }
#pragma unroll
                for (int s = 0; s < S; ++s) {
                  uint4 o;
                  o.x = (s == 0) ? W[s][0] : __vsub4(W[s][0], HALF4);
                  o.y = (s == 0) ? W[s][1] : __vsub4(W[s][1], HALF4);
                  o.z = (s == 0) ? W[s][2] : __vsub4(W[s][2], HALF4);
                  o.w = (s == 0) ? W[s][3] : __vsub4(W[s][3], HALF4);
                  *reinterpret_cast<uint4 *>(qrow + (size_t)s * pstride) = o;
                  // "gap" layout: second copy of a right-hand-side row, read by the full panels
                  if (gap && kind == KIND_RHS && n64 != p.n)
                    *reinterpret_cast<uint4 *>(qrow + (size_t)s * pstride -
                                               (size_t)(n64 - p.n) * ip.LDQ) = o;
                }
              }
            }
          }
          // the planes (generic-proxy stores) must be visible to the TMA engine before the producer is
          // told about them; the barrier also protects Ld / Dv against the next panel's potf2
          I8_PROF(sm, 5);
          fence_proxy_async_global();
          __threadfence_block();
          cbar256();
          if (tid == 0) {
            __threadfence_block();   // the barrier's view of the other warps' stores before the counter moves
            sm.stored = seq + 1;
          }
          I8_PROF(sm, 6);
        }
      }
      if (bad_range) sm.bad_range = 1;
      // ---- reductions -> lnlike
      const double quad = block_sum_i8(sm, quad_part);
      const double logdet = block_sum_i8(sm, logdet_part);
      I8_PROF(sm, 7);
#ifdef SPB_POTRF_PROF
      if (tid == 0) sm.prof[0][10] += 1;
#endif
      if (tid == 0) {
        const bool bad = sm.bad != 0;
        double ll = -0.5 * quad - (double)rm.M * logdet -
                    0.5 * (double)p.n * (double)rm.M * 1.8378770664093453;  // log(2 pi)
        const int prev = p.info ? (p.info[item] & ~(SPB_INFO_NOT_PD | SPB_INFO_I8_RANGE)) : 0;
        if (bad || (prev & (SPB_INFO_Z_RANGE | SPB_INFO_BOUNDS)) || ll != ll) ll = -INFINITY;
        if (p.lnlike) p.lnlike[item] = ll;
        if (p.logdet) p.logdet[item] = bad ? NAN : logdet;
        if (p.info)
          p.info[item] = prev | (bad ? SPB_INFO_NOT_PD : 0) | (sm.bad_range ? SPB_INFO_I8_RANGE : 0);
        sm.next_item = (int)gridDim.x + (int)atomicAdd(p.counter, 1u);
      }
    } else if (pw == I8_PRODUCER_WARP) {
      // ============================================================== TMA PRODUCER (one thread)
      if (lane == 0) {
        I8_PROF_DECL(1);
        unsigned known = sm.stored;
        for (int c0 = NB; c0 < p.n; c0 += NB) {
          const int nvirt = i8_nvirt(p.n, n64, M_, gap, c0);
          const int nchunks = c0 / I8_KCH;
          const bool last_partial = (c0 + NB > p.n);
          for (int v0 = 0; v0 < nvirt; v0 += I8_TM) {
            const int vlast = min(v0 + I8_TM, nvirt) - 1;   // last live virtual row of the tile
            for (int ch = 0; ch < nchunks; ++ch, ++xchunk) {
              // the k-columns of this chunk were written during panel P (a full panel), by the tile that
              // held the tile's last live row there.  Virtual row of that row in panel P: matrix rows and,
              // in full panels, right-hand sides simply shift by c0 - P; in the last (partial) panel the
              // right-hand side m = vlast - 64 sat at NB + nbel(P) + m.
              const int P = (ch * I8_KCH) & ~(NB - 1);
              int vP;
              if (last_partial) {
                const int m = vlast - NB;
                vP = (m >= 0) ? NB + i8_nbel(p.n, n64, gap, P) + m : (p.n - 1 - P);
              } else {
                vP = vlast + (c0 - P);
              }
              const unsigned need = sm.first[P / NB] + (unsigned)(vP / I8_TM) + 1u;
              I8_PROF(sm, 10);
              if ((int)(known - need) < 0) {
                while ((int)((known = sm.stored) - need) < 0) __nanosleep(100);
                __threadfence_block();
                fence_proxy_async_global();
              }
              I8_PROF(sm, 8);
              const unsigned st = xchunk % STAGES;
              if (xchunk >= (unsigned)STAGES) mbar_wait_sleep(&sm.empty[st], ((xchunk / STAGES) + 1u) & 1u, 100);
              I8_PROF(sm, 9);
              mbar_arrive_expect_tx(&sm.full[st], (unsigned)(S * (I8_TM + NB) * I8_KCH));
              tma_load_4d(&sm.A[st][0][0][0], &tmA, &sm.full[st], ch * I8_KCH, c0 + v0, 0, item);
              tma_load_4d(&sm.B[st][0][0][0], &tmB, &sm.full[st], ch * I8_KCH, c0, 0, item);
            }
          }
        }
      }
    } else if (pw == I8_ISSUER_WARP) {
      // ============================================================== MMA ISSUER (one thread)
      if (lane == 0) {
        I8_PROF_DECL(1);
        for (int c0 = NB; c0 < p.n; c0 += NB) {
          const int nvirt = i8_nvirt(p.n, n64, M_, gap, c0);
          const int nchunks = c0 / I8_KCH;
          for (int v0 = 0; v0 < nvirt; v0 += I8_TM, ++tq) {
            I8_PROF(sm, 2);
            if (tq > 0) mbar_wait_sleep(&sm.tmem_empty, (tq + 1u) & 1u, 100);   // the warps have drained TMEM
            tc_fence_after();
            I8_PROF(sm, 0);
            for (int ch = 0; ch < nchunks; ++ch, ++xchunk) {
              const unsigned st = xchunk % STAGES;
              I8_PROF(sm, 2);
              mbar_wait_sleep(&sm.full[st], (xchunk / STAGES) & 1u, 50);
              if (ip.gate & 1)
                while (*(volatile int *)&sm.quiet > 0) __nanosleep(40);
              tc_fence_after();
              I8_PROF(sm, 1);
#pragma unroll
              for (int s = 0; s < S; ++s) {
                const uint64_t da = i8_smem_desc(&sm.A[st][s][0][0]);
                const int nt = D - s + 1;                 // planes B_0 .. B_{nt-1}
#pragma unroll
                for (int t0 = 0; t0 < nt; t0 += 4) {
                  const int np = nt - t0 < 4 ? nt - t0 : 4;
                  // blocks 0 .. D are first written by s = 0 (accumulate = 0 on the first chunk)
#ifndef SPB_I8_NOMMA
                  i8_mma(tmem + (uint32_t)(NB * (s + t0)), da, i8_smem_desc(&sm.B[st][t0][0][0]),
                         i8_idesc(NB * np), (ch > 0 || s > 0) ? 1u : 0u);
#endif
                }
              }
              i8_commit(&sm.empty[st]);
            }
            i8_commit(&sm.tmem_full);
          }
        }
      }
    }
    __syncthreads();
    item = sm.next_item;
    seq_base = seq_end;
  }
  tc_fence_before();
  __syncthreads();
#ifdef SPB_POTRF_PROF
  if (tid < 32) atomicAdd(&g_potrf_prof[tid / 16][tid % 16], sm.prof[tid / 16][tid % 16]);
#endif
  if (pw == I8_ISSUER_WARP) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"(512u) : "memory");
  }
}
