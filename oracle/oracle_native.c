/*
 * TEST INFRASTRUCTURE ONLY -- never linked, imported or called by the product path.
 *
 * Plain-C (C99, no Eigen) CPU restatement of the reference's native numerics for the batched
 * log-likelihood hot path.  Each function cites the reference lines it follows (paths relative to
 * /root/reference/starry_process/ops/include/).  Value lanes only: the reference's derivative
 * outputs are not reproduced, except where they steer control flow (hyp2f1's stopping rule).
 *
 * Built with -ffp-contract=off so that, like the reference's `g++ -O2` x86-64 build
 * (ops/base_op.py:81-90), no multiply-add is fused.
 *
 * Validation: tests/test_oracle_vs_ref.py compares every entry point with oracle/_ref (the
 * reference headers themselves, compiled in place) -- see oracle/README.md.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

#define ORC_2F1_MAXITER 500 /* constants.h:46-48 */
#define ORC_2F1_MAXTOL 1e-15 /* constants.h:51-53 */
#define ORC_2F1_MAXDTOL 1e-13 /* constants.h:56-58 */
#define ORC_WIGNER_TOL 1.0e-14 /* constants.h:71-73 */

static int nwig(int l) { return ((l + 1) * (2 * l + 1) * (2 * l + 3)) / 3; } /* wigner.h:22-24 */

/* ------------------------------------------------------------------------------------------ */
/* special.h:173-232  Gauss 2F1 power series; b- and c-derivative lanes kept because the loop */
/* condition tests them.                                                                       */
/* ------------------------------------------------------------------------------------------ */
double orc_hyp2f1(double a, double b, double c, double z) {
  double term = a * b * z / c;
  double dtermdb = a * z / c;
  double dtermdc = -term / c;
  double value = 1.0 + term;
  double dfdb = dtermdb, dfdc = dtermdc;
  double fac1, fac2, fac3;
  int n = 1;
  while (((fabs(term / value) > ORC_2F1_MAXTOL) || (fabs(dtermdb / dfdb) > ORC_2F1_MAXDTOL) ||
          (fabs(dtermdc / dfdc) > ORC_2F1_MAXDTOL)) &&
         (n < ORC_2F1_MAXITER)) {
    a += 1;
    b += 1;
    c += 1;
    n += 1;
    fac1 = a * z / c / n;
    fac2 = fac1 * b;
    fac3 = -fac2 / c;
    dtermdb *= fac2;
    dtermdb += fac1 * term;
    dtermdc *= fac2;
    dtermdc += fac3 * term;
    term *= fac2;
    value += term;
    dfdb += dtermdb;
    dfdc += dtermdc;
  }
  return value; /* non-convergence is silent under -DSTARRY_NO_EXCEPTIONS (special.h:207-228) */
}

/* ------------------------------------------------------------------------------------------ */
/* latitude.h:22-173 (+ the >=0 clamp of ops/latitude/latitude.cc:47-48)                       */
/* q: (N), Q: (N,N) row-major, N = (ydeg+1)^2                                                  */
/* ------------------------------------------------------------------------------------------ */
void orc_latitude_integrals(int ydeg, double alpha, double beta, double *q, double *Q) {
  const int n = 4 * ydeg + 1;
  const int N = (ydeg + 1) * (ydeg + 1);
  double *B = (double *)malloc(sizeof(double) * n);
  double *F = (double *)malloc(sizeof(double) * n);
  double *term = (double *)calloc((size_t)n * n, sizeof(double));
  double c1, c2, c3;
  int i, j, k, k1, k2, i2, j2;

  alpha = alpha > 0.0 ? alpha : 0.0;
  beta = beta > 0.0 ? beta : 0.0;

  /* B functions, latitude.h:48-60 */
  B[0] = 1.0;
  for (k = 1; k < n; ++k) {
    c1 = 1.0 / (alpha + beta + k - 1.0);
    c2 = (alpha + k - 1.0) * c1;
    B[k] = c2 * B[k - 1];
  }

  /* F functions, latitude.h:63-109 */
  {
    double ab = alpha + beta;
    F[0] = sqrt(2.0) * orc_hyp2f1(-0.5, beta, ab, 0.5);
    F[1] = sqrt(2.0) * orc_hyp2f1(-0.5, beta, ab + 1.0, 0.5);
    for (k = 2; k < n; ++k) {
      c1 = (ab + k - 1.0) / ((alpha + k - 1.0) * (ab + k - 0.5));
      c2 = c1 * (ab + k - 2.0);
      c3 = c1 * (1.5 - beta);
      F[k] = c2 * F[k - 2] + c3 * F[k - 1];
    }
    for (k = 0; k < n; ++k) F[k] = F[k] * B[k];
  }

  /* Terms of the integrals, latitude.h:112-143 */
  for (i = 0; i < n; ++i) {
    const double *func = (i % 2 == 0) ? B : F;
    i2 = (i % 2 == 0) ? i / 2 : (i - 1) / 2;
    for (j = 0; j < n; j += 2) {
      double fac1 = 1.0, fac2, acc = 0.0;
      j2 = j / 2;
      for (k1 = 0; k1 < i2 + 1; ++k1) {
        fac2 = fac1;
        for (k2 = 0; k2 < j2 + 1; ++k2) {
          acc += fac2 * func[k1 + k2];
          fac2 *= (k2 - j2) / (k2 + 1.0);
        }
        fac1 *= (i2 - k1) / (k1 + 1.0);
      }
      term[i * n + j] = acc;
    }
  }

  /* Moments, latitude.h:146-172 */
  {
    int n1 = 0, n2, l1, m1, l2, m2, j1, i1;
    double inv_two_l1 = 1.0, inv_two_l1l2;
    for (l1 = 0; l1 < ydeg + 1; ++l1) {
      for (m1 = -l1; m1 < l1 + 1; ++m1) {
        j1 = m1 + l1;
        i1 = l1 - m1;
        q[n1] = term[j1 * n + i1] * inv_two_l1;
        n2 = 0;
        inv_two_l1l2 = inv_two_l1;
        for (l2 = 0; l2 < ydeg + 1; ++l2) {
          for (m2 = -l2; m2 < l2 + 1; ++m2) {
            j2 = m2 + l2;
            i2 = l2 - m2;
            Q[(size_t)n1 * N + n2] = term[(j1 + j2) * n + (i1 + i2)] * inv_two_l1l2;
            n2 += 1;
          }
          inv_two_l1l2 *= 0.5;
        }
        n1 += 1;
      }
      inv_two_l1 *= 0.5;
    }
  }
  free(B);
  free(F);
  free(term);
}

/* ------------------------------------------------------------------------------------------ */
/* wigner.h:37-139  complex Wigner d-matrix of degree l from degrees l-1, l-2 (value lane)     */
/* Dl is (2l+1)x(2l+1) row-major                                                               */
/* ------------------------------------------------------------------------------------------ */
static void orc_dlmn(int l, double c2, double s2, const double *Dlm2, const double *Dlm1,
                     double *Dl) {
  const int w = 2 * l + 1, w1 = 2 * l - 1, w2 = 2 * l - 3;
#define DL(r, c) Dl[(r) * w + (c)]
#define DM1(r, c) Dlm1[(r) * w1 + (c)]
#define DM2(r, c) Dlm2[(r) * w2 + (c)]
  int iinf = 1 - l, isup = -iinf, m, mp, al, al1, tal1, amp, laux, lbux, am, lauz, lbuz, sign;
  double ali, auz, aux, cux, fact, cuz, term, cosaux, tgbet2;

  if (fabs(s2) < ORC_WIGNER_TOL)
    tgbet2 = s2;
  else
    tgbet2 = (1.0 - c2) / s2;

  /* first row by recurrence, wigner.h:57-73 */
  DL(2 * l, 2 * l) = 0.5 * DM1(isup + l - 1, isup + l - 1) * (1.0 + c2);
  DL(2 * l, 0) = 0.5 * DM1(isup + l - 1, -isup + l - 1) * (1.0 - c2);
  for (m = isup; m > iinf - 1; --m)
    DL(2 * l, m + l) = -tgbet2 * sqrt((double)(l + m + 1) / (l - m)) * DL(2 * l, m + 1 + l);

  /* upper quarter triangle, wigner.h:77-110 */
  al = l;
  al1 = al - 1;
  tal1 = al + al1;
  ali = 1.0 / al1;
  cosaux = c2 * al * al1;
  for (mp = l - 1; mp > -1; --mp) {
    amp = mp;
    laux = l + mp;
    lbux = l - mp;
    aux = ali / sqrt((double)(laux * lbux));
    cux = sqrt((double)((laux - 1) * (lbux - 1))) * al;
    for (m = isup; m > iinf - 1; --m) {
      am = m;
      lauz = l + m;
      lbuz = l - m;
      auz = 1.0 / sqrt((double)(lauz * lbuz));
      fact = aux * auz;
      term = tal1 * (cosaux - (double)(am * amp)) * DM1(mp + l - 1, m + l - 1);
      if ((lbuz != 1) && (lbux != 1)) {
        cuz = sqrt((double)((lauz - 1) * (lbuz - 1)));
        term = term - DM2(mp + l - 2, m + l - 2) * cux * cuz;
      }
      DL(mp + l, m + l) = fact * term;
    }
    ++iinf;
    --isup;
  }

  /* reflection, wigner.h:117-129 */
  sign = 1;
  iinf = -l;
  isup = l - 1;
  for (m = l; m > 0; --m) {
    for (mp = iinf; mp < isup + 1; ++mp) {
      DL(mp + l, m + l) = sign * DL(m + l, mp + l);
      sign *= -1;
    }
    ++iinf;
    --isup;
  }

  /* inversion, wigner.h:131-142 */
  iinf = -l;
  isup = iinf;
  for (m = l - 1; m > -(l + 1); --m) {
    sign = -1;
    for (mp = isup; mp > iinf - 1; --mp) {
      DL(mp + l, m + l) = sign * DL(-mp + l, -m + l);
      sign *= -1;
    }
    ++isup;
  }
#undef DL
#undef DM1
#undef DM2
}

/* ------------------------------------------------------------------------------------------ */
/* wigner.h:146-284  rotar / computeRx: real Wigner matrices Rx(theta), packed, all l<=ydeg    */
/* ------------------------------------------------------------------------------------------ */
void orc_Rx(int ydeg, double theta, double *R) {
  const int NW = nwig(ydeg);
  double *D = (double *)calloc(NW > 10 ? NW : 10, sizeof(double));
  const double root_two = sqrt(2.0);
  const double c2 = cos(theta), s2 = sin(theta);
  int l;

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
  if (ydeg >= 1) {
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

  for (l = 2; l < ydeg + 1; ++l) {
    const int w = 2 * l + 1;
    double *Dl = D + nwig(l - 1);
    double *Rl = R + nwig(l - 1);
    int cosmal, sinmal, cosmga, sinmga, cosag, sinag, cosagm, sinagm, sign, aux, mp, m;
    double d1, d2;
    orc_dlmn(l, c2, s2, D + nwig(l - 3), D + nwig(l - 2), Dl);
#define DL(r, c) Dl[(r) * w + (c)]
#define RL(r, c) Rl[(r) * w + (c)]
    RL(l, l) = DL(l, l);
    cosmal = 0;
    sinmal = -1;
    sign = -1;
    for (mp = 1; mp < l + 1; ++mp) {
      cosmga = 0;
      sinmga = 1;
      RL(mp + l, l) = root_two * DL(l, mp + l) * cosmal;
      RL(-mp + l, l) = root_two * DL(l, mp + l) * sinmal;
      for (m = 1; m < l + 1; ++m) {
        d1 = DL(-mp + l, -m + l);
        d2 = sign * DL(mp + l, -m + l);
        cosag = cosmal * cosmga - sinmal * sinmga;
        cosagm = cosmal * cosmga + sinmal * sinmga;
        sinag = sinmal * cosmga + cosmal * sinmga;
        sinagm = sinmal * cosmga - cosmal * sinmga;
        RL(l, m + l) = root_two * DL(m + l, l) * cosmga;
        RL(l, -m + l) = -root_two * DL(m + l, l) * sinmga;
        RL(mp + l, m + l) = d1 * cosag + d2 * cosagm;
        RL(mp + l, -m + l) = -d1 * sinag + d2 * sinagm;
        RL(-mp + l, m + l) = d1 * sinag + d2 * sinagm;
        RL(-mp + l, -m + l) = d1 * cosag - d2 * cosagm;
        aux = -sinmga;
        sinmga = cosmga;
        cosmga = aux;
      }
      sign *= -1;
      aux = sinmal;
      sinmal = -cosmal;
      cosmal = aux;
    }
#undef DL
#undef RL
  }
  free(D);
}

/* cos(n theta), sin(n theta), n<=ydeg by the recurrences of wigner.h:307-316 */
static void orc_cs(int ydeg, double theta, double *cosnt, double *sinnt) {
  int n;
  cosnt[0] = 1.0;
  sinnt[0] = 0.0;
  if (ydeg >= 1) {
    cosnt[1] = cos(theta);
    sinnt[1] = sin(theta);
  }
  for (n = 2; n < ydeg + 1; ++n) {
    cosnt[n] = 2.0 * cosnt[n - 1] * cosnt[1] - cosnt[n - 2];
    sinnt[n] = 2.0 * sinnt[n - 1] * cosnt[1] - sinnt[n - 2];
  }
}

/* ------------------------------------------------------------------------------------------ */
/* wigner.h:290-339  f = M . Rz(theta_k) row by row;  M, f: (K, N) row-major                   */
/* ------------------------------------------------------------------------------------------ */
void orc_tensordotRz(int ydeg, const double *M, const double *theta, int K, double *f) {
  const int N = (ydeg + 1) * (ydeg + 1);
  double cosnt[64], sinnt[64];
  int k, l, j;
  for (k = 0; k < K; ++k) {
    const double *Mk = M + (size_t)k * N;
    double *fk = f + (size_t)k * N;
    orc_cs(ydeg, theta[k], cosnt, sinnt);
    for (l = 0; l < ydeg + 1; ++l) {
      for (j = 0; j < 2 * l + 1; ++j) {
        int m = j - l;
        double cm = cosnt[m < 0 ? -m : m];
        double sm = m < 0 ? -sinnt[-m] : sinnt[m];
        fk[l * l + j] = Mk[l * l + j] * cm + Mk[l * l + 2 * l - j] * sm;
      }
    }
  }
}

/* ------------------------------------------------------------------------------------------ */
/* wigner.h:410-459  f_k = sum_ij cosmt(k,i) TM1(i,j) + sinmt(k,i) TM2(i,j)                    */
/*   TM1 = T o M,  TM2(:,n) = T(:,n) o M(:, nbar),  nbar = same l, opposite m                  */
/* The reference evaluates (cosmt*TM1 + sinmt*TM2).rowwise().sum(): a K x N matrix product     */
/* followed by a row sum; the same association is kept here.                                   */
/* ------------------------------------------------------------------------------------------ */
void orc_special_tensordotRz(int ydeg, const double *T, const double *M, const double *theta,
                             int K, double *f) {
  const int N = (ydeg + 1) * (ydeg + 1);
  double cosnt[64], sinnt[64];
  double *cosmt = (double *)malloc(sizeof(double) * N);
  double *sinmt = (double *)malloc(sizeof(double) * N);
  double *TM1 = (double *)malloc(sizeof(double) * N * N);
  double *TM2 = (double *)malloc(sizeof(double) * N * N);
  int i, j, k, l, m, n;
  for (i = 0; i < N; ++i) {
    for (l = 0; l < ydeg + 1; ++l) {
      for (m = -l; m < l + 1; ++m) {
        int col = l * l + l + m, bar = l * l + l - m;
        TM1[(size_t)i * N + col] = T[(size_t)i * N + col] * M[(size_t)i * N + col];
        TM2[(size_t)i * N + col] = T[(size_t)i * N + col] * M[(size_t)i * N + bar];
      }
    }
  }
  for (k = 0; k < K; ++k) {
    double total = 0.0;
    orc_cs(ydeg, theta[k], cosnt, sinnt);
    n = 0;
    for (l = 0; l < ydeg + 1; ++l) {
      for (m = -l; m < 0; ++m) {
        cosmt[n] = cosnt[-m];
        sinmt[n] = -sinnt[-m];
        ++n;
      }
      for (m = 0; m < l + 1; ++m) {
        cosmt[n] = cosnt[m];
        sinmt[n] = sinnt[m];
        ++n;
      }
    }
    for (j = 0; j < N; ++j) {
      double a1 = 0.0, a2 = 0.0;
      for (i = 0; i < N; ++i) {
        a1 += cosmt[i] * TM1[(size_t)i * N + j];
        a2 += sinmt[i] * TM2[(size_t)i * N + j];
      }
      total += a1 + a2;
    }
    f[k] = total;
  }
  free(cosmt);
  free(sinmt);
  free(TM1);
  free(TM2);
}

/* ------------------------------------------------------------------------------------------ */
/* flux.h:23-68  rT: phase-curve solution vector in the polynomial basis, degree `deg`         */
/* ------------------------------------------------------------------------------------------ */
static void orc_rT(int deg, double *rT) {
  double amp0, amp, lfac1, lfac2;
  int l, m, mu, nu;
  memset(rT, 0, sizeof(double) * (deg + 1) * (deg + 1));
  amp0 = M_PI;
  lfac1 = 1.0;
  lfac2 = 2.0 / 3.0;
  for (l = 0; l < deg + 1; l += 4) {
    amp = amp0;
    for (m = 0; m < l + 1; m += 4) {
      mu = l - m;
      nu = l + m;
      rT[l * l + l + m] = amp * lfac1;
      rT[l * l + l - m] = amp * lfac1;
      if (l < deg) {
        rT[(l + 1) * (l + 1) + l + m + 1] = amp * lfac2;
        rT[(l + 1) * (l + 1) + l - m + 1] = amp * lfac2;
      }
      amp *= (nu + 2.0) / (mu - 2.0);
    }
    lfac1 /= (l / 2 + 2) * (l / 2 + 3);
    lfac2 /= (l / 2 + 2.5) * (l / 2 + 3.5);
    amp0 *= 0.0625 * (l + 2) * (l + 2);
  }
  amp0 = 0.5 * M_PI;
  lfac1 = 0.5;
  lfac2 = 4.0 / 15.0;
  for (l = 2; l < deg + 1; l += 4) {
    amp = amp0;
    for (m = 2; m < l + 1; m += 4) {
      mu = l - m;
      nu = l + m;
      rT[l * l + l + m] = amp * lfac1;
      rT[l * l + l - m] = amp * lfac1;
      if (l < deg) {
        rT[(l + 1) * (l + 1) + l + m + 1] = amp * lfac2;
        rT[(l + 1) * (l + 1) + l - m + 1] = amp * lfac2;
      }
      amp *= (nu + 2.0) / (mu - 2.0);
    }
    lfac1 /= (l / 2 + 2) * (l / 2 + 3);
    lfac2 /= (l / 2 + 2.5) * (l / 2 + 3.5);
    amp0 *= 0.0625 * l * (l + 4);
  }
}

/* flux.h:75-96  multiply a polynomial-basis vector (degree deg) by z; out has degree deg+1   */
static void orc_polymulz(int deg, const double *p, double *pz, int Nout) {
  int n = 0, l, m, lz, nz;
  memset(pz, 0, sizeof(double) * Nout);
  for (l = 0; l < deg + 1; ++l) {
    for (m = -l; m < l + 1; ++m) {
      lz = l + 1;
      nz = lz * lz + lz + m;
      if ((l + m) % 2 != 0) {
        pz[nz - 4 * lz + 2] += p[n];
        pz[nz - 2] -= p[n];
        pz[nz + 2] -= p[n];
      } else {
        pz[nz] += p[n];
      }
      ++n;
    }
  }
}

/* ------------------------------------------------------------------------------------------ */
/* flux.h:103-279  dense restatement of computeA1 (legendre, theta, amp, sparse product)       */
/* A1: (N,N) row-major, N = (deg+1)^2; column `col` = polynomial expansion of Ylm `col`        */
/* ------------------------------------------------------------------------------------------ */
static void orc_A1(int deg, double *A1) {
  const int N = (deg + 1) * (deg + 1);
  const double norm = 2.0 / sqrt(M_PI);
  /* Z[col*N + n]: coefficient of polynomial-basis term n in the P(z) factor (flux.h:103-152) */
  double *Z = (double *)calloc((size_t)N * N, sizeof(double));
  double *colvec = (double *)malloc(sizeof(double) * N);
  double *C = (double *)calloc(N, sizeof(double));
  double term = 1.0, fac = 1.0;
  int m, l, ip, im, col, n, j;
  for (m = 0; m < deg + 1; ++m) {
    ip = m * m + 2 * m;
    im = m * m;
    Z[(size_t)ip * N + 0] = fac;
    Z[(size_t)im * N + 0] = fac;
    /* flux.h:121-127 seeds the l = m+1 columns here; that assignment is dead because the    */
    /* recursion below rewrites the whole column for l = m+1, so it is not restated.          */
    for (l = m + 1; l < deg + 1; ++l) {
      ip = l * l + l + m;
      im = l * l + l - m;
      orc_polymulz(deg - 1, Z + (size_t)((l - 1) * (l - 1) + l - 1 + m) * N, colvec, N);
      for (n = 0; n < N; ++n) Z[(size_t)ip * N + n] = (2 * l - 1) * colvec[n] / (l - m);
      if (l > m + 1)
        for (n = 0; n < N; ++n)
          Z[(size_t)ip * N + n] -=
              (l + m - 1) * Z[(size_t)((l - 2) * (l - 2) + l - 2 + m) * N + n] / (l - m);
      for (n = 0; n < N; ++n) Z[(size_t)im * N + n] = Z[(size_t)ip * N + n];
    }
    fac *= -term;
    term += 2;
  }

  /* amplitudes (flux.h:190-203): one constant per column */
  {
    const double inv_root_two = sqrt(0.5);
    for (l = 0; l < deg + 1; ++l) {
      C[l * l + l] = sqrt((double)(2 * (2 * l + 1)));
      for (m = 1; m < l + 1; ++m) {
        C[l * l + l + m] = -C[l * l + l + m - 1] / sqrt((double)((l + m) * (l - m + 1)));
        C[l * l + l - m] = C[l * l + l + m];
      }
      C[l * l + l] *= inv_root_two;
    }
    for (n = 0; n < N; ++n) C[n] /= (2 * sqrt(M_PI));
  }

  memset(A1, 0, sizeof(double) * N * N);
  /* theta(x,y) terms (flux.h:159-183) multiplied into the z-polynomials (flux.h:211-236) and   */
  /* scattered into A1 (flux.h:262-276).  XY terms are generated on the fly in reference order. */
  for (col = 0; col < N; ++col) {
    int lc = (int)floor(sqrt((double)col));
    int mc = col - lc * lc - lc;
    int am = mc < 0 ? -mc : mc;
    double term1 = 1.0, term2 = am;
    for (j = 0; j < am + 1; j += 2) {
      int l2 = am, m2;
      double v2;
      int pass;
      if (j > 0) {
        term1 *= -(am - j + 1.0) * (am - j + 2.0) / (j * (j - 1.0));
        term2 *= -(am - j) * (am - j + 1.0) / (j * (j + 1.0));
      }
      /* m>=0 columns take (am, 2j-am, term1); m<0 columns take (am, 2(j+1)-am, term2), j<am */
      for (pass = 0; pass < 1; ++pass) {
        if (mc >= 0) {
          m2 = 2 * j - am;
          v2 = term1;
        } else {
          if (!(j < am)) continue;
          m2 = 2 * (j + 1) - am;
          v2 = term2;
        }
        {
          int odd2 = ((l2 + m2) % 2 != 0);
          int n1 = 0, l1, m1;
          for (l1 = 0; l1 < deg + 1; ++l1) {
            for (m1 = -l1; m1 < l1 + 1; ++m1, ++n1) {
              double v1 = Z[(size_t)col * N + n1], prod;
              int odd1, lo, mo;
              if (v1 == 0) continue;
              odd1 = ((l1 + m1) % 2 != 0);
              prod = v1 * v2;
              lo = l1 + l2;
              mo = m1 + m2;
              if (odd1 && odd2) {
                int r;
                r = (lo - 2) * (lo - 2) + (lo - 2) + mo;
                A1[(size_t)r * N + col] += prod * norm * C[col];
                r = lo * lo + lo + mo - 2;
                A1[(size_t)r * N + col] += -prod * norm * C[col];
                r = lo * lo + lo + mo + 2;
                A1[(size_t)r * N + col] += -prod * norm * C[col];
              } else {
                int r = lo * lo + lo + mo;
                A1[(size_t)r * N + col] += prod * norm * C[col];
              }
            }
          }
        }
      }
    }
  }
  free(Z);
  free(colvec);
  free(C);
}

/* flux.h:302-309  rTA1 = rT . A1  (no limb darkening) */
void orc_rTA1(int ydeg, double *out) {
  const int N = (ydeg + 1) * (ydeg + 1);
  double *rT = (double *)malloc(sizeof(double) * N);
  double *A1 = (double *)malloc(sizeof(double) * N * N);
  int r, c;
  orc_rT(ydeg, rT);
  orc_A1(ydeg, A1);
  for (c = 0; c < N; ++c) {
    double acc = 0.0;
    for (r = 0; r < N; ++r) acc += rT[r] * A1[(size_t)r * N + c];
    out[c] = acc;
  }
  free(rT);
  free(A1);
}

/* ------------------------------------------------------------------------------------------ */
/* flux.h:315-523  LimbDark: U1 (332-409), Lp (415-441), rTA1L forward (501-523)               */
/* YT is upper triangular, so the reference's HouseholderQR solve (flux.h:389-392) is restated */
/* as a back substitution.                                                                     */
/* ------------------------------------------------------------------------------------------ */
void orc_rTA1L(int ydeg, int udeg, const double *u, double *out) {
  const int lu = ydeg + udeg, n = lu + 1;
  const int N = (ydeg + 1) * (ydeg + 1), NLU = (lu + 1) * (lu + 1);
  const int NU = (udeg + 1) * (udeg + 1);
  const double norm = 2.0 / sqrt(M_PI);
  double *rT = (double *)malloc(sizeof(double) * NLU);
  double *A1 = (double *)malloc(sizeof(double) * NLU * NLU);
  double *LT = (double *)calloc((size_t)n * n, sizeof(double));
  double *YT = (double *)calloc((size_t)n * n, sizeof(double));
  double *U0 = (double *)calloc((size_t)n * n, sizeof(double));
  double *U1 = (double *)calloc((size_t)NU * (udeg + 1), sizeof(double));
  double *p = (double *)calloc(NU, sizeof(double));
  double *Lp = (double *)calloc((size_t)NLU * N, sizeof(double));
  double *rTLp = (double *)calloc(N, sizeof(double));
  double twol, amp, lfac, lchoosek, fac0, fac, dot;
  int l, k, r, c, j;

  if (udeg == 0) {
    orc_rTA1(ydeg, out);
    goto done;
  }
  orc_rT(lu, rT);
  orc_A1(lu, A1);

  /* L^T, flux.h:341-350 (column-major n x n, element (k,l)) */
  for (l = 0; l < n; ++l) {
    lchoosek = 1;
    for (k = 0; k < l + 1; ++k) {
      LT[k + l * n] = ((k + 1) % 2 == 0) ? lchoosek : -lchoosek;
      lchoosek *= (l - k) / (k + 1.0);
    }
  }
  /* Y^T, flux.h:352-386 */
  twol = 1.0;
  lfac = 1.0;
  fac0 = 1.0;
  for (l = 0; l < n; l += 2) {
    amp = twol * sqrt((2 * l + 1) / (4 * M_PI)) / lfac;
    lchoosek = 1;
    fac = fac0;
    for (k = 0; k < l + 1; k += 2) {
      YT[k + l * n] = amp * lchoosek * fac;
      fac *= (k + l + 1.0) / (k - l + 1.0);
      lchoosek *= (l - k) * (l - k - 1) / ((k + 1.0) * (k + 2.0));
    }
    fac0 *= -0.25 * (l + 1) * (l + 1);
    lfac *= (l + 1.0) * (l + 2.0);
    twol *= 4.0;
  }
  twol = 2.0;
  lfac = 1.0;
  fac0 = 0.5;
  for (l = 1; l < n; l += 2) {
    amp = twol * sqrt((2 * l + 1) / (4 * M_PI)) / lfac;
    lchoosek = l;
    fac = fac0;
    for (k = 1; k < l + 1; k += 2) {
      YT[k + l * n] = amp * lchoosek * fac;
      fac *= (k + l + 1.0) / (k - l + 1.0);
      lchoosek *= (l - k) * (l - k - 1) / ((k + 1.0) * (k + 2.0));
    }
    fac0 *= -0.25 * (l + 2) * l;
    lfac *= (l + 1.0) * (l + 2.0);
    twol *= 4.0;
  }
  /* U0 = YT^{-1} LT / norm  (flux.h:388-396) */
  for (c = 0; c < n; ++c) {
    for (r = n - 1; r >= 0; --r) {
      double acc = LT[r + c * n];
      for (j = r + 1; j < n; ++j) acc -= YT[r + j * n] * U0[j + c * n];
      U0[r + c * n] = acc / YT[r + r * n];
    }
  }
  for (r = 0; r < n * n; ++r) U0[r] /= norm;
  /* U1 = (A1 . X . U0)[:NU, :udeg+1], X(l(l+1), l) = 1   (flux.h:398-408) */
  for (r = 0; r < NU; ++r)
    for (c = 0; c < udeg + 1; ++c) {
      double acc = 0.0;
      for (l = 0; l < n; ++l) acc += A1[(size_t)r * NLU + l * (l + 1)] * U0[l + c * n];
      U1[r * (udeg + 1) + c] = acc;
    }

  /* limb darkening polynomial, flux.h:510-517 */
  for (r = 0; r < NU; ++r) {
    double acc = U1[r * (udeg + 1) + 0] * (-1.0);
    for (c = 1; c < udeg + 1; ++c) acc += U1[r * (udeg + 1) + c] * u[c - 1];
    p[r] = acc;
  }
  dot = 0.0;
  for (r = 0; r < NU; ++r) dot += rT[r] * p[r];
  {
    double nrm = 1.0 / dot;
    for (r = 0; r < NU; ++r) p[r] *= nrm * M_PI;
  }

  /* Lp, flux.h:415-441 */
  {
    int n1 = 0, n2, l1, m1, l2, m2, nn, odd1;
    for (l1 = 0; l1 < ydeg + 1; ++l1) {
      for (m1 = -l1; m1 < l1 + 1; ++m1) {
        odd1 = ((l1 + m1) % 2 != 0);
        n2 = 0;
        for (l2 = 0; l2 < udeg + 1; ++l2) {
          for (m2 = -l2; m2 < l2 + 1; ++m2) {
            l = l1 + l2;
            nn = l * l + l + m1 + m2;
            if (odd1 && ((l2 + m2) % 2 != 0)) {
              Lp[(size_t)(nn - 4 * l + 2) * N + n1] += p[n2];
              Lp[(size_t)(nn - 2) * N + n1] -= p[n2];
              Lp[(size_t)(nn + 2) * N + n1] -= p[n2];
            } else {
              Lp[(size_t)nn * N + n1] += p[n2];
            }
            ++n2;
          }
        }
        ++n1;
      }
    }
  }
  /* rTA1L = (rT . Lp) . A1[:N,:N]   flux.h:522 */
  for (c = 0; c < N; ++c) {
    double acc = 0.0;
    for (r = 0; r < NLU; ++r) acc += rT[r] * Lp[(size_t)r * N + c];
    rTLp[c] = acc;
  }
  for (c = 0; c < N; ++c) {
    double acc = 0.0;
    for (r = 0; r < N; ++r) acc += rTLp[r] * A1[(size_t)r * NLU + c];
    out[c] = acc;
  }
done:
  free(rT);
  free(A1);
  free(LT);
  free(YT);
  free(U0);
  free(U1);
  free(p);
  free(Lp);
  free(rTLp);
}
