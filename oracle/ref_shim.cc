// TEST INFRASTRUCTURE ONLY -- not part of the product path.
//
// extern "C" shim that exposes the reference's OWN native numerics (the header-only C++ under
// /root/reference/starry_process/ops/include, compiled where it lies) so that oracle/ can be
// validated against the real reference arithmetic.  No reference source is copied: this file
// only #includes the headers and forwards to their public entry points.
//
//   sp::wigner::computeRx                  ops/include/wigner.h:282
//   sp::wigner::computeTensordotRz         ops/include/wigner.h:290
//   sp::wigner::computeSpecialTensordotRz  ops/include/wigner.h:410
//   sp::flux::computerTA1                  ops/include/flux.h:302
//   sp::flux::LimbDark::computerTA1L       ops/include/flux.h:501
//   sp::latitude::computeLatitudeIntegrals ops/include/latitude.h:22
//
// Built by oracle/Makefile into oracle/_ref/libspref.so with the reference's own flags
// (ops/base_op.py:81-90).
#include <sstream>
#include "utils.h"
#include "special.h"
#include "latitude.h"
#include "wigner.h"
#include "flux.h"

using namespace sp::utils;

extern "C" {

int ref_ydeg() { return SP__LMAX; }
int ref_udeg() { return SP__UMAX; }

void ref_Rx(double theta, double *R, double *dR) {
  Map<Vector<double, SP__NWIG>> Rm(R), dRm(dR);
  sp::wigner::computeRx(theta, Rm, dRm);
}

void ref_tensordotRz(const double *M, const double *theta, int K, double *f) {
  Map<RowMatrix<double, Dynamic, SP__N>> Mm(const_cast<double *>(M), K, SP__N);
  Map<Vector<double, Dynamic>> th(const_cast<double *>(theta), K);
  Map<RowMatrix<double, Dynamic, SP__N>> fm(f, K, SP__N);
  sp::wigner::computeTensordotRz(Mm, th, fm);
}

void ref_special_tensordotRz(const double *T, const double *M, const double *theta, int K,
                             double *f) {
  RowMatrix<double, Dynamic, Dynamic> Tm =
      Map<const RowMatrix<double, SP__N, SP__N>>(T);
  RowMatrix<double, Dynamic, Dynamic> Mm =
      Map<const RowMatrix<double, SP__N, SP__N>>(M);
  Vector<double, Dynamic> tc = Map<const Vector<double, Dynamic>>(theta, K);
  Vector<double, Dynamic> fc(K);
  sp::wigner::computeSpecialTensordotRz(Tm, Mm, tc, fc);
  for (int k = 0; k < K; ++k) f[k] = fc(k);
}

void ref_rTA1(double *out) {
  Map<Vector<double, SP__N>> f(out);
  sp::flux::computerTA1(f);
}

void ref_rTA1L(const double *u, double *out) {
#if SP__UMAX > 0
  static sp::flux::LimbDark<double> *LD = NULL;
  if (LD == NULL) LD = new sp::flux::LimbDark<double>();
  Vector<double, SP__UMAX> uv = Map<const Vector<double, SP__UMAX>>(u);
  Map<RowVector<double, SP__N>> f(out);
  LD->computerTA1L(uv, f);
#else
  ref_rTA1(out);
#endif
}

// alpha/beta clamp as in ops/latitude/latitude.cc:47-48
void ref_latitude(double alpha, double beta, double *q, double *Q) {
  alpha = alpha > 0.0 ? alpha : 0.0;
  beta = beta > 0.0 ? beta : 0.0;
  // Outputs are mapped in place exactly as ops/latitude/latitude.cc:68-73 does; the derivative
  // lanes go to scratch heap buffers.
  std::vector<double> scratch(2 * SP__N + 2 * SP__N * SP__N);
  Map<Vector<double, SP__N>> qv(q), dqda(scratch.data()), dqdb(scratch.data() + SP__N);
  Map<RowMatrix<double, SP__N, SP__N>> Qm(Q), dQda(scratch.data() + 2 * SP__N),
      dQdb(scratch.data() + 2 * SP__N + SP__N * SP__N);
  sp::latitude::computeLatitudeIntegrals(alpha, beta, qv, dqda, dqdb, Qm, dQda, dQdb);
}

double ref_hyp2f1(double a, double b, double c, double z) {
  double dfdb, dfdc;
  return sp::special::hyp2f1(a, b, c, z, dfdb, dfdc);
}
}
