// psqrt_fused.cuh -- first-order Taylor (extended) linearization of the built-in bearings-only tracking model
// evaluated INSIDE the sweeps: the north star's "linearization fused into element construction"
// (parallel/_filtering.py:110-119 and _smoothing.py:50-55 call linearization_method per step; here step k's
// (F, b) comes from the nominal mean at k and (H, c) from the nominal mean at k + 1 on the fly, in registers).
//
//   transition   coordinated turn      tests/bearings/bearings_utils.py:7-46   (nx = 5)
//   observation  two bearings sensors  tests/bearings/bearings_utils.py:49-69  (ny = 2)
//   extended     F = df/dx(m), b = f(m) - F m + m_q                            linearization/_extended.py:59-70
//
// Nothing of the per-step model is written to or read from HBM (the unfused path stores and re-reads F [25], b [5],
// H [10], c [2] doubles per step and sweep); only the nominal means (5 doubles per step) are loaded.  The Jacobians
// are sparse with a known pattern: the accessors return the structural zeros and ones as literals, so after unrolling
// the sweeps' products skip them (F has 8 free entries of 25, H has 4 of 10).
// For f(x) = M(w) x (M the turn matrix) the offset is b = -w dM/dw x + m_q exactly, so b_i = m_q,i - w J_i4.
#pragma once
#include "psqrt_math.cuh"

namespace psq {

struct FusedCTBParams {  // by value in the kernel parameters (constant bank)
  double Q[25];          // cholQ, lower triangular
  double mq[5];
  double R[4];           // cholR
  double mr[2];
  double dt, s1x, s1y, s2x, s2y;
};

struct CTJac {
  double a, bq, cw, sw, j0, j1, j2, j3;
  double b[5];
  __device__ __forceinline__ void eval(const double* __restrict__ x, const FusedCTBParams& pr) {
    const double vx = ldg(x + 2), vy = ldg(x + 3), w = ldg(x + 4);
    const double dt = pr.dt;
    sincos(w * dt, &sw, &cw);
    double da, db;
    if (fabs(w) < 1e-6) {   // lax.cond branch: sin(wt)/w -> dt, (cos(wt) - 1)/w -> 0 as constants
      a = dt; bq = 0.0; da = 0.0; db = 0.0;
    } else {
      const double iw = 1.0 / w;
      a = sw * iw;
      bq = (cw - 1.0) * iw;
      da = (dt * cw * w - sw) * iw * iw;
      db = (-dt * sw * w - (cw - 1.0)) * iw * iw;
    }
    j0 = da * vx - db * vy;
    j1 = db * vx + da * vy;
    j2 = dt * (cw * vy - sw * vx);
    j3 = -dt * (cw * vx + sw * vy);
    b[0] = fma(-w, j0, pr.mq[0]);
    b[1] = fma(-w, j1, pr.mq[1]);
    b[2] = fma(-w, j2, pr.mq[2]);
    b[3] = fma(-w, j3, pr.mq[3]);
    b[4] = pr.mq[4];
  }
  __device__ __forceinline__ double F(int i, int j) const {
    if (i == j) return (i == 2 || i == 3) ? cw : 1.0;
    if (i == 4) return 0.0;
    if (j == 4) return i == 0 ? j0 : (i == 1 ? j1 : (i == 2 ? j2 : j3));
    if (i == 0) return j == 2 ? a : (j == 3 ? -bq : 0.0);
    if (i == 1) return j == 2 ? bq : (j == 3 ? a : 0.0);
    if (i == 2) return j == 3 ? sw : 0.0;
    return j == 2 ? -sw : 0.0;   // i == 3
  }
};

struct BearingsJac {
  double h00, h01, h10, h11;
  double c[2];
  __device__ __forceinline__ void eval(const double* __restrict__ x, const FusedCTBParams& pr) {
    const double px = ldg(x), py = ldg(x + 1);
    {
      const double dx = px - pr.s1x, dy = py - pr.s1y, ir = 1.0 / (dx * dx + dy * dy);
      h00 = -dy * ir;
      h01 = dx * ir;
      c[0] = atan2(dy, dx) - (h00 * px + h01 * py) + pr.mr[0];
    }
    {
      const double dx = px - pr.s2x, dy = py - pr.s2y, ir = 1.0 / (dx * dx + dy * dy);
      h10 = -dy * ir;
      h11 = dx * ir;
      c[1] = atan2(dy, dx) - (h10 * px + h11 * py) + pr.mr[1];
    }
  }
  __device__ __forceinline__ double H(int a, int k) const {
    if (k > 1) return 0.0;
    return a == 0 ? (k == 0 ? h00 : h01) : (k == 0 ? h10 : h11);
  }
};

// one step of the forward sweeps: transition at nominal[k], observation at nominal[k + 1]
struct StepCTB {
  CTJac t;
  BearingsJac o;
  const FusedCTBParams& pr;
  template <int N_> __device__ __forceinline__ double fF(int i, int j) const { return t.F(i, j); }
  template <int N_> __device__ __forceinline__ double fQ(int i, int j) const { return pr.Q[i * 5 + j]; }
  __device__ __forceinline__ double fb(int i) const { return t.b[i]; }
  template <int N_> __device__ __forceinline__ double fH(int a, int k) const { return o.H(a, k); }
  template <int NY_> __device__ __forceinline__ double fR(int a, int q) const { return pr.R[a * 2 + q]; }
  __device__ __forceinline__ double fc(int a) const { return o.c[a]; }
};
// one step of the backward sweep: transition only
struct StepCT {
  CTJac t;
  const FusedCTBParams& pr;
  template <int N_> __device__ __forceinline__ double fF(int i, int j) const { return t.F(i, j); }
  template <int N_> __device__ __forceinline__ double fQ(int i, int j) const { return pr.Q[i * 5 + j]; }
  __device__ __forceinline__ double fb(int i) const { return t.b[i]; }
};

struct SrcFusedCTB {
  FusedCTBParams pr;
  const double* nom;   // [B][T + 1][5] nominal means
  long long nbs;       // batch stride of nom (0 = shared)
  const double* y;
  long long ty, sy;
  __device__ __forceinline__ StepCTB at(long long seq, long long k) const {
    StepCTB s{CTJac(), BearingsJac(), pr};
    const double* x = nom + seq * nbs + k * 5;
    s.t.eval(x, pr);
    s.o.eval(x + 5, pr);
    return s;
  }
  __device__ __forceinline__ const double* yp(long long seq, long long k) const { return y + seq * sy + k * ty; }
};
struct SrcFusedCT {
  FusedCTBParams pr;
  const double* nom;
  long long nbs;
  __device__ __forceinline__ StepCT at(long long seq, long long k) const {
    StepCT s{CTJac(), pr};
    s.t.eval(nom + seq * nbs + k * 5, pr);
    return s;
  }
};

}  // namespace psq
