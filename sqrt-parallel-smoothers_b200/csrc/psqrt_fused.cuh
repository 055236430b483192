// psqrt_fused.cuh -- first-order Taylor (extended) linearization of the built-in bearings-only tracking model
// evaluated INSIDE the sweeps: the north star's "linearization fused into element construction"
// (parallel/_filtering.py:110-119 and _smoothing.py:50-55 call linearization_method per step; here step k's
// (F, b) comes from the nominal mean at k and (H, c) from the nominal mean at k + 1 on the fly, in registers).
//
//   transition   coordinated turn      tests/bearings/bearings_utils.py:7-46   (nx = 5)
//   observation  two bearings sensors  tests/bearings/bearings_utils.py:49-69  (ny = 2)
//   extended     F = df/dx(m), b = f(m) - F m + m_q                            linearization/_extended.py:59-70
//
// The per-step model is never written to or read from HBM (the unfused path stores and re-reads F [25], b [5], H [10],
// c [2] doubles per step and sweep).  What IS precomputed, by one light kernel per pass (k_fused_trig), are the four
// transcendental values per nominal point -- sin(w dt), cos(w dt) and the two bearings atan2 -- because libm's
// sincos / atan2 are calls with slow paths, and a call inside a 255-register sweep spills the whole state around it
// (first version, everything inline: 2.7x SLOWER than the unfused path).  The sweeps load those 4 doubles and the
// nominal mean (5 doubles) per step and build the Jacobians in registers with branch-free reciprocals.  The Jacobians
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
  // tr = (sin(w dt), cos(w dt), ., .) of this nominal point (k_fused_trig)
  __device__ __forceinline__ void eval(const double* __restrict__ x, const double* __restrict__ tr,
                                       const FusedCTBParams& pr) {
    const double vx = ldg(x + 2), vy = ldg(x + 3), w = ldg(x + 4);
    const double dt = pr.dt;
    sw = ldg(tr);
    cw = ldg(tr + 1);
    // lax.cond branch |w| < 1e-6: sin(wt)/w -> dt, (cos(wt) - 1)/w -> 0 as constants; both sides computed, one selected
    const bool small = fabs(w) < 1e-6;
    const double iw = rcp_nr(small ? 1.0 : w);
    a = small ? dt : sw * iw;
    bq = small ? 0.0 : (cw - 1.0) * iw;
    const double da = small ? 0.0 : (dt * cw * w - sw) * iw * iw;
    const double db = small ? 0.0 : (-dt * sw * w - (cw - 1.0)) * iw * iw;
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
  // tr = (., ., atan2 to sensor 1, atan2 to sensor 2) of this nominal point (k_fused_trig)
  __device__ __forceinline__ void eval(const double* __restrict__ x, const double* __restrict__ tr,
                                       const FusedCTBParams& pr) {
    const double px = ldg(x), py = ldg(x + 1);
    {
      const double dx = px - pr.s1x, dy = py - pr.s1y, ir = rcp_nr(fma(dx, dx, dy * dy));
      h00 = -dy * ir;
      h01 = dx * ir;
      c[0] = ldg(tr + 2) - (h00 * px + h01 * py) + pr.mr[0];
    }
    {
      const double dx = px - pr.s2x, dy = py - pr.s2y, ir = rcp_nr(fma(dx, dx, dy * dy));
      h10 = -dy * ir;
      h11 = dx * ir;
      c[1] = ldg(tr + 3) - (h10 * px + h11 * py) + pr.mr[1];
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
  const double* trig;  // [B][T + 1][4] (sin(w dt), cos(w dt), atan2 sensor 1, atan2 sensor 2) per nominal point
  long long tbs;
  const double* y;
  long long ty, sy;
  __device__ __forceinline__ StepCTB at(long long seq, long long k) const {
    StepCTB s{CTJac(), BearingsJac(), pr};
    const double* x = nom + seq * nbs + k * 5;
    const double* tr = trig + seq * tbs + k * 4;
    s.t.eval(x, tr, pr);
    s.o.eval(x + 5, tr + 4, pr);
    return s;
  }
  __device__ __forceinline__ const double* yp(long long seq, long long k) const { return y + seq * sy + k * ty; }
};
struct SrcFusedCT {
  FusedCTBParams pr;
  const double* nom;
  long long nbs;
  const double* trig;
  long long tbs;
  __device__ __forceinline__ StepCT at(long long seq, long long k) const {
    StepCT s{CTJac(), pr};
    s.t.eval(nom + seq * nbs + k * 5, trig + seq * tbs + k * 4, pr);
    return s;
  }
};

// one thread per nominal point: the four transcendental values the sweeps need (see the header comment)
__global__ void __launch_bounds__(128)
k_fused_trig(double dt, double s1x, double s1y, double s2x, double s2y, const double* __restrict__ nom, long long nbs,
             long long T1, double* __restrict__ trig) {
  const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long seq = blockIdx.y;
  if (k >= T1) return;
  const double* x = nom + seq * nbs + k * 5;
  double sw, cw;
  sincos(x[4] * dt, &sw, &cw);
  double* o = trig + (seq * T1 + k) * 4;
  o[0] = sw;
  o[1] = cw;
  o[2] = atan2(x[1] - s1y, x[0] - s1x);
  o[3] = atan2(x[1] - s2y, x[0] - s2x);
}

}  // namespace psq
