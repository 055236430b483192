// xla_ffi_shim.cc -- registers libpsqrt.so's whole-pass entry point as an XLA FFI custom call, so that
// parsmooth's `parallel=True, sqrt=True` path can dispatch to it inside jax.jit with no host round trip (XLA
// passes device buffers and its CUDA stream).  This is the binding a parsmooth maintainer would add next to
// parsmooth/parallel/_filtering.py:30-35 and _smoothing.py:30-34 (see INTEGRATION.md section 3 for the Python side).
//
// Compile-guarded: the XLA FFI headers ship with jaxlib >= 0.4.31 (`jaxlib/include/xla/ffi/api/ffi.h`).  This image
// has no jaxlib, so build.py only compiles this file when PSQRT_XLA_INCLUDE points at such an include directory:
//   g++ -std=c++17 -O2 -shared -fPIC -DPSQRT_HAVE_XLA_FFI -I$PSQRT_XLA_INCLUDE -I include \
//       sqrt-parallel-smoothers_b200/csrc/xla/xla_ffi_shim.cc -L sqrt-parallel-smoothers_b200/psqrt -lpsqrt \
//       -o sqrt-parallel-smoothers_b200/psqrt/libpsqrt_xla.so
// Without the macro the translation unit is empty (so that it is at least syntax-checked against psqrt.h below).
#include "psqrt.h"

#if defined(PSQRT_HAVE_XLA_FFI)
#include <cuda_runtime_api.h>

#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

namespace {

using F64 = ffi::Buffer<ffi::F64>;

// time stride of a per-step model array: 0 when the array is time-invariant (rank == core rank)
inline int64_t time_stride(const F64& a, int core, int64_t T) {
  return (int64_t)a.dimensions().size() == core ? 0 : (int64_t)(a.element_count() / T);
}

// Inputs: the linearised state-space model exactly as vmap(linearization_method) returns it
// (parallel/_filtering.py:117-119), observations y [T, ny], prior (m0 [nx], L0 [nx, nx] lower triangular).
// Results: filtered and smoothed trajectories [T + 1, ...], log-likelihood [1], scratch of psqrt_workspace_bytes.
ffi::Error FilterSmootherImpl(cudaStream_t stream, F64 F, F64 cholQ, F64 b, F64 H, F64 cholR, F64 c, F64 y, F64 m0,
                              F64 L0, ffi::Result<F64> fm, ffi::Result<F64> fL, ffi::Result<F64> sm,
                              ffi::Result<F64> sL, ffi::Result<F64> ell, ffi::Result<ffi::Buffer<ffi::U8>> ws) {
  const int64_t T = y.dimensions()[0];
  const int ny = (int)y.dimensions()[1];
  const int nx = (int)m0.dimensions()[0];
  psqrt_ssm s = {F.typed_data(), cholQ.typed_data(), b.typed_data(), H.typed_data(), cholR.typed_data(),
                 c.typed_data(),
                 time_stride(F, 2, T), time_stride(cholQ, 2, T), time_stride(b, 1, T), time_stride(H, 2, T),
                 time_stride(cholR, 2, T), time_stride(c, 1, T),
                 0, 0, 0, 0, 0, 0,                                       // one sequence: no batch strides
                 nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};  // no host mirrors inside XLA
  const int rc = psqrt_filter_smoother(&s, y.typed_data(), m0.typed_data(), L0.typed_data(), nx, ny, T, 1, 0,
                                       fm->typed_data(), fL->typed_data(), sm->typed_data(), sL->typed_data(),
                                       ell->typed_data(), ws->typed_data(), ws->element_count(), stream);
  if (rc != PSQRT_OK) return ffi::Error(ffi::ErrorCode::kInternal, psqrt_error_string(rc));
  return ffi::Error::Success();
}

// smoothing(...) on an existing filtered trajectory (parallel/_smoothing.py:14-44)
ffi::Error SmootherImpl(cudaStream_t stream, F64 F, F64 cholQ, F64 b, F64 fm, F64 fL, ffi::Result<F64> sm,
                        ffi::Result<F64> sL, ffi::Result<ffi::Buffer<ffi::U8>> ws) {
  const int64_t T = fm.dimensions()[0] - 1;
  const int nx = (int)fm.dimensions()[1];
  psqrt_ssm s = {F.typed_data(), cholQ.typed_data(), b.typed_data(), nullptr, nullptr, nullptr,
                 time_stride(F, 2, T), time_stride(cholQ, 2, T), time_stride(b, 1, T), 0, 0, 0,
                 0, 0, 0, 0, 0, 0, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  const int rc = psqrt_smoother(&s, fm.typed_data(), fL.typed_data(), nx, T, 1, 0, sm->typed_data(), sL->typed_data(),
                                ws->typed_data(), ws->element_count(), stream);
  if (rc != PSQRT_OK) return ffi::Error(ffi::ErrorCode::kInternal, psqrt_error_string(rc));
  return ffi::Error::Success();
}

}  // namespace

XLA_FFI_DEFINE_HANDLER_SYMBOL(PsqrtFilterSmoother, FilterSmootherImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>()   // F cholQ b H cholR c
                                  .Arg<F64>().Arg<F64>().Arg<F64>()                                      // y m0 L0
                                  .Ret<F64>().Ret<F64>().Ret<F64>().Ret<F64>().Ret<F64>()               // fm fL sm sL ell
                                  .Ret<ffi::Buffer<ffi::U8>>());                                        // workspace

XLA_FFI_DEFINE_HANDLER_SYMBOL(PsqrtSmoother, SmootherImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>()   // F cholQ b fm fL
                                  .Ret<F64>().Ret<F64>()                                     // sm sL
                                  .Ret<ffi::Buffer<ffi::U8>>());
#endif  // PSQRT_HAVE_XLA_FFI
