// Shared helpers for the re2nn_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/re2nn_b200.h"

namespace re2nn {

// ---- error reporting (C-ABI never aborts) -----------------------------------------------------
extern thread_local char g_err[512];
int set_error(const char* fmt, ...);

#define RE2NN_CHECK(cond, ...)                         \
  do {                                                 \
    if (!(cond)) return ::re2nn::set_error(__VA_ARGS__); \
  } while (0)

#define RE2NN_CUDA(expr)                                                                       \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess)                                                                     \
      return ::re2nn::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

#define RE2NN_LAUNCH_CHECK() RE2NN_CUDA(cudaGetLastError())

inline int cdiv(int a, int b) { return (a + b - 1) / b; }

// ---- per-device launch facts (a process may drive several GPUs; nothing here is a process-wide constant) ----
constexpr int kMaxDevices = 64;
inline int device_slot() {
  int d = 0;
  if (cudaGetDevice(&d) != cudaSuccess || d < 0) d = 0;
  return d % kMaxDevices;
}
inline int sm_count() {
  static int cache[kMaxDevices];      // 0 = not queried yet; a racing second query writes the same value
  const int d = device_slot();
  if (cache[d] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, d) != cudaSuccess || n <= 0) n = 148;
    cache[d] = n;
  }
  return cache[d];
}
// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device (per-context) attribute: `done` is the caller's
// per-kernel array of the largest size already configured on each device.
template <class Kernel>
inline cudaError_t ensure_dynamic_smem(Kernel kernel, int bytes, int (&done)[kMaxDevices]) {
  const int d = device_slot();
  if (done[d] >= bytes) return cudaSuccess;
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess) done[d] = bytes;
  return e;
}
inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ---- nonlinearities ---------------------------------------------------------------------------
__device__ __forceinline__ float apply_nl(float x, int kind) {
  switch (kind) {
    case RE2NN_NL_RELU: return fmaxf(x, 0.f);
    case RE2NN_NL_TANH: return tanhf(x);
    case RE2NN_NL_RELUTANH: return tanhf(fmaxf(x, 0.f));
    case RE2NN_NL_SIGMOID: return 1.f / (1.f + expf(-x));
    default: return x;
  }
}
// derivative expressed through the OUTPUT y = phi(x) (and x only where needed for relu masks)
__device__ __forceinline__ float nl_grad_from_out(float y, int kind) {
  switch (kind) {
    case RE2NN_NL_RELU: return y > 0.f ? 1.f : 0.f;
    case RE2NN_NL_TANH: return 1.f - y * y;
    case RE2NN_NL_RELUTANH: return y > 0.f ? 1.f - y * y : 0.f;
    case RE2NN_NL_SIGMOID: return y * (1.f - y);
    default: return 1.f;
  }
}
__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// hardware approximations (MUFU): ~2^-11 relative error, used only where operands are bf16 anyway
__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// Branch-free tanh for the parity-grade tensor-core epilogues: |x| < 0.55 an odd polynomial (1.2 ulp), otherwise
// 1 - 2 / (e^{2|x|} + 1) with ex2.approx / rcp.approx; <= 4e-7 relative overall (libm's tanhf is ~1e-7 but carries a
// divergent branch per call: ~45 instructions per element against ~16).  A five-instruction form without the
// polynomial (absolute error 2e-7, relative error unbounded near 0) was measured too: it breaks the 1e-5 score bound
// of the 128-step recurrence (cfg4, tf32x3) and raises the cfg5 tag differences against the fp32 path from 52 to
// 1092 of 4.2 M -- and buys nothing, the epilogue being bound by its L2 write traffic, not by issue slots.
__device__ __forceinline__ float sigmoid_ulp(float x);
__device__ __forceinline__ float tanh_bfree(float x) {
  const float t = fabsf(x), u = x * x;
  float q = fmaf(u, -6.615782622e-03f, 2.131276391e-02f);
  q = fmaf(q, u, -5.391006917e-02f);
  q = fmaf(q, u, 1.333311796e-01f);
  q = fmaf(q, u, -3.333333135e-01f);
  const float small = fmaf(x * u, q, x);
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(t * 2.885390082f));       // e^{2t}
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e + 1.f));
  const float big = copysignf(fmaf(-2.f, r, 1.f), x);
  return t < 0.55f ? small : big;
}
// MODE 0: libm (fp32 path, training), 1: MUFU tanh.approx (bf16 operands), 2: branch-free few-ulp (tf32x3 / fp16x3)
template <int MODE>
__device__ __forceinline__ float apply_nl_t(float x, int kind) {
  if (MODE == 0) return apply_nl(x, kind);
  switch (kind) {
    case RE2NN_NL_RELU: return fmaxf(x, 0.f);
    case RE2NN_NL_TANH: return MODE == 1 ? tanh_fast(x) : tanh_bfree(x);
    case RE2NN_NL_RELUTANH: return MODE == 1 ? tanh_fast(fmaxf(x, 0.f)) : tanh_bfree(fmaxf(x, 0.f));
    case RE2NN_NL_SIGMOID: return MODE == 1 ? 0.5f * tanh_fast(0.5f * x) + 0.5f : sigmoid_ulp(x);
    default: return x;
  }
}
template <bool FAST>
__device__ __forceinline__ float sigmoid_t(float x) {
  return FAST ? 0.5f * tanh_fast(0.5f * x) + 0.5f : sigmoidf_(x);
}
// Branch-free sigmoid for the split-format tensor-core epilogues: ex2.approx + rcp.approx (a few ulp, ~3e-7
// relative).  The IEEE division of sigmoidf_ carries a slow-path branch per element, which serialises the
// unrolled 16-row batches of the tcgen05 epilogue (the gate GEMM ran 3x slower than its operand stream).
__device__ __forceinline__ float sigmoid_ulp(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
// MODE 0: exact (fp32 path), 1: MUFU tanh (bf16), 2: few-ulp branch-free (tf32x3 / fp16x3)
template <int MODE>
__device__ __forceinline__ float sigmoid_m(float x) {
  return MODE == 1 ? 0.5f * tanh_fast(0.5f * x) + 0.5f : (MODE == 2 ? sigmoid_ulp(x) : sigmoidf_(x));
}

// ---- step <-> position mapping ----------------------------------------------------------------
// The reference runs the backward direction on reverse(input, lengths) and un-reverses the states
// afterwards (model_decompose_single.py:222,257-261).  We keep its step order but address tokens
// and outputs directly:
//   dir 0 (forward):  step k consumes token k, produces alpha[.,k]
//   dir 1 (backward): step k<n consumes token n-1-k, produces beta[., n-2-k] (k==n-1 -> beta_0, unused);
//                     step k>=n consumes pad token k, produces the reference's pad row k.
// `alive` = the row's state still matters at this step; `orow` = output row or -1.
__device__ __forceinline__ void step_pos(int dir, int k, int n, int full_pad, int& tpos, int& orow, bool& alive) {
  if (dir == 0) {
    tpos = k;
    alive = full_pad || (k < n);
    orow = alive ? k : -1;
  } else {
    if (k < n) {
      tpos = n - 1 - k;
      orow = n - 2 - k;               // -1 when k == n-1
      alive = full_pad || (k < n - 1);
    } else {
      tpos = k;
      alive = full_pad != 0;
      orow = alive ? k : -1;
    }
  }
}

// warp reductions
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// (value, index) arg-max with FIRST-index tie-break, as torch.max(dim) does
__device__ __forceinline__ void warp_argmax_first(float& v, int& i) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    float ov = __shfl_xor_sync(0xffffffffu, v, o);
    int oi = __shfl_xor_sync(0xffffffffu, i, o);
    if (ov > v || (ov == v && oi < i)) { v = ov; i = oi; }
  }
}

}  // namespace re2nn
