// Dense-transition semiring step of the FST (non-independent) variants and of max-product training.
// Reference: /root/reference/src_seq/utils.py:192-199 (_matmul = bmm, _maxmul = max over the source state of
// h[b,j] * Tr[b,j,s]); call sites farnn/model_onehot.py:96-103,274-290, model_decompose_independent.py:177-181,
// model_decompose_single.py:159-166, model_decompose.py:268-276.
// The FST forms multiply a label mask into the transition matrix element-wise, which destroys the rank-R structure:
// every (sequence, step) owns a dense S x S matrix Tr[b] that is used exactly once, so the step is a batched
// vector-matrix product bound by reading Tr (HBM / L2): one CTA per sequence, lanes along the contiguous index.
#include "common.cuh"

namespace re2nn {

constexpr int kVmThreads = 256;

// out[b,s] = (+|max)_j h[b,j] * T[b,j,s]      (TRANS = 0: T indexed [source j][target s], lanes along s)
// out[b,s] = (+|max)_j h[b,j] * T[b,s,j]      (TRANS = 1: the transposed matrix, a warp per target row s)
// idx[b,s] (MAXP, optional) = FIRST source j attaining the maximum (torch.max(dim) tie-break).
template <bool MAXP, bool TRANS>
__global__ void __launch_bounds__(kVmThreads) batched_vecmat_kernel(const float* __restrict__ h, const float* __restrict__ T,
                                                                    int S, float* __restrict__ out, int* __restrict__ idx) {
  extern __shared__ float hs[];
  const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const float* Tb = T + (size_t)b * S * S;
  for (int j = tid; j < S; j += kVmThreads) hs[j] = h[(size_t)b * S + j];
  __syncthreads();
  if (!TRANS) {
    for (int s = tid; s < S; s += kVmThreads) {
      float acc = MAXP ? -INFINITY : 0.f;
      int bi = 0;
      for (int j = 0; j < S; ++j) {
        const float v = hs[j] * __ldg(Tb + (size_t)j * S + s);
        if (MAXP) {
          if (v > acc) { acc = v; bi = j; }
        } else {
          acc += v;
        }
      }
      out[(size_t)b * S + s] = acc;
      if (MAXP && idx) idx[(size_t)b * S + s] = bi;
    }
  } else {
    for (int s = warp; s < S; s += kVmThreads / 32) {
      const float* row = Tb + (size_t)s * S;
      float acc = MAXP ? -INFINITY : 0.f;
      int bi = 0x7fffffff;
      for (int j = lane; j < S; j += 32) {
        const float v = hs[j] * __ldg(row + j);
        if (MAXP) {
          if (v > acc) { acc = v; bi = j; }
        } else {
          acc += v;
        }
      }
      if (MAXP) {
        if (bi == 0x7fffffff) bi = lane;
        warp_argmax_first(acc, bi);
        if (acc == -INFINITY) bi = 0;
      } else {
        acc = warp_sum(acc);
      }
      if (lane == 0) {
        out[(size_t)b * S + s] = acc;
        if (MAXP && idx) idx[(size_t)b * S + s] = bi;
      }
    }
  }
}

}  // namespace re2nn

using namespace re2nn;

extern "C" int re2nn_batched_vecmat(const float* h, const float* T, int B, int S, int transposed, int max_semiring,
                                    float* out, int32_t* argmax_out, void* stream) {
  RE2NN_CHECK(h && T && out && B > 0 && S > 0, "batched_vecmat: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t smem = (size_t)S * 4;
  RE2NN_CHECK(smem <= 48 * 1024, "batched_vecmat: S = %d too large", S);
  if (max_semiring) {
    if (transposed) batched_vecmat_kernel<true, true><<<B, kVmThreads, smem, st>>>(h, T, S, out, argmax_out);
    else batched_vecmat_kernel<true, false><<<B, kVmThreads, smem, st>>>(h, T, S, out, argmax_out);
  } else {
    if (transposed) batched_vecmat_kernel<false, true><<<B, kVmThreads, smem, st>>>(h, T, S, out, nullptr);
    else batched_vecmat_kernel<false, false><<<B, kVmThreads, smem, st>>>(h, T, S, out, nullptr);
  }
  RE2NN_LAUNCH_CHECK();
  return 0;
}
