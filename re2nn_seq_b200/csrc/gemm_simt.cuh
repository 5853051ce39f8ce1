// fp32 CUDA-core step GEMM with fused epilogue (RE2NN_PREC_FP32: the reference's own operand
// precision, every product an fp32 FFMA).  128x64 tile, 16-deep K slices staged in shared memory,
// 8x4 register micro-tile per thread.  Accumulation order is sequential in k => deterministic.
#pragma once
#include "gemm_common.cuh"

namespace re2nn {

struct ALoadPlain {
  __device__ __forceinline__ float operator()(const GemmSeg& s, int, int m, int k) const {
    return __ldg((const float*)s.A + (size_t)m * s.lda + k);
  }
};

// A[m,k] = alpha[m,k] * beta[m,k] for valid rows (m = b*L + t, t < len[b]), else 0
struct ALoadAlphaBeta {
  const float* alpha;
  const float* beta;
  const int64_t* len;
  int L, full_pad;
  __device__ __forceinline__ float operator()(const GemmSeg& s, int, int m, int k) const {
    int b = m / L, t = m - b * L;
    if (!full_pad && t >= (int)len[b]) return 0.f;
    size_t i = (size_t)m * s.lda + k;
    return __ldg(alpha + i) * __ldg(beta + i);
  }
};

// BM = 128 (8x4 micro-tile) for full grids, BM = 64 (4x4) when the 128-row tiling would leave SMs idle.
template <int BM, class Epi, class ALoad>
__global__ void __launch_bounds__(256) simt_gemm_kernel(const GemmProblem prob, const Epi epi_in, const ALoad aload) {
  constexpr int BN = 64, BK = 16, TM = BM / 16, TN = 4;
  const int z = blockIdx.z;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int M = prob.M, N = prob.N;
  const int rows = min(BM, M - m0);
  if (!epi_in.tile_alive(z, (int)(blockIdx.x * BM) / 128)) return;   // liveness is tracked per 128 rows
  const Epi epi = epi_in.for_dir(z);

  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  for (int sg = 0; sg < prob.nseg; ++sg) {
    const GemmSeg s = prob.seg[z][sg];
    const float* Bp = (const float*)s.B;
    for (int k0 = 0; k0 < s.K; k0 += BK) {
#pragma unroll
      for (int i = 0; i < (BM * BK) / 256; ++i) {
        int idx = tid + i * 256;
        int kk = idx & (BK - 1), mm = idx >> 4;
        float v = 0.f;
        if (mm < rows && k0 + kk < s.K) v = aload(s, z, m0 + mm, k0 + kk);
        As[kk][mm] = v;
      }
      if (s.b_nk) {
#pragma unroll
        for (int i = 0; i < (BN * BK) / 256; ++i) {
          int idx = tid + i * 256;
          int kk = idx & (BK - 1), nn = idx >> 4;
          float v = 0.f;
          if (n0 + nn < N && k0 + kk < s.K) v = __ldg(Bp + (size_t)(n0 + nn) * s.ldb + k0 + kk);
          Bs[kk][nn] = v;
        }
      } else {
#pragma unroll
        for (int i = 0; i < (BN * BK) / 256; ++i) {
          int idx = tid + i * 256;
          int nn = idx & (BN - 1), kk = idx >> 6;
          float v = 0.f;
          if (n0 + nn < N && k0 + kk < s.K) v = __ldg(Bp + (size_t)(k0 + kk) * s.ldb + n0 + nn);
          Bs[kk][nn] = v;
        }
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < BK; ++kk) {
        float a[TM], b[TN];
        const float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * TM]);
        const float4 b0 = *reinterpret_cast<const float4*>(&Bs[kk][tx * TN]);
        a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w;
        if constexpr (TM == 8) {
          const float4 a1 = *reinterpret_cast<const float4*>(&As[kk][ty * TM + 4]);
          a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
        }
        b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w;
#pragma unroll
        for (int i = 0; i < TM; ++i)
#pragma unroll
          for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
      __syncthreads();
    }
  }

#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int m = m0 + ty * TM + i;
    if (m >= M) continue;
    const RowCtx r = epi.row(m);
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int n = n0 + tx * TN + j;
      if (n < N) epi.apply(epi.col(n), r, m, n, acc[i][j], epi.prefetch(r, m, n));
    }
  }
}

template <class Epi, class ALoad>
inline cudaError_t launch_simt_gemm(const GemmProblem& prob, const Epi& epi, const ALoad& aload, cudaStream_t st) {
  const long ctas128 = (long)cdiv(prob.M, 128) * cdiv(prob.N, 64) * prob.ndir;
  if (ctas128 >= sm_count()) {
    dim3 grid(cdiv(prob.M, 128), cdiv(prob.N, 64), prob.ndir);
    simt_gemm_kernel<128, Epi, ALoad><<<grid, 256, 0, st>>>(prob, epi, aload);
  } else {
    dim3 grid(cdiv(prob.M, 64), cdiv(prob.N, 64), prob.ndir);
    simt_gemm_kernel<64, Epi, ALoad><<<grid, 256, 0, st>>>(prob, epi, aload);
  }
  return cudaGetLastError();
}

}  // namespace re2nn
