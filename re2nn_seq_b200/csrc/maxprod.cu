// Decompose i-FST recurrence under the max-product semiring (train_mode == 'max').
// Reference: /root/reference/src_seq/farnn/model_decompose_single.py:159-166 + utils.py:192-195 (_maxmul):
//   Tr_b = S1 diag(v_t) S2^T + W  (S x S, per sequence and step);  fwd h'[j] = max_s hbar[s] Tr[s][j],
//   bwd h'[s] = max_j hbar[j] Tr[s][j];  gates / mask / nonlinearity as in the sum semiring.
// The reference materialises B x S x S transition tensors per step; here one CTA owns a (sequence, direction),
// keeps the state in shared memory and forms Tr[s][j] in registers: warps stride the "from" rows s, lanes the
// "to" columns j (S2 is read through a transposed copy so every load is a contiguous 128 B line).
// O(S^2 R) per position by construction of the semiring; inference only (no backward).
#include "common.cuh"

namespace re2nn {

constexpr int kMxThreads = 512;
constexpr int kMxWarps = kMxThreads / 32;

__global__ void transpose_kernel(const float* __restrict__ src, int rows, int cols, float* __restrict__ dst) {
  __shared__ float t[32][33];
  int c = blockIdx.x * 32 + threadIdx.x, r = blockIdx.y * 32 + threadIdx.y;
  if (r < rows && c < cols) t[threadIdx.y][threadIdx.x] = src[(size_t)r * cols + c];
  __syncthreads();
  int rr = blockIdx.x * 32 + threadIdx.y, cc = blockIdx.y * 32 + threadIdx.x;
  if (rr < cols && cc < rows) dst[(size_t)rr * rows + cc] = t[threadIdx.x][threadIdx.y];
}

// y[j] = sum_s x[s] * M[s][j]  (lanes along j, warps stride s; result reduced through `part`)
__device__ void matvec_rows(const float* x, const float* __restrict__ M, int S, float* part, float* y, int tid) {
  const int warp = tid >> 5, lane = tid & 31;
  for (int jb = 0; jb < S; jb += 32) {
    const int j = jb + lane;
    if (j < S) {
      float acc = 0.f;
      for (int s = warp; s < S; s += kMxWarps) acc = fmaf(x[s], __ldg(M + (size_t)s * S + j), acc);
      part[warp * S + j] = acc;
    }
  }
  __syncthreads();
  for (int j = tid; j < S; j += kMxThreads) {
    float acc = 0.f;
#pragma unroll
    for (int w = 0; w < kMxWarps; ++w) acc += part[w * S + j];
    y[j] = acc;
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kMxThreads) decompose_max_kernel(const re2nn_recurrence_args a, const float* __restrict__ S1T,
                                                                   const float* __restrict__ S2T) {
  extern __shared__ float smem[];
  const int S = a.S, R = a.R;
  float* h = smem;                  // S  state
  float* hb = h + S;                // S  operand (after reset gate / * o)
  float* zt = hb + S;               // S
  float* tmp = zt + S;              // S
  float* v = tmp + S;               // R
  float* part = v + R;              // kMxWarps * S
  const int b = blockIdx.x, z = blockIdx.y;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n = (int)a.lengths[b];
  const float* __restrict__ o = a.o;
  const float* hinit = z == 0 ? a.h0 : a.hT;
  float* out = z == 0 ? a.alpha : a.beta;
  // fwd: Tr[s][j] = sum_r S1[s][r] v[r] S2[j][r] + W[s][j]; row factor F = S1, column factor G^T = S2T
  // bwd: h'[s] = max_j hb[j] Tr[s][j]: we iterate the CONTRACTED index over warps and the OUTPUT index over lanes
  //      in both directions, i.e. bwd uses row factor S2 (index j), column factor S1T (index s), W transposed.
  const float* __restrict__ F = z == 0 ? a.S1 : a.S2;          // [contracted][R]
  const float* __restrict__ GT = z == 0 ? S2T : S1T;           // [R][output]
  for (int s = tid; s < S; s += kMxThreads) {
    float x = hinit[s];
    h[s] = x;
    if (z == 1 && n >= 1 && n <= a.L) out[((size_t)b * a.L + (n - 1)) * S + s] = x;
  }
  __syncthreads();
  const int gw = S * a.farnn;
  for (int k = 0; k < a.L; ++k) {
    int tpos, orow;
    bool alive;
    step_pos(z, k, n, a.full_pad, tpos, orow, alive);
    if (!alive) break;
    const size_t vrow = a.v_mode == RE2NN_V_TOKEN ? (size_t)a.x[(size_t)b * a.Lpad + tpos] : (size_t)b * a.Lpad + tpos;
    for (int r = tid; r < R; r += kMxThreads) v[r] = __ldg(a.vtab + vrow * R + r);
    __syncthreads();
    // gates
    if (a.farnn >= 1) {
      matvec_rows(h, a.Wss1, S, part, zt, tid);
      for (int s = tid; s < S; s += kMxThreads) zt[s] = sigmoidf_((zt[s] + __ldg(a.gtab + vrow * gw + s)) * a.sigmoid_exponent);
      __syncthreads();
    }
    if (a.farnn == 2) {
      matvec_rows(h, a.Wss2, S, part, tmp, tid);
      for (int s = tid; s < S; s += kMxThreads) {
        float rt = sigmoidf_((tmp[s] + __ldg(a.gtab + vrow * gw + S + s)) * a.sigmoid_exponent);
        hb[s] = (1.f - rt) * hinit[s] + rt * h[s];
      }
    } else {
      for (int s = tid; s < S; s += kMxThreads) hb[s] = h[s];
    }
    __syncthreads();
    if (z == 1) {
      for (int s = tid; s < S; s += kMxThreads) hb[s] *= o[s];
      __syncthreads();
    }
    // max-product: out[q] = max_c hb[c] * (sum_r F[c][r] v[r] GT[r][q] + Wcq),  Wcq = W[c][q] (fwd) | W[q][c] (bwd)
    for (int qb = 0; qb < S; qb += 32) {
      const int q = qb + lane;
      float best = -INFINITY;
      if (q < S) {
        for (int c = warp; c < S; c += kMxWarps) {
          float acc = 0.f;
          const float* __restrict__ Fc = F + (size_t)c * R;
          for (int r = 0; r < R; ++r) acc = fmaf(__ldg(Fc + r) * v[r], __ldg(GT + (size_t)r * S + q), acc);
          acc += z == 0 ? __ldg(a.W + (size_t)c * S + q) : __ldg(a.W + (size_t)q * S + c);
          best = fmaxf(best, hb[c] * acc);
        }
        part[warp * S + q] = best;
      }
    }
    __syncthreads();
    for (int q = tid; q < S; q += kMxThreads) {
      float m = -INFINITY;
#pragma unroll
      for (int w = 0; w < kMxWarps; ++w) m = fmaxf(m, part[w * S + q]);
      if (z == 0) m *= o[q];
      m = apply_nl(m, a.update_nonlinear);
      float hn = a.farnn >= 1 ? (1.f - zt[q]) * h[q] + zt[q] * m : m;
      tmp[q] = hn;
      if (orow >= 0) out[((size_t)b * a.L + orow) * S + q] = hn;
    }
    __syncthreads();
    for (int s = tid; s < S; s += kMxThreads) h[s] = tmp[s];
    __syncthreads();
  }
}

}  // namespace re2nn

using namespace re2nn;

extern "C" size_t re2nn_decompose_max_workspace(int S, int R) { return (size_t)2 * S * R * sizeof(float) + 512; }

extern "C" int re2nn_decompose_max_recurrence(const re2nn_recurrence_args* a, void* stream) {
  RE2NN_CHECK(a != nullptr, "decompose_max_recurrence: null args");
  RE2NN_CHECK(a->B > 0 && a->L > 0 && a->S > 0 && a->R > 0 && a->L <= a->Lpad, "decompose_max_recurrence: bad dims");
  RE2NN_CHECK(a->lengths && a->vtab && a->S1 && a->S2 && a->W && a->o && a->h0 && a->hT && a->alpha && a->beta && a->ws,
              "decompose_max_recurrence: null tensor");
  RE2NN_CHECK(a->v_mode == RE2NN_V_DENSE || a->x, "decompose_max_recurrence: token mode needs x");
  RE2NN_CHECK(a->farnn == 0 || (a->gtab && a->Wss1), "decompose_max_recurrence: farnn>=1 needs gtab and Wss1");
  RE2NN_CHECK(a->farnn < 2 || a->Wss2, "decompose_max_recurrence: farnn==2 needs Wss2");
  RE2NN_CHECK(!a->save_for_backward, "decompose_max_recurrence: the max-product semiring has no backward");
  RE2NN_CHECK(a->ws_bytes >= re2nn_decompose_max_workspace(a->S, a->R), "decompose_max_recurrence: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  float* S1T = (float*)a->ws;
  float* S2T = S1T + (size_t)a->S * a->R;
  dim3 tb(32, 32), tg(cdiv(a->R, 32), cdiv(a->S, 32));
  transpose_kernel<<<tg, tb, 0, st>>>(a->S1, a->S, a->R, S1T);
  RE2NN_LAUNCH_CHECK();
  transpose_kernel<<<tg, tb, 0, st>>>(a->S2, a->S, a->R, S2T);
  RE2NN_LAUNCH_CHECK();
  const size_t smem = ((size_t)(4 + kMxWarps) * a->S + a->R) * sizeof(float);
  RE2NN_CHECK(smem <= 220 * 1024, "decompose_max_recurrence: S=%d too large", a->S);
  RE2NN_CUDA(cudaFuncSetAttribute(decompose_max_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  decompose_max_kernel<<<dim3(a->B, 2), kMxThreads, smem, st>>>(*a, S1T, S2T);
  RE2NN_LAUNCH_CHECK();
  return 0;
}
