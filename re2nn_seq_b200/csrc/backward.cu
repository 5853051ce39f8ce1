// Hand-written backward of the decompose i-FST path (fp32): label scores -> BPTT through both
// recurrences (gates included) -> parameter gradients.  Replaces torch.autograd over
// FARNN_S_D_W_I_S.forward_local / FARNN_S_SF.forward
// (/root/reference/src_seq/farnn/model_decompose_single.py:138-269, train_decompose.py:192).
//
// Per step k = L-1 .. 0 (both directions in every launch):
//   E1   G = g + dOut ; gate blend grads ; d pre-activation ; DA[k]                    (elementwise)
//   GEMM dq  = DA[k] @ S2|S1
//   E2   DU[k] = dq*v ; Q[k] = u*v ; dvtab[token] += dq*u                               (elementwise)
//   GEMM dhb = DU[k] @ S1^T|S2^T + DA[k] @ W^T|W
//   E3   reset-gate grads ; g <- carry + dh~ ; o / h_init products                      (elementwise)
//   GEMM g  += [dz|dr] @ [Wss1|Wss2]^T                                                  (farnn >= 1)
// Without gates E3 of step k and E1 of step k-1 are one launch (bwd_e31_kernel), the step GEMMs run on tcgen05 in
// 3xTF32, and the elementwise kernels between them are launched with programmatic stream serialization like the GEMMs.
// After the sweep every weight gradient is ONE large transposed GEMM over all (step, sequence) rows
// (K = 2*L*B) -- on tcgen05 with the operands transposed on the fly (gemm_tn_tc.cuh), CUDA cores for short reductions --
// reduced deterministically (split-K partials + ordered sum); bias / vector gradients are deterministic column sums
// of the per-step slabs.  The whole sweep also exists as ONE resident launch (ResidentBackward below): parity-
// identical, slower at the benchmarked batch, off by default.
#include <algorithm>
#include <memory>

#include "gemm_simt.cuh"
#include "gemm_tc.cuh"
#include "recurrence_resident.cuh"
#include "gemm_tn_tc.cuh"

namespace re2nn {

// ---- epilogues local to the backward ---------------------------------------------------------------------
struct EpiStore2 {          // C_z = acc (per direction output pointers)
  float* C[2];
  int ldc;
  int accumulate;           // C += acc
  __device__ __forceinline__ EpiStore2 for_dir(int z) const { EpiStore2 e = *this; e.C[0] = C[z]; return e; }
  __device__ __forceinline__ bool tile_alive(int, int) const { return true; }
  __device__ __forceinline__ RowCtx row(int) const { return RowCtx{0, 0, true}; }
  __device__ __forceinline__ Col col(int) const { return Col{0.f, 0.f}; }
  __device__ __forceinline__ Pre prefetch(const RowCtx&, int m, int n) const {
    return Pre{accumulate ? C[0][(size_t)m * ldc + n] : 0.f, 0.f};
  }
  __device__ __forceinline__ float compute(const Col&, float acc, const Pre& pre) const { return acc + pre.a; }
  __device__ __forceinline__ void store(const Col&, const RowCtx&, int m, int n, float v, float, const Pre&) const {
    C[0][(size_t)m * ldc + n] = v;
  }
  __device__ __forceinline__ void apply(const Col&, const RowCtx&, int m, int n, float acc, const Pre& pre) const {
    C[0][(size_t)m * ldc + n] = acc + pre.a;
  }
};

// dAB = draw @ C ; dAlpha = dAB * beta ; dBeta = dAB * alpha  (valid rows only)
struct EpiDAB {
  const float* alpha; const float* beta; const int64_t* len;
  float* dalpha; float* dbeta;
  int L, S, full_pad;
  __device__ __forceinline__ EpiDAB for_dir(int) const { return *this; }
  __device__ __forceinline__ bool tile_alive(int, int) const { return true; }
  __device__ __forceinline__ RowCtx row(int m) const {
    int b = m / L, t = m - b * L;
    return RowCtx{0, (full_pad || t < (int)len[b]) ? 0 : -1, true};
  }
  __device__ __forceinline__ Col col(int) const { return Col{0.f, 0.f}; }
  __device__ __forceinline__ Pre prefetch(const RowCtx& r, int m, int n) const {
    if (r.orow < 0) return Pre{0.f, 0.f};
    size_t i = (size_t)m * S + n;
    return Pre{alpha[i], beta[i]};
  }
  __device__ __forceinline__ void apply(const Col&, const RowCtx& r, int m, int n, float acc, const Pre& pre) const {
    size_t i = (size_t)m * S + n;
    dalpha[i] = r.orow < 0 ? 0.f : acc * pre.b;
    dbeta[i] = r.orow < 0 ? 0.f : acc * pre.a;
  }
  // compute / store split of the tcgen05 epilogue (batches of rows: all arithmetic, then all stores)
  static constexpr bool kRows = false;
  __device__ __forceinline__ float compute(const Col&, float acc, const Pre&) const { return acc; }
  __device__ __forceinline__ void store(const Col& c, const RowCtx& r, int m, int n, float v, float, const Pre& pre) const {
    apply(c, r, m, n, v, pre);
  }
};

// ---- transposed ("TN") GEMM: C[P x Q] = sum over pairs, rows m:  X[m, p] * Y[m, q]  ---------------------------
// (TnPair / TnProblem: gemm_tn_tc.cuh, which holds the tcgen05 version of this kernel)

// XLoad hook lets dC form (alpha*beta) on the fly
struct YLoadPlain {
  static constexpr bool kPlain = true;
  __device__ __forceinline__ float operator()(const TnPair& pr, size_t m, int q) const { return __ldg(pr.Y + m * pr.ldy + q); }
};
struct YLoadAlphaBeta {
  static constexpr bool kPlain = false;      // CUDA-core kernel: forms the masked alpha * beta itself (the tcgen05 one reads Y2 / mask_len)
  const float* beta; const int64_t* len; int L, full_pad;
  __device__ __forceinline__ float operator()(const TnPair& pr, size_t m, int q) const {
    int b = (int)(m / L), t = (int)(m - (size_t)b * L);
    if (!full_pad && t >= (int)len[b]) return 0.f;
    return __ldg(pr.Y + m * pr.ldy + q) * __ldg(beta + m * pr.ldy + q);
  }
};

constexpr int kTnSplit = 64;

template <class YLoad>
__global__ void __launch_bounds__(256) tn_gemm_kernel(const TnProblem prob, float* __restrict__ partial, const YLoad yload) {
  constexpr int BP = 64, BQ = 64, BK = 16;
  __shared__ __align__(16) float Xs[BK][BP + 4];
  __shared__ __align__(16) float Ys[BK][BQ + 4];
  const int p0 = blockIdx.x * BP, q0 = blockIdx.y * BQ, split = blockIdx.z;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int pi = 0; pi < prob.npairs; ++pi) {
    const TnPair pr = prob.pair[pi];
    const size_t chunk = (pr.rows + kTnSplit - 1) / kTnSplit;
    const size_t r0 = chunk * split, r1 = min(pr.rows, r0 + chunk);
    for (size_t m0 = r0; m0 < r1; m0 += BK) {
#pragma unroll
      for (int i = 0; i < (BP * BK) / 256; ++i) {
        int idx = tid + i * 256;
        int pp = idx & (BP - 1), kk = idx >> 6;
        float v = 0.f;
        if (m0 + kk < r1 && p0 + pp < prob.P) v = __ldg(pr.X + (m0 + kk) * pr.ldx + p0 + pp);
        Xs[kk][pp] = v;
      }
#pragma unroll
      for (int i = 0; i < (BQ * BK) / 256; ++i) {
        int idx = tid + i * 256;
        int qq = idx & (BQ - 1), kk = idx >> 6;
        float v = 0.f;
        if (m0 + kk < r1 && q0 + qq < prob.Q) v = yload(pr, m0 + kk, q0 + qq);
        Ys[kk][qq] = v;
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < BK; ++kk) {
        const float4 a = *reinterpret_cast<const float4*>(&Xs[kk][ty * 4]);
        const float4 b = *reinterpret_cast<const float4*>(&Ys[kk][tx * 4]);
        const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
      }
      __syncthreads();
    }
  }
  float* out = partial + (size_t)split * prob.P * prob.Q;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int pp = p0 + ty * 4 + i;
    if (pp >= prob.P) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int qq = q0 + tx * 4 + j;
      if (qq < prob.Q) out[(size_t)pp * prob.Q + qq] = acc[i][j];
    }
  }
}

// out[i] (+)= sum_s partial[s][i]   (ordered => deterministic)
__global__ void split_reduce_kernel(const float* __restrict__ partial, int nsplit, size_t n, float* __restrict__ out,
                                    int accumulate) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    float acc = accumulate ? out[i] : 0.f;
    for (int s = 0; s < nsplit; ++s) acc += partial[(size_t)s * n + i];
    out[i] = acc;
  }
}

static bool g_tn_tc = true;      // long reductions on tcgen05 (gemm_tn_tc.cuh); re2nn_debug_set_tn_tc
extern "C" int re2nn_has_tcgen05(void);

template <class YLoad>
static cudaError_t run_tn(const TnProblem& prob, float* partial, float* out, int accumulate, const YLoad& yl, cudaStream_t st) {
  const size_t n = (size_t)prob.P * prob.Q;
  size_t rows = 0;
  for (int i = 0; i < prob.npairs; ++i) rows += prob.pair[i].rows;
#ifdef RE2NN_HAVE_TC
  // tensor cores once the reduction is long enough to fill a pipeline per CTA (the partial buffer holds kTnSplit slices)
  if ((YLoad::kPlain || prob.Y2 != nullptr) && g_tn_tc && re2nn_has_tcgen05() != 0 && rows >= 4096 && prob.P >= 16 && prob.Q >= 16) {
    const TnTcPlan pl = tn_tc_plan(prob, kTnSplit);
    if (pl.stages >= 2) {
      if (cudaError_t e = launch_tn_tc(prob, partial, pl, st)) return e;
      split_reduce_kernel<<<(unsigned)std::min<size_t>((n + 255) / 256, 1184), 256, 0, st>>>(partial, pl.nsplit, n, out, accumulate);
      return cudaGetLastError();
    }
  }
#endif
  dim3 grid(cdiv(prob.P, 64), cdiv(prob.Q, 64), kTnSplit);
  tn_gemm_kernel<YLoad><<<grid, 256, 0, st>>>(prob, partial, yl);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  split_reduce_kernel<<<(unsigned)std::min<size_t>((n + 255) / 256, 1184), 256, 0, st>>>(partial, kTnSplit, n, out, accumulate);
  return cudaGetLastError();
}

// deterministic column sums: out[c] (+)= sum_r X[r, c].  grid (column blocks of 32, row splits): every CTA sums its
// contiguous row range for 32 columns (32 x 32 threads, fixed order); with more than one split the partials go to
// `part[split][c]` and colsum_finish_kernel adds them in split order, so the result does not depend on scheduling.
__global__ void __launch_bounds__(1024) colsum_rows_kernel(const float* __restrict__ X, size_t rows, int cols, int ld,
                                                           float* __restrict__ out, int accumulate,
                                                           float* __restrict__ part) {
  __shared__ float sh[32][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  const size_t per = (rows + gridDim.y - 1) / gridDim.y;
  const size_t r0 = (size_t)blockIdx.y * per, r1 = r0 + per < rows ? r0 + per : rows;
  float acc = 0.f;
  if (c < cols)
    for (size_t r = r0 + threadIdx.y; r < r1; r += 32) acc += X[r * ld + c];
  sh[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && c < cols) {
    float t = (part == nullptr && accumulate) ? out[c] : 0.f;
    for (int i = 0; i < 32; ++i) t += sh[i][threadIdx.x];
    if (part) part[(size_t)blockIdx.y * cols + c] = t;
    else out[c] = t;
  }
}
__global__ void colsum_finish_kernel(const float* __restrict__ part, int nsplit, int cols, float* __restrict__ out,
                                     int accumulate) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  float t = accumulate ? out[c] : 0.f;
  for (int i = 0; i < nsplit; ++i) t += part[(size_t)i * cols + c];
  out[c] = t;
}
// ws: optional scratch of at least 64 * cols floats; tall inputs are then split over up to 64 CTAs per column block
static cudaError_t colsum_rows(const float* X, size_t rows, int cols, int ld, float* out, int accumulate, cudaStream_t st,
                               float* ws = nullptr) {
  const int nsplit = ws ? (int)std::min<size_t>(64, (rows + 2047) / 2048) : 1;
  if (nsplit <= 1) {
    colsum_rows_kernel<<<dim3(cdiv(cols, 32), 1), dim3(32, 32), 0, st>>>(X, rows, cols, ld, out, accumulate, nullptr);
    return cudaGetLastError();
  }
  colsum_rows_kernel<<<dim3(cdiv(cols, 32), nsplit), dim3(32, 32), 0, st>>>(X, rows, cols, ld, out, accumulate, ws);
  colsum_finish_kernel<<<cdiv(cols, 256), 256, 0, st>>>(ws, nsplit, cols, out, accumulate);
  return cudaGetLastError();
}

// ---- per-step elementwise kernels ------------------------------------------------------------------------------
struct BwdCtx {
  int B, Lpad, L, S, R, k, farnn, nl, v_mode, full_pad;
  float kappa;
  const int64_t* x; const int64_t* len;
  const float *vtab, *o, *h0, *hT;
  const float *dalpha, *dbeta;                       // B x L x S
  const float *hst_save, *u_save, *a_save, *zsave, *rsave;
  float *g, *GA;                                     // 2 x B x S
  float *DA, *DZR, *DOprod, *Pinit;                  // slabs 2 x L x B x (S | S*farnn)
  float *DU, *Qs;                                    // slabs 2 x L x B x R
  float *DQ, *DHb;                                   // 2 x B x R, 2 x B x S
  float *dvtab, *dgtab;
  // tensor-core sweep: 3xTF32 operand-format copies of this step's DA / DU (A operands of the two step GEMMs)
  void* DAop[2]; void* DUop[2];
  void* DZop[2]; void* DRop[2];                      // farnn >= 1: dz / dr of this step (A operands of the gate GEMM)
  int ldS, ldR;
  size_t da_plane, du_plane;
  // resident sweep: second DA operand buffer (ping-pong between steps) and the per-tile last live step
  void* DAop1[2];
  int* tile_last[2];
  void* DrawOp; void* CmatOp;                        // tensor-core label-score backward: operand-format draw / C_mat^T
};

__device__ __forceinline__ size_t slab(const BwdCtx& c, int z, int k, int width) {
  return ((size_t)z * c.L + k) * c.B * width;
}
__device__ __forceinline__ size_t slab1(const BwdCtx& c, int z, int k, int width) {   // (L+1)-deep save slabs
  return ((size_t)z * (c.L + 1) + k) * c.B * width;
}

__global__ void bwd_e1_kernel(const BwdCtx c) {
  const size_t total = (size_t)2 * c.B * c.S;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int s = (int)(i % c.S);
    const size_t rr = i / c.S;
    const int b = (int)(rr % c.B), z = (int)(rr / c.B);
    int tpos, orow;
    bool alive;
    step_pos(z, c.k, (int)c.len[b], c.full_pad, tpos, orow, alive);
    const size_t e = (size_t)b * c.S + s;
    const size_t sl = slab(c, z, c.k, c.S) + e;
    const int gw = c.S * c.farnn;
    if (!alive) {
      c.DA[sl] = 0.f;
      if (c.DAop[0]) OperandFmt<RE2NN_PREC_TF32X3>::store(c.DAop[z], (size_t)b * c.ldS + s, c.da_plane, 0.f);
      if (z == 0) c.DOprod[sl] = 0.f;
      if (c.farnn >= 1) {
        c.DZR[slab(c, z, c.k, gw) + (size_t)b * gw + s] = 0.f;
        if (c.DZop[0]) OperandFmt<RE2NN_PREC_TF32X3>::store(c.DZop[z], (size_t)b * c.ldS + s, c.da_plane, 0.f);
      }
      c.GA[(size_t)z * c.B * c.S + e] = 0.f;
      continue;
    }
    const float* dout = z == 0 ? c.dalpha : c.dbeta;
    const float G = c.g[(size_t)z * c.B * c.S + e] + (orow >= 0 ? dout[((size_t)b * c.L + orow) * c.S + s] : 0.f);
    const float a = c.a_save[sl];
    const float on = c.o[s];
    const float hhat = apply_nl(z == 0 ? a * on : a, c.nl);
    float dhhat = G, gA = 0.f;
    if (c.farnn >= 1) {
      const float zt = c.zsave[sl];
      const float hk = c.hst_save[slab1(c, z, c.k, c.S) + e];
      dhhat = G * zt;
      gA = G * (1.f - zt);
      const float dzpre = G * (hhat - hk) * zt * (1.f - zt) * c.kappa;
      c.DZR[slab(c, z, c.k, gw) + (size_t)b * gw + s] = dzpre;
      if (c.DZop[0]) OperandFmt<RE2NN_PREC_TF32X3>::store(c.DZop[z], (size_t)b * c.ldS + s, c.da_plane, dzpre);
    }
    const float dpre = dhhat * nl_grad_from_out(hhat, c.nl);
    const float da = z == 0 ? dpre * on : dpre;
    c.DA[sl] = da;
    if (z == 0) c.DOprod[sl] = dpre * a;
    if (c.DAop[0]) OperandFmt<RE2NN_PREC_TF32X3>::store(c.DAop[z], (size_t)b * c.ldS + s, c.da_plane, da);
    c.GA[(size_t)z * c.B * c.S + e] = gA;
  }
}

// bwd_e2 / bwd_e31 sit between two tcgen05 step GEMMs 35 times per sweep: they are launched with programmatic stream
// serialization like the GEMMs, let the next GEMM begin its prologue (barrier init, TMEM allocation, descriptor
// prefetch) right away and wait for the previous GEMM's results themselves.
__global__ void bwd_e2_kernel(const BwdCtx c) {
#ifdef RE2NN_HAVE_TC
  griddep_launch_dependents();
  griddep_wait();
#endif
  const size_t total = (size_t)2 * c.B * c.R;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int r = (int)(i % c.R);
    const size_t rr = i / c.R;
    const int b = (int)(rr % c.B), z = (int)(rr / c.B);
    int tpos, orow;
    bool alive;
    step_pos(z, c.k, (int)c.len[b], c.full_pad, tpos, orow, alive);
    const size_t e = (size_t)b * c.R + r;
    const size_t sl = slab(c, z, c.k, c.R) + e;
    if (!alive) {
      c.DU[sl] = 0.f;
      c.Qs[sl] = 0.f;
      if (c.DUop[0]) OperandFmt<RE2NN_PREC_TF32X3>::store(c.DUop[z], (size_t)b * c.ldR + r, c.du_plane, 0.f);
      continue;
    }
    const size_t vrow = c.v_mode == RE2NN_V_TOKEN ? (size_t)c.x[(size_t)b * c.Lpad + tpos] : (size_t)b * c.Lpad + tpos;
    const float v = c.vtab[vrow * c.R + r];
    const float u = c.u_save[sl];
    const float dq = c.DQ[(size_t)z * c.B * c.R + e];
    c.DU[sl] = dq * v;
    if (c.DUop[0]) OperandFmt<RE2NN_PREC_TF32X3>::store(c.DUop[z], (size_t)b * c.ldR + r, c.du_plane, dq * v);
    c.Qs[sl] = u * v;
    const float dv = dq * u;
    if (dv != 0.f) atomicAdd(c.dvtab + vrow * c.R + r, dv);
  }
}

__global__ void bwd_e3_kernel(const BwdCtx c) {
  const size_t total = (size_t)2 * c.B * c.S;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int s = (int)(i % c.S);
    const size_t rr = i / c.S;
    const int b = (int)(rr % c.B), z = (int)(rr / c.B);
    int tpos, orow;
    bool alive;
    step_pos(z, c.k, (int)c.len[b], c.full_pad, tpos, orow, alive);
    const size_t e = (size_t)b * c.S + s;
    const size_t sl = slab(c, z, c.k, c.S) + e;
    const int gw = c.S * c.farnn;
    if (!alive) {   // g keeps whatever it had (0 until the row's last live step)
      if (z == 1) c.DOprod[sl] = 0.f;
      if (c.farnn == 2) {
        c.Pinit[sl] = 0.f;
        c.DZR[slab(c, z, c.k, gw) + (size_t)b * gw + c.S + s] = 0.f;
        if (c.DRop[0]) OperandFmt<RE2NN_PREC_TF32X3>::store(c.DRop[z], (size_t)b * c.ldS + s, c.da_plane, 0.f);
      }
      continue;
    }
    const float dhb = c.DHb[(size_t)z * c.B * c.S + e];
    const float on = c.o[s];
    const float hk = c.hst_save[slab1(c, z, c.k, c.S) + e];
    const float hinit = z == 0 ? c.h0[s] : c.hT[s];
    float rt = 1.f;
    if (c.farnn == 2) rt = c.rsave[sl];
    const float htil = c.farnn == 2 ? (1.f - rt) * hinit + rt * hk : hk;
    float dht = dhb;
    if (z == 1) {
      dht = dhb * on;
      c.DOprod[sl] = dhb * htil;
    }
    float gB = dht;
    if (c.farnn == 2) {
      gB = dht * rt;
      c.Pinit[sl] = dht * (1.f - rt);
      const float drpre = dht * (hk - hinit) * rt * (1.f - rt) * c.kappa;
      c.DZR[slab(c, z, c.k, gw) + (size_t)b * gw + c.S + s] = drpre;
      if (c.DRop[0]) OperandFmt<RE2NN_PREC_TF32X3>::store(c.DRop[z], (size_t)b * c.ldS + s, c.da_plane, drpre);
    }
    c.g[(size_t)z * c.B * c.S + e] = c.GA[(size_t)z * c.B * c.S + e] + gB;
  }
}

// farnn = 0: E3 of step k and E1 of step k-1 touch the same element and nothing runs between them (no gate GEMM), so
// they are one launch: the carry g stays in a register, the GA / g round trip disappears.  Same arithmetic, same order.
__global__ void bwd_e31_kernel(const BwdCtx c) {
#ifdef RE2NN_HAVE_TC
  griddep_launch_dependents();
  griddep_wait();
#endif
  const size_t total = (size_t)2 * c.B * c.S;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int s = (int)(i % c.S);
    const size_t rr = i / c.S;
    const int b = (int)(rr % c.B), z = (int)(rr / c.B);
    const int n = (int)c.len[b];
    int tpos, orow;
    bool alive;
    step_pos(z, c.k, n, c.full_pad, tpos, orow, alive);
    const size_t e = (size_t)b * c.S + s;
    const size_t gi = (size_t)z * c.B * c.S + e;
    const float on = c.o[s];
    // ---- E3 of step k
    float g;
    if (!alive) {
      g = c.g[gi];                       // keeps whatever it had (0 until the row's last live step)
      if (z == 1) c.DOprod[slab(c, z, c.k, c.S) + e] = 0.f;
    } else {
      const float dhb = c.DHb[gi];
      g = dhb;
      if (z == 1) {
        g = dhb * on;
        c.DOprod[slab(c, z, c.k, c.S) + e] = dhb * c.hst_save[slab1(c, z, c.k, c.S) + e];
      }
      g = 0.f + g;                       // the carry is GA + gB with GA = 0 without gates
      if (c.k == 0) c.g[gi] = g;         // the final carry feeds dh0 / dhT
    }
    if (c.k == 0) continue;
    // ---- E1 of step k-1
    const int k1 = c.k - 1;
    step_pos(z, k1, n, c.full_pad, tpos, orow, alive);
    const size_t sl = slab(c, z, k1, c.S) + e;
    if (!alive) {
      c.DA[sl] = 0.f;
      if (c.DAop[0]) OperandFmt<RE2NN_PREC_TF32X3>::store(c.DAop[z], (size_t)b * c.ldS + s, c.da_plane, 0.f);
      if (z == 0) c.DOprod[sl] = 0.f;
      continue;
    }
    const float* dout = z == 0 ? c.dalpha : c.dbeta;
    const float G = g + (orow >= 0 ? dout[((size_t)b * c.L + orow) * c.S + s] : 0.f);
    const float a = c.a_save[sl];
    const float hhat = apply_nl(z == 0 ? a * on : a, c.nl);
    const float dpre = G * nl_grad_from_out(hhat, c.nl);
    const float da = z == 0 ? dpre * on : dpre;
    c.DA[sl] = da;
    if (z == 0) c.DOprod[sl] = dpre * a;
    if (c.DAop[0]) OperandFmt<RE2NN_PREC_TF32X3>::store(c.DAop[z], (size_t)b * c.ldS + s, c.da_plane, da);
  }
}

// dgtab[token] += [dz | dr] of this step (after the gate slab is complete)
__global__ void bwd_gate_scatter_kernel(const BwdCtx c) {
  const int gw = c.S * c.farnn;
  const size_t total = (size_t)2 * c.B * gw;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int col = (int)(i % gw);
    const size_t rr = i / gw;
    const int b = (int)(rr % c.B), z = (int)(rr / c.B);
    int tpos, orow;
    bool alive;
    step_pos(z, c.k, (int)c.len[b], c.full_pad, tpos, orow, alive);
    if (!alive) continue;
    const float v = c.DZR[slab(c, z, c.k, gw) + (size_t)b * gw + col];
    if (v == 0.f) continue;
    const size_t vrow = c.v_mode == RE2NN_V_TOKEN ? (size_t)c.x[(size_t)b * c.Lpad + tpos] : (size_t)b * c.Lpad + tpos;
    atomicAdd(c.dgtab + vrow * gw + col, v);
  }
}

// dhT += sum_b dbeta[b, n_b - 1, :]   (beta_n = hT is emitted directly)
__global__ void bwd_direct_hT_kernel(const float* __restrict__ dbeta, const int64_t* __restrict__ len, int B, int L,
                                     int S, float* __restrict__ dhT) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S) return;
  float acc = dhT[s];
  for (int b = 0; b < B; ++b) {
    int n = (int)len[b];
    if (n >= 1 && n <= L) acc += dbeta[((size_t)b * L + (n - 1)) * S + s];
  }
  dhT[s] = acc;
}

// token table backward, elementwise part: gen = phi(E@G) arrives as the GEMM accumulator
struct EpiTokenBwd {
  const float* dvtab; const float* V_embed; const float* beta_vec;
  float* dV; float* dbeta_prod; float* dGpre;
  int R, nl;
  __device__ __forceinline__ EpiTokenBwd for_dir(int) const { return *this; }
  __device__ __forceinline__ bool tile_alive(int, int) const { return true; }
  __device__ __forceinline__ RowCtx row(int) const { return RowCtx{0, 0, true}; }
  __device__ __forceinline__ Col col(int n) const { return Col{__ldg(beta_vec + n), 0.f}; }
  __device__ __forceinline__ Pre prefetch(const RowCtx&, int m, int n) const {
    size_t i = (size_t)m * R + n;
    return Pre{dvtab[i], V_embed[i]};
  }
  __device__ __forceinline__ void apply(const Col& c, const RowCtx&, int m, int n, float acc, const Pre& pre) const {
    const size_t i = (size_t)m * R + n;
    const float gen = apply_nl(acc, nl);
    if (dV) dV[i] = pre.a * c.a;
    dbeta_prod[i] = pre.a * (pre.b - gen);
    dGpre[i] = pre.a * (1.f - c.a) * nl_grad_from_out(gen, nl);
  }
};

// ---- resident BPTT sweep (farnn = 0): the kernel of recurrence_resident.cuh run backwards -----------------------------
// The sweep has the shape of the forward recurrence with the weights transposed: per step a GEMM of width R
// (dq = DA[k] @ S2|S1) and one of width S with two K segments (dhb = DU[k] @ S1^T|S2^T + DA[k] @ W^T|W); what the
// three elementwise kernels E1 / E2 / E3 do between them becomes the two fused epilogues below, and a CTA pair keeps
// one 128-row tile for all steps: one launch instead of five per step, no grid-wide dependency between steps.
//   epilogue of dq   (step k)   = E2(k):            DU[k], Q[k], dvtab, the DU operand of this step
//   epilogue of dhb  (step k)   = E3(k) + E1(k-1):  carry g, d_o products, DA[k-1] and the DA operand of step k-1
// E1 of a tile's first step (its last live step) runs in bwd_resident_init_kernel, which also zeroes the slabs of the
// steps the tile never executes (the weight-gradient GEMMs read every step of every row).
struct EpiBwdQ {
  static constexpr bool kRows = false;
  const float* vtab; const float* u;          // token table; u = hbar @ S of this step (forward save)
  float* DU; float* Qs; float* dvtab;
  void* DUop;
  int R, ldR;
  size_t du_plane;
  __device__ __forceinline__ Col col(int) const { return Col{0.f, 0.f}; }
  __device__ __forceinline__ Pre prefetch(const RowCtx& r, int m, int n) const {
    return Pre{__ldg(vtab + (size_t)((uint32_t)r.vrow * (uint32_t)R + (uint32_t)n)), u[(uint32_t)m * (uint32_t)R + (uint32_t)n]};
  }
  __device__ __forceinline__ float compute(const Col&, float acc, const Pre& pre) const { return acc * pre.a; }
  __device__ __forceinline__ void store(const Col&, const RowCtx& r, int m, int n, float du, float dq, const Pre& pre) const {
    const uint32_t i = (uint32_t)m * (uint32_t)R + (uint32_t)n;
    DU[i] = du;
    Qs[i] = pre.b * pre.a;
    OperandFmt<RE2NN_PREC_TF32X3>::store(DUop, (uint32_t)m * (uint32_t)ldR + (uint32_t)n, du_plane, du);
    const float dv = dq * pre.b;
    if (dv != 0.f) atomicAdd(dvtab + (size_t)((uint32_t)r.vrow * (uint32_t)R + (uint32_t)n), dv);
  }
  __device__ __forceinline__ void apply(const Col& c, const RowCtx& r, int m, int n, float acc, const Pre& pre) const {
    store(c, r, m, n, compute(c, acc, pre), acc, pre);
  }
};

// RowCtx here: orow = output row of step k-1 (-1: the row is not alive at k-1, or k = 0)
template <int NL> struct EpiBwdH {
  static constexpr bool kRows = false;
  const float* o; const float* dout;          // d alpha (z = 0) / d beta (z = 1), B x L x S
  const float* second;                        // z = 0: pre-activation a of step k-1 ; z = 1: state after step k-1 (= h~ of step k)
  float* DAprev; float* DOprod; float* gout;  // DA slab of step k-1 ; d_o products (z = 0: step k-1, z = 1: step k) ; carry at k = 0
  void* DAop;
  int L, S, ldS, z, has_prev, nl_rt;
  size_t da_plane;
  __device__ __forceinline__ int nl() const { return NL >= 0 ? NL : nl_rt; }
  __device__ __forceinline__ Col col(int n) const { return Col{__ldg(o + n), 0.f}; }
  __device__ __forceinline__ Pre prefetch(const RowCtx& r, int m, int n) const {
    Pre q{0.f, 0.f};
    if (r.orow >= 0) q.a = __ldg(dout + ((size_t)m * (uint32_t)L + (uint32_t)r.orow) * (uint32_t)S + (uint32_t)n);
    if (second) q.b = second[(uint32_t)m * (uint32_t)S + (uint32_t)n];
    return q;
  }
  // -> d pre-activation of step k-1 (or the final carry when there is no step k-1)
  __device__ __forceinline__ float compute(const Col& c, float dhb, const Pre& pre) const {
    const float g = z == 1 ? dhb * c.a : dhb;
    if (!has_prev) return g;
    const float G = g + pre.a;
    const float hhat = z == 0 ? apply_nl(pre.b * c.a, nl()) : pre.b;
    return G * nl_grad_from_out(hhat, nl());
  }
  __device__ __forceinline__ void store(const Col& c, const RowCtx&, int m, int n, float v, float dhb, const Pre& pre) const {
    const uint32_t si = (uint32_t)m * (uint32_t)S + (uint32_t)n;
    if (z == 1) DOprod[si] = dhb * pre.b;
    if (!has_prev) {
      gout[si] = v;
      return;
    }
    const float da = z == 0 ? v * c.a : v;
    DAprev[si] = da;
    if (z == 0) DOprod[si] = v * pre.b;
    OperandFmt<RE2NN_PREC_TF32X3>::store(DAop, (uint32_t)m * (uint32_t)ldS + (uint32_t)n, da_plane, da);
  }
  __device__ __forceinline__ void apply(const Col& c, const RowCtx& r, int m, int n, float acc, const Pre& pre) const {
    store(c, r, m, n, compute(c, acc, pre), acc, pre);
  }
};

// Both epilogues read slabs nobody has touched since the forward pass wrote them (HBM latency), from an SM whose L1 is
// all but given away to shared memory: the loads of a chunk then trickle in at the few lines L1 can keep in flight
// (measured: 12k cycles per 16 rows).  Pulling the lines into L2 while the step's MMAs run cuts the latency each of
// those slots is held for.  Rows are not 128-byte aligned (pitch 4*S): a 32-column segment straddles two lines.
template <> struct EpiL2Prefetch<EpiBwdQ> {
  static constexpr bool kOn = true;
  __device__ static __forceinline__ void issue(const EpiBwdQ& e, const RowCtx&, int m, int n) {
    const float* u = e.u + (uint32_t)m * (uint32_t)e.R + (uint32_t)n;
    l2_prefetch(u);
    l2_prefetch(u + min(31, e.R - 1 - n));
  }
};
template <int NL> struct EpiL2Prefetch<EpiBwdH<NL>> {
  static constexpr bool kOn = true;
  __device__ static __forceinline__ void issue(const EpiBwdH<NL>& e, const RowCtx& r, int m, int n) {
    const int last = min(31, e.S - 1 - n);
    if (r.orow >= 0) {
      const float* d = e.dout + ((size_t)m * (uint32_t)e.L + (uint32_t)r.orow) * (uint32_t)e.S + (uint32_t)n;
      l2_prefetch(d);
      l2_prefetch(d + last);
    }
    if (e.second) {
      const float* s2 = e.second + (uint32_t)m * (uint32_t)e.S + (uint32_t)n;
      l2_prefetch(s2);
      l2_prefetch(s2 + last);
    }
  }
};

template <int NL> struct ResidentBackward {
  static constexpr int kPrec = RE2NN_PREC_TF32X3, kFarnn = 0, kEpiWarps = 8;
  using E1 = EpiBwdQ;
  using E2 = EpiBwdH<NL>;
  using EG = EpiBwdQ;     // no gate phases
  BwdCtx c;
  struct Tile { int z, k, par; };
  __device__ __forceinline__ int M() const { return c.B; }
  __device__ __forceinline__ int S() const { return c.S; }
  __device__ __forceinline__ int R() const { return c.R; }
  __device__ __forceinline__ int nsteps(int z, int mt, int steps) const { return min(steps, __ldg(c.tile_last[z] + mt) + 1); }
  __device__ __forceinline__ void begin(Tile& t, int z) const { t.z = z; }
  __device__ __forceinline__ void step(Tile& t, int i, int ns) const {      // the i-th executed step is step ns-1-i
    t.k = ns - 1 - i;
    t.par = i & 1;
  }
  __device__ __forceinline__ RowCtx row(const Tile& t, int m) const {
    const int n = (int)c.len[m];
    int tpos, orow, tp2, oprev = -1;
    bool alive, al2;
    step_pos(t.z, t.k, n, 0, tpos, orow, alive);
    if (t.k >= 1) step_pos(t.z, t.k - 1, n, 0, tp2, oprev, al2);
    const int vrow = c.v_mode == RE2NN_V_TOKEN ? (int)c.x[(size_t)m * c.Lpad + tpos] : m * c.Lpad + tpos;
    return RowCtx{vrow, oprev, oprev >= 0};
  }
  __device__ __forceinline__ EG eg(const Tile& t) const { return e1(t); }
  __device__ __forceinline__ E1 e1(const Tile& t) const {
    const size_t sl = slab(c, t.z, t.k, c.R);
    return E1{c.vtab, c.u_save + sl, c.DU + sl, c.Qs + sl, c.dvtab, c.DUop[t.z], c.R, c.ldR, c.du_plane};
  }
  __device__ __forceinline__ E2 e2(const Tile& t) const {
    const bool prev = t.k >= 1;
    const size_t sk = slab(c, t.z, t.k, c.S), sp = prev ? slab(c, t.z, t.k - 1, c.S) : 0;
    E2 e;
    e.o = c.o;
    e.dout = t.z == 0 ? c.dalpha : c.dbeta;
    e.second = t.z == 0 ? (prev ? c.a_save + sp : nullptr) : c.hst_save + slab1(c, t.z, t.k, c.S);
    e.DAprev = prev ? c.DA + sp : nullptr;
    e.DOprod = t.z == 0 ? (prev ? c.DOprod + sp : nullptr) : c.DOprod + sk;
    e.gout = c.g + (size_t)t.z * c.B * c.S;
    e.DAop = t.par == 0 ? c.DAop1[t.z] : c.DAop[t.z];      // this step reads parity par, the next one par ^ 1
    e.L = c.L; e.S = c.S; e.ldS = c.ldS; e.z = t.z; e.has_prev = prev ? 1 : 0; e.nl_rt = c.nl;
    e.da_plane = c.da_plane;
    return e;
  }
};

// grid (2 * m-tiles, L): block (tile, k).  k above the tile's last live step: zero the rows of the step slabs the sweep
// never writes; k = last live step: E1 with a zero carry (DA, d_o product, the parity-0 DA operand); below: nothing.
__global__ void __launch_bounds__(256) bwd_resident_init_kernel(const BwdCtx c) {
  const int m_tiles = (c.B + 127) / 128;
  const int z = blockIdx.x / m_tiles, mt = blockIdx.x - z * m_tiles, k = blockIdx.y;
  const int ks = c.tile_last[z][mt];
  if (k < ks) return;
  const int m0 = mt * 128, nrows = min(128, c.B - m0);
  if (k > ks) {
    float* da = c.DA + slab(c, z, k, c.S) + (size_t)m0 * c.S;
    float* dp = c.DOprod + slab(c, z, k, c.S) + (size_t)m0 * c.S;
    float* du = c.DU + slab(c, z, k, c.R) + (size_t)m0 * c.R;
    float* qs = c.Qs + slab(c, z, k, c.R) + (size_t)m0 * c.R;
    for (int i = threadIdx.x; i < nrows * c.S; i += blockDim.x) { da[i] = 0.f; dp[i] = 0.f; }
    for (int i = threadIdx.x; i < nrows * c.R; i += blockDim.x) { du[i] = 0.f; qs[i] = 0.f; }
    return;
  }
  for (int i = threadIdx.x; i < nrows * c.S; i += blockDim.x) {
    const int r = i / c.S, s = i - r * c.S, b = m0 + r;
    int tpos, orow;
    bool alive;
    step_pos(z, k, (int)c.len[b], 0, tpos, orow, alive);
    const size_t e = (size_t)b * c.S + s;
    const size_t sl = slab(c, z, k, c.S) + e;
    float da = 0.f, dop = 0.f;
    if (alive) {
      const float* dout = z == 0 ? c.dalpha : c.dbeta;
      const float G = orow >= 0 ? dout[((size_t)b * c.L + orow) * c.S + s] : 0.f;
      const float a = c.a_save[sl];
      const float on = c.o[s];
      const float hhat = apply_nl(z == 0 ? a * on : a, c.nl);
      const float dpre = G * nl_grad_from_out(hhat, c.nl);
      da = z == 0 ? dpre * on : dpre;
      dop = dpre * a;
    }
    c.DA[sl] = da;
    if (z == 0) c.DOprod[sl] = dop;
    OperandFmt<RE2NN_PREC_TF32X3>::store(c.DAop[z], (size_t)b * c.ldS + s, c.da_plane, da);
  }
}

// The two GEMMs of every BPTT step (dq = DA @ S, dhbar = DU @ S^T + DA @ W^T) run on the tensor cores in 3xTF32
// (fp32-grade products, 8-bit exponent: safe for gradients of any magnitude) when tcgen05 is there.
static bool g_bwd_tc = true;
// this translation unit's copy of the phase-trace pointer (re2nn_debug_set_tc_trace sets both)
cudaError_t backward_set_trace(unsigned long long* device_buf) {
  return cudaMemcpyToSymbol(g_tc_trace, &device_buf, sizeof(device_buf));
}
extern int g_resident_train_on;       // recurrence.cu (re2nn_debug_set_resident_train)
cudaError_t launch_tile_last(const int64_t* len, int B, int* fwd, int* bwd, cudaStream_t st);      // recurrence.cu
extern "C" int re2nn_has_tcgen05(void);
static bool bwd_uses_tc() { return g_bwd_tc && re2nn_has_tcgen05() != 0; }

static size_t bwd_carve(const re2nn_backward_args& a, char* base, BwdCtx* c, float** draw, float** partial,
                        float** dalpha, float** dbeta, WeightPrep* wp = nullptr) {
  size_t off = 0;
  auto take = [&](size_t floats) -> float* {
    float* p = base ? (float*)(base + off) : nullptr;
    off += align_up(floats * 4, 256);
    return p;
  };
  const size_t B = a.B, L = a.L, S = a.S, R = a.R, gw = (size_t)a.S * a.farnn;
  BwdCtx x;
  memset(&x, 0, sizeof(x));
  float* dA = take(B * L * S);
  float* dBt = take(B * L * S);
  x.g = take(2 * B * S);
  x.GA = take(2 * B * S);
  x.DA = take(2 * L * B * S);
  x.DOprod = take(2 * L * B * S);
  x.DU = take(2 * L * B * R);
  x.Qs = take(2 * L * B * R);
  x.DQ = take(2 * B * R);
  x.DHb = take(2 * B * S);
  if (a.farnn >= 1) {
    x.DZR = take(2 * L * B * gw);
    x.dgtab = take((size_t)a.table_rows * gw);
  }
  if (a.farnn == 2) x.Pinit = take(2 * L * B * S);
  float* dr = a.priority_mat ? take(B * L * (size_t)a.C) : nullptr;
  size_t pmax = std::max({S * R, S * S, (size_t)a.C * S, gw ? S * S : (size_t)0, gw ? R * S : (size_t)0});
  float* part = take((size_t)kTnSplit * pmax);
  if (bwd_uses_tc()) {
    constexpr int P = RE2NN_PREC_TF32X3;
    x.ldS = operand_ld(P, (int)S); x.ldR = operand_ld(P, (int)R);
    x.da_plane = B * x.ldS; x.du_plane = B * x.ldR;
    for (int z = 0; z < 2; ++z) {
      x.DAop[z] = take(operand_bytes(P, B, (int)S) / 4);
      x.DUop[z] = take(operand_bytes(P, B, (int)R) / 4);
      x.DAop1[z] = take(operand_bytes(P, B, (int)S) / 4);
      x.tile_last[z] = (int*)take((B + 127) / 128);
    }
    if (a.C > 0) {        // label-score backward on tensor cores: draw (B*L x C) and C_mat^T (S x C) in operand format
      x.DrawOp = take(operand_bytes(P, B * L, a.C) / 4);
      x.CmatOp = take(operand_bytes(P, S, a.C) / 4);
    }
    for (int z = 0; z < 2 && a.farnn >= 1; ++z) {
      x.DZop[z] = take(operand_bytes(P, B, (int)S) / 4);
      if (a.farnn == 2) x.DRop[z] = take(operand_bytes(P, B, (int)S) / 4);
    }
    WeightPrep w;
    off += weight_prep_carve(P, (int)S, (int)R, 0, base ? base + off : nullptr, &w);
    if (a.farnn >= 1) {      // K-major copies of Wss1 / Wss2 themselves (the forward holds their transposes)
      w.buf[6] = take(operand_bytes(P, S, (int)S) / 4);
      if (a.farnn == 2) w.buf[7] = take(operand_bytes(P, S, (int)S) / 4);
    }
    if (wp) *wp = w;
  }
  if (c) *c = x;
  if (draw) *draw = dr;
  if (partial) *partial = part;
  if (dalpha) *dalpha = dA;
  if (dbeta) *dbeta = dBt;
  return off;
}

// launch with programmatic stream serialization (the kernel calls griddepcontrol itself)
static cudaError_t launch_pdl(void (*kernel)(BwdCtx), int grid, cudaStream_t st, const BwdCtx& c) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(256);
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, c);
}

static int grid_for(size_t total) { return (int)std::min<size_t>((total + 255) / 256, (size_t)sm_count() * 16); }

static int run_backward(const re2nn_backward_args& a, cudaStream_t st) {
  BwdCtx c;
  float *draw_ws, *partial, *dalpha, *dbeta;
  WeightPrep wp;
  memset(&wp, 0, sizeof(wp));
  const bool tc = bwd_uses_tc();
  const size_t need = bwd_carve(a, (char*)a.ws, &c, &draw_ws, &partial, &dalpha, &dbeta, &wp);
  RE2NN_CHECK(a.ws && a.ws_bytes >= need, "decompose_backward: workspace too small (%zu < %zu)", a.ws_bytes, need);
  const int B = a.B, L = a.L, S = a.S, R = a.R, C = a.C, gw = a.S * a.farnn;
  const size_t M = (size_t)B * L;
  c.B = B; c.Lpad = a.Lpad; c.L = L; c.S = S; c.R = R; c.farnn = a.farnn; c.nl = a.update_nonlinear;
  c.v_mode = a.v_mode; c.full_pad = a.full_pad; c.kappa = a.sigmoid_exponent;
  c.x = a.x; c.len = a.lengths; c.vtab = a.vtab; c.o = a.o; c.h0 = a.h0; c.hT = a.hT;
  c.dalpha = dalpha; c.dbeta = dbeta;
  c.hst_save = a.hst_save; c.u_save = a.u_save; c.a_save = a.a_save; c.zsave = a.zsave; c.rsave = a.rsave;
  c.dvtab = a.dvtab;

  // 0. undo the priority layer: draw = dscores @ P^T
  const float* draw = a.dscores;
  GemmProblem g;
  const bool direct = a.dalpha_in != nullptr && a.dbeta_in != nullptr;   // caller differentiated its own score stage
  if (direct) {
    c.dalpha = a.dalpha_in; c.dbeta = a.dbeta_in;
    dbeta = const_cast<float*>(a.dbeta_in);
  }
  if (!direct && a.priority_mat) {
    memset(&g, 0, sizeof(g));
    g.M = (int)M; g.N = C; g.nseg = 1; g.ndir = 1;
    g.seg[0][0] = GemmSeg{a.dscores, a.priority_mat, C, C, C, 1, 0, 0};
    RE2NN_CUDA(launch_simt_gemm(g, EpiStore{draw_ws, C, nullptr}, ALoadPlain{}, st));
    draw = draw_ws;
  }
  // 1. dAB = draw @ C  ->  dAlpha, dBeta
  if (!direct) {
    const EpiDAB edab{a.alpha, a.beta, a.lengths, dalpha, dbeta, L, S, a.full_pad};
    if (tc && c.DrawOp != nullptr) {
      // 3xTF32 on tcgen05: the GEMM is tiny (K = C), what matters is the coalesced epilogue over alpha / beta / dalpha / dbeta
      constexpr int P = RE2NN_PREC_TF32X3;
      const int ldc = operand_ld(P, C);
      const size_t pa = M * ldc, pb = (size_t)S * ldc;
      convert_weight_kernel<P><<<(unsigned)std::min<size_t>((M * ldc + 255) / 256, (size_t)sm_count() * 32), 256, 0, st>>>(
          draw, (int)M, C, C, 0, c.DrawOp, ldc, pa, 0);
      RE2NN_LAUNCH_CHECK();
      convert_weight_kernel<P><<<(unsigned)(((size_t)S * ldc + 255) / 256), 256, 0, st>>>(a.C_mat, S, C, S, 1, c.CmatOp, ldc, pb, 0);
      RE2NN_LAUNCH_CHECK();
      memset(&g, 0, sizeof(g));
      g.M = (int)M; g.N = S; g.nseg = 1; g.ndir = 1;
      g.seg[0][0] = GemmSeg{c.DrawOp, c.CmatOp, ldc, ldc, C, 1, pa, pb};
      std::unique_ptr<TcLaunch> tl(new TcLaunch);
      if (int rc = tc_make_launch<P>(g, tl.get())) return rc;
      RE2NN_CUDA((launch_tc_gemm<P>(g, edab, tl.get(), st)));
    } else {
      memset(&g, 0, sizeof(g));
      g.M = (int)M; g.N = S; g.nseg = 1; g.ndir = 1;
      g.seg[0][0] = GemmSeg{draw, a.C_mat, C, S, C, 0, 0, 0};
      RE2NN_CUDA(launch_simt_gemm(g, edab, ALoadPlain{}, st));
    }
  }
  // 2. dC = draw^T @ (alpha * beta)   (the tensor-core kernel forms the masked product itself: Y2 / mask_len)
  if (!direct && a.dC) {
    TnProblem t;
    memset(&t, 0, sizeof(t));
    t.P = C; t.Q = S; t.npairs = 1;
    t.pair[0] = TnPair{draw, a.alpha, C, S, M};
    t.Y2 = a.beta;
    if (!a.full_pad) { t.mask_len = a.lengths; t.mask_L = L; }
    RE2NN_CUDA(run_tn(t, partial, a.dC, 0, YLoadAlphaBeta{a.beta, a.lengths, L, a.full_pad}, st));
  }
  // 3. sweep
  RE2NN_CUDA(cudaMemsetAsync(c.g, 0, (size_t)2 * B * S * 4, st));
  RE2NN_CUDA(cudaMemsetAsync(a.dvtab, 0, (size_t)a.table_rows * R * 4, st));
  if (a.farnn >= 1) RE2NN_CUDA(cudaMemsetAsync(c.dgtab, 0, (size_t)a.table_rows * gw * 4, st));
  std::unique_ptr<TcLaunch> tq, th, tg;
  if (tc) {
    // K-major operand-format copies of S1, S2, W (the same layouts the forward uses) + the tensor maps of the two
    // step GEMMs; the A operands are the per-step DAop / DUop buffers, rewritten every step
    constexpr int P = RE2NN_PREC_TF32X3;
    re2nn_recurrence_args ra;
    memset(&ra, 0, sizeof(ra));
    ra.S = S; ra.R = R; ra.farnn = 0; ra.S1 = a.S1; ra.S2 = a.S2; ra.W = a.W;
    if (int rc = weight_prep_run<P>(ra, wp, st)) return rc;
    GemmProblem gq, gh;
    memset(&gq, 0, sizeof(gq));
    memset(&gh, 0, sizeof(gh));
    gq.M = B; gq.N = R; gq.nseg = 1; gq.ndir = 2;
    gh.M = B; gh.N = S; gh.nseg = 2; gh.ndir = 2;
    for (int z = 0; z < 2; ++z) {
      gq.seg[z][0] = wp.seg_g1(1 - z, c.DAop[z], c.ldS, c.da_plane);      // DA @ S2 (fwd) | DA @ S1 (bwd)
      gh.seg[z][0] = wp.seg_g2q(1 - z, c.DUop[z], c.ldR, c.du_plane);     // DU @ S1^T (fwd) | DU @ S2^T (bwd)
      gh.seg[z][1] = wp.seg_g2w(1 - z, c.DAop[z], c.ldS, c.da_plane);     // DA @ W^T (fwd) | DA @ W (bwd)
    }
    tq.reset(new TcLaunch);
    th.reset(new TcLaunch);
    if (int rc = tc_make_launch<P>(gq, tq.get())) return rc;
    if (int rc = tc_make_launch<P>(gh, th.get())) return rc;
    if (a.farnn >= 1) {
      const size_t pl = (size_t)S * wp.ldS;
      const int blocks = (int)(((size_t)S * wp.ldS + 255) / 256);
      convert_weight_kernel<P><<<blocks, 256, 0, st>>>(a.Wss1, S, S, S, 0, wp.buf[6], wp.ldS, pl, 0);
      if (a.farnn == 2) convert_weight_kernel<P><<<blocks, 256, 0, st>>>(a.Wss2, S, S, S, 0, wp.buf[7], wp.ldS, pl, 0);
      RE2NN_LAUNCH_CHECK();
      GemmProblem gg;
      memset(&gg, 0, sizeof(gg));
      gg.M = B; gg.N = S; gg.nseg = a.farnn; gg.ndir = 2;
      for (int z = 0; z < 2; ++z) {
        gg.seg[z][0] = GemmSeg{c.DZop[z], wp.buf[6], c.ldS, wp.ldS, S, 1, c.da_plane, pl};
        if (a.farnn == 2) gg.seg[z][1] = GemmSeg{c.DRop[z], wp.buf[7], c.ldS, wp.ldS, S, 1, c.da_plane, pl};
      }
      tg.reset(new TcLaunch);
      if (int rc = tc_make_launch<P>(gg, tg.get())) return rc;
    }
  }
  // farnn = 0 on tensor cores: the whole sweep in one resident launch (see ResidentBackward)
  const bool resident = tc && a.farnn == 0 && !a.full_pad && (g_resident_train_on & 2) && resident_supported(2, S, R, true);
  if (resident) {
    constexpr int P = RE2NN_PREC_TF32X3;
    RE2NN_CUDA(launch_tile_last(a.lengths, B, c.tile_last[0], c.tile_last[1], st));
    bwd_resident_init_kernel<<<dim3(2 * cdiv(B, 128), L), 256, 0, st>>>(c);
    RE2NN_LAUNCH_CHECK();
    std::unique_ptr<ResidentLaunch> rl(new ResidentLaunch);
    memset(rl.get(), 0, sizeof(ResidentLaunch));
    const int bn1 = resident_part(R), bn2 = resident_part(S);
    for (int par = 0; par < 2; ++par) {
      GemmProblem gq, gh;
      memset(&gq, 0, sizeof(gq));
      memset(&gh, 0, sizeof(gh));
      gq.M = B; gq.N = R; gq.nseg = 1; gq.ndir = 2;
      gh.M = B; gh.N = S; gh.nseg = 2; gh.ndir = 2;
      for (int z = 0; z < 2; ++z) {
        void* da = par == 0 ? c.DAop[z] : c.DAop1[z];
        gq.seg[z][0] = wp.seg_g1(1 - z, da, c.ldS, c.da_plane);               // DA @ S2 (fwd) | DA @ S1 (bwd)
        gh.seg[z][0] = wp.seg_g2q(1 - z, c.DUop[z], c.ldR, c.du_plane);       // DU @ S1^T (fwd) | DU @ S2^T (bwd)
        gh.seg[z][1] = wp.seg_g2w(1 - z, da, c.ldS, c.da_plane);              // DA @ W^T (fwd) | DA @ W (bwd)
      }
      if (int rc = tc_make_launch<P>(gq, &rl->g1[par], bn1)) return rc;
      if (int rc = tc_make_launch<P>(gh, &rl->g2[par], bn2)) return rc;
    }
    rl->steps = L;
    rl->q_first = 1;
    rl->alias_tbuf = resident_alias(2, true) ? 1 : 0;
    rl->stage_bytes = resident_stage_bytes(2, S, R);
    rl->stages = resident_stages(2, S, R, true);
    if (a.update_nonlinear == RE2NN_NL_TANH) RE2NN_CUDA(launch_resident_policy(*rl, ResidentBackward<RE2NN_NL_TANH>{c}, B, st));
    else RE2NN_CUDA(launch_resident_policy(*rl, ResidentBackward<-1>{c}, B, st));
  }
  for (int k = L - 1; k >= 0 && !resident; --k) {
    c.k = k;
    if (k == L - 1 || a.farnn >= 1) {      // without gates E1 of the later steps rides in bwd_e31_kernel
      bwd_e1_kernel<<<grid_for((size_t)2 * B * S), 256, 0, st>>>(c);
      RE2NN_LAUNCH_CHECK();
    }
    // dq = DA[k] @ S2 (fwd) | S1 (bwd)
    memset(&g, 0, sizeof(g));
    g.M = B; g.N = R; g.nseg = 1; g.ndir = 2;
    for (int z = 0; z < 2; ++z)
      g.seg[z][0] = GemmSeg{c.DA + ((size_t)z * L + k) * B * S, z == 0 ? a.S2 : a.S1, S, R, S, 0, 0, 0};
    if (tc)
      RE2NN_CUDA((launch_tc_gemm<RE2NN_PREC_TF32X3>(g, EpiStore2{{c.DQ, c.DQ + (size_t)B * R}, R, 0}, tq.get(), st)));
    else
      RE2NN_CUDA(launch_simt_gemm(g, EpiStore2{{c.DQ, c.DQ + (size_t)B * R}, R, 0}, ALoadPlain{}, st));
    RE2NN_CUDA(launch_pdl(bwd_e2_kernel, grid_for((size_t)2 * B * R), st, c));
    // dhb = DU[k] @ S1^T + DA[k] @ W^T (fwd) | DU[k] @ S2^T + DA[k] @ W (bwd)
    memset(&g, 0, sizeof(g));
    g.M = B; g.N = S; g.nseg = 2; g.ndir = 2;
    for (int z = 0; z < 2; ++z) {
      g.seg[z][0] = GemmSeg{c.DU + ((size_t)z * L + k) * B * R, z == 0 ? a.S1 : a.S2, R, R, R, 1, 0, 0};
      g.seg[z][1] = GemmSeg{c.DA + ((size_t)z * L + k) * B * S, a.W, S, S, S, z == 0 ? 1 : 0, 0, 0};
    }
    if (tc)
      RE2NN_CUDA((launch_tc_gemm<RE2NN_PREC_TF32X3>(g, EpiStore2{{c.DHb, c.DHb + (size_t)B * S}, S, 0}, th.get(), st)));
    else
      RE2NN_CUDA(launch_simt_gemm(g, EpiStore2{{c.DHb, c.DHb + (size_t)B * S}, S, 0}, ALoadPlain{}, st));
    if (a.farnn == 0) RE2NN_CUDA(launch_pdl(bwd_e31_kernel, grid_for((size_t)2 * B * S), st, c));
    else {
      bwd_e3_kernel<<<grid_for((size_t)2 * B * S), 256, 0, st>>>(c);
      RE2NN_LAUNCH_CHECK();
    }
    if (a.farnn >= 1) {
      // g += [dz | dr] @ [Wss1 | Wss2]^T
      memset(&g, 0, sizeof(g));
      g.M = B; g.N = S; g.nseg = a.farnn; g.ndir = 2;
      for (int z = 0; z < 2; ++z) {
        const float* dzr = c.DZR + ((size_t)z * L + k) * B * gw;
        g.seg[z][0] = GemmSeg{dzr, a.Wss1, gw, S, S, 1, 0, 0};
        if (a.farnn == 2) g.seg[z][1] = GemmSeg{dzr + S, a.Wss2, gw, S, S, 1, 0, 0};
      }
      if (tc)
        RE2NN_CUDA((launch_tc_gemm<RE2NN_PREC_TF32X3>(g, EpiStore2{{c.g, c.g + (size_t)B * S}, S, 1}, tg.get(), st)));
      else
        RE2NN_CUDA(launch_simt_gemm(g, EpiStore2{{c.g, c.g + (size_t)B * S}, S, 1}, ALoadPlain{}, st));
      bwd_gate_scatter_kernel<<<grid_for((size_t)2 * B * gw), 256, 0, st>>>(c);
      RE2NN_LAUNCH_CHECK();
    }
  }
  // 4. weight gradients: one transposed GEMM each over all (step, sequence) rows
  const size_t rows1 = (size_t)L * B;                    // per-direction slab rows
  const float* Hb0 = a.hbar_save;                        // (L+1)-deep slabs: direction stride (L+1)*B*S
  const float* Hb1 = a.hbar_save + (size_t)(L + 1) * B * S;
  const float *DA0 = c.DA, *DA1 = c.DA + rows1 * S, *DU0 = c.DU, *DU1 = c.DU + rows1 * R;
  const float *Q0 = c.Qs, *Q1 = c.Qs + rows1 * R;
  TnProblem t;
  if (a.dS1) {   // fwd: hbar^T du ; bwd: da^T q
    memset(&t, 0, sizeof(t));
    t.P = S; t.Q = R; t.npairs = 2;
    t.pair[0] = TnPair{Hb0, DU0, S, R, rows1};
    t.pair[1] = TnPair{DA1, Q1, S, R, rows1};
    RE2NN_CUDA(run_tn(t, partial, a.dS1, 0, YLoadPlain{}, st));
  }
  if (a.dS2) {   // fwd: da^T q ; bwd: hbar^T du
    memset(&t, 0, sizeof(t));
    t.P = S; t.Q = R; t.npairs = 2;
    t.pair[0] = TnPair{DA0, Q0, S, R, rows1};
    t.pair[1] = TnPair{Hb1, DU1, S, R, rows1};
    RE2NN_CUDA(run_tn(t, partial, a.dS2, 0, YLoadPlain{}, st));
  }
  if (a.dW) {    // fwd: hbar^T da ; bwd: da^T hbar
    memset(&t, 0, sizeof(t));
    t.P = S; t.Q = S; t.npairs = 2;
    t.pair[0] = TnPair{Hb0, DA0, S, S, rows1};
    t.pair[1] = TnPair{DA1, Hb1, S, S, rows1};
    RE2NN_CUDA(run_tn(t, partial, a.dW, 0, YLoadPlain{}, st));
  }
  if (a.farnn >= 1) {
    const float* Hs0 = a.hst_save;
    const float* Hs1 = a.hst_save + (size_t)(L + 1) * B * S;
    const float* Z0 = c.DZR;
    const float* Z1 = c.DZR + rows1 * gw;
    if (a.dWss1) {
      memset(&t, 0, sizeof(t));
      t.P = S; t.Q = S; t.npairs = 2;
      t.pair[0] = TnPair{Hs0, Z0, S, gw, rows1};
      t.pair[1] = TnPair{Hs1, Z1, S, gw, rows1};
      RE2NN_CUDA(run_tn(t, partial, a.dWss1, 0, YLoadPlain{}, st));
    }
    if (a.farnn == 2 && a.dWss2) {
      memset(&t, 0, sizeof(t));
      t.P = S; t.Q = S; t.npairs = 2;
      t.pair[0] = TnPair{Hs0, Z0 + S, S, gw, rows1};
      t.pair[1] = TnPair{Hs1, Z1 + S, S, gw, rows1};
      RE2NN_CUDA(run_tn(t, partial, a.dWss2, 0, YLoadPlain{}, st));
    }
    // token-side gate parameters through the gate table: Wrs = vtab^T dgtab, bs = colsum(dgtab), dvtab += dgtab Wrs^T
    if (a.dWrs1) {
      memset(&t, 0, sizeof(t));
      t.P = R; t.Q = S; t.npairs = 1;
      t.pair[0] = TnPair{a.vtab, c.dgtab, R, gw, (size_t)a.table_rows};
      RE2NN_CUDA(run_tn(t, partial, a.dWrs1, 0, YLoadPlain{}, st));
    }
    if (a.farnn == 2 && a.dWrs2) {
      memset(&t, 0, sizeof(t));
      t.P = R; t.Q = S; t.npairs = 1;
      t.pair[0] = TnPair{a.vtab, c.dgtab + S, R, gw, (size_t)a.table_rows};
      RE2NN_CUDA(run_tn(t, partial, a.dWrs2, 0, YLoadPlain{}, st));
    }
    if (a.dbs1) RE2NN_CUDA(colsum_rows(c.dgtab, a.table_rows, S, gw, a.dbs1, 0, st, partial));
    if (a.farnn == 2 && a.dbs2) RE2NN_CUDA(colsum_rows(c.dgtab + S, a.table_rows, S, gw, a.dbs2, 0, st, partial));
    memset(&g, 0, sizeof(g));
    g.M = a.table_rows; g.N = R; g.nseg = a.farnn; g.ndir = 1;
    g.seg[0][0] = GemmSeg{c.dgtab, a.Wrs1, gw, S, S, 1, 0, 0};
    if (a.farnn == 2) g.seg[0][1] = GemmSeg{c.dgtab + S, a.Wrs2, gw, S, S, 1, 0, 0};
    RE2NN_CUDA(launch_simt_gemm(g, EpiStore2{{a.dvtab, a.dvtab}, R, 1}, ALoadPlain{}, st));
  }
  // 5. vector gradients
  if (a.d_o) RE2NN_CUDA(colsum_rows(c.DOprod, (size_t)2 * rows1, S, S, a.d_o, 0, st, partial));
  if (a.dh0) {
    RE2NN_CUDA(colsum_rows(c.g, B, S, S, a.dh0, 0, st));
    if (a.farnn == 2) RE2NN_CUDA(colsum_rows(c.Pinit, rows1, S, S, a.dh0, 1, st, partial));
  }
  if (a.dhT) {
    RE2NN_CUDA(colsum_rows(c.g + (size_t)B * S, B, S, S, a.dhT, 0, st));
    if (a.farnn == 2) RE2NN_CUDA(colsum_rows(c.Pinit + rows1 * S, rows1, S, S, a.dhT, 1, st, partial));
    bwd_direct_hT_kernel<<<cdiv(S, 128), 128, 0, st>>>(dbeta, a.lengths, B, L, S, a.dhT);
    RE2NN_LAUNCH_CHECK();
  }
  return 0;
}

}  // namespace re2nn

using namespace re2nn;

extern "C" {

int re2nn_debug_set_backward_tc(int on) {
  g_bwd_tc = on != 0;
  return 0;
}

int re2nn_debug_set_tn_tc(int on) {
  g_tn_tc = on != 0;
  return 0;
}

size_t re2nn_decompose_backward_workspace(const re2nn_backward_args* a) {
  if (!a) return 0;
  return bwd_carve(*a, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
}

int re2nn_decompose_backward(const re2nn_backward_args* a, void* stream) {
  RE2NN_CHECK(a != nullptr, "decompose_backward: null args");
  RE2NN_CHECK(a->B > 0 && a->L > 0 && a->S > 0 && a->R > 0 && (a->C > 0 || (a->dalpha_in && a->dbeta_in)), "decompose_backward: bad dims");
  RE2NN_CHECK(a->farnn >= 0 && a->farnn <= 2, "decompose_backward: farnn must be 0, 1 or 2");
  const bool direct = a->dalpha_in && a->dbeta_in;
  RE2NN_CHECK((direct || (a->dscores && a->C_mat)) && a->lengths && a->vtab && a->S1 && a->S2 && a->W && a->o && a->h0 &&
                  a->hT && a->alpha && a->beta && a->hbar_save && a->hst_save && a->u_save && a->a_save && a->dvtab,
              "decompose_backward: null tensor");
  RE2NN_CHECK(a->v_mode == RE2NN_V_DENSE || a->x, "decompose_backward: token mode needs x");
  RE2NN_CHECK(a->farnn == 0 || (a->zsave && a->Wss1 && a->Wrs1), "decompose_backward: farnn>=1 needs gate tensors");
  RE2NN_CHECK(a->farnn < 2 || (a->rsave && a->Wss2 && a->Wrs2), "decompose_backward: farnn==2 needs reset-gate tensors");
  return run_backward(*a, (cudaStream_t)stream);
}

int re2nn_label_scores_backward(const float* dscores, const float* alpha, const float* beta, const int64_t* lengths,
                                int B, int L, int S, const float* C_mat, int C, const float* priority_mat,
                                int full_pad, float* dalpha, float* dbeta, float* ws, void* stream) {
  RE2NN_CHECK(dscores && alpha && beta && lengths && C_mat && dalpha && dbeta, "label_scores_backward: null tensor");
  RE2NN_CHECK(!priority_mat || ws, "label_scores_backward: priority needs a workspace");
  cudaStream_t st = (cudaStream_t)stream;
  const float* draw = dscores;
  GemmProblem g;
  if (priority_mat) {
    memset(&g, 0, sizeof(g));
    g.M = B * L; g.N = C; g.nseg = 1; g.ndir = 1;
    g.seg[0][0] = GemmSeg{dscores, priority_mat, C, C, C, 1, 0, 0};
    RE2NN_CUDA(launch_simt_gemm(g, EpiStore{ws, C, nullptr}, ALoadPlain{}, st));
    draw = ws;
  }
  memset(&g, 0, sizeof(g));
  g.M = B * L; g.N = S; g.nseg = 1; g.ndir = 1;
  g.seg[0][0] = GemmSeg{draw, C_mat, C, S, C, 0, 0, 0};
  RE2NN_CUDA(launch_simt_gemm(g, EpiDAB{alpha, beta, lengths, dalpha, dbeta, L, S, full_pad}, ALoadPlain{}, st));
  return 0;
}

size_t re2nn_token_table_backward_workspace(int rows, int D, int R) {
  return align_up((size_t)rows * R * 4, 256) * 2 + align_up((size_t)kTnSplit * D * R * 4, 256) + 256;
}

int re2nn_token_table_backward(const float* dvtab, const float* V_embed, const float* E, const float* G,
                               const float* beta_vec, int rows, int D, int R, int additional_nonlinear,
                               float* dV_embed, float* dbeta_vec, float* dG, float* dE, void* ws, size_t ws_bytes,
                               void* stream) {
  RE2NN_CHECK(dvtab && V_embed && E && G && beta_vec && ws, "token_table_backward: null tensor");
  RE2NN_CHECK(ws_bytes >= re2nn_token_table_backward_workspace(rows, D, R), "token_table_backward: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  char* w = (char*)ws;
  float* dbeta_prod = (float*)w;
  w += align_up((size_t)rows * R * 4, 256);
  float* dGpre = (float*)w;
  w += align_up((size_t)rows * R * 4, 256);
  float* partial = (float*)w;
  GemmProblem g;
  memset(&g, 0, sizeof(g));
  g.M = rows; g.N = R; g.nseg = 1; g.ndir = 1;
  g.seg[0][0] = GemmSeg{E, G, D, R, D, 0, 0, 0};
  RE2NN_CUDA(launch_simt_gemm(g, EpiTokenBwd{dvtab, V_embed, beta_vec, dV_embed, dbeta_prod, dGpre, R, additional_nonlinear},
                              ALoadPlain{}, st));
  if (dbeta_vec) RE2NN_CUDA(colsum_rows(dbeta_prod, rows, R, R, dbeta_vec, 0, st));
  if (dG) {
    TnProblem t;
    memset(&t, 0, sizeof(t));
    t.P = D; t.Q = R; t.npairs = 1;
    t.pair[0] = TnPair{E, dGpre, D, R, (size_t)rows};
    RE2NN_CUDA(run_tn(t, partial, dG, 0, YLoadPlain{}, st));
  }
  if (dE) {
    memset(&g, 0, sizeof(g));
    g.M = rows; g.N = D; g.nseg = 1; g.ndir = 1;
    g.seg[0][0] = GemmSeg{dGpre, G, R, R, R, 1, 0, 0};
    RE2NN_CUDA(launch_simt_gemm(g, EpiStore{dE, D, nullptr}, ALoadPlain{}, st));
  }
  return 0;
}

}  // extern "C"
