// Training forward of the decompose recurrence in ONE resident launch: the kernel of recurrence_resident.cuh with the
// epilogue functors that also write what BPTT needs (u, pre-activation, state and operand slabs per step; the layout
// is the one the per-step launches of recurrence.cu produce: model_decompose_single.py:138-269 under autograd).
// Its own translation unit: twelve more instantiations of the resident kernel would double recurrence.cu's build.
#include "gemm_tc.cuh"
#include "recurrence_resident.cuh"

namespace re2nn {

template <int PREC, int NL>
static cudaError_t resident_train_farnn(int farnn, const ResidentLaunch& RL, const StepParams& p, size_t sS, size_t sR, int B,
                                        cudaStream_t st) {
  if (farnn == 0) return launch_resident_policy(RL, ResidentForward<PREC, NL, 0, true>{p, sS, sR}, B, st);
  if (farnn == 1) return launch_resident_policy(RL, ResidentForward<PREC, NL, 1, true>{p, sS, sR}, B, st);
  return launch_resident_policy(RL, ResidentForward<PREC, NL, 2, true>{p, sS, sR}, B, st);
}

cudaError_t launch_resident_train(int prec, int nl, int farnn, const ResidentLaunch& RL, const StepParams& p, size_t sS,
                                  size_t sR, int B, cudaStream_t st) {
#ifdef RE2NN_HAVE_TC
  const bool th = nl == RE2NN_NL_TANH;
  if (prec == RE2NN_PREC_TF32X3)
    return th ? resident_train_farnn<RE2NN_PREC_TF32X3, RE2NN_NL_TANH>(farnn, RL, p, sS, sR, B, st)
              : resident_train_farnn<RE2NN_PREC_TF32X3, -1>(farnn, RL, p, sS, sR, B, st);
  if (prec == RE2NN_PREC_FP16X3)
    return th ? resident_train_farnn<RE2NN_PREC_FP16X3, RE2NN_NL_TANH>(farnn, RL, p, sS, sR, B, st)
              : resident_train_farnn<RE2NN_PREC_FP16X3, -1>(farnn, RL, p, sS, sR, B, st);
#endif
  return cudaErrorInvalidValue;
}

}  // namespace re2nn
