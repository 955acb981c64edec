// Geometry of a gate pass and the register-blocked task shared by the streaming kernel (apply.cu)
// and the TMA-staged innermost-axis kernel (inner.cu).
#pragma once
#include "blocks.cuh"

namespace b200 {

struct Geometry {
  // slice s (0 <= s < n_slices) -> element offset
  //   (s / (mid*inner)) * outer_step + ((s / inner) % mid) * mid_step + (s % inner)
  unsigned n_slices;
  unsigned inner, mid;
  long long outer_step, mid_step;
  long long stride1, stride2;  // element strides of the gate axes (stride2 = 0 for SINGLE)
  long long state_batch_stride;
  long long coef_batch_stride;
  int coef_count;  // packed entries to stage in shared memory
  int conj;
};

// one task = two blocks of C0 + C1 = D amplitudes (C1 = 0: a single block)
template <int C0, int C1>
__device__ __forceinline__ void task_apply(cplx* __restrict__ p0, cplx* __restrict__ p1, long long step,
                                           const cplx* __restrict__ M0, const cplx* __restrict__ M1) {
  cplx x0[C0];
  cplx x1[C1 > 0 ? C1 : 1];
#pragma unroll
  for (int j = 0; j < C0; ++j) x0[j] = p0[j * step];
#pragma unroll
  for (int j = 0; j < C1; ++j) x1[j] = p1[j * step];
  rows_apply<C0>(p0, step, M0, x0);
  if constexpr (C1 > 0) rows_apply<C1>(p1, step, M1, x1);
}

template <int D>
__device__ __forceinline__ void task_dispatch(int c0, cplx* p0, cplx* p1, long long step, const cplx* M0,
                                              const cplx* M1) {
#define B200_CASE(N) \
  case N:            \
    if constexpr (N <= D) task_apply<N, D - N>(p0, p1, step, M0, M1); \
    break;
  switch (c0) {
    B200_CASE(1) B200_CASE(2) B200_CASE(3) B200_CASE(4) B200_CASE(5) B200_CASE(6) B200_CASE(7) B200_CASE(8)
    B200_CASE(9) B200_CASE(10) B200_CASE(11) B200_CASE(12) B200_CASE(13) B200_CASE(14) B200_CASE(15)
    B200_CASE(16)
    default: break;
  }
#undef B200_CASE
}

// innermost-axis gates staged through shared memory by the bulk-copy engine (inner.cu); returns
// false when the geometry / cutoff is not handled (the caller then uses its own kernels)
bool launch_inner_tma(int D, cplx* state, const cplx* coef, const Geometry& g, const TaskTable& tt, int nbatch,
                      cudaStream_t st, int* status);

}  // namespace b200
