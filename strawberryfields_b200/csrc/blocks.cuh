// Dense c x c block applied in place to c amplitudes that form an arithmetic progression
// in memory (global or shared): the inner loop shared by the streaming kernel (apply.cu)
// and the tile kernel (tile.cu).
#pragma once
#include "common.cuh"

namespace b200 {

// y = M x for a C x C block with x held in registers; rows are produced two at a time and each
// row keeps FOUR independent FMA chains (re*re, im*im, re*im, im*re), i.e. eight chains in
// flight: the FP64 pipe's fixed dependent-issue latency ("wait" stalls in ncu) is covered by
// instruction-level parallelism instead of by occupancy.
template <int C, typename IdxT>
__device__ __forceinline__ void rows_apply(cplx* __restrict__ p, IdxT step, const cplx* __restrict__ M,
                                           const cplx (&x)[C]) {
#pragma unroll 1
  for (int a = 0; a + 1 < C; a += 2) {
    double axx = 0.0, ayy = 0.0, axy = 0.0, ayx = 0.0;
    double bxx = 0.0, byy = 0.0, bxy = 0.0, byx = 0.0;
    const cplx* __restrict__ Ma = M + a * C;
    const cplx* __restrict__ Mb = Ma + C;
#pragma unroll
    for (int j = 0; j < C; ++j) {
      const cplx ma = Ma[j], mb = Mb[j], v = x[j];
      axx = fma(ma.x, v.x, axx);
      ayy = fma(ma.y, v.y, ayy);
      axy = fma(ma.x, v.y, axy);
      ayx = fma(ma.y, v.x, ayx);
      bxx = fma(mb.x, v.x, bxx);
      byy = fma(mb.y, v.y, byy);
      bxy = fma(mb.x, v.y, bxy);
      byx = fma(mb.y, v.x, byx);
    }
    p[a * step] = make_double2(axx - ayy, axy + ayx);
    p[(a + 1) * step] = make_double2(bxx - byy, bxy + byx);
  }
  if (C & 1) {
    double axx = 0.0, ayy = 0.0, axy = 0.0, ayx = 0.0;
    const cplx* __restrict__ Ma = M + (C - 1) * C;
#pragma unroll
    for (int j = 0; j < C; ++j) {
      const cplx ma = Ma[j], v = x[j];
      axx = fma(ma.x, v.x, axx);
      ayy = fma(ma.y, v.y, ayy);
      axy = fma(ma.x, v.y, axy);
      ayx = fma(ma.y, v.x, ayx);
    }
    p[(C - 1) * step] = make_double2(axx - ayy, axy + ayx);
  }
}

template <int C, typename IdxT>
__device__ __forceinline__ void block_apply(cplx* __restrict__ p, IdxT step, const cplx* __restrict__ M) {
  cplx x[C];
#pragma unroll
  for (int j = 0; j < C; ++j) x[j] = p[j * step];
  rows_apply<C>(p, step, M, x);
}

// any block size (cutoffs above B200_MAX_FAST_CUTOFF): amplitudes staged in local memory
template <typename IdxT>
__device__ __noinline__ void block_apply_dyn(int c, cplx* __restrict__ p, IdxT step, const cplx* __restrict__ M) {
  cplx x[B200_MAX_CUTOFF];
  for (int j = 0; j < c; ++j) x[j] = p[j * step];
  for (int a = 0; a < c; ++a) {
    cplx acc = make_double2(0.0, 0.0);
    for (int j = 0; j < c; ++j) cfma(acc, M[a * c + j], x[j]);
    p[a * step] = acc;
  }
}

template <typename IdxT>
__device__ __forceinline__ void block_dispatch(int c, cplx* p, IdxT step, const cplx* M) {
  switch (c) {
    case 1: block_apply<1>(p, step, M); break;
    case 2: block_apply<2>(p, step, M); break;
    case 3: block_apply<3>(p, step, M); break;
    case 4: block_apply<4>(p, step, M); break;
    case 5: block_apply<5>(p, step, M); break;
    case 6: block_apply<6>(p, step, M); break;
    case 7: block_apply<7>(p, step, M); break;
    case 8: block_apply<8>(p, step, M); break;
    case 9: block_apply<9>(p, step, M); break;
    case 10: block_apply<10>(p, step, M); break;
    case 11: block_apply<11>(p, step, M); break;
    case 12: block_apply<12>(p, step, M); break;
    case 13: block_apply<13>(p, step, M); break;
    case 14: block_apply<14>(p, step, M); break;
    case 15: block_apply<15>(p, step, M); break;
    case 16: block_apply<16>(p, step, M); break;
    default: block_apply_dyn(c, p, step, M); break;
  }
}

// same, with the switch capped at a compile-time maximum block size (smaller code, fewer registers)
template <int CMAX, typename IdxT>
__device__ __forceinline__ void block_dispatch_upto(int c, cplx* p, IdxT step, const cplx* M) {
#define B200_CASE(N) \
  case N:            \
    if constexpr (N <= CMAX) block_apply<N>(p, step, M); \
    break;
  switch (c) {
    B200_CASE(1) B200_CASE(2) B200_CASE(3) B200_CASE(4) B200_CASE(5) B200_CASE(6) B200_CASE(7) B200_CASE(8)
    B200_CASE(9) B200_CASE(10) B200_CASE(11) B200_CASE(12) B200_CASE(13) B200_CASE(14) B200_CASE(15)
    B200_CASE(16)
    default: break;
  }
#undef B200_CASE
}

// ---- task tables: a two-axis operator cut into tasks of exactly D amplitudes per slice ----
constexpr int MAX_TASKS = B200_MAX_CUTOFF;

struct SubBlock {
  int c;        // members (0 = unused)
  int coef;     // offset of the c x c matrix in the packed table
  int start_k;  // first member's index on axis 1
  int start_l;  // first member's index on axis 2
};
struct TaskTable {
  int ntasks;
  int dl;  // per-member step on axis 2: -1 (SUM), +1 (DIFF), 0 (SINGLE)
  SubBlock sub[MAX_TASKS][2];
};

inline void build_tasks(int rule, int D, TaskTable& tt) {
  memset(&tt, 0, sizeof(tt));
  if (rule == B200_RULE_SINGLE) {
    tt.ntasks = 1;
    tt.dl = 0;
    tt.sub[0][0] = SubBlock{D, 0, 0, 0};
    return;
  }
  tt.dl = (rule == B200_RULE_SUM) ? -1 : +1;
  auto make = [&](int b) {
    int lo = blk_lo(b, D);
    SubBlock s;
    s.c = blk_size(b, D);
    s.coef = blk_off(b, D);
    s.start_k = lo;
    s.start_l = (rule == B200_RULE_SUM) ? (b - lo) : (lo - (b - (D - 1)));
    return s;
  };
  // blocks sorted by size: lower half b (size b+1) and upper half 2D-2-b (same size);
  // the i-th smallest is paired with the i-th largest so that every task holds D amplitudes.
  int order[2 * B200_MAX_CUTOFF];
  int n = 0;
  for (int b = 0; b <= D - 2; ++b) {
    order[n++] = b;
    order[n++] = 2 * D - 2 - b;
  }
  int t = 0;
  tt.sub[t++][0] = make(D - 1);  // the middle block, D members
  for (int i = 0; i < n / 2; ++i) {
    tt.sub[t][0] = make(order[n - 1 - i]);
    tt.sub[t][1] = make(order[i]);
    ++t;
  }
  tt.ntasks = t;
}

}  // namespace b200
