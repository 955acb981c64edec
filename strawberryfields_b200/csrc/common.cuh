// Shared device/host helpers for libb200fock.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/b200fock.h"

typedef double2 cplx;  // same layout as b200_c128 / numpy complex128

namespace b200 {

// ---- error plumbing ---------------------------------------------------------------
extern thread_local char g_err[512];
extern long long g_launches;

inline int fail(int code, const char* fmt, const char* a = "") {
  snprintf(g_err, sizeof(g_err), fmt, a);
  return code;
}
inline int cuda_status(const char* where) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    snprintf(g_err, sizeof(g_err), "%s: %s", where, cudaGetErrorString(e));
    return (int)e;
  }
  ++g_launches;
  return 0;
}
#define B200_CHECK_ARG(cond, msg) \
  do {                            \
    if (!(cond)) return b200::fail(B200_EINVAL, "%s", msg); \
  } while (0)

// ---- complex arithmetic (explicit FMAs: 4 DFMA per complex multiply-add) ------------
__device__ __forceinline__ cplx cmul(cplx a, cplx b) {
  return make_double2(fma(a.x, b.x, -a.y * b.y), fma(a.x, b.y, a.y * b.x));
}
__device__ __forceinline__ void cfma(cplx& acc, cplx m, cplx x) {
  acc.x = fma(m.x, x.x, acc.x);
  acc.x = fma(-m.y, x.y, acc.x);
  acc.y = fma(m.x, x.y, acc.y);
  acc.y = fma(m.y, x.x, acc.y);
}
__device__ __forceinline__ cplx cconj(cplx a) { return make_double2(a.x, -a.y); }
__device__ __forceinline__ cplx cadd(cplx a, cplx b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ cplx cscale(cplx a, double s) { return make_double2(a.x * s, a.y * s); }

// ---- block-packed two-axis operators ----------------------------------------------------
// Block b = 0 .. 2D-2 has c_b = min(b, 2D-2-b) + 1 members and a dense c_b x c_b matrix
// M_b[a][j] (row = output member, column = input member) stored at packed offset off_b.
//   SUM  rule (BS, MZ): b = i + j.      lo = max(0, b-D+1), member m -> (k, l) = (lo+m, b-lo-m)
//   DIFF rule (S2, loss): b = i - j + D-1. lo = max(0, b-D+1), member m -> (k, l) = (lo+m, lo+m-(b-D+1))
__host__ __device__ inline int blk_size(int b, int D) {
  int r = 2 * D - 2 - b;
  return (b < r ? b : r) + 1;
}
__host__ __device__ inline int blk_lo(int b, int D) { return b - D + 1 > 0 ? b - D + 1 : 0; }
__host__ __device__ inline long long sum_sq(long long n) { return n * (n + 1) * (2 * n + 1) / 6; }
__host__ __device__ inline int blk_off(int b, int D) {
  // sum_{b' < b} c_{b'}^2 ; c = 1..D for b' = 0..D-1, then D-1..1
  if (b <= D) return (int)sum_sq(b);
  long long r = 2 * D - 1 - b;  // remaining blocks have sizes r .. 1
  return (int)(sum_sq(D) + sum_sq(D - 1) - sum_sq(r));
}
__host__ __device__ inline int packed_size(int D) { return blk_off(2 * D - 1, D); }

}  // namespace b200
