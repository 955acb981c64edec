// Single-mode reduced density matrix of a ket in ONE read of the state:
//     rho[a][b] = sum_r psi[.., a, ..] * conj(psi[.., b, ..])      (r = every other index)
// -- the D x D Gram matrix of the [D x N/D] matricisation of psi along one axis.  It is what homodyne
// measurement, mean_photon, quad_expectation and wigner reduce to (reference: backend.py:219-253 +
// states.py:613-642 via np.einsum / np.tensordot on the host), and its diagonal is the photon-number
// marginal of the mode (circuit.py:675-677).  The generic gather/reduce kernel reads the state once per OUTPUT
// element (D x D times: 7.3 ms for 1e8 amplitudes); here every thread streams whole columns (the D amplitudes
// of one slice, front-loaded like the gate kernels) and keeps the upper triangle of the Gram matrix --
// D (D + 1) / 2 complex accumulators, independent FMA chains -- in registers; lanes, warps and CTAs are
// reduced once at the end (shuffles, shared memory, a deterministic finish kernel).
#include "blocks.cuh"

namespace b200 {

constexpr int GRAM_THREADS = 128;
constexpr int GRAM_CTAS = 148 * 3;

// the upper triangle is D (D + 1) accumulator registers: cutoffs up to 12 fit the 255-register budget of two
// CTAs per SM (cutoff 10: 110 accumulators + 20 amplitudes); larger cutoffs use the generic reduction
constexpr int GRAM_MAX_CUTOFF = 12;
template <int D, bool DIAG>
__global__ void __launch_bounds__(GRAM_THREADS, (DIAG || D <= 7) ? 3 : 2)
k_gram1(const cplx* __restrict__ psi, unsigned n_slices, unsigned inner, long long state_batch_stride,
        double* __restrict__ part /* [batch][cta][NACC] */) {
  constexpr int NP = DIAG ? D : D * (D + 1) / 2;   // pairs (a <= b)
  constexpr int NACC = DIAG ? D : 2 * NP;          // doubles per partial
  double acc[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) acc[i] = 0.0;
  const cplx* base = psi + (size_t)blockIdx.z * state_batch_stride;
  const unsigned stride = gridDim.x * GRAM_THREADS;
  for (unsigned s = blockIdx.x * GRAM_THREADS + threadIdx.x; s < n_slices; s += stride) {
    const unsigned i_in = s % inner, i_out = s / inner;
    const cplx* p = base + (size_t)i_out * D * inner + i_in;
    cplx x[D];
#pragma unroll
    for (int j = 0; j < D; ++j) x[j] = p[(size_t)j * inner];
    if constexpr (DIAG) {
#pragma unroll
      for (int a = 0; a < D; ++a) acc[a] = fma(x[a].x, x[a].x, fma(x[a].y, x[a].y, acc[a]));
    } else {
      int k = 0;
#pragma unroll
      for (int a = 0; a < D; ++a)
#pragma unroll
        for (int b = a; b < D; ++b) {
          // x[a] * conj(x[b])
          acc[2 * k] = fma(x[a].x, x[b].x, fma(x[a].y, x[b].y, acc[2 * k]));
          acc[2 * k + 1] = fma(x[a].y, x[b].x, fma(-x[a].x, x[b].y, acc[2 * k + 1]));
          ++k;
        }
    }
  }
  // lanes -> warp -> CTA
  __shared__ double ws[GRAM_THREADS / 32][NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) {
    double v = acc[i];
#pragma unroll
    for (int sft = 16; sft > 0; sft >>= 1) v += __shfl_down_sync(0xffffffffu, v, sft);
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5][i] = v;
  }
  __syncthreads();
  double* out = part + ((size_t)blockIdx.z * gridDim.x + blockIdx.x) * NACC;
  for (int i = threadIdx.x; i < NACC; i += GRAM_THREADS) {
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < GRAM_THREADS / 32; ++w) t += ws[w][i];
    out[i] = t;
  }
}

// partials of every CTA -> rho[batch][D][D] (Hermitian completion) or probs[batch][D]
template <bool DIAG>
__global__ void k_gram1_finish(const double* __restrict__ part, int n_ctas, int D, void* __restrict__ out) {
  const int NP = DIAG ? D : D * (D + 1) / 2, NACC = DIAG ? D : 2 * NP;
  const double* pb = part + (size_t)blockIdx.x * n_ctas * NACC;
  for (int k = threadIdx.x; k < NP; k += blockDim.x) {
    double re = 0.0, im = 0.0;
    for (int c = 0; c < n_ctas; ++c) {
      if (DIAG) {
        re += pb[(size_t)c * NACC + k];
      } else {
        re += pb[(size_t)c * NACC + 2 * k];
        im += pb[(size_t)c * NACC + 2 * k + 1];
      }
    }
    if (DIAG) {
      reinterpret_cast<double*>(out)[(size_t)blockIdx.x * D + k] = re;
    } else {
      // k -> (a, b), a <= b, row-major over the upper triangle
      int a = 0, rem = k;
      while (rem >= D - a) {
        rem -= D - a;
        ++a;
      }
      const int b = a + rem;
      cplx* rho = reinterpret_cast<cplx*>(out) + (size_t)blockIdx.x * D * D;
      rho[a * D + b] = make_double2(re, im);
      if (a != b) rho[b * D + a] = make_double2(re, -im);
    }
  }
}

template <int D>
static void launch_gram(const cplx* psi, unsigned n_slices, unsigned inner, long long sbs, int nbatch, int diag,
                        double* part, void* out, unsigned ctas, cudaStream_t st) {
  dim3 grid(ctas, 1, nbatch);
  if (diag) {
    k_gram1<D, true><<<grid, GRAM_THREADS, 0, st>>>(psi, n_slices, inner, sbs, part);
    k_gram1_finish<true><<<nbatch, 128, 0, st>>>(part, (int)ctas, D, out);
  } else if constexpr (D <= GRAM_MAX_CUTOFF) {
    k_gram1<D, false><<<grid, GRAM_THREADS, 0, st>>>(psi, n_slices, inner, sbs, part);
    k_gram1_finish<false><<<nbatch, 128, 0, st>>>(part, (int)ctas, D, out);
  }
}

}  // namespace b200

using namespace b200;

extern "C" {

int64_t b200_gram1_part_doubles(int D, int nbatch) {
  if (D < 1 || nbatch < 1) return 0;
  return (int64_t)nbatch * GRAM_CTAS * D * (D + 1);
}

int b200_gram1(const b200_c128* psi_dev, int64_t outer, int D, int64_t inner, int diag_only, void* out_dev,
               double* part_dev, int nbatch, int64_t state_batch_stride, void* stream) {
  B200_CHECK_ARG(psi_dev && out_dev && part_dev, "gram1: null pointer");
  B200_CHECK_ARG(D >= 1 && D <= (diag_only ? B200_MAX_FAST_CUTOFF : GRAM_MAX_CUTOFF),
                 "gram1: cutoff outside the compiled range (matrix: 1..12, marginal: 1..16)");
  B200_CHECK_ARG(outer >= 1 && inner >= 1 && nbatch >= 1 && nbatch <= 65535, "gram1: bad geometry");
  B200_CHECK_ARG(outer * inner < (1ll << 32), "gram1: too many slices for one launch");
  const unsigned n_slices = (unsigned)(outer * inner);
  unsigned ctas = (n_slices + GRAM_THREADS - 1) / GRAM_THREADS;
  if (ctas > (unsigned)GRAM_CTAS) ctas = GRAM_CTAS;
  cudaStream_t st = (cudaStream_t)stream;
#define B200_LAUNCH(N) \
  case N:              \
    launch_gram<N>((const cplx*)psi_dev, n_slices, (unsigned)inner, state_batch_stride, nbatch, diag_only, part_dev, \
                   out_dev, ctas, st);                                                                             \
    break;
  switch (D) {
    B200_LAUNCH(1) B200_LAUNCH(2) B200_LAUNCH(3) B200_LAUNCH(4) B200_LAUNCH(5) B200_LAUNCH(6) B200_LAUNCH(7)
    B200_LAUNCH(8) B200_LAUNCH(9) B200_LAUNCH(10) B200_LAUNCH(11) B200_LAUNCH(12) B200_LAUNCH(13) B200_LAUNCH(14)
    B200_LAUNCH(15) B200_LAUNCH(16)
    default: break;
  }
#undef B200_LAUNCH
  int rc = cuda_status("gram1");
  if (rc == 0) ++g_launches;  // two kernels per call
  return rc;
}

}  // extern "C"
