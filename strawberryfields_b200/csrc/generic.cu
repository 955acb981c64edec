// Strided gather / product / reduce and elementwise helpers (K8, K10-K14 in SURVEY.md).
//
//   C[c(o)] = sum_r A[a(o) + ta(r)] * op(B[b(o) + tb(r)])
//
// One kernel family covers what the reference does with np.einsum / np.transpose /
// np.tensordot on the host: pure -> mixed outer product (fockbackend/ops.py:110-120),
// partial trace (ops.py:144-157), diagonal / marginal photon-number distributions
// (ops.py:123-131, circuit.py:675-677, states.py:596-608), reduced density matrices
// (backend.py:219-253, states.py:613-642), project-and-reset (ops.py:179-198) and the
// mode permutations of state preparation (circuit.py:441-473).
#include "common.cuh"

namespace b200 {

constexpr int GR_THREADS = 256;
constexpr int GR_MAX_SPLIT = 64;

struct OutOffsets {
  long long a, b, c;
};

// Mixed-radix decode of a linear index.  Indices below 2^32 (every single-GPU state) take the
// 32-bit divide path: a 64-bit divide is ~5x the instructions and made this kernel ALU-bound.
__device__ __forceinline__ OutOffsets decode_out(const b200_gather_desc& d, unsigned long long o) {
  OutOffsets r{d.base_a, d.base_b, d.base_c};
  if ((o >> 32) == 0) {
    unsigned o32 = (unsigned)o;
    for (int j = d.n_out_axes - 1; j >= 0; --j) {
      unsigned e = (unsigned)d.out_ext[j];
      unsigned q = o32 / e;
      long long dig = (long long)(o32 - q * e);
      o32 = q;
      r.a += dig * d.out_sa[j];
      r.b += dig * d.out_sb[j];
      r.c += dig * d.out_sc[j];
    }
    return r;
  }
  for (int j = d.n_out_axes - 1; j >= 0; --j) {
    unsigned long long q = o / (unsigned)d.out_ext[j];
    long long dig = (long long)(o - q * (unsigned)d.out_ext[j]);
    o = q;
    r.a += dig * d.out_sa[j];
    r.b += dig * d.out_sb[j];
    r.c += dig * d.out_sc[j];
  }
  return r;
}
__device__ __forceinline__ void decode_red(const b200_gather_desc& d, unsigned long long r, long long& ta,
                                           long long& tb) {
  ta = 0;
  tb = 0;
  if ((r >> 32) == 0) {
    unsigned r32 = (unsigned)r;
    for (int j = d.n_red_axes - 1; j >= 0; --j) {
      unsigned e = (unsigned)d.red_ext[j];
      unsigned q = r32 / e;
      long long dig = (long long)(r32 - q * e);
      r32 = q;
      ta += dig * d.red_ta[j];
      tb += dig * d.red_tb[j];
    }
    return;
  }
  for (int j = d.n_red_axes - 1; j >= 0; --j) {
    unsigned long long q = r / (unsigned)d.red_ext[j];
    long long dig = (long long)(r - q * (unsigned)d.red_ext[j]);
    r = q;
    ta += dig * d.red_ta[j];
    tb += dig * d.red_tb[j];
  }
}

__device__ __forceinline__ cplx load_term(const void* A, const cplx* B, long long ia, long long ib, int flags) {
  cplx v;
  if (flags & 4) v = make_double2(reinterpret_cast<const double*>(A)[ia], 0.0);
  else v = reinterpret_cast<const cplx*>(A)[ia];
  if (B) {
    cplx w = B[ib];
    if (flags & 1) w.y = -w.y;
    v = cmul(v, w);
  }
  return v;
}
__device__ __forceinline__ void store_out(void* C, long long ic, cplx v, int flags) {
  if (flags & 2) reinterpret_cast<double*>(C)[ic] = v.x;
  else reinterpret_cast<cplx*>(C)[ic] = v;
}

// pure copy / product (no reduction): four output elements per thread, all loads issued before
// the first store, so ~128 KB are in flight per SM (one element per thread left the kernel
// latency-bound at a third of the HBM rate -- it is the pack/unpack of the multi-GPU exchange)
constexpr int GC_UNROLL = 4;
__global__ void __launch_bounds__(GR_THREADS)
k_gather_copy(const b200_gather_desc d, const void* __restrict__ A, const cplx* __restrict__ B,
              void* __restrict__ C, unsigned long long n_out, int flags) {
  const unsigned long long tile = (unsigned long long)GR_THREADS * GC_UNROLL;
  const unsigned long long o0 = (unsigned long long)blockIdx.x * tile + threadIdx.x;
  OutOffsets off[GC_UNROLL];
  cplx v[GC_UNROLL];
#pragma unroll
  for (int u = 0; u < GC_UNROLL; ++u) {
    unsigned long long o = o0 + (unsigned long long)u * GR_THREADS;
    if (o < n_out) off[u] = decode_out(d, o);
  }
#pragma unroll
  for (int u = 0; u < GC_UNROLL; ++u) {
    unsigned long long o = o0 + (unsigned long long)u * GR_THREADS;
    if (o < n_out) v[u] = load_term(A, B, off[u].a, off[u].b, flags);
  }
#pragma unroll
  for (int u = 0; u < GC_UNROLL; ++u) {
    unsigned long long o = o0 + (unsigned long long)u * GR_THREADS;
    if (o < n_out) store_out(C, off[u].c, v[u], flags);
  }
}

// one thread per output element, sequential reduction (small n_red)
__global__ void __launch_bounds__(GR_THREADS)
k_gather_thread(const b200_gather_desc d, const void* __restrict__ A, const cplx* __restrict__ B,
                void* __restrict__ C, unsigned long long n_out, unsigned long long n_red, int flags) {
  unsigned long long o = (unsigned long long)blockIdx.x * GR_THREADS + threadIdx.x;
  if (o >= n_out) return;
  OutOffsets off = decode_out(d, o);
  cplx acc = make_double2(0.0, 0.0);
  if (d.n_red_axes == 0) {
    acc = load_term(A, B, off.a, off.b, flags);
  } else {
    for (unsigned long long r = 0; r < n_red; ++r) {
      long long ta, tb;
      decode_red(d, r, ta, tb);
      acc = cadd(acc, load_term(A, B, off.a + ta, off.b + tb, flags));
    }
  }
  store_out(C, off.c, acc, flags);
}

// one CTA per (output element, reduction chunk): lanes stride over r, shuffle + smem reduce.
// grid (n_out, split).  split > 1 writes partials to part[chunk * n_out + o].
__global__ void __launch_bounds__(GR_THREADS)
k_gather_block(const b200_gather_desc d, const void* __restrict__ A, const cplx* __restrict__ B,
               void* __restrict__ C, cplx* __restrict__ part, unsigned long long n_out,
               unsigned long long n_red, int flags) {
  __shared__ cplx warp_sum[GR_THREADS / 32];
  unsigned long long o = blockIdx.x;
  const unsigned split = gridDim.y, chunk = blockIdx.y;
  unsigned long long per = (n_red + split - 1) / split;
  unsigned long long r0 = chunk * per, r1 = r0 + per < n_red ? r0 + per : n_red;
  OutOffsets off = decode_out(d, o);
  cplx acc = make_double2(0.0, 0.0);
  for (unsigned long long r = r0 + threadIdx.x; r < r1; r += GR_THREADS) {
    long long ta, tb;
    decode_red(d, r, ta, tb);
    acc = cadd(acc, load_term(A, B, off.a + ta, off.b + tb, flags));
  }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) {
    acc.x += __shfl_down_sync(0xffffffffu, acc.x, s);
    acc.y += __shfl_down_sync(0xffffffffu, acc.y, s);
  }
  if ((threadIdx.x & 31) == 0) warp_sum[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    cplx t = warp_sum[0];
    for (int w = 1; w < GR_THREADS / 32; ++w) t = cadd(t, warp_sum[w]);
    if (split == 1) store_out(C, off.c, t, flags);
    else part[(unsigned long long)chunk * n_out + o] = t;
  }
}

__global__ void __launch_bounds__(GR_THREADS)
k_gather_finish(const b200_gather_desc d, void* __restrict__ C, const cplx* __restrict__ part,
                unsigned long long n_out, int split, int flags) {
  unsigned long long o = (unsigned long long)blockIdx.x * GR_THREADS + threadIdx.x;
  if (o >= n_out) return;
  cplx t = make_double2(0.0, 0.0);
  for (int s = 0; s < split; ++s) t = cadd(t, part[(unsigned long long)s * n_out + o]);
  store_out(C, decode_out(d, o).c, t, flags);
}

// ---- elementwise ---------------------------------------------------------------------------
__global__ void k_fill_zero(cplx* __restrict__ p, long long n) {
  long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    p[i] = make_double2(0.0, 0.0);
}
__global__ void k_set_element(cplx* p, long long idx, double re, double im) { p[idx] = make_double2(re, im); }

__global__ void k_abs2(const cplx* __restrict__ psi, double* __restrict__ out, long long n) {
  long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    cplx v = psi[i];
    out[i] = fma(v.x, v.x, v.y * v.y);
  }
}

constexpr int NORM_BLOCKS = 148 * 8;
__global__ void __launch_bounds__(256) k_norm2_partial(const cplx* __restrict__ psi, long long n,
                                                        double* __restrict__ part) {
  __shared__ double ws[8];
  double acc = 0.0;
  long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    cplx v = psi[i];
    acc = fma(v.x, v.x, fma(v.y, v.y, acc));
  }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, s);
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += ws[w];
    part[blockIdx.x] = t;
  }
}
__global__ void __launch_bounds__(256) k_norm2_final(const double* __restrict__ part, int n, double* out) {
  __shared__ double ws[8];
  double acc = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) acc += part[i];
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, s);
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += ws[w];
    *out = t;
  }
}

__global__ void k_scale(cplx* __restrict__ p, long long n, double re, double im, const double* divisor,
                        int sqrt_div) {
  cplx f = make_double2(re, im);
  if (divisor) {
    double dv = *divisor;
    if (sqrt_div) dv = sqrt(dv);
    f.x /= dv;
    f.y /= dv;
  }
  long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) p[i] = cmul(p[i], f);
}

// out[b][j][i] = f[b][j] * in[b][i]: a product factor becomes the new OUTERMOST axis of the tensor (lazy vacuum,
// DESIGN 4.7).  One coalesced read of the old tensor, nf coalesced writes; two inputs per thread in flight.  The
// generic gather kernel decodes a mixed-radix index per output element and wrote this at 1.2 TB/s.
__global__ void __launch_bounds__(256)
k_outer_axis(const cplx* __restrict__ in, const cplx* __restrict__ f, cplx* __restrict__ out, long long n_in, int nf,
             long long in_bs, long long f_bs) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cplx* F = reinterpret_cast<cplx*>(smem_raw);
  const int b = blockIdx.z;
  for (int j = threadIdx.x; j < nf; j += blockDim.x) F[j] = f[(size_t)b * f_bs + j];
  __syncthreads();
  const cplx* ib = in + (size_t)b * in_bs;
  cplx* ob = out + (size_t)b * n_in * nf;
  const long long stride = (long long)gridDim.x * blockDim.x * 2;
  for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 2; i < n_in; i += stride) {
    const cplx v0 = ib[i];
    const bool two = i + 1 < n_in;
    const cplx v1 = two ? ib[i + 1] : make_double2(0.0, 0.0);
    for (int j = 0; j < nf; ++j) {
      const cplx w = F[j];
      cplx* o = ob + (size_t)j * n_in + i;
      if (two && ((reinterpret_cast<size_t>(o) & 31) == 0)) {
        // two amplitudes = one 32-byte sector per thread
        double4 pk = make_double4(fma(v0.x, w.x, -v0.y * w.y), fma(v0.x, w.y, v0.y * w.x),
                                  fma(v1.x, w.x, -v1.y * w.y), fma(v1.x, w.y, v1.y * w.x));
        *reinterpret_cast<double4*>(o) = pk;
      } else {
        o[0] = cmul(v0, w);
        if (two) o[1] = cmul(v1, w);
      }
    }
  }
}

static unsigned ew_blocks(long long n, int thr) {
  long long want = (n + thr - 1) / thr;
  long long cap = 148ll * 16;
  return (unsigned)(want < 1 ? 1 : (want < cap ? want : cap));
}

}  // namespace b200

using namespace b200;

extern "C" {

int b200_gather_reduce(const b200_gather_desc* desc, const void* A_dev, const b200_c128* B_dev, void* C_dev,
                       int flags, b200_c128* part_dev, void* stream) {
  B200_CHECK_ARG(desc && A_dev && C_dev, "gather_reduce: null pointer");
  B200_CHECK_ARG(desc->n_out_axes >= 0 && desc->n_out_axes <= B200_MAX_AXES && desc->n_red_axes >= 0 &&
                     desc->n_red_axes <= B200_MAX_AXES,
                 "gather_reduce: rank out of range");
  B200_CHECK_ARG(!((flags & 4) && B_dev), "gather_reduce: real input cannot be combined with B");
  unsigned long long n_out = 1, n_red = 1;
  for (int j = 0; j < desc->n_out_axes; ++j) {
    B200_CHECK_ARG(desc->out_ext[j] >= 1, "gather_reduce: bad output extent");
    n_out *= (unsigned long long)desc->out_ext[j];
  }
  for (int j = 0; j < desc->n_red_axes; ++j) {
    B200_CHECK_ARG(desc->red_ext[j] >= 1, "gather_reduce: bad reduction extent");
    n_red *= (unsigned long long)desc->red_ext[j];
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (desc->n_red_axes == 0) {
    unsigned long long per = (unsigned long long)GR_THREADS * GC_UNROLL;
    unsigned long long blocks = (n_out + per - 1) / per;
    B200_CHECK_ARG(blocks < (1ull << 31), "gather_reduce: output too large for one launch");
    k_gather_copy<<<(unsigned)blocks, GR_THREADS, 0, st>>>(*desc, A_dev, (const cplx*)B_dev, C_dev, n_out, flags);
    return cuda_status("gather_copy");
  }
  if (n_red <= 64 || n_out >= (1ull << 16)) {
    unsigned long long blocks = (n_out + GR_THREADS - 1) / GR_THREADS;
    B200_CHECK_ARG(blocks < (1ull << 31), "gather_reduce: output too large for one launch");
    k_gather_thread<<<(unsigned)blocks, GR_THREADS, 0, st>>>(*desc, A_dev, (const cplx*)B_dev, C_dev, n_out,
                                                             n_red, flags);
    return cuda_status("gather_thread");
  }
  // CTA per output; split the reduction so that the grid has >= ~4 CTAs per SM
  int split = 1;
  if (part_dev) {
    unsigned long long want = (148ull * 4 + n_out - 1) / n_out;
    unsigned long long max_by_work = n_red / (GR_THREADS * 4ull);
    if (want > max_by_work) want = max_by_work;
    if (want > GR_MAX_SPLIT) want = GR_MAX_SPLIT;
    if (want > 1) split = (int)want;
  }
  dim3 grid((unsigned)n_out, split);
  k_gather_block<<<grid, GR_THREADS, 0, st>>>(*desc, A_dev, (const cplx*)B_dev, C_dev, (cplx*)part_dev, n_out,
                                              n_red, flags);
  int rc = cuda_status("gather_block");
  if (rc || split == 1) return rc;
  k_gather_finish<<<(unsigned)((n_out + GR_THREADS - 1) / GR_THREADS), GR_THREADS, 0, st>>>(
      *desc, C_dev, (const cplx*)part_dev, n_out, split, flags);
  return cuda_status("gather_finish");
}

int b200_fill_zero(b200_c128* dev, int64_t n, void* stream) {
  B200_CHECK_ARG(dev && n >= 0, "fill_zero: bad args");
  if (n == 0) return 0;
  k_fill_zero<<<ew_blocks(n, 256), 256, 0, (cudaStream_t)stream>>>((cplx*)dev, n);
  return cuda_status("fill_zero");
}

int b200_outer_axis(const b200_c128* in_dev, const b200_c128* f_dev, b200_c128* out_dev, int64_t n_in, int nf,
                    int nbatch, int64_t in_batch_stride, int64_t f_batch_stride, void* stream) {
  B200_CHECK_ARG(in_dev && f_dev && out_dev, "outer_axis: null pointer");
  B200_CHECK_ARG(n_in >= 1 && nf >= 1 && nf <= B200_MAX_CUTOFF * B200_MAX_CUTOFF && nbatch >= 1 && nbatch <= 65535,
                 "outer_axis: bad geometry");
  dim3 grid(ew_blocks((n_in + 1) / 2, 256), 1, nbatch);
  k_outer_axis<<<grid, 256, (size_t)nf * sizeof(cplx), (cudaStream_t)stream>>>(
      (const cplx*)in_dev, (const cplx*)f_dev, (cplx*)out_dev, n_in, nf, in_batch_stride, f_batch_stride);
  return cuda_status("outer_axis");
}

int b200_set_element(b200_c128* dev, int64_t index, double re, double im, void* stream) {
  B200_CHECK_ARG(dev && index >= 0, "set_element: bad args");
  k_set_element<<<1, 1, 0, (cudaStream_t)stream>>>((cplx*)dev, index, re, im);
  return cuda_status("set_element");
}

int b200_abs2(const b200_c128* psi_dev, double* probs_dev, int64_t n, void* stream) {
  B200_CHECK_ARG(psi_dev && probs_dev && n >= 1, "abs2: bad args");
  k_abs2<<<ew_blocks(n, 256), 256, 0, (cudaStream_t)stream>>>((const cplx*)psi_dev, probs_dev, n);
  return cuda_status("abs2");
}

int b200_norm2(const b200_c128* psi_dev, int64_t n, double* out_dev, double* part_dev, void* stream) {
  B200_CHECK_ARG(psi_dev && out_dev && part_dev && n >= 1, "norm2: bad args");
  unsigned blocks = ew_blocks(n, 256);
  if (blocks > NORM_BLOCKS) blocks = NORM_BLOCKS;
  k_norm2_partial<<<blocks, 256, 0, (cudaStream_t)stream>>>((const cplx*)psi_dev, n, part_dev);
  int rc = cuda_status("norm2_partial");
  if (rc) return rc;
  k_norm2_final<<<1, 256, 0, (cudaStream_t)stream>>>(part_dev, (int)blocks, out_dev);
  return cuda_status("norm2_final");
}

int b200_scale(b200_c128* dev, int64_t n, double re, double im, const double* divisor_dev, int sqrt_div,
               void* stream) {
  B200_CHECK_ARG(dev && n >= 1, "scale: bad args");
  k_scale<<<ew_blocks(n, 256), 256, 0, (cudaStream_t)stream>>>((cplx*)dev, n, re, im, divisor_dev, sqrt_div);
  return cuda_status("scale");
}

}  // extern "C"
