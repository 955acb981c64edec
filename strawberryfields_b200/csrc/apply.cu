// Gate application kernels (K1-K7 in SURVEY.md section 2): one universal "blocked
// arithmetic-progression contraction".
//
// Every gate on the path is block-diagonal over index sets that are arithmetic
// progressions in memory:
//   * dense one-mode gate on axis a:            one block of D members, step = stride(a)
//   * BSgate / MZgate (i+j = k+l conserved):     blocks b = i+j, members (lo+m, b-lo-m),
//                                                step = stride1 - stride2
//   * S2gate, loss superoperator (i-j conserved): blocks b = i-j+D-1, members (lo+m, lo+m-d),
//                                                step = stride1 + stride2
// A "slice" fixes all other indices.  Work is cut into tasks of exactly D amplitudes per
// slice (the middle block alone, every other block paired with its complement), so a
// two-mode gate has the same shape as a one-mode gate: N/D tasks, each loading D
// amplitudes into registers (all D loads in flight before the first FMA), multiplying by
// small dense matrices broadcast from shared memory, and storing D amplitudes back in
// place.  Lanes of a warp run over consecutive slices (the fastest remaining index), which
// makes every load/store a run of consecutive 16-byte amplitudes whenever the inner stride
// allows.  The cutoff is a template parameter (2..16) so that the register file holds
// exactly D amplitudes; larger cutoffs take a local-memory path.
//
// Replaces Circuit.apply_gate_BLAS (fockbackend/circuit.py:118-217),
// Circuit.apply_twomode_gate + numba kernels (circuit.py:219-365) and, for the loss
// channel, Circuit._apply_channel (circuit.py:65-87) of the reference.
#include "tasks.cuh"

namespace b200 {

constexpr int APPLY_THREADS = 256;

__device__ __forceinline__ void stage_coef(cplx* M, const cplx* cg, int n, int conj) {
  for (int i = threadIdx.x; i < n; i += APPLY_THREADS) {
    cplx v = cg[i];
    if (conj) v.y = -v.y;
    M[i] = v;
  }
  __syncthreads();
}

// grid: (slice groups, 1, nbatch).  CTA = 8 warps; warp-task = (32 consecutive slices, task).
// A CTA owns `groups_per_cta` slice groups x all tasks and its warps stride over them.
// D = 0: any cutoff (local-memory blocks).
template <int D>
__global__ void __launch_bounds__(APPLY_THREADS, (D >= 1 && D <= 12) ? 3 : 2)
k_apply_blocks(cplx* __restrict__ state, const cplx* __restrict__ coef, const Geometry g, const TaskTable tt,
               int groups_per_cta) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cplx* M = reinterpret_cast<cplx*>(smem_raw);
  const int batch = blockIdx.z;
  stage_coef(M, coef + (size_t)batch * g.coef_batch_stride, g.coef_count, g.conj);

  cplx* base = state + (size_t)batch * g.state_batch_stride;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long step = g.stride1 + (long long)tt.dl * g.stride2;
  const unsigned group0 = blockIdx.x * (unsigned)groups_per_cta;
  const int n_wt = groups_per_cta * tt.ntasks;
  for (int wt = warp; wt < n_wt; wt += APPLY_THREADS / 32) {
    const int gi = wt / tt.ntasks, task = wt - gi * tt.ntasks;
    const unsigned s = (group0 + gi) * 32u + lane;
    if (s >= g.n_slices) continue;
    const unsigned i_in = s % g.inner, rest = s / g.inner;
    const unsigned i_mid = rest % g.mid, i_out = rest / g.mid;
    cplx* ps = base + (long long)i_out * g.outer_step + (long long)i_mid * g.mid_step + i_in;
    const SubBlock sb0 = tt.sub[task][0], sb1 = tt.sub[task][1];
    cplx* p0 = ps + sb0.start_k * g.stride1 + sb0.start_l * g.stride2;
    cplx* p1 = ps + sb1.start_k * g.stride1 + sb1.start_l * g.stride2;
    if constexpr (D > 0) {
      task_dispatch<D>(sb0.c, p0, p1, step, M + sb0.coef, M + sb1.coef);
    } else {
      block_apply_dyn(sb0.c, p0, step, M + sb0.coef);
      if (sb1.c) block_apply_dyn(sb1.c, p1, step, M + sb1.coef);
    }
  }
}

// ---- gates that touch the INNERMOST axis: staged through shared memory ----------------------
// When a gate axis has stride 1, consecutive slices are D (one-mode gate) or D*stride apart, so
// the lanes of a warp cannot read neighbouring 16-byte words and the streaming kernel drops to
// 3-4.6 TB/s.  Here a CTA copies the slices it owns -- whole runs of D consecutive amplitudes --
// into shared memory with cp.async (coalesced), runs the same register-blocked tasks on the
// staged copy (slice stride padded to an odd number of 16-byte words: conflict-free LDS.128 with
// lanes over slices), and writes the runs back.
__device__ __forceinline__ void cp_async16_(void* smem, const void* gmem) {
  unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gmem));
}

__host__ __device__ constexpr int inner_threads(int D) { return 32 * (D < 4 ? 4 : D); }

template <int D>
__global__ void __launch_bounds__(inner_threads(D), (D <= 10 ? 3 : (D <= 12 ? 2 : 1)))
k_apply_inner(cplx* __restrict__ state, const cplx* __restrict__ coef, const Geometry g, const TaskTable tt,
              int rows /* D: pair gate, 1: one-mode gate */, int slices_per_cta) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cplx* M = reinterpret_cast<cplx*>(smem_raw);
  cplx* tile = M + g.coef_count;
  const int nthr = inner_threads(D), tid = threadIdx.x;
  const int batch = blockIdx.z;
  const int SL = (rows * D) | 1;  // padded slice stride
  const long long row_stride = g.stride1 > g.stride2 ? g.stride1 : g.stride2;  // outer gate axis (pair)
  cplx* base = state + (size_t)batch * g.state_batch_stride;
  const unsigned s0 = blockIdx.x * (unsigned)slices_per_cta;

  // stage: a thread owns (slice, l) and walks the rows of that slice
  for (int idx = tid; idx < slices_per_cta * D; idx += nthr) {
    const int sg = idx / D, l = idx - sg * D;
    const unsigned s = s0 + sg;
    if (s < g.n_slices) {
      const unsigned i_mid = s % g.mid, i_out = s / g.mid;
      const cplx* src = base + (long long)i_out * g.outer_step + (long long)i_mid * g.mid_step + l;
      cplx* dst = tile + sg * SL + l;
      for (int k = 0; k < rows; ++k) cp_async16_(dst + k * D, src + (long long)k * row_stride);
    }
  }
  {
    const cplx* cgp = coef + (size_t)batch * g.coef_batch_stride;
    for (int i = tid; i < g.coef_count; i += nthr) {
      cplx v = cgp[i];
      if (g.conj) v.y = -v.y;
      M[i] = v;
    }
  }
  asm volatile("cp.async.wait_all;\n" ::: "memory");
  __syncthreads();

  // compute on the staged copy: lanes over slices, warps over (lane group, task)
  const int warp = tid >> 5, lane = tid & 31, nwarps = nthr >> 5;
  const int sk = rows == 1 ? 1 : (g.stride1 > g.stride2 ? D : 1);  // staged stride of gate index 1
  const int sl = rows == 1 ? 0 : (g.stride1 > g.stride2 ? 1 : D);  // staged stride of gate index 2
  const long long step = sk + tt.dl * sl;
  const int ngroups = slices_per_cta / 32;
  for (int wt = warp; wt < ngroups * tt.ntasks; wt += nwarps) {
    const int gi = wt / tt.ntasks, task = wt - gi * tt.ntasks;
    const int sg = gi * 32 + lane;
    if (s0 + sg < g.n_slices) {
      cplx* ps = tile + sg * SL;
      const SubBlock sb0 = tt.sub[task][0], sb1 = tt.sub[task][1];
      task_dispatch<D>(sb0.c, ps + sb0.start_k * sk + sb0.start_l * sl, ps + sb1.start_k * sk + sb1.start_l * sl,
                       step, M + sb0.coef, M + sb1.coef);
    }
  }
  __syncthreads();

  // write back the runs
  for (int idx = tid; idx < slices_per_cta * D; idx += nthr) {
    const int sg = idx / D, l = idx - sg * D;
    const unsigned s = s0 + sg;
    if (s < g.n_slices) {
      const unsigned i_mid = s % g.mid, i_out = s / g.mid;
      cplx* dst = base + (long long)i_out * g.outer_step + (long long)i_mid * g.mid_step + l;
      const cplx* src = tile + sg * SL + l;
      for (int k = 0; k < rows; ++k) dst[(long long)k * row_stride] = src[k * D];
    }
  }
}

// ---- diagonal gates ---------------------------------------------------------------------
// Index arithmetic is 32-bit whenever the state has fewer than 2^32 elements (always true for
// one launch on one GPU: 180 GB / 16 B = 1.1e10 would need it, 1e9-1e10-element shards do not).
template <typename I>
struct DiagAxes {
  int naxes;
  int conj[B200_MAX_AXES];
  I stride[B200_MAX_AXES];
};

// state[e] *= prod_k tabs[k][digit_k(e)] -- every pending diagonal gate in one pass.
// No per-element divide: the index is split as e = (sr * SUP + d) * L + i with L = D^j (<= DIAG_LOW_MAX)
// the extent of the innermost axes and SUP = D rows per super-row.  The product over the axes inside a
// row (stride < L) is a table of L factors built once per CTA in shared memory; the axis of stride L (if it
// has a gate) contributes tab[d]; the product over the outer axes is one factor per super-row (the only
// integer divides, amortised over SUP * L elements).  Per element: one LDS.128 and two complex multiplies,
// four elements per thread in flight.
constexpr int DIAG_LOW_MAX = 1024;
template <typename I>
__global__ void __launch_bounds__(256)
k_apply_diag_multi(cplx* __restrict__ state, I nsuper, int SUP, int L, int D, int axisL, const DiagAxes<I> md,
                   const cplx* __restrict__ tabs, long long state_batch_stride, long long tab_batch_stride) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cplx* T = reinterpret_cast<cplx*>(smem_raw);   // [naxes][D]
  cplx* low = T + md.naxes * D;                  // [L]
  const int batch = blockIdx.z;
  const cplx* tg = tabs + (size_t)batch * tab_batch_stride;
  for (int i = threadIdx.x; i < md.naxes * D; i += blockDim.x) {
    cplx v = tg[i];
    if (md.conj[i / D]) v.y = -v.y;
    T[i] = v;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < L; i += blockDim.x) {
    cplx f = make_double2(1.0, 0.0);
    for (int k = 0; k < md.naxes; ++k)
      if (md.stride[k] < (I)L) f = cmul(f, T[k * D + (i / (int)md.stride[k]) % D]);
    low[i] = f;
  }
  __syncthreads();
  cplx* base = state + (size_t)batch * state_batch_stride;
  const I hiL = (I)L * (I)SUP;
  for (I sr = blockIdx.x; sr < nsuper; sr += gridDim.x) {
    cplx fhh = make_double2(1.0, 0.0);
    for (int k = 0; k < md.naxes; ++k)
      if (md.stride[k] >= hiL) fhh = cmul(fhh, T[k * D + (int)((sr / (md.stride[k] / hiL)) % (I)D)]);
    for (int d = 0; d < SUP; ++d) {
      const cplx fh = axisL >= 0 ? cmul(fhh, T[axisL * D + d]) : fhh;
      cplx* p = base + ((size_t)sr * SUP + d) * L;
      for (int i0 = threadIdx.x; i0 < L; i0 += 4 * 256) {
        cplx v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (i0 + u * 256 < L) v[u] = p[i0 + u * 256];
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (i0 + u * 256 < L) p[i0 + u * 256] = cmul(v[u], cmul(low[i0 + u * 256], fh));
      }
    }
  }
}

// state[e] *= tab[digit1(e) * D + digit2(e)]  (two-axis diagonal: cross-Kerr)
template <typename I>
__global__ void __launch_bounds__(256)
k_apply_diag_pair(cplx* __restrict__ state, I total, int D, I stride1, I stride2, const cplx* __restrict__ tab,
                  int conj, long long state_batch_stride, long long tab_batch_stride) {
  const int batch = blockIdx.z;
  const cplx* tg = tab + (size_t)batch * tab_batch_stride;
  cplx* base = state + (size_t)batch * state_batch_stride;
  const I nthreads = (I)gridDim.x * blockDim.x;
  const I uD = (I)D;
  for (I e = (I)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += nthreads) {
    int idx = (int)((e / stride1) % uD);
    if (stride2) idx = idx * D + (int)((e / stride2) % uD);
    cplx f = __ldg(&tg[idx]);
    if (conj) f.y = -f.y;
    base[e] = cmul(base[e], f);
  }
}

__global__ void k_mul_tables(long long n, const cplx* __restrict__ a, const cplx* __restrict__ b, int conj_b,
                             cplx* __restrict__ out) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  cplx w = b[i];
  if (conj_b) w.y = -w.y;
  out[i] = cmul(a[i], w);
}

// ---- host side ------------------------------------------------------------------------------
template <int D>
static cudaError_t launch_blocks_d(cplx* state, const cplx* coef, const Geometry& g, const TaskTable& tt,
                                   dim3 grid, size_t smem, int gpc, cudaStream_t st) {
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(k_apply_blocks<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  k_apply_blocks<D><<<grid, APPLY_THREADS, smem, st>>>(state, coef, g, tt, gpc);
  return cudaSuccess;
}

template <int D>
static cudaError_t launch_inner_d(cplx* state, const cplx* coef, const Geometry& g, const TaskTable& tt, int rows,
                                  int nbatch, cudaStream_t st) {
  const int spc = rows == 1 ? inner_threads(D) : 32;  // slices per CTA: one per thread / one lane group
  const int SL = (rows * D) | 1;
  size_t smem = ((size_t)g.coef_count + (size_t)spc * SL) * sizeof(cplx);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(k_apply_inner<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  dim3 grid((g.n_slices + spc - 1) / spc, 1, nbatch);
  k_apply_inner<D><<<grid, inner_threads(D), smem, st>>>(state, coef, g, tt, rows, spc);
  return cudaSuccess;
}

static int launch_blocks(int D, cplx* state, const cplx* coef, Geometry& g, const TaskTable& tt, int nbatch,
                         cudaStream_t st) {
  // a gate axis is the innermost one: TMA-staged persistent kernel (inner.cu); B200_INNER_LEGACY=1 keeps the
  // round-1 kernels (cp.async staging for pair gates, register streaming for one-mode gates)
  static const bool legacy_inner = [] {
    const char* v = getenv("B200_INNER_LEGACY");
    return v && v[0] == '1';
  }();
  if (g.inner == 1 && !legacy_inner) {
    int status = 0;
    if (launch_inner_tma(D, state, coef, g, tt, nbatch, st, &status)) return status;
  }
  if (g.inner == 1 && g.stride2 != 0 && D >= 2 && D <= B200_MAX_FAST_CUTOFF) {
    // a PAIR gate with one axis innermost: staged kernel (measured 3.97 vs 2.97 TB/s streaming at D = 10;
    // the one-mode gate on the innermost axis stays on the streaming kernel: 4.6 vs 3.4 TB/s staged)
    const int rows = g.stride2 == 0 ? 1 : D;
    cudaError_t e = cudaSuccess;
#define B200_LAUNCH(N) \
  case N:              \
    e = launch_inner_d<N>(state, coef, g, tt, rows, nbatch, st); \
    break;
    switch (D) {
      B200_LAUNCH(2) B200_LAUNCH(3) B200_LAUNCH(4) B200_LAUNCH(5) B200_LAUNCH(6) B200_LAUNCH(7) B200_LAUNCH(8)
      B200_LAUNCH(9) B200_LAUNCH(10) B200_LAUNCH(11) B200_LAUNCH(12) B200_LAUNCH(13) B200_LAUNCH(14)
      B200_LAUNCH(15) B200_LAUNCH(16)
      default: break;
    }
#undef B200_LAUNCH
    if (e != cudaSuccess) return fail((int)e, "apply: shared memory opt-in failed: %s", cudaGetErrorString(e));
    return cuda_status("apply_inner");
  }
  size_t smem = (size_t)g.coef_count * sizeof(cplx);
  unsigned n_groups = (g.n_slices + 31u) / 32u;
  // ~16 warp-tasks per CTA: two per warp, enough to amortise staging the gate table
  int gpc = tt.ntasks >= 16 ? 1 : (16 + tt.ntasks - 1) / tt.ntasks;
  unsigned n_cta = (n_groups + gpc - 1) / gpc;
  dim3 grid(n_cta, 1, nbatch);
  cudaError_t e = cudaSuccess;
#define B200_LAUNCH(N) \
  case N:              \
    e = launch_blocks_d<N>(state, coef, g, tt, grid, smem, gpc, st); \
    break;
  switch (D <= B200_MAX_FAST_CUTOFF ? D : 0) {
    B200_LAUNCH(1) B200_LAUNCH(2) B200_LAUNCH(3) B200_LAUNCH(4) B200_LAUNCH(5) B200_LAUNCH(6) B200_LAUNCH(7)
    B200_LAUNCH(8) B200_LAUNCH(9) B200_LAUNCH(10) B200_LAUNCH(11) B200_LAUNCH(12) B200_LAUNCH(13)
    B200_LAUNCH(14) B200_LAUNCH(15) B200_LAUNCH(16)
    default: e = launch_blocks_d<0>(state, coef, g, tt, grid, smem, gpc, st); break;
  }
#undef B200_LAUNCH
  if (e != cudaSuccess) return fail((int)e, "apply: shared memory opt-in failed: %s", cudaGetErrorString(e));
  return cuda_status("apply_blocks");
}

static unsigned diag_blocks(long long total) {
  long long want = (total + 255) / 256;
  return (unsigned)(want < 148 * 16 ? want : 148 * 16);
}

}  // namespace b200

using namespace b200;

extern "C" {

int b200_apply_gate1(b200_c128* state_dev, int64_t outer, int D, int64_t inner, const b200_c128* U_dev,
                     int conj, int nbatch, int64_t state_batch_stride, int64_t gate_batch_stride,
                     void* stream) {
  B200_CHECK_ARG(state_dev && U_dev, "apply_gate1: null pointer");
  B200_CHECK_ARG(D >= 1 && D <= B200_MAX_CUTOFF, "apply_gate1: cutoff out of range");
  B200_CHECK_ARG(outer >= 1 && inner >= 1 && nbatch >= 1, "apply_gate1: bad geometry");
  B200_CHECK_ARG(outer * inner < (1ll << 32) && inner < (1ll << 32), "apply_gate1: too many slices for one launch");
  Geometry g;
  g.n_slices = (unsigned)(outer * inner);
  g.inner = (unsigned)inner;
  g.mid = 1;
  g.outer_step = (long long)D * inner;
  g.mid_step = 0;
  g.stride1 = inner;
  g.stride2 = 0;
  g.state_batch_stride = state_batch_stride;
  g.coef_batch_stride = gate_batch_stride;
  g.coef_count = D * D;
  g.conj = conj;
  TaskTable tt;
  build_tasks(B200_RULE_SINGLE, D, tt);
  return launch_blocks(D, (cplx*)state_dev, (const cplx*)U_dev, g, tt, nbatch, (cudaStream_t)stream);
}

int b200_apply_gate2(b200_c128* state_dev, int64_t total, int D, int64_t stride1, int64_t stride2, int rule,
                     const b200_c128* packed_dev, int conj, int nbatch, int64_t state_batch_stride,
                     int64_t gate_batch_stride, void* stream) {
  B200_CHECK_ARG(state_dev && packed_dev, "apply_gate2: null pointer");
  B200_CHECK_ARG(rule == B200_RULE_SUM || rule == B200_RULE_DIFF, "apply_gate2: bad rule");
  B200_CHECK_ARG(D >= 1 && D <= B200_MAX_CUTOFF, "apply_gate2: cutoff out of range");
  B200_CHECK_ARG(stride1 >= 1 && stride2 >= 1 && stride1 != stride2 && nbatch >= 1, "apply_gate2: bad strides");
  int64_t hi = stride1 > stride2 ? stride1 : stride2, lo = stride1 > stride2 ? stride2 : stride1;
  B200_CHECK_ARG(hi % ((int64_t)D * lo) == 0 && total % ((int64_t)D * hi) == 0, "apply_gate2: strides do not tile the state");
  int64_t slices = total / ((int64_t)D * D);
  B200_CHECK_ARG(slices < (1ll << 32) && lo < (1ll << 32), "apply_gate2: too many slices for one launch");
  Geometry g;
  g.n_slices = (unsigned)slices;
  g.inner = (unsigned)lo;
  g.mid = (unsigned)(hi / ((int64_t)D * lo));
  g.outer_step = (long long)D * hi;
  g.mid_step = (long long)D * lo;
  g.stride1 = stride1;
  g.stride2 = stride2;
  g.state_batch_stride = state_batch_stride;
  g.coef_batch_stride = gate_batch_stride;
  g.coef_count = packed_size(D);
  g.conj = conj;
  TaskTable tt;
  build_tasks(rule, D, tt);
  return launch_blocks(D, (cplx*)state_dev, (const cplx*)packed_dev, g, tt, nbatch, (cudaStream_t)stream);
}

int b200_apply_diag(b200_c128* state_dev, int64_t total, int D, int64_t stride1, int64_t stride2,
                    const b200_c128* tab_dev, int conj, int nbatch, int64_t state_batch_stride,
                    int64_t tab_batch_stride, void* stream) {
  B200_CHECK_ARG(state_dev && tab_dev, "apply_diag: null pointer");
  B200_CHECK_ARG(D >= 1 && D <= B200_MAX_CUTOFF && total >= 1 && stride1 >= 1 && stride2 >= 0 && nbatch >= 1,
                 "apply_diag: bad geometry");
  dim3 grid(diag_blocks(total), 1, nbatch);
  if (total < (1ll << 32))
    k_apply_diag_pair<unsigned><<<grid, 256, 0, (cudaStream_t)stream>>>(
        (cplx*)state_dev, (unsigned)total, D, (unsigned)stride1, (unsigned)stride2, (const cplx*)tab_dev, conj,
        state_batch_stride, tab_batch_stride);
  else
    k_apply_diag_pair<unsigned long long><<<grid, 256, 0, (cudaStream_t)stream>>>(
        (cplx*)state_dev, (unsigned long long)total, D, (unsigned long long)stride1, (unsigned long long)stride2,
        (const cplx*)tab_dev, conj, state_batch_stride, tab_batch_stride);
  return cuda_status("apply_diag");
}

int b200_apply_diag_multi(b200_c128* state_dev, int64_t total, int D, int naxes, const int64_t* strides,
                          const int* conj_flags, const b200_c128* tabs_dev, int nbatch,
                          int64_t state_batch_stride, int64_t tab_batch_stride, void* stream) {
  B200_CHECK_ARG(state_dev && tabs_dev && strides && conj_flags, "apply_diag_multi: null pointer");
  B200_CHECK_ARG(naxes >= 1 && naxes <= B200_MAX_AXES, "apply_diag_multi: bad axis count");
  B200_CHECK_ARG(D >= 1 && D <= B200_MAX_CUTOFF && total >= 1 && nbatch >= 1, "apply_diag_multi: bad geometry");
  for (int k = 0; k < naxes; ++k) B200_CHECK_ARG(strides[k] >= 1, "apply_diag_multi: bad stride");
  // L = D^j: the largest power of the cutoff that divides the state and fits the row table; SUP = D rows
  // form a super-row when the state allows
  long long Lr = 1;
  while (D > 1 && Lr * D <= DIAG_LOW_MAX && total % (Lr * D) == 0) Lr *= D;
  const int SUP = (D > 1 && total % (Lr * D) == 0) ? D : 1;
  int axisL = -1;
  for (int k = 0; k < naxes; ++k) {
    if (SUP > 1 && strides[k] == Lr) axisL = k;
    B200_CHECK_ARG(strides[k] < Lr ? Lr % (strides[k] * D) == 0
                                   : (strides[k] == Lr && SUP > 1) || strides[k] % (Lr * SUP) == 0,
                   "apply_diag_multi: axis stride does not tile the state");
  }
  const long long nsuper = total / (Lr * SUP);
  long long want = nsuper < 148 * 8 ? nsuper : 148 * 8;
  dim3 grid((unsigned)want, 1, nbatch);
  size_t smem = ((size_t)naxes * D + (size_t)Lr) * sizeof(cplx);
  if (total < (1ll << 32)) {
    DiagAxes<unsigned> md;
    md.naxes = naxes;
    for (int k = 0; k < naxes; ++k) {
      md.stride[k] = (unsigned)strides[k];
      md.conj[k] = conj_flags[k];
    }
    k_apply_diag_multi<unsigned><<<grid, 256, smem, (cudaStream_t)stream>>>(
        (cplx*)state_dev, (unsigned)nsuper, SUP, (int)Lr, D, axisL, md, (const cplx*)tabs_dev, state_batch_stride,
        tab_batch_stride);
  } else {
    DiagAxes<unsigned long long> md;
    md.naxes = naxes;
    for (int k = 0; k < naxes; ++k) {
      md.stride[k] = (unsigned long long)strides[k];
      md.conj[k] = conj_flags[k];
    }
    k_apply_diag_multi<unsigned long long><<<grid, 256, smem, (cudaStream_t)stream>>>(
        (cplx*)state_dev, (unsigned long long)nsuper, SUP, (int)Lr, D, axisL, md, (const cplx*)tabs_dev,
        state_batch_stride, tab_batch_stride);
  }
  return cuda_status("apply_diag_multi");
}

int b200_mul_tables(int64_t n, const b200_c128* a_dev, const b200_c128* b_dev, int conj_b, b200_c128* out_dev,
                    void* stream) {
  B200_CHECK_ARG(n >= 1 && a_dev && b_dev && out_dev, "mul_tables: bad args");
  k_mul_tables<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(n, (const cplx*)a_dev,
                                                                            (const cplx*)b_dev, conj_b,
                                                                            (cplx*)out_dev);
  return cuda_status("mul_tables");
}

}  // extern "C"
