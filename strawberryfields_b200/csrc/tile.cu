// Tile pass: several gates per HBM round trip.
//
// A tile is the D x D x D sub-tensor spanned by two arbitrary axes (strides stride0 >
// stride1 > 1) and the innermost axis (stride 1) with every other index fixed:
// D^3 amplitudes that sit in memory as D^2 runs of D consecutive complex128 (160 B runs at
// D = 10, sector aligned), so loading / storing a tile is coalesced wherever its two outer
// axes are.  The kernel is PERSISTENT (one CTA per SM) and DOUBLE-BUFFERED: while the CTA
// applies the pass's operator list to group i of G tiles in shared memory, the cp.async
// copies of group i+1 are in flight into the other buffer, and the write-back of group i
// is fire-and-forget.  Operators are dense one-axis gates, diagonal gates and
// block-structured two-axis gates (BS / MZ / S2 / loss) on any of the three tile axes; the
// tiles are written back with the three axes optionally permuted, which is what lets the
// host scheduler rotate which logical mode occupies the innermost position so that any
// three modes can share a tile on the next pass (scheduler.py).
//
// Shared-memory layout: tile axis strides (s0, s1, 1) in 16-byte units are padded to odd
// numbers, so 8 consecutive lanes that differ along ANY axis hit 8 distinct 16-byte bank
// groups (LDS.128 is conflict-free per quarter warp).  Lanes run over the tile indices an
// operator does not touch; the warp index selects the task, so coefficient reads are
// warp-uniform shared-memory broadcasts and all lanes of a warp do identical work.  The
// operators' tables are staged once per CTA and reused for every group it processes.
#include "blocks.cuh"

namespace b200 {

constexpr int TILE_MAX_OPS = 16;
constexpr int TILE_KIND_DIAG = 3;

struct TileOpDev {
  int kind;  // B200_RULE_SINGLE / SUM / DIFF / TILE_KIND_DIAG
  int a1, a2;
  int conj;
  int coef;   // offset into the staged coefficient arena
  int table;  // index of the task table of this rule
};

struct TilePass {
  int nops;
  TileOpDev ops[TILE_MAX_OPS];
  int G;
  int s[3];        // padded shared-memory strides of the tile axes
  int tile_elems;  // padded elements per tile (odd)
  int out_perm[3];
  unsigned ntiles, ngroups, LO, MID;  // tile t -> (hi, mid, lo) = (t / (MID*LO), (t / LO) % MID, t % LO)
  long long stride0, stride1;
  long long state_batch_stride, coef_batch_stride;
  int coef_count;
};

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_le1() { asm volatile("cp.async.wait_group 1;\n" ::: "memory"); }

// tiles per group: as many as give <= 32 lanes for a pair operator, while two groups fit in smem
__host__ __device__ constexpr int tile_groups(int D) { return D >= 14 ? 1 : (32 / D < 1 ? 1 : 32 / D); }
// one warp per task of a pair operator; every thread owns at most one (tile, i1, i2) column
__host__ __device__ constexpr int tile_threads(int D) {
  int cols = tile_groups(D) * D * D;
  int w = D < 4 ? 4 : D;
  while (32 * w < cols) ++w;
  return 32 * w;
}

// tables[0] = SINGLE, [1] = SUM, [2] = DIFF.  The cutoff is a template parameter: every index
// computation divides by a constant and the block switch only holds sizes that can occur.
template <int D>
__global__ void __launch_bounds__(tile_threads(D), 1)
k_tile_pass(cplx* __restrict__ state, const cplx* __restrict__ coef, const TilePass P,
            const TaskTable* __restrict__ tables) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int D2 = D * D, G = tile_groups(D);
  cplx* buf0 = reinterpret_cast<cplx*>(smem_raw);
  cplx* buf1 = buf0 + (size_t)G * P.tile_elems;
  cplx* M = buf1 + (size_t)G * P.tile_elems;
  const int nthr = blockDim.x, tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31, nwarps = nthr >> 5;
  const int batch = blockIdx.z;
  cplx* base = state + (size_t)batch * P.state_batch_stride;

  // ---- stage coefficients once per CTA, one range per operator (conjugated for bra-side ops) ----
  const cplx* cg = coef + (size_t)batch * P.coef_batch_stride;
  for (int o = 0; o < P.nops; ++o) {
    const int off = P.ops[o].coef, cj = P.ops[o].conj;
    const int sz = P.ops[o].kind == TILE_KIND_DIAG ? D : (P.ops[o].kind == B200_RULE_SINGLE ? D2 : packed_size(D));
    for (int i = tid; i < sz; i += nthr) {
      cplx v = cg[off + i];
      if (cj) v.y = -v.y;
      M[off + i] = v;
    }
  }

  // ---- this thread's column: tile g of the group, indices (c1, c2) on tile positions 1 and 2 ----
  const bool has_col = tid < G * D2;
  const int cg_ = tid / D2, cr = tid - cg_ * D2;
  const int c1 = cr / D, c2 = cr - c1 * D;
  int ss[3];  // shared stride of the tile axis that lands on global tile position j
  for (int k = 0; k < 3; ++k) ss[P.out_perm[k]] = P.s[k];

  // global element offset of this thread's column in group `gr` (or -1 past the end)
  auto column_offset = [&](unsigned gr) -> long long {
    unsigned t = gr * (unsigned)G + (unsigned)cg_;
    if (!has_col || t >= P.ntiles) return -1;
    unsigned lo = t % P.LO, rest = t / P.LO;
    unsigned mid = rest % P.MID, hi = rest / P.MID;
    return (long long)hi * (D * P.stride0) + (long long)mid * (D * P.stride1) + (long long)lo * D +
           (long long)c1 * P.stride1 + c2;
  };
  // D^2 runs of D consecutive amplitudes per tile: consecutive lanes take consecutive c2, so every
  // cp.async instruction covers whole 16*D-byte runs; the address arithmetic is paid once per D copies.
  auto issue_load = [&](cplx* buf, long long off) {
    if (off >= 0) {
      const cplx* src = base + off;
      cplx* dst = &buf[cg_ * P.tile_elems + c1 * P.s[1] + c2];
#pragma unroll
      for (int i0 = 0; i0 < D; ++i0) cp_async16(dst + i0 * P.s[0], src + (long long)i0 * P.stride0);
    }
    cp_async_commit();
  };

  unsigned gr = blockIdx.x;
  long long off_cur = gr < P.ngroups ? column_offset(gr) : -1;
  issue_load(buf0, off_cur);
  for (int it = 0; gr < P.ngroups; gr += gridDim.x, ++it) {
    cplx* tile = (it & 1) ? buf1 : buf0;
    const unsigned gnext = gr + gridDim.x;
    const long long off_next = gnext < P.ngroups ? column_offset(gnext) : -1;
    issue_load((it & 1) ? buf0 : buf1, off_next);  // prefetch the next group into the other buffer
    cp_async_wait_le1();                             // everything but that prefetch has landed
    __syncthreads();
    const int ntile_here = (int)min((unsigned)G, P.ntiles - gr * (unsigned)G);

    for (int o = 0; o < P.nops; ++o) {
      const TileOpDev op = P.ops[o];
      if (op.kind == TILE_KIND_DIAG) {
        if (has_col && cg_ < ntile_here) {
          cplx* p = &tile[cg_ * P.tile_elems + c1 * P.s[1] + c2];
          const cplx fixed = M[op.coef + (op.a1 == 1 ? c1 : c2)];
#pragma unroll
          for (int i0 = 0; i0 < D; ++i0) {
            cplx f = op.a1 == 0 ? M[op.coef + i0] : fixed;
            p[i0 * P.s[0]] = cmul(p[i0 * P.s[0]], f);
          }
        }
      } else if (op.kind == B200_RULE_SINGLE) {
        // slices = (tile, two other axes); one task; lanes over slices, warps over lane groups
        const int oa = op.a1 == 0 ? 1 : 0, ob = op.a1 == 2 ? 1 : 2;  // the two untouched axes, oa slower
        const int nsl = ntile_here * D2;
        const int step = P.s[op.a1];
        for (int s0 = warp * 32; s0 < nsl; s0 += nwarps * 32) {
          int s = s0 + lane;
          if (s < nsl) {
            int g = s / D2, r = s - g * D2;
            cplx* p = &tile[g * P.tile_elems + (r / D) * P.s[oa] + (r % D) * P.s[ob]];
            block_apply<D>(p, step, M + op.coef);
          }
        }
      } else {
        // pair operator on (a1, a2); third axis a3 + tile index give the lanes, warps take tasks
        const int a3 = 3 - op.a1 - op.a2;
        const TaskTable& tt = tables[op.table];
        const int nsl = ntile_here * D;
        const int ngroups = (nsl + 31) / 32;
        const int step = P.s[op.a1] + tt.dl * P.s[op.a2];
        for (int wt = warp; wt < ngroups * tt.ntasks; wt += nwarps) {
          int gi = wt / tt.ntasks, task = wt - gi * tt.ntasks;
          int s = gi * 32 + lane;
          if (s < nsl) {
            int g = s / D, i3 = s - g * D;
            cplx* ps = &tile[g * P.tile_elems + i3 * P.s[a3]];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const SubBlock sb = tt.sub[task][h];
              if (sb.c == 0) continue;
              block_dispatch_upto<D>(sb.c, ps + sb.start_k * P.s[op.a1] + sb.start_l * P.s[op.a2], step,
                                     M + op.coef + sb.coef);
            }
          }
        }
      }
      __syncthreads();
    }

    // ---- write back (fire and forget), tile axis k going to global tile position out_perm[k].
    //      The thread's column keeps the same (tile, position-1, position-2) indices, so the
    //      global offset computed for the load is reused. ----
    if (off_cur >= 0) {
      cplx* dst = base + off_cur;
      const cplx* src = &tile[cg_ * P.tile_elems + c1 * ss[1] + c2 * ss[2]];
      cplx v[D];
#pragma unroll
      for (int j0 = 0; j0 < D; ++j0) v[j0] = src[j0 * ss[0]];
#pragma unroll
      for (int j0 = 0; j0 < D; ++j0) dst[(long long)j0 * P.stride0] = v[j0];
    }
    off_cur = off_next;
    __syncthreads();  // every read of this buffer is done before the next prefetch overwrites it
  }
}

// [SINGLE, SUM, DIFF] task tables per (device, cutoff), built once and kept for the process
static TaskTable* g_tables[16][B200_MAX_FAST_CUTOFF + 1] = {{nullptr}};
static int g_sm_count[16] = {0};

static int get_tables(int D, TaskTable** out, int* sms) {
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 16) return fail(B200_EUNSUPPORTED, "%s", "tile pass: device index out of range");
  if (g_sm_count[dev] == 0) {
    cudaError_t e = cudaDeviceGetAttribute(&g_sm_count[dev], cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess || g_sm_count[dev] <= 0) g_sm_count[dev] = 148;
  }
  *sms = g_sm_count[dev];
  if (g_tables[dev][D] == nullptr) {
    TaskTable* host = new TaskTable[3];
    build_tasks(B200_RULE_SINGLE, D, host[0]);
    build_tasks(B200_RULE_SUM, D, host[1]);
    build_tasks(B200_RULE_DIFF, D, host[2]);
    TaskTable* devp = nullptr;
    cudaError_t e = cudaMalloc(&devp, 3 * sizeof(TaskTable));
    // synchronous pageable copy: happens once per cutoff, before the first launch that reads it
    if (e == cudaSuccess) e = cudaMemcpy(devp, host, 3 * sizeof(TaskTable), cudaMemcpyHostToDevice);
    delete[] host;
    if (e != cudaSuccess) return fail((int)e, "tile pass: task table upload: %s", cudaGetErrorString(e));
    g_tables[dev][D] = devp;
  }
  *out = g_tables[dev][D];
  return 0;
}

static int tile_elems_of(int D) {
  int s1 = D | 1, s0 = (D * s1) | 1;
  return (D * s0) | 1;
}

}  // namespace b200

using namespace b200;

extern "C" {

int b200_tile_groups(int D) {
  if (D < 2 || D > B200_MAX_FAST_CUTOFF) return 0;
  return tile_groups(D);
}

int64_t b200_tile_smem_bytes(int D, int64_t coef_count) {
  if (D < 2 || D > B200_MAX_FAST_CUTOFF) return -1;
  return (2 * (int64_t)tile_groups(D) * tile_elems_of(D) + coef_count) * (int64_t)sizeof(cplx);
}

int b200_apply_tile_pass(b200_c128* state_dev, int64_t total, int D, int64_t stride0, int64_t stride1,
                         const b200_tile_op* ops, int nops, const int* out_perm, const b200_c128* coef_dev,
                         int64_t coef_count, int nbatch, int64_t state_batch_stride,
                         int64_t coef_batch_stride, void* stream) {
  B200_CHECK_ARG(state_dev && ops && out_perm, "tile_pass: null pointer");
  B200_CHECK_ARG(D >= 2 && D <= B200_MAX_FAST_CUTOFF, "tile_pass: cutoff outside 2..16");
  B200_CHECK_ARG(nops >= 0 && nops <= TILE_MAX_OPS && nbatch >= 1, "tile_pass: bad op count");
  B200_CHECK_ARG(stride1 >= D && stride1 % D == 0 && stride0 >= (int64_t)D * stride1 &&
                     stride0 % ((int64_t)D * stride1) == 0 && total % ((int64_t)D * stride0) == 0,
                 "tile_pass: strides do not tile the state");
  B200_CHECK_ARG(coef_count == 0 || coef_dev, "tile_pass: missing coefficients");
  int seen = 0;
  for (int k = 0; k < 3; ++k) {
    B200_CHECK_ARG(out_perm[k] >= 0 && out_perm[k] < 3, "tile_pass: bad permutation");
    seen |= 1 << out_perm[k];
  }
  B200_CHECK_ARG(seen == 7, "tile_pass: bad permutation");
  int64_t ntiles = total / ((int64_t)D * D * D);
  B200_CHECK_ARG(ntiles < (1ll << 31), "tile_pass: too many tiles for one launch");

  TilePass P;
  memset(&P, 0, sizeof(P));
  P.nops = nops;
  P.G = tile_groups(D);
  P.s[2] = 1;
  P.s[1] = D | 1;
  P.s[0] = (D * P.s[1]) | 1;
  P.tile_elems = tile_elems_of(D);
  for (int k = 0; k < 3; ++k) P.out_perm[k] = out_perm[k];
  P.ntiles = (unsigned)ntiles;
  P.ngroups = (unsigned)((ntiles + P.G - 1) / P.G);
  P.LO = (unsigned)(stride1 / D);
  P.MID = (unsigned)(stride0 / ((int64_t)D * stride1));
  P.stride0 = stride0;
  P.stride1 = stride1;
  P.state_batch_stride = state_batch_stride;
  P.coef_batch_stride = coef_batch_stride;
  P.coef_count = (int)coef_count;
  for (int o = 0; o < nops; ++o) {
    const b200_tile_op& u = ops[o];
    TileOpDev& d = P.ops[o];
    B200_CHECK_ARG(u.kind >= 0 && u.kind <= TILE_KIND_DIAG, "tile_pass: bad op kind");
    B200_CHECK_ARG(u.axis1 >= 0 && u.axis1 < 3, "tile_pass: bad op axis");
    bool pair = (u.kind == B200_RULE_SUM || u.kind == B200_RULE_DIFF);
    B200_CHECK_ARG(!pair || (u.axis2 >= 0 && u.axis2 < 3 && u.axis2 != u.axis1), "tile_pass: bad op axes");
    int64_t sz = u.kind == TILE_KIND_DIAG ? D : (u.kind == B200_RULE_SINGLE ? D * D : packed_size(D));
    B200_CHECK_ARG(u.coef_offset >= 0 && u.coef_offset + sz <= coef_count, "tile_pass: coefficient range");
    d.kind = u.kind;
    d.a1 = u.axis1;
    d.a2 = u.axis2;
    d.conj = u.conj;
    d.coef = (int)u.coef_offset;
    d.table = u.kind == B200_RULE_SUM ? 1 : (u.kind == B200_RULE_DIFF ? 2 : 0);
  }
  TaskTable* tables = nullptr;
  int sms = 148;
  int rc = get_tables(D, &tables, &sms);
  if (rc) return rc;
  size_t smem = (size_t)b200_tile_smem_bytes(D, coef_count);
  if (smem > 227 * 1024) return fail(B200_EUNSUPPORTED, "%s", "tile_pass: operators do not fit in shared memory");
  // persistent: one CTA per SM (shared between batch entries), each looping over its groups
  unsigned per_batch = (unsigned)((sms + nbatch - 1) / nbatch);
  if (per_batch < 1) per_batch = 1;
  unsigned gx = P.ngroups < per_batch ? P.ngroups : per_batch;
  dim3 grid(gx, 1, nbatch);
  cudaError_t e = cudaSuccess;
#define B200_LAUNCH(N)                                                                                         \
  case N:                                                                                                      \
    if (smem > 48 * 1024)                                                                                      \
      e = cudaFuncSetAttribute(k_tile_pass<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);        \
    if (e == cudaSuccess)                                                                                      \
      k_tile_pass<N><<<grid, tile_threads(N), smem, (cudaStream_t)stream>>>((cplx*)state_dev,                  \
                                                                            (const cplx*)coef_dev, P, tables); \
    break;
  switch (D) {
    B200_LAUNCH(2) B200_LAUNCH(3) B200_LAUNCH(4) B200_LAUNCH(5) B200_LAUNCH(6) B200_LAUNCH(7) B200_LAUNCH(8)
    B200_LAUNCH(9) B200_LAUNCH(10) B200_LAUNCH(11) B200_LAUNCH(12) B200_LAUNCH(13) B200_LAUNCH(14)
    B200_LAUNCH(15) B200_LAUNCH(16)
    default: return fail(B200_EUNSUPPORTED, "%s", "tile_pass: cutoff outside 2..16");
  }
#undef B200_LAUNCH
  if (e != cudaSuccess) return fail((int)e, "tile_pass: shared memory opt-in failed: %s", cudaGetErrorString(e));
  return cuda_status("tile_pass");
}

}  // extern "C"
