// Gates that touch the INNERMOST tensor axis, staged through shared memory by the bulk-copy
// engine (TMA) in a persistent, warp-specialised pipeline.
//
// When a gate axis has stride 1, consecutive slices are D (one-mode gate) or D*stride apart, so the
// lanes of a warp cannot read neighbouring 16-byte words and the register-streaming kernel
// (apply.cu) drops to 3-4.6 TB/s.  Round 1 staged the slices with per-thread cp.async (LDGSTS.128):
// 3.99 TB/s, every 32-byte sector requested twice from L2 (two 16-byte requests), and load, compute
// and write-back serialised by CTA-wide barriers.  Here
//   * two DMA warps move whole tiles with cp.async.bulk: a loader (global -> shared, completion on an
//     mbarrier) and a storer (shared -> global, bulk groups), the handful of copies of a tile dealt to
//     their lanes: full-sector requests, no per-element address arithmetic, and the write-back is
//     asynchronous too;
//   * ten consumer warps run the same register-blocked tasks as the streaming kernel on the staged
//     tile, in place (different tasks of a slice touch disjoint elements), and hand the stage on
//     through mbarriers -- there is no __syncthreads in the steady state;
//   * a CTA per SM walks a contiguous range of tiles with a ring of 3-4 stages, so 2-3 tiles
//     (100-150 KB per SM) are in flight while one is being computed: the FP64 work (~1400 of the
//     ~3500 cycles a tile's bytes take at the HBM rate) is off the critical path;
//   * the staged layout is conflict-free.  Lanes own slices, and a slice is SS = D or D*D 16-byte words
//     long: for even SS the lanes of a quarter warp would hit 2-4 of the 8 16-byte bank groups (ncu,
//     first version: 86 M of 154 M shared-memory wavefronts were bank conflicts and the LSU pipe was 76 %
//     busy -- the kernel was shared-memory bound at 4.6 TB/s).  With r = 8 / gcd(SS mod 8, 8), a row of
//     slices is therefore staged in granules of 4r slices, each followed by ONE pad word (one bulk copy
//     per granule), and lane l of a group of 32 slices takes slice 4r * (x mod 8/r) + x / (8/r) + r * (l / 8),
//     x = l mod 8: the word offsets of a quarter warp are then a + (SS mod 8) * b with a = 0 .. 8/r - 1 and r
//     consecutive b -- eight different bank groups.
//
// Tile shapes (state viewed as [outer][k: D][mid][l: D] for a pair gate on (axis k, innermost),
// [slices][l: D] for a one-mode gate on the innermost axis):
//   contiguous -- one-mode gates, and pair gates with mid < 32: a tile is a contiguous chunk of
//                 whole slices / whole outer blocks;
//   rows       -- pair gates with mid >= 32: D rows (pitch stride_k) of 32 slices each.
//
// Replaces, for these geometries, the same reference sites as apply.cu:
// Circuit.apply_gate_BLAS (fockbackend/circuit.py:118-217) and apply_twomode_gate (219-365).
#include "tasks.cuh"
#include "tma.cuh"

namespace b200 {

constexpr int IN_CONSUMERS = 10;                       // consumer warps
constexpr int IN_THREADS = 32 * (IN_CONSUMERS + 2);    // + a loader and a storer warp
constexpr int IN_MAX_STAGES = 4;
constexpr int IN_COPY_MAX = 16 * 1024;                 // bytes per bulk copy
constexpr int IN_SMEM_LIMIT = 227 * 1024 - 1024;       // dynamic shared memory budget
constexpr int IN_TILE_TARGET = 48 * 1024;              // small cutoffs stage several lane groups per tile

struct InnerPlan {
  int rows_mode;            // 0: contiguous tiles, 1: D rows of slices_per_tile slices
  int groups;               // lane groups (32 slices each) per tile
  int slices_per_tile;
  int slice_elems;          // elements of a slice: D (one-mode gate) or D * D
  // staged origin of slice s of a tile: (s / block_slices) * block_elems + (s % block_slices) * row_ss
  //                                     + s / gran   (one pad word per granule of `gran` slices; gran = 0: none)
  int block_slices, block_elems, row_ss;
  int gran, r;              // padding granule (slices) and r = 8 / gcd(row_ss mod 8, 8) of the lane map
  int sk, sl;               // staged element strides of gate index 1 / 2 inside a slice
  int rows, row_pitch;      // rows mode: D rows, staged row pitch in elements
  int chunks_per_outer;     // rows mode: ceil(mid / slices_per_tile)
  long long hi_stride;      // rows mode: element stride of the outer gate axis
  unsigned long long tiles_per_state;
  int stages;
  int tile_elems;           // staged elements per stage
};

// tile -> number of valid slices (the last tile of a state / of a row run may be partial)
__device__ __forceinline__ int tile_slices(const InnerPlan& p, const Geometry& g, unsigned long long t) {
  if (p.rows_mode) {
    const unsigned c = (unsigned)(t % (unsigned)p.chunks_per_outer);
    const unsigned left = g.mid - c * (unsigned)p.slices_per_tile;
    return (int)(left < (unsigned)p.slices_per_tile ? left : (unsigned)p.slices_per_tile);
  }
  const unsigned long long s0 = t * (unsigned long long)p.slices_per_tile;
  const unsigned long long left = (unsigned long long)g.n_slices - s0;
  return (int)(left < (unsigned long long)p.slices_per_tile ? left : (unsigned long long)p.slices_per_tile);
}

// a row of `nsl` slices of `ss` elements, contiguous in memory, staged in granules with one pad word each.
// The copies of a tile are dealt round-robin to the 32 lanes of the DMA warp (`turn` counts them): one
// thread issuing every operation sustains ~2.1-2.8 TB/s of copies, several reach the HBM rate
// (tools/probes/tma_probe.cu).
template <bool LOAD>
__device__ __forceinline__ void row_copy(const InnerPlan& p, cplx* gp, unsigned sa, int nsl, int ss, unsigned bar,
                                         int lane, int& turn) {
  const int gran = p.gran > 0 ? p.gran : nsl;
  const unsigned gbytes = (unsigned)gran * (unsigned)ss * 16u;
  for (int s0 = 0, gi = 0; s0 < nsl; s0 += gran, ++gi) {
    const int n = nsl - s0 < gran ? nsl - s0 : gran;
    const unsigned bytes = (unsigned)n * (unsigned)ss * 16u;
    const unsigned so = (unsigned)gi * (gbytes + (p.gran > 0 ? 16u : 0u));
    const char* gsrc = reinterpret_cast<const char*>(gp) + (size_t)gi * gbytes;
    for (unsigned off = 0; off < bytes; off += IN_COPY_MAX) {
      if ((turn++ & 31) != lane) continue;
      const unsigned m = bytes - off < (unsigned)IN_COPY_MAX ? bytes - off : (unsigned)IN_COPY_MAX;
      if (LOAD) bulk_load(sa + so + off, gsrc + off, m, bar);
      else bulk_store(const_cast<char*>(gsrc) + off, sa + so + off, m);
    }
  }
}

// the bulk copies of one tile: LOAD global -> stage (arming the stage's mbarrier), else stage -> global
template <bool LOAD>
__device__ __forceinline__ void tile_copy(const InnerPlan& p, const Geometry& g, cplx* base, unsigned long long t,
                                          unsigned stage_smem, unsigned bar, int lane) {
  const int nsl = tile_slices(p, g, t);
  int turn = 0;
  if (LOAD) {
    // the expected byte count is armed before any lane's copy can complete
    if (lane == 0) mbar_expect_tx(bar, (unsigned)nsl * (unsigned)p.slice_elems * 16u);
    __syncwarp();
  }
  if (!p.rows_mode) {
    // whole slices / whole outer blocks: contiguous in memory
    cplx* gp = base + t * (unsigned long long)p.slices_per_tile * (unsigned long long)p.slice_elems;
    row_copy<LOAD>(p, gp, stage_smem, nsl, p.slice_elems, bar, lane, turn);
  } else {
    const unsigned long long o = t / (unsigned)p.chunks_per_outer;
    const unsigned c = (unsigned)(t % (unsigned)p.chunks_per_outer);
    cplx* gp = base + (long long)o * g.outer_step + (long long)c * p.slices_per_tile * p.rows;
    for (int k = 0; k < p.rows; ++k)
      row_copy<LOAD>(p, gp + (long long)k * p.hi_stride, stage_smem + (unsigned)k * (unsigned)p.row_pitch * 16u, nsl,
                     p.rows, bar, lane, turn);
  }
}

template <int D>
__global__ void __launch_bounds__(IN_THREADS, 1)
k_apply_inner_tma(cplx* __restrict__ state, const cplx* __restrict__ coef, const Geometry g, const TaskTable tt,
                  const InnerPlan p, unsigned long long n_tiles /* over all batch entries */) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) unsigned long long full_bar[IN_MAX_STAGES], done_bar[IN_MAX_STAGES], free_bar[IN_MAX_STAGES];
  cplx* tiles = reinterpret_cast<cplx*>(smem_raw);
  cplx* M = tiles + (size_t)p.stages * p.tile_elems;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&done_bar[s]), IN_CONSUMERS);
      mbar_init(smem_u32(&free_bar[s]), 1);
    }
    mbar_fence_init();
  }
  __syncthreads();

  // this CTA's contiguous range of tiles
  const unsigned long long per = (n_tiles + gridDim.x - 1) / gridDim.x;
  const unsigned long long t_lo = (unsigned long long)blockIdx.x * per;
  const unsigned long long t_hi = t_lo + per < n_tiles ? t_lo + per : n_tiles;
  if (t_lo >= t_hi) return;
  const unsigned long long n_my = t_hi - t_lo;
  const unsigned S = (unsigned)p.stages;

  if (warp >= IN_CONSUMERS) {
    // ---------------- DMA warps: a loader and a storer, their 32 lanes share a tile's copies ---------
    auto where = [&](unsigned long long i, cplx*& base, unsigned long long& t) {
      const unsigned long long gt = t_lo + i;
      const unsigned long long b = gt / p.tiles_per_state;
      t = gt - b * p.tiles_per_state;
      base = state + (size_t)b * g.state_batch_stride;
    };
    cplx* base;
    unsigned long long t;
    if (warp == IN_CONSUMERS) {
      for (unsigned long long i = 0; i < n_my; ++i) {
        const unsigned st = (unsigned)(i % S);
        if (i >= S) mbar_wait(smem_u32(&free_bar[st]), (unsigned)(((i / S) - 1) & 1));  // tile i-S has left the stage
        where(i, base, t);
        tile_copy<true>(p, g, base, t, smem_u32(tiles + (size_t)st * p.tile_elems), smem_u32(&full_bar[st]), lane);
      }
    } else {
      for (unsigned long long i = 0; i < n_my; ++i) {
        const unsigned st = (unsigned)(i % S);
        mbar_wait(smem_u32(&done_bar[st]), (unsigned)((i / S) & 1));  // consumers are done with tile i
        fence_async_smem();
        where(i, base, t);
        tile_copy<false>(p, g, base, t, smem_u32(tiles + (size_t)st * p.tile_elems), 0, lane);
        bulk_commit();  // every lane commits its own (possibly empty) group per tile
        if (i >= 1) {
          bulk_wait_read<1>();  // this lane's stores of tile i-1 have read the stage
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&free_bar[(unsigned)((i - 1) % S)]));
        }
      }
      bulk_wait_all();  // every store is complete before the kernel ends
    }
    return;
  }

  // ---------------- consumer warps ------------------------------------------------------------
  const long long step = p.sk + tt.dl * p.sl;
  const int n_wt = p.groups * tt.ntasks;
  // conflict-free lane -> slice map inside a group of 32 slices (identity when no padding is needed)
  int lane_slice = lane;
  if (p.gran > 0) {
    const int x = lane & 7, w = 8 / p.r;
    lane_slice = 4 * p.r * (x % w) + x / w + p.r * (lane >> 3);
  }
  long long cur_batch = -1;
  for (unsigned long long i = 0; i < n_my; ++i) {
    const unsigned long long gt = t_lo + i;
    const long long b = (long long)(gt / p.tiles_per_state);
    const unsigned long long t = gt - (unsigned long long)b * p.tiles_per_state;
    if (b != cur_batch && (cur_batch < 0 || g.coef_batch_stride != 0)) {
      // (re)load the gate table: every consumer has finished the previous batch entry's tiles
      asm volatile("bar.sync 1, %0;\n" ::"n"(IN_CONSUMERS * 32) : "memory");
      const cplx* cg = coef + (size_t)b * g.coef_batch_stride;
      for (int e = threadIdx.x; e < g.coef_count; e += IN_CONSUMERS * 32) {
        cplx v = cg[e];
        if (g.conj) v.y = -v.y;
        M[e] = v;
      }
      asm volatile("bar.sync 1, %0;\n" ::"n"(IN_CONSUMERS * 32) : "memory");
    }
    cur_batch = b;
    const unsigned st = (unsigned)(i % S);
    mbar_wait(smem_u32(&full_bar[st]), (unsigned)((i / S) & 1));
    cplx* tile = tiles + (size_t)st * p.tile_elems;
    const int nsl = tile_slices(p, g, t);
    for (int wt = warp; wt < n_wt; wt += IN_CONSUMERS) {
      const int gi = wt / tt.ntasks, task = wt - gi * tt.ntasks;
      const int sg = gi * 32 + lane_slice;
      if (sg < nsl) {
        cplx* ps = tile + (sg / p.block_slices) * p.block_elems + (sg % p.block_slices) * p.row_ss +
                   (p.gran > 0 ? sg / p.gran : 0);
        const SubBlock sb0 = tt.sub[task][0], sb1 = tt.sub[task][1];
        task_dispatch<D>(sb0.c, ps + sb0.start_k * p.sk + sb0.start_l * p.sl,
                         ps + sb1.start_k * p.sk + sb1.start_l * p.sl, step, M + sb0.coef, M + sb1.coef);
      }
    }
    fence_async_smem();  // generic-proxy writes to the stage become visible to the bulk store
    __syncwarp();
    if (lane == 0) mbar_arrive(smem_u32(&done_bar[st]));
  }
}

template <int D>
static cudaError_t launch_d(cplx* state, const cplx* coef, const Geometry& g, const TaskTable& tt,
                            const InnerPlan& p, unsigned long long n_tiles, size_t smem, cudaStream_t st) {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
  }
  cudaError_t e = cudaFuncSetAttribute(k_apply_inner_tma<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  const unsigned grid = (unsigned)(n_tiles < (unsigned long long)sms ? n_tiles : (unsigned long long)sms);
  k_apply_inner_tma<D><<<grid, IN_THREADS, smem, st>>>(state, coef, g, tt, p, n_tiles);
  return cudaSuccess;
}

static int gcd_int(int a, int b) { return b == 0 ? a : gcd_int(b, a % b); }

bool launch_inner_tma(int D, cplx* state, const cplx* coef, const Geometry& g, const TaskTable& tt, int nbatch,
                      cudaStream_t st, int* status) {
  if (D < 2 || D > B200_MAX_FAST_CUTOFF || g.inner != 1) return false;
  const bool pair = g.stride2 != 0;
  InnerPlan p;
  memset(&p, 0, sizeof(p));
  const int group_bytes = 32 * D * D * 16;                 // one lane group of a pair gate
  int q = IN_TILE_TARGET / group_bytes;
  if (q < 1) q = 1;
  auto pad_plan = [&](int ss) {  // bank-conflict-free staging of a row of slices of `ss` 16-byte words
    p.r = 8 / gcd_int(ss % 8, 8);
    p.gran = p.r == 8 ? 0 : 4 * p.r;
  };
  if (!pair) {
    // one-mode gate on the innermost axis: [slices][D], a tile = q * D lane groups of whole slices
    p.groups = q * D;
    p.slices_per_tile = 32 * p.groups;
    p.slice_elems = D;
    p.block_slices = p.slices_per_tile;
    p.block_elems = 0;
    p.row_ss = D;
    pad_plan(D);
    p.sk = 1;
    p.sl = 0;
    p.tile_elems = p.slices_per_tile * D + (p.gran ? p.slices_per_tile / p.gran : 0);
    p.tiles_per_state = ((unsigned long long)g.n_slices + p.slices_per_tile - 1) / p.slices_per_tile;
  } else {
    const long long hi = g.stride1 > g.stride2 ? g.stride1 : g.stride2;
    const unsigned mid = g.mid;
    p.slice_elems = D * D;
    int hi_staged;
    if (mid >= 32u * (unsigned)q) {
      p.rows_mode = 1;
      p.groups = q;
      p.slices_per_tile = 32 * q;
      p.rows = D;
      p.row_ss = D;
      pad_plan(D);
      p.row_pitch = p.slices_per_tile * D + (p.gran ? p.slices_per_tile / p.gran : 0);
      p.chunks_per_outer = (int)((mid + p.slices_per_tile - 1) / p.slices_per_tile);
      p.hi_stride = hi;
      p.block_slices = p.slices_per_tile;
      p.block_elems = 0;
      hi_staged = p.row_pitch;
      p.tile_elems = D * p.row_pitch;
      p.tiles_per_state = (unsigned long long)(g.n_slices / mid) * p.chunks_per_outer;
    } else if (mid == 1) {
      // adjacent axes: every slice is D * D contiguous elements, a tile a row of 32 * q slices
      p.groups = q;
      p.slices_per_tile = 32 * q;
      p.block_slices = p.slices_per_tile;
      p.block_elems = 0;
      p.row_ss = D * D;
      pad_plan(D * D);
      hi_staged = D;
      p.tile_elems = p.slices_per_tile * D * D + (p.gran ? p.slices_per_tile / p.gran : 0);
      p.tiles_per_state = ((unsigned long long)g.n_slices + p.slices_per_tile - 1) / p.slices_per_tile;
    } else {
      // [outer][k][mid][l] with a short mid: whole outer blocks (D * mid * D contiguous elements), unpadded
      int n_o = (32 * q) / (int)mid;
      p.slices_per_tile = n_o * (int)mid;
      p.groups = (p.slices_per_tile + 31) / 32;
      p.block_slices = (int)mid;
      p.block_elems = D * (int)mid * D;
      p.row_ss = D;
      p.gran = 0;
      p.r = 8;
      hi_staged = (int)mid * D;
      p.tile_elems = n_o * p.block_elems;
      p.tiles_per_state = ((unsigned long long)g.n_slices + p.slices_per_tile - 1) / p.slices_per_tile;
    }
    p.sk = g.stride1 > g.stride2 ? hi_staged : 1;
    p.sl = g.stride1 > g.stride2 ? 1 : hi_staged;
  }
  const size_t tile_bytes = (((size_t)p.tile_elems * 16) + 127) / 128 * 128;
  p.tile_elems = (int)(tile_bytes / 16);
  const size_t coef_bytes = (size_t)g.coef_count * 16;
  if (coef_bytes + 3 * tile_bytes > (size_t)IN_SMEM_LIMIT) return false;  // fewer than three stages: not worth it
  p.stages = (int)((IN_SMEM_LIMIT - coef_bytes) / tile_bytes);
  if (p.stages > IN_MAX_STAGES) p.stages = IN_MAX_STAGES;
  const size_t smem = (size_t)p.stages * tile_bytes + coef_bytes;
  const unsigned long long n_tiles = p.tiles_per_state * (unsigned long long)nbatch;
  cudaError_t e = cudaSuccess;
#define B200_LAUNCH(N) \
  case N:              \
    e = launch_d<N>(state, coef, g, tt, p, n_tiles, smem, st); \
    break;
  switch (D) {
    B200_LAUNCH(2) B200_LAUNCH(3) B200_LAUNCH(4) B200_LAUNCH(5) B200_LAUNCH(6) B200_LAUNCH(7) B200_LAUNCH(8)
    B200_LAUNCH(9) B200_LAUNCH(10) B200_LAUNCH(11) B200_LAUNCH(12) B200_LAUNCH(13) B200_LAUNCH(14)
    B200_LAUNCH(15) B200_LAUNCH(16)
    default: return false;
  }
#undef B200_LAUNCH
  if (e != cudaSuccess) {
    *status = fail((int)e, "apply_inner_tma: shared memory opt-in failed: %s", cudaGetErrorString(e));
    return true;
  }
  *status = cuda_status("apply_inner_tma");
  return true;
}

}  // namespace b200
