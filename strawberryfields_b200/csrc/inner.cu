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
//   * teams of ten consumer warps (two teams on alternate tiles for cutoffs up to 12) run the same
//     register-blocked tasks as the streaming kernel on the staged tile, in place (different tasks of a
//     slice touch disjoint elements), and hand the stage on through mbarriers -- there is no
//     __syncthreads in the steady state;
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
#include <cuda.h>  // CUtensorMap and its enums (types only: the encoder is fetched through the runtime)

#include <map>
#include <vector>

#include "tasks.cuh"
#include "tma.cuh"

namespace b200 {

constexpr int IN_TEAM = 10;                            // consumer warps that share one tile
// Two teams work on alternate tiles when the register file allows (cutoffs up to 12): twenty warps keep
// the FP64 pipe fed through the dependent-issue latency of the DFMA chains (ncu, one team: "wait" was the
// top stall and the pipe 31 % busy; the dense one-mode task was compute-latency bound at 3.8 TB/s).
__host__ __device__ constexpr int in_teams(int D) { return D <= 12 ? 2 : 1; }
__host__ __device__ constexpr int in_threads(int teams) { return 32 * (IN_TEAM * teams + 2); }  // + loader, storer
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
  int teams;                // consumer teams launched (1 or 2, at most in_teams(D))
  int use_tmap;             // 1: rows geometry, ONE 4-D box per tile; 2: contiguous tile, `boxes` 3-D boxes
  int box_units;            // 128-byte units of a box (box dim 1)
  int boxes;                // use_tmap == 2: boxes per tile (<= 256 units each)
  int swz;                  // the tensor map stages with SWIZZLE_128B: element word c lives at c ^ ((c >> 3) & 7)
  unsigned char lane_map[32];  // swz: slice (inside a group of 32) owned by each lane, conflict free by search
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
                                          unsigned stage_smem, unsigned bar, int lane, const CUtensorMap* tmap,
                                          int batch) {
  const int nsl = tile_slices(p, g, t);
  int turn = 0;
  if (p.use_tmap == 1) {
    // tensor [batch][outer * D rows][row run in 128-byte units][16 doubles]; box = [1][D][box_units][16]
    if (lane == 0) {
      const unsigned long long o = t / (unsigned)p.chunks_per_outer;
      const int c = (int)(t % (unsigned)p.chunks_per_outer);
      if (LOAD) {
        mbar_expect_tx(bar, (unsigned)p.box_units * 128u * (unsigned)p.rows);  // the whole box, zero fill included
        tensor_load_4d(stage_smem, tmap, 0, c * p.box_units, (int)(o * (unsigned)p.rows), batch, bar);
      } else {
        tensor_store_4d(tmap, 0, c * p.box_units, (int)(o * (unsigned)p.rows), batch, stage_smem);
      }
    }
    return;
  }
  if (p.use_tmap == 2) {
    // contiguous tile: tensor [batch][1][state in 128-byte units][16 doubles], `boxes` boxes of box_units units;
    // boxes that start behind the end of the state are not issued, one that straddles it is clipped
    if (lane == 0) {
      const long long u0 = (long long)t * p.boxes * p.box_units;
      const long long total_units = (long long)g.n_slices * p.slice_elems / 8;
      int nb = 0;
      for (int b = 0; b < p.boxes; ++b) nb += (u0 + (long long)b * p.box_units < total_units) ? 1 : 0;
      if (LOAD) mbar_expect_tx(bar, (unsigned)nb * (unsigned)p.box_units * 128u);
      for (int b = 0; b < nb; ++b) {
        const unsigned sa = stage_smem + (unsigned)b * (unsigned)p.box_units * 128u;
        if (LOAD) tensor_load_4d(sa, tmap, 0, (int)(u0 + (long long)b * p.box_units), 0, batch, bar);
        else tensor_store_4d(tmap, 0, (int)(u0 + (long long)b * p.box_units), 0, batch, sa);
      }
    }
    return;
  }
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

// ---- tasks on a SWIZZLE_128B stage ---------------------------------------------------------------------
// The tensor-map copies stage a tile densely but with the hardware's 128-byte swizzle: 16-byte word c of the
// stage is stored at c ^ ((c >> 3) & 7) (stage bases are 1024-byte aligned).  With the right lane -> slice map
// (searched on the host, InnerPlan::lane_map) the eight lanes of a quarter warp then hit eight different bank
// groups although their slices are an even number of words apart -- conflict free WITHOUT padding, so a tile is
// still one or two bulk operations.
__device__ __forceinline__ cplx* swz_at(cplx* base, int c) { return base + (c ^ ((c >> 3) & 7)); }

template <int C>
__device__ __forceinline__ void rows_apply_swz(cplx* __restrict__ base, int i0, int step, const cplx* __restrict__ M,
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
    *swz_at(base, i0 + a * step) = make_double2(axx - ayy, axy + ayx);
    *swz_at(base, i0 + (a + 1) * step) = make_double2(bxx - byy, bxy + byx);
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
    *swz_at(base, i0 + (C - 1) * step) = make_double2(axx - ayy, axy + ayx);
  }
}

template <int C0, int C1>
__device__ __forceinline__ void task_apply_swz(cplx* __restrict__ base, int i0, int i1, int step,
                                               const cplx* __restrict__ M0, const cplx* __restrict__ M1) {
  cplx x0[C0];
  cplx x1[C1 > 0 ? C1 : 1];
#pragma unroll
  for (int j = 0; j < C0; ++j) x0[j] = *swz_at(base, i0 + j * step);
#pragma unroll
  for (int j = 0; j < C1; ++j) x1[j] = *swz_at(base, i1 + j * step);
  rows_apply_swz<C0>(base, i0, step, M0, x0);
  if constexpr (C1 > 0) rows_apply_swz<C1>(base, i1, step, M1, x1);
}

template <int D>
__device__ __forceinline__ void task_dispatch_swz(int c0, cplx* base, int i0, int i1, int step, const cplx* M0,
                                                  const cplx* M1) {
#define B200_CASE(N) \
  case N:            \
    if constexpr (N <= D) task_apply_swz<N, D - N>(base, i0, i1, step, M0, M1); \
    break;
  switch (c0) {
    B200_CASE(1) B200_CASE(2) B200_CASE(3) B200_CASE(4) B200_CASE(5) B200_CASE(6) B200_CASE(7) B200_CASE(8)
    B200_CASE(9) B200_CASE(10) B200_CASE(11) B200_CASE(12) B200_CASE(13) B200_CASE(14) B200_CASE(15)
    B200_CASE(16)
    default: break;
  }
#undef B200_CASE
}

// TEAMS is a template parameter because it sets the register budget: one team (384 threads) compiles to
// ~133 registers, two teams (704 threads) are capped at 80 -- running one team under the 80-register cap
// cost the pair-gate kernel a quarter of its rate (4.6 against 5.8-6.1 TB/s).
template <int D, int TEAMS>
__global__ void __launch_bounds__(in_threads(TEAMS), 1)
k_apply_inner_tma(cplx* __restrict__ state, const cplx* __restrict__ coef, const Geometry g, const TaskTable tt,
                  const InnerPlan p, unsigned long long n_tiles /* over all batch entries */,
                  const __grid_constant__ CUtensorMap tmap) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];  // swizzled stages need 1024-byte alignment
  __shared__ __align__(8) unsigned long long full_bar[IN_MAX_STAGES], done_bar[IN_MAX_STAGES], free_bar[IN_MAX_STAGES];
  cplx* tiles = reinterpret_cast<cplx*>(smem_raw);
  cplx* M = tiles + (size_t)p.stages * p.tile_elems;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&done_bar[s]), IN_TEAM);
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

  if (warp >= IN_TEAM * TEAMS) {
    // ---------------- DMA warps: a loader and a storer, their 32 lanes share a tile's copies ---------
    int batch = 0;
    auto where = [&](unsigned long long i, cplx*& base, unsigned long long& t) {
      const unsigned long long gt = t_lo + i;
      const unsigned long long b = gt / p.tiles_per_state;
      t = gt - b * p.tiles_per_state;
      base = state + (size_t)b * g.state_batch_stride;
      batch = (int)b;
    };
    cplx* base;
    unsigned long long t;
    if (warp == IN_TEAM * TEAMS) {
      for (unsigned long long i = 0; i < n_my; ++i) {
        const unsigned st = (unsigned)(i % S);
        if (i >= S) mbar_wait(smem_u32(&free_bar[st]), (unsigned)(((i / S) - 1) & 1));  // tile i-S has left the stage
        where(i, base, t);
        tile_copy<true>(p, g, base, t, smem_u32(tiles + (size_t)st * p.tile_elems), smem_u32(&full_bar[st]), lane, &tmap,
                        batch);
      }
    } else {
      for (unsigned long long i = 0; i < n_my; ++i) {
        const unsigned st = (unsigned)(i % S);
        mbar_wait(smem_u32(&done_bar[st]), (unsigned)((i / S) & 1));  // consumers are done with tile i
        fence_async_smem();
        where(i, base, t);
        tile_copy<false>(p, g, base, t, smem_u32(tiles + (size_t)st * p.tile_elems), 0, lane, &tmap, batch);
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

  // ---------------- consumer warps: team `team` takes the tiles i = team, team + TEAMS, ... ----------
  const int team = warp / IN_TEAM, wteam = warp - team * IN_TEAM;
  cplx* Mt = M + (size_t)team * g.coef_count;   // every team keeps its own copy of the gate table
  const long long step = p.sk + tt.dl * p.sl;
  const int n_wt = p.groups * tt.ntasks;
  // conflict-free lane -> slice map inside a group of 32 slices (identity when no padding is needed)
  // padded granules (gran = 4r): lane l of a group takes slice 4r * (x mod w) + x / w + r * (l / 8), x = l mod 8,
  // w = 8 / r; swizzled stages: the searched map; else the identity
  int lane_slice = p.swz ? (int)p.lane_map[lane] : lane;
  if (p.gran > 0) {
    const int x = lane & 7, lw = 8 / p.r;
    lane_slice = 4 * p.r * (x % lw) + x / lw + p.r * (lane >> 3);
  }
  long long cur_batch = -1;
  for (unsigned long long i = team; i < n_my; i += TEAMS) {
    const unsigned long long gt = t_lo + i;
    const long long b = (long long)(gt / p.tiles_per_state);
    const unsigned long long t = gt - (unsigned long long)b * p.tiles_per_state;
    if (b != cur_batch && (cur_batch < 0 || g.coef_batch_stride != 0)) {
      // (re)load the team's gate table: the whole team has finished the previous batch entry's tiles
      asm volatile("bar.sync %0, %1;\n" ::"r"(1 + team), "n"(IN_TEAM * 32) : "memory");
      const cplx* cg = coef + (size_t)b * g.coef_batch_stride;
      for (int e = wteam * 32 + lane; e < g.coef_count; e += IN_TEAM * 32) {
        cplx v = cg[e];
        if (g.conj) v.y = -v.y;
        Mt[e] = v;
      }
      asm volatile("bar.sync %0, %1;\n" ::"r"(1 + team), "n"(IN_TEAM * 32) : "memory");
    }
    cur_batch = b;
    const unsigned st = (unsigned)(i % S);
    mbar_wait(smem_u32(&full_bar[st]), (unsigned)((i / S) & 1));
    cplx* tile = tiles + (size_t)st * p.tile_elems;
    const int nsl = tile_slices(p, g, t);
    for (int wt = wteam; wt < n_wt; wt += IN_TEAM) {
      const int gi = wt / tt.ntasks, task = wt - gi * tt.ntasks;
      const int sg = gi * 32 + lane_slice;
      if (sg < nsl) {
        const int org = (sg / p.block_slices) * p.block_elems + (sg % p.block_slices) * p.row_ss +
                        (p.gran > 0 ? sg / p.gran : 0);
        const SubBlock sb0 = tt.sub[task][0], sb1 = tt.sub[task][1];
        if (p.swz) {
          task_dispatch_swz<D>(sb0.c, tile, org + sb0.start_k * p.sk + sb0.start_l * p.sl,
                               org + sb1.start_k * p.sk + sb1.start_l * p.sl, (int)step, Mt + sb0.coef, Mt + sb1.coef);
        } else {
          cplx* ps = tile + org;
          task_dispatch<D>(sb0.c, ps + sb0.start_k * p.sk + sb0.start_l * p.sl,
                           ps + sb1.start_k * p.sk + sb1.start_l * p.sl, step, Mt + sb0.coef, Mt + sb1.coef);
        }
      }
    }
    fence_async_smem();  // generic-proxy writes to the stage become visible to the bulk store
    __syncwarp();
    if (lane == 0) mbar_arrive(smem_u32(&done_bar[st]));
  }
}

template <int D, int TEAMS>
static cudaError_t launch_dt(cplx* state, const cplx* coef, const Geometry& g, const TaskTable& tt,
                             const InnerPlan& p, unsigned long long n_tiles, size_t smem, cudaStream_t st,
                             const CUtensorMap& tmap) {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
  }
  cudaError_t e =
      cudaFuncSetAttribute(k_apply_inner_tma<D, TEAMS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  const unsigned grid = (unsigned)(n_tiles < (unsigned long long)sms ? n_tiles : (unsigned long long)sms);
  k_apply_inner_tma<D, TEAMS><<<grid, in_threads(TEAMS), smem, st>>>(state, coef, g, tt, p, n_tiles, tmap);
  return cudaSuccess;
}

template <int D>
static cudaError_t launch_d(cplx* state, const cplx* coef, const Geometry& g, const TaskTable& tt,
                            const InnerPlan& p, unsigned long long n_tiles, size_t smem, cudaStream_t st,
                            const CUtensorMap& tmap) {
  if constexpr (in_teams(D) == 2) {
    if (p.teams == 2) return launch_dt<D, 2>(state, coef, g, tt, p, n_tiles, smem, st, tmap);
  }
  return launch_dt<D, 1>(state, coef, g, tt, p, n_tiles, smem, st, tmap);
}

// cuTensorMapEncodeTiled through the runtime (no link against libcuda)
typedef CUresult (*encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static encode_tiled_fn tensor_map_encoder() {
  static encode_tiled_fn fn = [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      ptr = nullptr;
    return (encode_tiled_fn)ptr;
  }();
  return fn;
}

// rows geometry as a tensor of doubles [batch][outer * D rows][row run / 128 B][16]: the D rows of a tile are ONE box
static bool rows_tensor_map(CUtensorMap* tm, cplx* state, int D, const Geometry& g, long long hi, int nbatch,
                            int box_units, bool swizzle) {
  encode_tiled_fn enc = tensor_map_encoder();
  if (!enc || (hi * 2) % 16 != 0 || box_units > 256 || D > 256) return false;
  const unsigned long long rows_total = (unsigned long long)(g.n_slices / g.mid) * (unsigned long long)D;
  cuuint64_t dims[4] = {16, (cuuint64_t)(hi * 2 / 16), (cuuint64_t)rows_total, (cuuint64_t)nbatch};
  cuuint64_t strides[3] = {128, (cuuint64_t)hi * 16,
                           (cuuint64_t)(nbatch > 1 ? g.state_batch_stride : (long long)rows_total * hi) * 16};
  cuuint32_t box[4] = {16, (cuuint32_t)box_units, (cuuint32_t)D, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  if (dims[1] >= (1ull << 32) || dims[2] >= (1ull << 32) || strides[1] >= (1ull << 40) || strides[2] >= (1ull << 40))
    return false;
  return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, state, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             swizzle ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// a state (per batch entry) as a tensor of doubles [batch][1][state / 128 B][16]: contiguous tiles are boxes of
// box_units units
static bool flat_tensor_map(CUtensorMap* tm, cplx* state, long long state_elems, long long batch_stride, int nbatch,
                            int box_units, bool swizzle) {
  encode_tiled_fn enc = tensor_map_encoder();
  if (!enc || state_elems % 8 != 0 || box_units > 256 || box_units < 1) return false;
  cuuint64_t dims[4] = {16, (cuuint64_t)(state_elems / 8), 1, (cuuint64_t)nbatch};
  cuuint64_t strides[3] = {128, (cuuint64_t)state_elems * 16,
                           (cuuint64_t)(nbatch > 1 ? batch_stride : state_elems) * 16};
  cuuint32_t box[4] = {16, (cuuint32_t)box_units, 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  if (dims[1] >= (1ull << 32) || strides[1] >= (1ull << 40) || strides[2] >= (1ull << 40)) return false;
  return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, state, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             swizzle ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// Lane -> slice map for a swizzled stage: among the maps slice = (a * x + b * q) mod 32 (x = lane mod 8,
// q = lane / 8) that are bijections, the one with the fewest bank-group collisions inside a quarter warp,
// over every element offset a task touches.  Pure host arithmetic, cached per geometry.
static int swizzled_lane_map(const InnerPlan& p, int D, bool pair, unsigned char* out) {
  struct Key {
    int D, bs, be, ss, sk, sl, pair;
    bool operator<(const Key& o) const {
      return memcmp(this, &o, sizeof(Key)) < 0;
    }
  };
  struct Val {
    unsigned char map[32];
    int worst;
  };
  static std::map<Key, Val> cache;
  Key key{D, p.block_slices, p.block_elems, p.row_ss, p.sk, p.sl, pair ? 1 : 0};
  auto it = cache.find(key);
  if (it == cache.end()) {
    std::vector<int> offs;
    if (pair) {
      for (int k = 0; k < D; ++k)
        for (int l = 0; l < D; ++l) offs.push_back(k * (p.sk > p.sl ? p.sk : p.sl) + l);
    } else {
      for (int l = 0; l < D; ++l) offs.push_back(l);
    }
    auto origin = [&](int sg) { return (sg / p.block_slices) * p.block_elems + (sg % p.block_slices) * p.row_ss; };
    Val best;
    best.worst = 1 << 30;
    for (int a = 1; a < 32; ++a)
      for (int b = 1; b < 32; ++b) {
        unsigned char m[32];
        unsigned seen = 0;
        for (int lane = 0; lane < 32; ++lane) {
          m[lane] = (unsigned char)((a * (lane & 7) + b * (lane >> 3)) & 31);
          seen |= 1u << m[lane];
        }
        if (seen != 0xffffffffu) continue;
        int worst = 0;
        for (int gi = 0; gi < p.groups && gi < 4; ++gi)
          for (int q = 0; q < 4; ++q)
            for (int off : offs) {
              int cnt[8] = {0, 0, 0, 0, 0, 0, 0, 0};
              for (int x = 0; x < 8; ++x) {
                const int c = origin(gi * 32 + m[q * 8 + x]) + off;
                const int w = ++cnt[(c ^ (c >> 3)) & 7];
                if (w > worst) worst = w;
              }
            }
        if (worst < best.worst) {
          best.worst = worst;
          memcpy(best.map, m, 32);
        }
      }
    it = cache.emplace(key, best).first;
  }
  memcpy(out, it->second.map, 32);
  return it->second.worst;
}

static int gcd_int(int a, int b) { return b == 0 ? a : gcd_int(b, a % b); }

bool launch_inner_tma(int D, cplx* state, const cplx* coef, const Geometry& g, const TaskTable& tt, int nbatch,
                      cudaStream_t st, int* status) {
  if (D < 2 || D > B200_MAX_FAST_CUTOFF || g.inner != 1) return false;
  const bool pair = g.stride2 != 0;
  InnerPlan p;
  memset(&p, 0, sizeof(p));
  CUtensorMap tmap;
  memset(&tmap, 0, sizeof(tmap));
  // Tensor-map copies (tools/inner_probe.py, D = 10, TB/s; bulk row / granule copies in brackets):
  //   rows geometry, plain box            5.25   [4.4-4.8]   with SWIZZLE_128B + searched lane map 5.21-5.28
  //   one-mode gate, swizzled boxes       5.24   [5.0-5.3 unpadded, 2-way conflicts]
  //   adjacent pairs, swizzled boxes      4.93   [5.64 padded granules: the swizzle's address arithmetic costs more
  //                                               than the two extra bulk copies]
  // Default: plain box for rows, swizzled boxes for one-mode gates, padded granules for adjacent pairs.
  // B200_INNER_TMAP = 0: bulk copies only; "swizzle": swizzled tensor maps wherever they apply.
  static const char* tm_env = getenv("B200_INNER_TMAP");
  const bool want_tmap = !(tm_env && tm_env[0] == '0');
  const bool swizzle_all = want_tmap && tm_env && tm_env[0] == 's';
  const bool want_swizzle = want_tmap && (swizzle_all || !pair);
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
    // Measured (B200, D = 10): the one-mode tile is ONE contiguous 51 KB run, and every extra bulk copy costs
    // more than the bank conflicts it removes -- unpadded (4 copies of <= 16 KB, 2-way conflicts) 5.0-5.3 TB/s,
    // granules of 32 slices (10 copies) 4.4, of 16 (20 copies) 3.7.  The tile stays unpadded; the swizzled
    // tensor-map boxes below make it conflict free without more copies.
    p.gran = 0;
    p.r = 8;
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
      // a row run of `mid` slices is cut into equal chunks of at most 32 * q slices (mid = 100: 4 x 25, not
      // 3 x 32 + 4)
      p.chunks_per_outer = (int)((mid + 32u * q - 1) / (32u * q));
      p.slices_per_tile = (int)((mid + p.chunks_per_outer - 1) / p.chunks_per_outer);
      p.rows = D;
      p.row_ss = D;
      // Measured (tools/inner_probe.py, D = 10): padding the rows (two granules of 16 slices per row, 20 + 20
      // bulk copies per tile) gives 3.5 TB/s, unpadded rows (10 + 10 copies, 2-way bank conflicts) 4.8 TB/s:
      // rows stay unpadded.
      static const char* rows_mode = getenv("B200_INNER_ROWS");
      p.gran = 0;
      p.r = 8;
      // One tensor-map copy per tile and direction instead of D row copies (each bulk operation costs ~40 ns
      // of tile time): chunks of exactly 32 * q slices, the tail of a row run is clipped by the
      // tensor bounds (zero fill on load, skipped on store).  B200_INNER_ROWS=bulk keeps the row copies.
      // With SWIZZLE_128B (rows of a multiple of 1024 bytes) the stage is conflict free as well.
      const bool rows_swz = swizzle_all && (32 * q * D * 16) % 1024 == 0;
      if (!(rows_mode && rows_mode[0] == 'b') && want_tmap &&
          rows_tensor_map(&tmap, state, D, g, hi, nbatch, 32 * q * D * 2 / 16, rows_swz)) {
        p.use_tmap = 1;
        p.swz = rows_swz ? 1 : 0;
        p.slices_per_tile = 32 * q;
        p.chunks_per_outer = (int)((mid + p.slices_per_tile - 1) / p.slices_per_tile);
        p.box_units = p.slices_per_tile * D * 2 / 16;
      }
      p.row_pitch = p.slices_per_tile * D + (p.gran ? p.slices_per_tile / p.gran : 0);
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
  // Contiguous tiles (one-mode gates, adjacent pairs, short-mid blocks that fit one box) through a tensor map with
  // SWIZZLE_128B: unpadded, one or two bulk operations per tile, and conflict free with the searched lane map.
  if (want_swizzle && !p.rows_mode && D % 2 == 0) {
    const int flat_elems = pair && p.block_elems ? (p.slices_per_tile / p.block_slices) * p.block_elems
                                                 : p.slices_per_tile * p.slice_elems;
    const long long state_elems = (long long)g.n_slices * p.slice_elems;
    if (flat_elems % 8 == 0) {
      const int units = flat_elems / 8;
      int boxes = (units + 255) / 256;
      while (boxes <= 8 && (units % boxes != 0 || (boxes > 1 && (units / boxes) % 8 != 0))) ++boxes;
      if (boxes <= 8 && flat_tensor_map(&tmap, state, state_elems, g.state_batch_stride, nbatch, units / boxes, true)) {
        p.use_tmap = 2;
        p.swz = 1;
        p.boxes = boxes;
        p.box_units = units / boxes;
        p.gran = 0;
        p.r = 8;
        p.tile_elems = flat_elems;
      }
    }
  }
  if (p.swz) {
    swizzled_lane_map(p, D, pair, p.lane_map);
  } else {
    for (int l = 0; l < 32; ++l) p.lane_map[l] = (unsigned char)l;
  }
  // Two consumer teams for the dense one-mode task (400 DFMA per warp-task: one team was compute-latency
  // bound, 3.7-4.9 TB/s); one team for pair gates, where a second team would take the ring stage that keeps
  // a second tile of loads in flight (measured 5.2-5.5 against 5.8-6.1 TB/s).  B200_INNER_TEAMS overrides.
  static const char* teams_env = getenv("B200_INNER_TEAMS");
  p.teams = pair ? 1 : 2;
  if (teams_env && (teams_env[0] == '1' || teams_env[0] == '2')) p.teams = teams_env[0] - '0';
  if (p.teams > in_teams(D)) p.teams = in_teams(D);
  const size_t tile_align = p.swz ? 1024 : 128;
  const size_t tile_bytes = (((size_t)p.tile_elems * 16) + tile_align - 1) / tile_align * tile_align;
  p.tile_elems = (int)(tile_bytes / 16);
  const size_t coef_bytes = (size_t)g.coef_count * 16 * p.teams;  // one copy of the table per team
  if (coef_bytes + 3 * tile_bytes > (size_t)IN_SMEM_LIMIT) return false;  // fewer than three stages: not worth it
  p.stages = (int)((IN_SMEM_LIMIT - coef_bytes) / tile_bytes);
  if (p.stages > IN_MAX_STAGES) p.stages = IN_MAX_STAGES;
  const size_t smem = (size_t)p.stages * tile_bytes + coef_bytes;
  const unsigned long long n_tiles = p.tiles_per_state * (unsigned long long)nbatch;
  cudaError_t e = cudaSuccess;
#define B200_LAUNCH(N) \
  case N:              \
    e = launch_d<N>(state, coef, g, tt, p, n_tiles, smem, st, tmap); \
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
