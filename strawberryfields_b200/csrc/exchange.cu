// Multi-GPU axis exchange: ONE launch moves this rank's whole share of the
// all-to-all over NVLink, and a device-side flag barrier replaces the host-side
// synchronize + barrier pairs.
//
// The exchange swaps the g sharded leading axes of the state with g local ones.  For every
// (source rank, destination rank) pair that is a strided block copy
//     dst[dst_base + sum_j i_j * ds_j + r] = src[src_base + sum_j i_j * ss_j + r],  0 <= r < run
// whose innermost `run` elements are contiguous on both sides.  One of the two sides lives in a
// peer's HBM (mapped with CUDA IPC): "pull" reads the peers' old shards, "push" writes the
// peers' new shards.
//
// Data path: no thread ever touches the payload.  A warp's elected lane drives a ring of
// shared-memory stages with the bulk-copy engine (TMA): cp.async.bulk global -> shared
// (completion on an mbarrier), then cp.async.bulk shared -> global, so a CTA keeps
// 4 warps x 2 stages x 16 KB of NVLink reads in flight with a handful of instructions --
// remote loads have ~2 us latency (B300_MICROARCH: 1834-2036 cycles) and the thread-level
// gather kernel this replaces ran out of registers long before it ran out of link.
//
// The reference has no distributed path; this is the scale-out row of the scope table.
#include "tma.cuh"

namespace b200 {

constexpr int XC_WARPS = 4;               // bulk-copy issuers per CTA (one elected lane each)
constexpr int XC_THREADS = 512;           // 16 warps: all of them copy in the CTAs that move the local block
constexpr int XC_STAGES = 3;
constexpr int XC_STAGE_BYTES = 16 * 1024;
constexpr int XC_SMEM = XC_WARPS * XC_STAGES * XC_STAGE_BYTES;
constexpr int XC_BULK_CTAS = 32;          // CTAs that drive NVLink next to a local block (probe: 32 CTAs reach
                                          // 667-689 GB/s, 64: 660-712, all 148: 676-712); the rest stay free for the
                                          // gate kernels that overlap the exchange
constexpr int XC_SEG = 256;               // elements of a run one warp copies at a time (8 per lane in flight)

// offsets of run q (0 <= q < n_outer) of the block copy: mixed-radix decode over the outer axes
__device__ __forceinline__ void run_offsets(const b200_xchg_desc& d, unsigned long long o, long long& so,
                                            long long& dof) {
  so = 0;
  dof = 0;
  for (int j = d.n_axes - 1; j >= 0; --j) {
    const unsigned long long e = (unsigned)d.ext[j];
    const unsigned long long q = o / e;
    const long long dig = (long long)(o - q * e);
    o = q;
    so += dig * d.ss[j];
    dof += dig * d.ds[j];
  }
}

// A *unit* is what one ring stage holds: `m` consecutive runs (short runs are batched so that a
// stage is always ~16 KB) or one piece of a long run.  unit -> (source, first run, piece).
struct UnitPlan {
  unsigned long long runs_per_src;   // n_outer * pieces_per_run
  unsigned long long units_per_src;
  unsigned pieces_per_run;           // > 1: runs longer than a stage
  unsigned runs_per_unit;            // > 1: runs shorter than a stage (then pieces_per_run == 1)
  unsigned piece_elems;              // elements per (full) piece
};

// The block whose source AND destination are local (source == this rank) never touches NVLink; the CTAs
// behind the first `bulk_ctas` copy it with plain 16-byte loads and stores, a warp per run segment, eight
// elements per lane in flight.  (The bulk-copy engine needs ~240 cycles per operation and SM -- fine for
// the link, 5x too slow for an HBM-rate copy of 1.6 KB runs.)
__device__ __forceinline__ void copy_local_block(const b200_xchg_desc& d, int src, unsigned cta, unsigned n_ctas) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = XC_THREADS / 32;
  const cplx* sp = reinterpret_cast<const cplx*>(d.src[src]) + d.src_base[src];
  cplx* dp = reinterpret_cast<cplx*>(d.dst[src]) + d.dst_base[src];
  unsigned long long n_outer = 1;
  for (int j = 0; j < d.n_axes; ++j) n_outer *= (unsigned)d.ext[j];
  const unsigned long long segs = ((unsigned long long)d.run + XC_SEG - 1) / XC_SEG;
  const unsigned long long total = n_outer * segs;
  for (unsigned long long w = (unsigned long long)cta * nwarps + warp; w < total;
       w += (unsigned long long)n_ctas * nwarps) {
    const unsigned long long o = w / segs;
    const long long r0 = (long long)(w - o * segs) * XC_SEG;
    long long so, dof;
    run_offsets(d, o, so, dof);
    const cplx* a = sp + so + r0;
    cplx* b = dp + dof + r0;
    const int left = (int)(d.run - r0 < XC_SEG ? d.run - r0 : XC_SEG);
    cplx v[XC_SEG / 32];
#pragma unroll
    for (int u = 0; u < XC_SEG / 32; ++u)
      if (lane + 32 * u < left) v[u] = a[lane + 32 * u];
#pragma unroll
    for (int u = 0; u < XC_SEG / 32; ++u)
      if (lane + 32 * u < left) b[lane + 32 * u] = v[u];
  }
}

// d.local_src >= 0: that source is copied by the CTAs [bulk_ctas, gridDim.x) with threads; the bulk path
// (CTAs [0, bulk_ctas)) then walks the other n_src - 1 sources only.
__global__ void __launch_bounds__(XC_THREADS, 1)
k_exchange_bulk(const b200_xchg_desc d, const UnitPlan up, unsigned long long n_units, int local_src,
                unsigned bulk_ctas) {
  extern __shared__ __align__(128) unsigned char xsmem[];
  __shared__ __align__(8) unsigned long long bars[XC_WARPS * XC_STAGES];
  if (blockIdx.x >= bulk_ctas) {
    copy_local_block(d, local_src, blockIdx.x - bulk_ctas, gridDim.x - bulk_ctas);
    return;
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0 && warp < XC_WARPS)
    for (int s = 0; s < XC_STAGES; ++s) mbar_init(smem_u32(&bars[warp * XC_STAGES + s]), 1);
  mbar_fence_init();
  __syncthreads();
  if (lane != 0 || warp >= XC_WARPS) return;

  const unsigned stage0 = smem_u32(xsmem) + (unsigned)warp * XC_STAGES * XC_STAGE_BYTES;
  const unsigned bar0 = smem_u32(&bars[warp * XC_STAGES]);
  const unsigned long long wid = (unsigned long long)blockIdx.x * XC_WARPS + warp;
  const unsigned long long nw = (unsigned long long)bulk_ctas * XC_WARPS;
  const unsigned long long my_units = wid < n_units ? (n_units - wid + nw - 1) / nw : 0;
  const unsigned n_bulk_src = (unsigned)d.n_src - (local_src >= 0 ? 1u : 0u);

  // geometry of unit number i of this warp
  auto unit_geometry = [&](unsigned long long i, int& src, unsigned long long& q0, unsigned& nruns) {
    const unsigned long long u = wid + i * nw;
    // sources interleaved, starting behind this rank so that every peer link is busy at once
    src = (int)((u % n_bulk_src + (unsigned)d.first_src + (local_src >= 0 ? 1u : 0u)) % (unsigned)d.n_src);
    const unsigned long long v = u / n_bulk_src;
    q0 = v * up.runs_per_unit;
    const unsigned long long left = up.runs_per_src - q0;
    nruns = (unsigned)(left < up.runs_per_unit ? left : up.runs_per_unit);
  };
  // Offsets of the runs of a unit.  A unit is either ONE piece of a long run (one decode) or up to
  // runs_per_unit consecutive whole runs: the first is decoded (a few 32-bit divides), the rest follow by
  // odometer increments -- the issuing thread must not spend more time on addresses than the copy
  // engine needs per operation (~30 cycles per SM, tools/probes/tma_probe.cu).
  struct Walk {
    int dig[B200_XCHG_MAX_AXES];
    long long so, dof;
  };
  auto walk_start = [&](unsigned long long o, Walk& w) {
    w.so = 0;
    w.dof = 0;
    unsigned o32 = (unsigned)o;  // n_outer < 2^32 (checked by the host)
    for (int j = d.n_axes - 1; j >= 0; --j) {
      const unsigned e = (unsigned)d.ext[j];
      const unsigned q = o32 / e;
      const int dg = (int)(o32 - q * e);
      o32 = q;
      w.dig[j] = dg;
      w.so += (long long)dg * d.ss[j];
      w.dof += (long long)dg * d.ds[j];
    }
  };
  auto walk_next = [&](Walk& w) {
    for (int j = d.n_axes - 1; j >= 0; --j) {
      w.so += d.ss[j];
      w.dof += d.ds[j];
      if (++w.dig[j] < d.ext[j]) return;
      w.so -= (long long)d.ext[j] * d.ss[j];
      w.dof -= (long long)d.ext[j] * d.ds[j];
      w.dig[j] = 0;
    }
  };
  auto issue = [&](unsigned long long i, bool load) {
    const int st = (int)(i % XC_STAGES);
    int src;
    unsigned long long q0;
    unsigned nruns;
    unit_geometry(i, src, q0, nruns);
    const cplx* sp = reinterpret_cast<const cplx*>(d.src[src]) + d.src_base[src];
    cplx* dp = reinterpret_cast<cplx*>(d.dst[src]) + d.dst_base[src];
    const unsigned sbase = stage0 + (unsigned)st * XC_STAGE_BYTES, bar = bar0 + 8u * st;
    Walk w;
    if (up.pieces_per_run > 1) {
      // one piece of a long run
      const unsigned long long o = q0 / up.pieces_per_run;
      const long long off = (long long)(q0 - o * up.pieces_per_run) * up.piece_elems;
      const long long left = d.run - off;
      const unsigned bytes = (unsigned)((left < (long long)up.piece_elems ? left : (long long)up.piece_elems) * 16);
      walk_start(o, w);
      if (load) {
        mbar_expect_tx(bar, bytes);
        bulk_load(sbase, sp + w.so + off, bytes, bar);
      } else {
        bulk_store(dp + w.dof + off, sbase, bytes);
      }
      return;
    }
    const unsigned bytes = (unsigned)d.run * 16u;
    walk_start(q0, w);
    if (load) mbar_expect_tx(bar, bytes * nruns);
    for (unsigned r = 0; r < nruns; ++r) {
      if (load) bulk_load(sbase + r * bytes, sp + w.so, bytes, bar);
      else bulk_store(dp + w.dof, sbase + r * bytes, bytes);
      walk_next(w);
    }
  };
  auto issue_load = [&](unsigned long long i) { issue(i, true); };
  auto issue_store = [&](unsigned long long i) {
    issue(i, false);
    bulk_commit();
  };

  for (unsigned long long i = 0; i < XC_STAGES - 1 && i < my_units; ++i) issue_load(i);
  for (unsigned long long i = 0; i < my_units; ++i) {
    const int st = (int)(i % XC_STAGES);
    mbar_wait(bar0 + 8u * st, (unsigned)((i / XC_STAGES) & 1));
    fence_async_smem();
    issue_store(i);
    // the stage of unit i-1 is refilled with unit i+STAGES-1 once its store has read it
    if (i + XC_STAGES - 1 < my_units) {
      bulk_wait_read<1>();
      issue_load(i + XC_STAGES - 1);
    }
  }
  bulk_wait_all();  // the stores are complete (and visible) before the kernel ends
}

// ---- device-side barrier over peer-mapped flag words -------------------------------------------
// flags[r] on every rank is a counter written by rank r: thread t publishes `epoch` into peer t's
// flags[rank] and waits until its own flags[t] has reached `epoch`.  Everything issued on the stream
// before the barrier (the gates that made this shard final) is visible to the peers' kernels issued
// after it: the kernel boundary orders the earlier kernels, the release/acquire pair orders the flag.
__global__ void k_peer_barrier(b200_peer_flags pf, unsigned long long epoch, unsigned long long timeout_ns) {
  const int t = threadIdx.x;
  if (t >= pf.n_ranks) return;
  __threadfence_system();
  unsigned long long* remote = reinterpret_cast<unsigned long long*>(pf.flags[t]) + pf.rank;
  asm volatile("st.release.sys.global.u64 [%0], %1;\n" ::"l"(remote), "l"(epoch) : "memory");
  const unsigned long long* mine = reinterpret_cast<const unsigned long long*>(pf.flags[pf.rank]) + t;
  unsigned long long t0;
  asm volatile("mov.u64 %0, %%globaltimer;\n" : "=l"(t0));
  for (;;) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];\n" : "=l"(v) : "l"(mine) : "memory");
    if (v >= epoch) break;
    unsigned long long now;
    asm volatile("mov.u64 %0, %%globaltimer;\n" : "=l"(now));
    if (now - t0 > timeout_ns) {
      // a peer never arrived: fail loudly instead of hanging the GPU
      __trap();
    }
    __nanosleep(200);
  }
  __threadfence_system();
}

}  // namespace b200

using namespace b200;

extern "C" {

int b200_exchange_copy(const b200_xchg_desc* desc, int local_src, int bulk_ctas, void* stream) {
  B200_CHECK_ARG(desc, "exchange_copy: null descriptor");
  B200_CHECK_ARG(desc->n_axes >= 0 && desc->n_axes <= B200_XCHG_MAX_AXES, "exchange_copy: too many axes");
  B200_CHECK_ARG(desc->n_src >= 1 && desc->n_src <= B200_XCHG_MAX_PEERS, "exchange_copy: bad source count");
  B200_CHECK_ARG(desc->run >= 1 && desc->first_src >= 0 && desc->first_src < desc->n_src, "exchange_copy: bad run");
  B200_CHECK_ARG(local_src >= -1 && local_src < desc->n_src, "exchange_copy: bad local source");
  B200_CHECK_ARG(local_src < 0 || local_src == desc->first_src, "exchange_copy: the local source must be first_src");
  unsigned long long n_outer = 1;
  for (int j = 0; j < desc->n_axes; ++j) {
    B200_CHECK_ARG(desc->ext[j] >= 1, "exchange_copy: bad extent");
    n_outer *= (unsigned long long)desc->ext[j];
  }
  B200_CHECK_ARG(n_outer < (1ull << 32), "exchange_copy: too many runs for one launch");
  for (int s = 0; s < desc->n_src; ++s)
    B200_CHECK_ARG(desc->src[s] && desc->dst[s] && desc->src_base[s] >= 0 && desc->dst_base[s] >= 0,
                   "exchange_copy: null source / destination");
  UnitPlan up;
  const long long stage_elems = XC_STAGE_BYTES / 16;
  if (desc->run >= stage_elems) {
    up.pieces_per_run = (unsigned)((desc->run + stage_elems - 1) / stage_elems);
    up.runs_per_unit = 1;
    up.piece_elems = (unsigned)stage_elems;
  } else {
    up.pieces_per_run = 1;
    up.runs_per_unit = (unsigned)(stage_elems / desc->run);
    up.piece_elems = (unsigned)desc->run;
  }
  up.runs_per_src = n_outer * up.pieces_per_run;
  up.units_per_src = (up.runs_per_src + up.runs_per_unit - 1) / up.runs_per_unit;
  const int n_bulk_src = desc->n_src - (local_src >= 0 ? 1 : 0);
  const unsigned long long n_units = up.units_per_src * (unsigned long long)n_bulk_src;
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
  }
  // CTAs [0, nb): bulk-copy engine over the n_bulk_src (remote) sources; CTAs [nb, grid): the local block
  unsigned nb = 0, grid;
  if (n_bulk_src > 0) {
    unsigned long long want = (n_units + XC_WARPS - 1) / XC_WARPS;
    unsigned long long cap = bulk_ctas > 0 ? (unsigned long long)bulk_ctas
                                           : (unsigned long long)(local_src >= 0 ? XC_BULK_CTAS : sms);
    if (cap > (unsigned long long)sms) cap = (unsigned long long)sms;
    nb = (unsigned)(want < cap ? want : cap);
  }
  grid = nb;
  if (local_src >= 0) {
    unsigned long long segs = n_outer * (((unsigned long long)desc->run + XC_SEG - 1) / XC_SEG);
    unsigned long long want = (segs + XC_THREADS / 32 - 1) / (XC_THREADS / 32);
    unsigned long long room = (unsigned long long)(sms > (int)nb ? sms - (int)nb : 1);
    grid += (unsigned)(want < room ? want : room);
  }
  B200_CHECK_ARG(grid >= 1, "exchange_copy: nothing to copy");
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(k_exchange_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, XC_SMEM);
    if (e != cudaSuccess) return fail((int)e, "exchange_copy: shared memory opt-in failed: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  k_exchange_bulk<<<grid, XC_THREADS, XC_SMEM, (cudaStream_t)stream>>>(*desc, up, n_units, local_src, nb);
  return cuda_status("exchange_bulk");
}

int b200_peer_barrier(const b200_peer_flags* pf, uint64_t epoch, double timeout_s, void* stream) {
  B200_CHECK_ARG(pf && pf->n_ranks >= 1 && pf->n_ranks <= B200_XCHG_MAX_PEERS && pf->rank >= 0 &&
                     pf->rank < pf->n_ranks,
                 "peer_barrier: bad rank geometry");
  for (int r = 0; r < pf->n_ranks; ++r) B200_CHECK_ARG(pf->flags[r], "peer_barrier: null flag pointer");
  B200_CHECK_ARG(timeout_s > 0, "peer_barrier: timeout must be positive");
  k_peer_barrier<<<1, 32, 0, (cudaStream_t)stream>>>(*pf, (unsigned long long)epoch,
                                                     (unsigned long long)(timeout_s * 1e9));
  return cuda_status("peer_barrier");
}

}  // extern "C"
