// Microbenchmark: throughput of cp.async.bulk (global -> shared -> global) as a function of the
// operation size and of how many threads issue operations.  Answers "what does one bulk operation
// cost?" for the exchange and the innermost-axis kernels.   nvcc -arch=sm_100a -O3 -o tma_probe tma_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "../../strawberryfields_b200/csrc/tma.cuh"
using namespace b200;

constexpr int STAGE = 16 * 1024;
constexpr int STAGES = 3;

// every issuing thread owns a ring of STAGES stages of STAGE bytes; a stage is filled with STAGE/op
// operations of `op` bytes, then written back with as many stores
__global__ void __launch_bounds__(1024, 1) k_probe(const char* src, char* dst, size_t bytes, int op, int lanes, int warps) {
  extern __shared__ __align__(128) unsigned char sm[];
  __shared__ __align__(8) unsigned long long bars[256 * STAGES];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool active = warp < warps && lane < lanes;
  const int slot = warp * lanes + lane;           // issuing thread index inside the CTA
  const int nslots = warps * lanes;
  if (active) for (int s = 0; s < STAGES; ++s) mbar_init(smem_u32(&bars[slot * STAGES + s]), 1);
  mbar_fence_init();
  __syncthreads();
  if (!active) return;
  const int stage_bytes = STAGE / (nslots > 4 ? nslots / 4 : 1);   // keep total smem at 4 * STAGES * STAGE
  const int ops = stage_bytes / op > 0 ? stage_bytes / op : 1;
  const size_t unit = (size_t)ops * op;
  const unsigned st0 = smem_u32(sm) + (unsigned)slot * STAGES * stage_bytes;
  const unsigned b0 = smem_u32(&bars[slot * STAGES]);
  const size_t nthreads = (size_t)gridDim.x * nslots, me = (size_t)blockIdx.x * nslots + slot;
  const size_t n_units = bytes / unit;
  const size_t mine = me < n_units ? (n_units - me + nthreads - 1) / nthreads : 0;
  auto load = [&](size_t i) {
    const int st = i % STAGES;
    const char* p = src + (me + i * nthreads) * unit;
    mbar_expect_tx(b0 + 8 * st, (unsigned)unit);
    for (int k = 0; k < ops; ++k) bulk_load(st0 + st * stage_bytes + k * op, p + (size_t)k * op, op, b0 + 8 * st);
  };
  for (size_t i = 0; i < STAGES - 1 && i < mine; ++i) load(i);
  for (size_t i = 0; i < mine; ++i) {
    const int st = i % STAGES;
    mbar_wait(b0 + 8 * st, (i / STAGES) & 1);
    fence_async_smem();
    char* q = dst + (me + i * nthreads) * unit;
    for (int k = 0; k < ops; ++k) bulk_store(q + (size_t)k * op, st0 + st * stage_bytes + k * op, op);
    bulk_commit();
    if (i + STAGES - 1 < mine) {
      bulk_wait_read<1>();
      load(i + STAGES - 1);
    }
  }
  bulk_wait_all();
}

int main() {
  const size_t bytes = 2ull << 30;
  char *a, *b;
  cudaMalloc(&a, bytes);
  cudaMalloc(&b, bytes);
  cudaMemset(a, 1, bytes);
  cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * STAGES * STAGE);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  printf("{\"what\": \"cp.async.bulk copy rate (GB/s of payload, read+write = 2x) on 148 CTAs\", \"rows\": [\n");
  const int sizes[] = {160, 640, 1600, 3200, 5120, 16384};
  const int cfg[][2] = {{4, 1}, {4, 4}, {4, 8}, {4, 32}, {8, 32}, {1, 1}, {1, 32}};  // {warps, lanes}
  bool first = true;
  for (int op : sizes)
    for (auto& c : cfg) {
      if (c[0] * c[1] > 4 && STAGE / (c[0] * c[1] / 4) < op) continue;
      for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        k_probe<<<148, 1024, 4 * STAGES * STAGE>>>(a, b, bytes, op, c[1], c[0]);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
      }
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      cudaError_t err = cudaGetLastError();
      printf("%s{\"op_bytes\": %d, \"warps\": %d, \"lanes\": %d, \"GBps\": %.0f, \"err\": \"%s\"}", first ? "" : ",\n", op, c[0],
             c[1], bytes / ms / 1e6, err == cudaSuccess ? "" : cudaGetErrorString(err));
      first = false;
    }
  printf("\n]}\n");
  return 0;
}
