// FP64 peaks of the GPU this runs on: dependent-chain-free DFMA throughput (the pipe the gate kernels use)
// and DMMA (mma.sync.m8n8k4.f64, the FP64 tensor-core path).  BASELINE.md section 2 leaves both "unmeasured";
// DESIGN.md's FP64 budget argument uses the numbers printed here.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_probe fp64_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(256) k_dfma(double* out, int iters) {
  double a[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 1e-9 + i;
  const double x = 1.0000001, y = 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = fma(a[i], x, y);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) k_dmma(double* out, int iters) {
  double c[4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i) c[i][0] = c[i][1] = 0.0;
  double a = threadIdx.x * 1e-9, b = 1.0000001;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                   : "+d"(c[i][0]), "+d"(c[i][1])
                   : "d"(a), "d"(b));
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = c[0][0] + c[1][1] + c[2][0] + c[3][1];
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  double* out;
  cudaMalloc(&out, sizeof(double) * sms * 8 * 256);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int iters = 200000;
  float ms_fma = 0, ms_mma = 0;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0);
    k_dfma<<<sms * 8, 256>>>(out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms_fma, e0, e1);
  }
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0);
    k_dmma<<<sms * 8, 256>>>(out, iters / 4);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms_mma, e0, e1);
  }
  const double fma_flops = 2.0 * 16 * iters * (double)sms * 8 * 256;
  const double mma_flops = 2.0 * 8 * 8 * 4 * 4 * (iters / 4) * (double)sms * 8 * (256 / 32);
  printf("{\"sms\": %d, \"dfma_tflops\": %.2f, \"dfma_ms\": %.2f, \"dmma_m8n8k4_tflops\": %.2f, \"dmma_ms\": %.2f, \"err\": \"%s\"}\n",
         sms, fma_flops / ms_fma / 1e9, ms_fma, mma_flops / ms_mma / 1e9, ms_mma, cudaGetErrorString(cudaGetLastError()));
  return 0;
}
