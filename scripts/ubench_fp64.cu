// FP64 pipe micro-benchmark for sm_100a (B200): what a SINGLE warp sees -- the regime of the latency kernels.
//   dependent DFMA chain (latency), ILP-k independent chains (issue cadence), rsqrt() / sqrt+div / sincos / log latency,
//   64-bit shuffle round trip, shared-memory store->load round trip.  Cycles per operation from clock64().
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o gpurun_out/ubench_fp64 scripts/ubench_fp64.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void k_dfma(double* out, long long* cyc, int iters, double x, double y) {
  double a[ILP];
  for (int i = 0; i < ILP; ++i) a[i] = threadIdx.x * 1e-9 + i;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) a[i] = fma(a[i], x, y);
  }
  long long t1 = clock64();
  double s = 0;
  for (int i = 0; i < ILP; ++i) s += a[i];
  out[threadIdx.x + blockIdx.x * blockDim.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP>
__global__ void k_op(double* out, long long* cyc, int iters, double x) {
  double a = 1.5 + threadIdx.x * 1e-3;
  __shared__ double sh[64];
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (OP == 0) a = rsqrt(a) + x;
    if (OP == 1) a = 1.0 / sqrt(a) + x;
    if (OP == 2) { double s, c; sincos(a, &s, &c); a = s + c + x; }
    if (OP == 3) a = log(a) + x;
    if (OP == 4) a = __shfl_xor_sync(0xffffffffu, a, 1) + x;
    if (OP == 5) { sh[threadIdx.x] = a; __syncwarp(); a = sh[threadIdx.x ^ 1] + x; __syncwarp(); }
    if (OP == 6) a = a / (x + 2.0) + x;
  }
  long long t1 = clock64();
  out[threadIdx.x] = a;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 4096 * 8);
  long long h[4096];
  const int iters = 20000;
#define RUN_DFMA(ILP, WARPS)                                                                         \
  { k_dfma<ILP><<<1, 32 * WARPS>>>(out, cyc, iters, 0.999999, 1e-7); cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost); \
    printf("{\"bench\": \"dfma\", \"ilp\": %d, \"warps_per_sm\": %d, \"cycles_per_dfma_per_warp\": %.3f}\n", ILP, WARPS, (double)h[0] / iters / ILP); }
  RUN_DFMA(1, 1) RUN_DFMA(2, 1) RUN_DFMA(4, 1) RUN_DFMA(8, 1) RUN_DFMA(16, 1)
  RUN_DFMA(1, 4) RUN_DFMA(4, 4) RUN_DFMA(8, 4) RUN_DFMA(1, 8) RUN_DFMA(4, 8) RUN_DFMA(1, 16) RUN_DFMA(8, 16)
  const char* names[] = {"rsqrt+add", "1/sqrt+add", "sincos+2add", "log+add", "shfl64+add", "smem st-sync-ld+add", "div+add"};
#define RUN_OP(OP) { k_op<OP><<<1, 32>>>(out, cyc, 2000, 0.25); cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost); \
    printf("{\"bench\": \"%s\", \"cycles_per_iteration\": %.1f}\n", names[OP], (double)h[0] / 2000); }
  RUN_OP(0) RUN_OP(1) RUN_OP(2) RUN_OP(3) RUN_OP(4) RUN_OP(5) RUN_OP(6)
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
  return 0;
}
