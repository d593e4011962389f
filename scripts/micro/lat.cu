// Latency microbenchmarks on the target GPU: dependent DFMA, double shuffle, shared load, FP64 reciprocal, block barrier.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void lat(double* out, long long* t, double x0) {
  __shared__ double sm[1024];
  const int tid = threadIdx.x;
  sm[tid] = x0 + tid * 1e-3;
  sm[tid + 512] = 1.0;
  __syncthreads();
  double a = x0, b = 1.0000001, c = 1e-9;
  long long t0, t1;
  // 1: dependent DFMA chain
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 256; ++i) a = fma(a, b, c);
  t1 = clock64();
  if (tid == 0) t[0] = (t1 - t0);
  // 2: dependent double shuffle chain
  double s = a;
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 64; ++i) s = __shfl_sync(0xffffffffu, s, (i * 7 + 1) & 31);
  t1 = clock64();
  if (tid == 0) t[1] = (t1 - t0);
  // 3: dependent shared loads (pointer chase through indices)
  int idx = tid & 31;
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 64; ++i) idx = (int)sm[(idx + i) & 511] & 31;
  t1 = clock64();
  if (tid == 0) t[2] = (t1 - t0);
  // 4: dependent reciprocal chain
  double r = a + 3.0;
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 64; ++i) r = 1.0 / r + 1.5;
  t1 = clock64();
  if (tid == 0) t[3] = (t1 - t0);
  // 5: block barriers
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 64; ++i) __syncthreads();
  t1 = clock64();
  if (tid == 0) t[4] = (t1 - t0);
  // 6: shfl + fma alternating (triangular-solve pattern)
  double q = s;
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 64; ++i) { const double z = __shfl_sync(0xffffffffu, q, i & 31); q = fma(z, b, q); }
  t1 = clock64();
  if (tid == 0) t[5] = (t1 - t0);
  // 7: independent DFMA throughput per warp (8 chains)
  double e[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) e[k] = a + k;
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 64; ++i)
#pragma unroll
    for (int k = 0; k < 8; ++k) e[k] = fma(e[k], b, c);
  t1 = clock64();
  if (tid == 0) t[6] = (t1 - t0);
  // 8: rsqrt / sqrt chain
  double g = a + 2.0;
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 64; ++i) g = sqrt(g) + 1.5;
  t1 = clock64();
  if (tid == 0) t[7] = (t1 - t0);
  double acc = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) acc += e[k];
  out[blockIdx.x * blockDim.x + tid] = a + s + idx + r + q + acc + g;
}
int main() {
  double* out; long long* t;
  cudaMalloc(&out, 8 * 1024 * 8); cudaMalloc(&t, 64 * 8);
  for (int threads : {32, 128}) {
    lat<<<1, threads>>>(out, t, 1.25);
    cudaDeviceSynchronize();
    lat<<<1, threads>>>(out, t, 1.25);
    cudaDeviceSynchronize();
    long long h[8]; cudaMemcpy(h, t, 64, cudaMemcpyDeviceToHost);
    printf("threads %d: dfma dep %.1f cyc | dshfl dep %.1f | lds dep %.1f | rcp+add dep %.1f | bar %.1f | shfl+fma %.1f | dfma 8 indep chains: %.2f cyc/inst | sqrt+add %.1f\n", threads,
           h[0] / 256.0, h[1] / 64.0, h[2] / 64.0, h[3] / 64.0, h[4] / 64.0, h[5] / 64.0, h[6] / 512.0, h[7] / 64.0);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
