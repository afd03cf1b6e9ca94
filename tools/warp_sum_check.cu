// Standalone check (run on the GPU box): warp_sum_16 of wsi_warp.cuh gives bit-identical sums to the
// plain shfl_down tree it replaced in wsi_mb_kernel.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o gpurun_out/warp_sum_check tools/warp_sum_check.cu && gpurun_out/warp_sum_check
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "../wsi_b200/csrc/wsi_warp.cuh"

__global__ void check(const double* in, double* tree, double* fast) {
  const int lane = threadIdx.x & 31, warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  double a[16];
  for (int p = 0; p < 16; p++) a[p] = in[((size_t)warp * 32 + lane) * 16 + p];
  for (int p = 0; p < 16; p++) {
    double x = a[p];
    for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
    if (lane == 0) tree[warp * 16 + p] = x;
  }
  int p_mine;
  const double x = warp_sum_16(a, lane, p_mine);
  if ((lane & 1) == 0) fast[warp * 16 + p_mine] = x;
}

int main() {
  const int warps = 4096, n = warps * 32 * 16;
  double* h = (double*)malloc(sizeof(double) * n);
  srand(7);
  for (int i = 0; i < n; i++) {
    const double u = rand() / (double)RAND_MAX, e = (rand() % 60) - 30;
    h[i] = (rand() % 7 == 0) ? 0.0 : (u - 0.3) * exp2(e);
  }
  double *d_in, *d_tree, *d_fast;
  cudaMalloc(&d_in, sizeof(double) * n);
  cudaMalloc(&d_tree, sizeof(double) * warps * 16);
  cudaMalloc(&d_fast, sizeof(double) * warps * 16);
  cudaMemcpy(d_in, h, sizeof(double) * n, cudaMemcpyHostToDevice);
  cudaMemset(d_fast, 0xff, sizeof(double) * warps * 16);
  check<<<warps / 4, 128>>>(d_in, d_tree, d_fast);
  if (cudaDeviceSynchronize() != cudaSuccess) { printf("kernel failed\n"); return 2; }
  double* t = (double*)malloc(sizeof(double) * warps * 16);
  double* f = (double*)malloc(sizeof(double) * warps * 16);
  cudaMemcpy(t, d_tree, sizeof(double) * warps * 16, cudaMemcpyDeviceToHost);
  cudaMemcpy(f, d_fast, sizeof(double) * warps * 16, cudaMemcpyDeviceToHost);
  const int bad = memcmp(t, f, sizeof(double) * warps * 16) != 0;
  printf("warp_sum_16 vs shfl_down tree over %d sums: %s\n", warps * 16, bad ? "MISMATCH" : "bit-identical");
  return bad;
}
