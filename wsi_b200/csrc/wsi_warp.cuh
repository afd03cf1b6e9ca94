// Warp-level reductions shared by the object-parallel kernels (wsi_kernels.cuh).
#pragma once

// Warp sums of 16 per-lane values with 16 shuffles instead of 80: at each offset a lane keeps one half
// of its values and hands the other half to its partner lane, so the values spread over the lanes while
// the lanes fold.  The additions pair up exactly as in a shfl_down tree (lane l with l + 16, then l + 8,
// ...; addition commutes bit for bit), so every sum is bit-identical to the plain reduction.  On return
// the even lane `lane` holds the sum of value `p_mine` (the odd lane next to it holds the same).
__device__ __forceinline__ double warp_sum_16(const double (&a)[16], const int lane, int& p_mine) {
  constexpr unsigned kAll = 0xffffffffu;
  const bool h4 = lane & 16, h3 = lane & 8, h2 = lane & 4, h1 = lane & 2;
  double b[8], c4[4], c2[2];
#pragma unroll
  for (int i = 0; i < 8; i++)
    b[i] = (h4 ? a[2 * i + 1] : a[2 * i]) + __shfl_xor_sync(kAll, h4 ? a[2 * i] : a[2 * i + 1], 16);
#pragma unroll
  for (int i = 0; i < 4; i++)
    c4[i] = (h3 ? b[2 * i + 1] : b[2 * i]) + __shfl_xor_sync(kAll, h3 ? b[2 * i] : b[2 * i + 1], 8);
#pragma unroll
  for (int i = 0; i < 2; i++)
    c2[i] = (h2 ? c4[2 * i + 1] : c4[2 * i]) + __shfl_xor_sync(kAll, h2 ? c4[2 * i] : c4[2 * i + 1], 4);
  double x = (h1 ? c2[1] : c2[0]) + __shfl_xor_sync(kAll, h1 ? c2[0] : c2[1], 2);
  x += __shfl_xor_sync(kAll, x, 1);
  p_mine = (h1 ? 8 : 0) + (h2 ? 4 : 0) + (h3 ? 2 : 0) + (h4 ? 1 : 0);
  return x;
}
