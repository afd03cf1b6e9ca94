// libwsi_b200.so: sm_100a kernels and the C ABI of include/wsi_b200.h.
//
// Kernels
//   wsi_step_kernel<CFG>  one thread per replica runs wsi::replica_step (wsi_device.cuh):
//                         schedule, mass balance, recording, end_timestep of one timestep.
//                         The per-block mass-balance partials (violation count, system
//                         totals) are reduced in shared memory / with warp shuffles.
//   wsi_mb_finalize       fixed-order reduction of the per-block partials -> mb_sums[t].
//
// The path is HBM-bound fp64 element-wise work: no tensor cores.  Grid sizing,
// coalescing and the byte accounting are described in DESIGN.md.
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>

#include <atomic>

#include "wsi_device.cuh"

using namespace wsi;

namespace {

thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};

int fail(int code, const char* fmt, const char* detail = "") {
  snprintf(g_err, sizeof g_err, fmt, detail);
  return code;
}

constexpr int kBlock = 64;

template <class CFG>
__global__ void __launch_bounds__(kBlock)
    wsi_step_kernel(const __grid_constant__ WsiModelDesc d, const CFG cfg, const int t, const int sched_begin,
                    const int sched_end, const uint32_t flags) {
  extern __shared__ double s_acc[];  // [3*V][blockDim.x] system-total accumulators
  constexpr int V = CFG::P + 1;
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = r < d.R;
  const bool want_sums = (flags & WSI_RUN_MASS_BALANCE) && d.mb_sums != nullptr;
  Ctx<CFG> c = make_ctx(&d, cfg, active ? r : 0, t, flags);
  if (want_sums) {
    for (int j = 0; j < 3 * V; j++) s_acc[j * blockDim.x + threadIdx.x] = 0.0;
    c.mb_acc = s_acc + threadIdx.x;
    c.mb_stride = blockDim.x;
  }
  if (active) replica_step(c, sched_begin, sched_end);

  if (flags & WSI_RUN_MASS_BALANCE) {
    // block-reduced mass-balance check: violations via warp shuffle + one atomic per warp
    int v = active ? c.viol : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0 && v != 0 && d.mb_count) atomicAdd(&d.mb_count[t], v);
    if (want_sums) {
      __syncthreads();
      // one thread per (term, lane) row sums the block's columns in a fixed order
      for (int j = threadIdx.x; j < 3 * V; j += blockDim.x) {
        double acc = 0.0;
        for (int k = 0; k < (int)blockDim.x; k++) acc += s_acc[j * blockDim.x + k];
        d.mb_partial[(size_t)blockIdx.x * 3 * V + j] = acc;
      }
    }
  }
}

__global__ void wsi_mb_finalize(const double* __restrict__ partial, double* __restrict__ sums, int n_blocks,
                                int n) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  double acc = 0.0;
  for (int b = 0; b < n_blocks; b++) acc += partial[(size_t)b * n + j];
  sums[j] = acc;
}

template <int P>
CfgRuntime<P> runtime_cfg(const WsiModelDesc* d) {
  CfgRuntime<P> c;
  c.mask = d->additive_mask;
  c.i_no3 = d->idx[0];
  c.i_no2 = d->idx[1];
  c.i_nh4 = d->idx[2];
  c.i_po4 = d->idx[3];
  c.i_orgp = d->idx[4];
  c.i_orgn = d->idx[5];
  c.i_solids = d->idx[6];
  c.i_temp = d->idx[7];
  return c;
}

bool is_default12(const WsiModelDesc* d) {
  static const int def_idx[8] = {10, 11, 4, 3, 2, 12, 5, 9};
  if (d->P != 12 || d->additive_mask != 0x1CFCu) return false;
  for (int i = 0; i < 8; i++)
    if (d->idx[i] != def_idx[i]) return false;
  return true;
}

int n_blocks(const WsiModelDesc* d) { return (d->R + kBlock - 1) / kBlock; }

template <class CFG>
int launch(const WsiModelDesc* d, const CFG& cfg, int64_t t, int b, int e, uint32_t flags, cudaStream_t s) {
  constexpr int V = CFG::P + 1;
  const bool want_sums = (flags & WSI_RUN_MASS_BALANCE) && d->mb_sums != nullptr;
  const size_t smem = want_sums ? (size_t)3 * V * kBlock * sizeof(double) : 0;
  const int nb = n_blocks(d);
  wsi_step_kernel<CFG><<<nb, kBlock, smem, s>>>(*d, cfg, (int)t, b, e, flags);
  g_launches++;
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) return fail(WSI_ERR_LAUNCH, "wsi_step_kernel launch failed: %s", cudaGetErrorString(err));
  if (want_sums) {
    const int n = 3 * V;
    wsi_mb_finalize<<<1, 64, 0, s>>>(d->mb_partial, d->mb_sums + (size_t)t * n, nb, n);
    g_launches++;
    err = cudaGetLastError();
    if (err != cudaSuccess) return fail(WSI_ERR_LAUNCH, "wsi_mb_finalize launch failed: %s", cudaGetErrorString(err));
  }
  return WSI_OK;
}

int dispatch(const WsiModelDesc* d, int64_t t, int b, int e, uint32_t flags, cudaStream_t s) {
  if (is_default12(d)) return launch(d, CfgDefault12(), t, b, e, flags, s);
#ifndef WSI_ONLY_DEFAULT12  // development builds instantiate a single configuration
  if (d->P == 12) return launch(d, runtime_cfg<12>(d), t, b, e, flags, s);
  if (d->P == 2) return launch(d, runtime_cfg<2>(d), t, b, e, flags, s);
#endif
  return fail(WSI_ERR_UNSUPPORTED, "no kernel instantiated for this number of pollutants%s");
}

}  // namespace

extern "C" {

int wsi_version(void) { return WSI_ABI_VERSION; }

const char* wsi_last_error(void) { return g_err; }

int wsi_validate(const WsiModelDesc* d) {
  if (!d) return fail(WSI_ERR_BAD_DESC, "null descriptor%s");
  if (d->abi_version != WSI_ABI_VERSION) return fail(WSI_ERR_BAD_DESC, "ABI version mismatch%s");
  if (d->V != d->P + 1 || d->P < 1 || d->P > WSI_MAXP) return fail(WSI_ERR_BAD_DESC, "bad P/V%s");
  if (d->R < 1 || d->T < 1) return fail(WSI_ERR_BAD_DESC, "R and T must be positive%s");
  if (d->nq < 2 || d->nq > 16) return fail(WSI_ERR_BAD_DESC, "nq out of range%s");
  if (d->n_nodes < 0 || d->n_arcs < 0 || d->n_sched < 0) return fail(WSI_ERR_BAD_DESC, "negative sizes%s");
  if (d->n_nodes > 0 && (!d->node_type || !d->node_gidx || !d->out_ptr || !d->in_ptr || !d->cand_ptr))
    return fail(WSI_ERR_BAD_DESC, "graph arrays missing%s");
  if (d->n_sched > 0 && (!d->sched_op || !d->sched_node)) return fail(WSI_ERR_BAD_DESC, "schedule missing%s");
  if ((d->n_arcs > 0 && !d->arc) || (d->n_land > 0 && !d->land) || (d->n_surf > 0 && !d->surf) ||
      (d->n_tank > 0 && !d->tank) || (d->n_river > 0 && !d->river) || (d->n_sewer > 0 && !d->sewer) ||
      (d->n_catch > 0 && !d->catchm))
    return fail(WSI_ERR_BAD_DESC, "group tensor missing%s");
  if (!d->node_forcing || !d->month_idx || !d->dayofyear || !d->doy_adj)
    return fail(WSI_ERR_BAD_DESC, "forcing arrays missing%s");
  if (d->n_probe > 0 && (!d->probe_group || !d->probe_plane || !d->probe_obj || !d->out_probe))
    return fail(WSI_ERR_BAD_DESC, "probe arrays missing%s");
  if (d->record_mb && !d->out_mb) return fail(WSI_ERR_BAD_DESC, "record_mb without out_mb%s");
  if (d->mb_sums && !d->mb_partial) return fail(WSI_ERR_BAD_DESC, "mb_sums without mb_partial%s");
  if (!d->node_flags) return fail(WSI_ERR_BAD_DESC, "node_flags missing%s");
  if (!(is_default12(d) || d->P == 12 || d->P == 2))
    return fail(WSI_ERR_UNSUPPORTED, "kernels are instantiated for P in {2, 12}%s");
  return WSI_OK;
}

int wsi_num_blocks(const WsiModelDesc* d) { return d ? n_blocks(d) : 0; }

int wsi_sched_range(const WsiModelDesc* d, int64_t t, int32_t begin, int32_t end, uint32_t flags, void* stream) {
  int rc = wsi_validate(d);
  if (rc) return rc;
  if (t < 0 || t >= d->T) return fail(WSI_ERR_BAD_DESC, "timestep outside the forcing tables%s");
  if (begin < 0 || end > d->n_sched || begin > end) return fail(WSI_ERR_BAD_DESC, "bad schedule range%s");
  return dispatch(d, t, begin, end, flags, (cudaStream_t)stream);
}

int wsi_step(const WsiModelDesc* d, int64_t t, void* stream) {
  if (!d) return fail(WSI_ERR_BAD_DESC, "null descriptor%s");
  return wsi_sched_range(d, t, 0, d->n_sched, WSI_RUN_ALL, stream);
}

int wsi_run_steps(const WsiModelDesc* d, int64_t t0, int64_t t1, void* stream) {
  int rc = wsi_validate(d);
  if (rc) return rc;
  if (t0 < 0 || t1 > d->T || t0 > t1) return fail(WSI_ERR_BAD_DESC, "timestep range outside the forcing tables%s");
  for (int64_t t = t0; t < t1; t++) {
    rc = dispatch(d, t, 0, d->n_sched, WSI_RUN_ALL, (cudaStream_t)stream);
    if (rc) return rc;
  }
  return WSI_OK;
}

int64_t wsi_launch_count(void) { return (int64_t)g_launches.load(); }

}  // extern "C"
