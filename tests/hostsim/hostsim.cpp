// TEST-ONLY host build of the kernel logic (wsi_b200/csrc/wsi_device.cuh).
//
// The CUDA library instantiates wsi::replica_step inside __global__ kernels; this file
// instantiates the very same templates for the CPU so that the CPU-only test run
// (`pytest -m "not gpu"`) can check the kernel logic against the golden vectors without a
// GPU.  It is built by tests/hostsim/build.py into tests/hostsim/libwsi_hostsim.so, loaded
// only by tests/test_hostsim.py, and is NOT part of the product: wsi_b200 never loads it and
// has no CPU execution path.
#include <stdlib.h>
#include <string.h>

#include "../../wsi_b200/csrc/wsi_device.cuh"

using namespace wsi;

template <bool LEAN, class CFG>
static int run_cfg_as(const WsiModelDesc* d, const CFG& cfg, int64_t t0, int64_t t1) {
  const int V = d->V;
  double* acc = (double*)calloc((size_t)3 * V, sizeof(double));
  for (int64_t t = t0; t < t1; t++) {
    memset(acc, 0, sizeof(double) * 3 * V);
    int viol = 0;
    for (int r = 0; r < d->R; r++) {
      Ctx<CFG, false, LEAN> c = make_ctx<false, LEAN>(d, 0, cfg, r, (int)t, WSI_RUN_ALL);
      c.mb_acc = d->mb_sums ? acc : nullptr;
      c.mb_stride = 1;
      replica_step<true>(c, 0, d->n_sched);
      viol += c.viol;
    }
    if (d->mb_sums) memcpy(d->mb_sums + (size_t)t * 3 * V, acc, sizeof(double) * 3 * V);
    if (d->mb_count) d->mb_count[t] = viol;
  }
  free(acc);
  return 0;
}

// the library's choice (wsi_kernels.cuh launch()): the lean instantiation when desc.features == 0
template <class CFG>
static int run_cfg(const WsiModelDesc* d, const CFG& cfg, int64_t t0, int64_t t1) {
  return d->features == 0 ? run_cfg_as<true>(d, cfg, t0, t1) : run_cfg_as<false>(d, cfg, t0, t1);
}

template <int P>
static CfgRuntime<P> runtime_cfg(const WsiModelDesc* d) {
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

extern "C" int wsi_hostsim_run(const WsiModelDesc* d, int64_t t0, int64_t t1, int force_runtime_cfg) {
  const int def_idx[8] = {10, 11, 4, 3, 2, 12, 5, 9};
  bool is_default = d->P == 12 && d->additive_mask == 0x1CFCu;
  for (int i = 0; i < 8; i++) is_default = is_default && d->idx[i] == def_idx[i];
  if (is_default && !force_runtime_cfg) return run_cfg(d, CfgDefault12(), t0, t1);
  if (d->P == 12) return run_cfg(d, runtime_cfg<12>(d), t0, t1);
  if (d->P == 2) return run_cfg(d, runtime_cfg<2>(d), t0, t1);
  if (d->P == 15) return run_cfg(d, runtime_cfg<15>(d), t0, t1);
  return WSI_ERR_UNSUPPORTED;
}
