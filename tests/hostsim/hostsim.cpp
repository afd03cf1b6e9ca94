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

// per-process switch of surface_phase (csrc/wsi_land.cuh WSI_PROC): all on unless a known-answer test says otherwise
#define WSI_HOSTSIM_PROCESS_MASK
namespace wsi {
unsigned g_wsi_process_mask = ~0u;
}
extern "C" void wsi_hostsim_process_mask(unsigned mask) { wsi::g_wsi_process_mask = mask; }

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

// ---- per-function entry points for the known-answer tests (tests/test_device_functions.py) --------------
// The reference's unit tests hold known answers for the VQIP algebra (tests/test_core.py:32-130), the tanks
// (tests/test_tanks.py:35-200) and the decay step; they use two pollutants, "phosphate" (additive) and
// "temperature" (not).  These calls run the very device functions the kernels are made of (wsi_vq.cuh,
// wsi_device.cuh, wsi_land.cuh) on that configuration: VQIP = {volume, phosphate, temperature}.
namespace {
struct Kat {
  WsiModelDesc d;
  Ctx<CfgRuntime<2>> c;
  Kat() {
    memset(&d, 0, sizeof d);
    d.P = 2;
    d.V = 3;
    d.R = 1;
    d.float_accuracy = 1e-6;  // constants.FLOAT_ACCURACY as the reference's tests run
    CfgRuntime<2> cfg;
    cfg.mask = 1u << 1;  // plane 1 = phosphate is additive, plane 2 = temperature is not
    cfg.i_no3 = cfg.i_no2 = cfg.i_nh4 = cfg.i_orgp = cfg.i_orgn = cfg.i_solids = -1;
    cfg.i_po4 = 1;
    cfg.i_temp = 2;
    c = make_ctx(&d, 0, cfg, 0, 0, 0u);
  }
};
using V2 = VQ<2>;
V2 ld(const double* p) { V2 v; for (int i = 0; i < 3; i++) v.x[i] = p[i]; return v; }
void st(double* p, const V2& v) { for (int i = 0; i < 3; i++) p[i] = v.x[i]; }
}  // namespace

// op: 0 sum_vqip, 1 extract_vqip, 2 v_change_vqip(a, s), 3 compare_vqip (out[0] = 1/0), 4 is-empty
extern "C" void wsi_hostsim_vq(int op, const double* a, const double* b, double s, double* out) {
  Kat k;
  switch (op) {
    case 0: st(out, k.c.sum(ld(a), ld(b))); break;
    case 1: st(out, k.c.extract(ld(a), ld(b))); break;
    case 2: st(out, k.c.vchange(ld(a), s)); break;
    case 3: out[0] = vq_compare(k.c, ld(a), ld(b)) ? 1.0 : 0.0; break;
    default: out[0] = vq_is_empty(k.c, ld(a)) ? 1.0 : 0.0; break;
  }
}
// Tank with `capacity` holding storage[3] (updated in place).
// op: 0 push_storage(v) -> reply, 1 push_storage(v, force=True) -> reply (empty), 2 pull_storage({volume: s}),
//     3 pull_ponded, 4 get_excess({volume: s}) volume (s < 0: no argument), 5 pull_pollutants(v),
//     6 ResidenceTank.pull_outflow (residence_time = s), 7 evaporate(s) -> out[0] = evaporated
extern "C" void wsi_hostsim_tank(int op, double capacity, double* storage, const double* v, double s, double* out) {
  Kat k;
  switch (op) {
    case 0: st(out, tank_push(k.c, storage, 1, capacity, ld(v))); break;
    case 1: tank_push_force(k.c, storage, 1, ld(v)); st(out, vq_zero<2>()); break;
    case 2: st(out, tank_pull(k.c, storage, 1, s)); break;
    case 3: st(out, tank_pull(k.c, storage, 1, pymax(storage[0] - capacity, 0.0))); break;
    case 4: out[0] = tank_excess_vol(capacity, storage[0], s >= 0, s); break;
    case 5: {
      V2 sv = ld(storage), req = ld(v);
      vq_pull_pollutants(k.c, sv, req);
      st(storage, sv);
      st(out, req);
      break;
    }
    case 6: {
      V2 sv = ld(storage);
      st(out, residence_pull(k.c, sv, s));
      st(storage, sv);
      break;
    }
    default: {  // Tank.evaporate tanks.py:303-317 as surface_phase applies it
      const double e = pymin(s, storage[0]);
      storage[0] -= e;
      out[0] = e;
      break;
    }
  }
}
// DecayTank.end_timestep_decay (tanks.py:442-447, core.py:474-521) of one decaying pollutant (plane `pol`):
// storage[3] in place, out = {total_decayed[pol], storage_[pol]}
extern "C" void wsi_hostsim_decay(double* storage, int pol, double constant, double exponent, double temperature, double* out) {
  Kat k;
  double planes[1 + 5 * WSI_MAX_DECAY] = {0};
  planes[0] = 1;
  planes[1] = pol;
  planes[1 + WSI_MAX_DECAY] = constant;
  planes[1 + 2 * WSI_MAX_DECAY] = exponent;
  auto plane = [&](int f) { return planes + f; };
  V2 sv = ld(storage);
  decay_end_timestep(k.c, sv, plane, 0, temperature);
  st(storage, sv);
  out[0] = planes[1 + 3 * WSI_MAX_DECAY];
  out[1] = planes[1 + 4 * WSI_MAX_DECAY];
}
