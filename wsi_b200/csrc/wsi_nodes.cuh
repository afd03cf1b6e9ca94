// Storage / Groundwater / River / Reservoir / Sewer / Catchment operations, per-node mass
// balance, recording, end_timestep and the per-replica step driver
// (reference: wsimod/nodes/storage.py, sewer.py, catchment.py, waste.py, core/core.py,
// orchestration/model.py:993-1176).  Included by wsi_device.cuh.
#pragma once

namespace wsi {

// ---- Groundwater / Storage / Reservoir ----------------------------------------------------
// A schedule entry that pushes is split in two: *_pre pulls what it wants to send out of its tank
// (returning false when there is nothing to push), the push itself is the one instance in run_op,
// and tank_post forces whatever comes back into the tank again.
template <class C>
WSI_HD void gw_infiltrate_pre(C& c, int node, typename C::Vq& v) {  // storage.py:231-245
  constexpr int V = C::V;
  const int g = c.D().node_gidx[node];
  double* st = c.tank(WSI_TANK_STORAGE(V), g);
  double avail = st[0];
  avail = pymax(avail - *c.tank(WSI_TANK_CAPACITY, g) * *c.tank(WSI_TANK_INFILTRATION_THRESHOLD, g), 0.0);
  avail = pow(avail * *c.tank(WSI_TANK_INFILTRATION_PCT, g), 0.5);
  v = tank_pull(c, st, c.s_tank(), avail);
}
template <class C>
WSI_HD void gw_distribute_pre(C& c, int node, typename C::Vq& v) {  // storage.py:222-229
  constexpr int V = C::V;
  const int g = c.D().node_gidx[node];
  double* st = c.tank(WSI_TANK_STORAGE(V), g);
  const double avail = st[0] / *c.tank(WSI_TANK_RESIDENCE_TIME, g);
  v = tank_pull(c, st, c.s_tank(), avail);
}
template <class C>
WSI_HD void storage_distribute_pre(C& c, int node, typename C::Vq& v) {  // storage.py:120-127
  constexpr int V = C::V;
  const int g = c.D().node_gidx[node];
  double* st = c.tank(WSI_TANK_STORAGE(V), g);
  v = tank_pull(c, st, c.s_tank(), st[0]);
}
// ---- River --------------------------------------------------------------------------------
// Python's sum() over a list slice (CPython >= 3.12: Neumaier compensated summation) on a
// ring buffer stored as planes; entries [first, first+n) counted from the oldest.
WSI_HD double ring_sum(const double* ring, size_t stride, int cap, int head, int first, int n) {
  if (n <= 0) return 0.0;
  double f = ring[(size_t)((head + first) % cap) * stride], comp = 0.0;
  for (int k = 1; k < n; k++) {
    const double x = ring[(size_t)((head + first + k) % cap) * stride];
    const double t = f + x;
    if (fabs(f) >= fabs(x))
      comp += (f - t) + x;
    else
      comp += (x - t) + f;
    f = t;
  }
  if (comp != 0.0 && isfinite(comp)) f += comp;
  return f;
}

// River.calculate_discharge -> biochemical_processes storage.py:738-898, 917-924
template <class C>
WSI_HD void river_calc(C& c, int node) {
  using Vq = typename C::Vq;
  constexpr int P = C::P;
  constexpr int V = C::V;
  const auto& cfg = c.cfg;
  if (cfg.no3() < 0) return;
  const int g = c.D().node_gidx[node];
  const Obj Rv = c.river_obj(g);
  const int tk = (int)*Rv(WSI_RIVER_TANK);
  const size_t sr = c.s_river();
  double* stp = c.tank(WSI_TANK_STORAGE(V), tk);
  Vq st = c.ldv(stp, c.s_tank());
  Vq in_ = vq_zero<P>(), out_ = vq_zero<P>();
  const double area = *Rv(WSI_RIVER_AREA);
  const double current_depth = st.x[0] / area;  // update_depth storage.py:723-725
  const double T_wdays = *Rv(WSI_RIVER_T_WDAYS);
  const double airT = c.forcing((int)*c.tank(WSI_TANK_FORCING_ID, tk), 2);
  double T = (1 - 1 / T_wdays) * vget<P>(st, cfg.temp()) + (1 / T_wdays) * airT;
  vset<P>(st, cfg.temp(), T);
  // lagged lists storage.py:753-763
  int nT = (int)*Rv(WSI_RIVER_N_LAG_T), hT = (int)*Rv(WSI_RIVER_HEAD_T);
  if (nT > (int)*Rv(WSI_RIVER_MAX_TEMP_LAG)) {
    hT = (hT + 1) % WSI_LAG_T;
    nT--;
  }
  *Rv(WSI_RIVER_LAG_T + (hT + nT) % WSI_LAG_T) = T;
  nT++;
  *Rv(WSI_RIVER_N_LAG_T) = (double)nT;
  *Rv(WSI_RIVER_HEAD_T) = (double)hT;
  int nTP = (int)*Rv(WSI_RIVER_N_LAG_TP), hTP = (int)*Rv(WSI_RIVER_HEAD_TP);
  const double max_p_lag = *Rv(WSI_RIVER_MAX_PHOSPHORUS_LAG);
  if (nTP > (int)max_p_lag) {
    hTP = (hTP + 1) % WSI_LAG_TP;
    nTP--;
  }
  *Rv(WSI_RIVER_LAG_TP + (hTP + nTP) % WSI_LAG_TP) = vget<P>(st, cfg.po4()) + vget<P>(st, cfg.orgp());
  nTP++;
  *Rv(WSI_RIVER_N_LAG_TP) = (double)nTP;
  *Rv(WSI_RIVER_HEAD_TP) = (double)hTP;

  if (!(st.x[0] < c.FA) && !(T <= 0)) {
    const int dc[2] = {cfg.nh4(), cfg.no3()};  // din_components storage.py:566
    double tempfcn = pow(2.0, (T - 20) / 10);
    if (T < 5) tempfcn *= T / 5;
    double din = 0.0 + vget<P>(st, dc[0]) + vget<P>(st, dc[1]);
    const double din_concentration = din / st.x[0];
    const double confcn = din_concentration / (din_concentration + *Rv(WSI_RIVER_HALFSATINWATER));
    const double denitri_water = *Rv(WSI_RIVER_DENPAR_W) * area * tempfcn * confcn;
    const double river_denitrification = pymin(denitri_water, 0.5 * din);
    const double din_ = din - river_denitrification;
    if (din > 0) {
#pragma unroll
      for (int k = 0; k < 2; k++) {
        const double loss = (din - din_) / din * vget<P>(st, dc[k]);
        vadd<P>(out_, dc[k], loss);
        vsub<P>(st, dc[k], loss);
      }
    }
    din = 0.0 + vget<P>(st, dc[0]) + vget<P>(st, dc[1]);
    const int n10 = nT < 10 ? nT : 10, n20 = nT < 20 ? nT : 20;
    const double temp_10_day = ring_sum(Rv(WSI_RIVER_LAG_T), sr, WSI_LAG_T, hT, nT - n10, n10) / 10;
    const double temp_20_day = ring_sum(Rv(WSI_RIVER_LAG_T), sr, WSI_LAG_T, hT, nT - n20, n20) / 20;
    const double total_phos_365_day = ring_sum(Rv(WSI_RIVER_LAG_TP), sr, WSI_LAG_TP, hTP, 0, nTP) / max_p_lag;
    tempfcn = T / 20 * (temp_10_day - temp_20_day) / 5;
    const double limpppar = *Rv(WSI_RIVER_LIMPPPAR), hsatTP = *Rv(WSI_RIVER_HSATTP);
    double totalphosfcn = 0.0;
    if (total_phos_365_day - limpppar + hsatTP > 0)
      totalphosfcn = (total_phos_365_day - limpppar) / (total_phos_365_day - limpppar + hsatTP);
    double minprodN = *Rv(WSI_RIVER_PRODNPAR) * totalphosfcn * tempfcn * area * current_depth;
    double minprodP =
        *Rv(WSI_RIVER_PRODPPAR) * totalphosfcn * tempfcn * area * current_depth * *Rv(WSI_RIVER_UPTAKE_PNRATIO);
    if (minprodN > 0) {  // production (inorganic -> organic)
      minprodN = pymin(0.5 * din, minprodN);
      minprodP = pymin(0.5 * vget<P>(st, cfg.po4()), minprodP);
      vset<P>(out_, cfg.po4(), minprodP);
      vsub<P>(st, cfg.po4(), minprodP);
      vset<P>(in_, cfg.orgp(), minprodP);
      vadd<P>(st, cfg.orgp(), minprodP);
      if (din > 0) {
#pragma unroll
        for (int k = 0; k < 2; k++) {
          const double loss = minprodN * vget<P>(st, dc[k]) / din;
          vadd<P>(out_, dc[k], loss);
          vsub<P>(st, dc[k], loss);
        }
      }
      vset<P>(in_, cfg.orgn(), minprodN);
      vadd<P>(st, cfg.orgn(), minprodN);
    } else {  // mineralisation (organic -> inorganic)
      minprodN = pymin(0.5 * vget<P>(st, cfg.orgn()), -minprodN);
      minprodP = pymin(0.5 * vget<P>(st, cfg.orgp()), -minprodP);
      vset<P>(in_, cfg.po4(), minprodP);
      vadd<P>(st, cfg.po4(), minprodP);
      vset<P>(out_, cfg.orgp(), minprodP);
      vsub<P>(st, cfg.orgp(), minprodP);
      if (din > 0) {
#pragma unroll
        for (int k = 0; k < 2; k++) {
          const double gain = minprodN * vget<P>(st, dc[k]) / din;
          vadd<P>(in_, dc[k], gain);
          vadd<P>(st, dc[k], gain);
        }
      }
      vset<P>(out_, cfg.orgn(), minprodN);
      vsub<P>(st, cfg.orgn(), minprodN);
    }
    din = 0.0 + vget<P>(st, dc[0]) + vget<P>(st, dc[1]);
    const double tempfcn1 = pow(pymax(0.0, T) / 20, 0.3);
    const double tempfcn2 = (T - temp_20_day) / 5;
    tempfcn = pymax(0.0, tempfcn1 * tempfcn2);
    const double macrouptN = *Rv(WSI_RIVER_MUPTNPAR) * tempfcn * area;
    const double mN = pymin(0.5 * din, macrouptN);
    if (din > 0) {
#pragma unroll
      for (int k = 0; k < 2; k++) {
        const double loss = mN * vget<P>(st, dc[k]) / din;
        vadd<P>(out_, dc[k], loss);
        vsub<P>(st, dc[k], loss);
      }
    }
    const double macrouptP = *Rv(WSI_RIVER_MUPTPPAR) * tempfcn * pymax(0.0, totalphosfcn) * area;
    const double mP = pymin(0.5 * vget<P>(st, cfg.po4()), macrouptP);
    vadd<P>(out_, cfg.po4(), mP);
    vsub<P>(st, cfg.po4(), mP);
  }
  c.stv(stp, c.s_tank(), st);
  c.stv(Rv(WSI_RIVER_BIO_IN(V)), sr, in_);
  c.stv(Rv(WSI_RIVER_BIO_OUT(V)), sr, out_);
}

template <class C>
WSI_HD void river_distribute_pre(C& c, int node, typename C::Vq& v) {  // storage.py:926-936
  constexpr int V = C::V;
  const int g = c.D().node_gidx[node];
  const int tk = (int)*c.river(WSI_RIVER_TANK, g);
  double* st = c.tank(WSI_TANK_STORAGE(V), tk);
  v = tank_pull(c, st, c.s_tank(), st[0] * *c.river(WSI_RIVER_RIVERRC, g));
}

// ---- Sewer --------------------------------------------------------------------------------
template <class C>
WSI_HD void sewer_discharge_pre(C& c, int node, typename C::Vq& v) {  // sewer.py:201-239, up to the push
  constexpr int V = C::V;
  const int w = c.D().node_gidx[node];
  sewer_update_queue(c, w);
  v = c.ldv(c.sewer(WSI_SEWER_ACTIVE_STORAGE(V), w), c.s_sewer());
}
template <class C>
WSI_HD void sewer_discharge_post(C& c, int node, const typename C::Vq& remaining) {
  using Vq = typename C::Vq;
  constexpr int V = C::V;
  const int w = c.D().node_gidx[node];
  const size_t s = c.s_sewer();
  double* act = c.sewer(WSI_SEWER_ACTIVE_STORAGE(V), w);
  double* st = c.sewer(WSI_SEWER_STORAGE(V), w);
  Vq active = c.ldv(act, s);
  Vq sent = c.extract(active, remaining);
  // pull_storage_exact tanks.py:566-586
  Vq reply = sent;
  reply.x[0] = pymin(reply.x[0], active.x[0]);
#pragma unroll
  for (int p = 1; p <= C::P; p++)
    if (bit(c.addm(), p)) reply.x[p] = pymin(reply.x[p], active.x[p]);
  active = c.extract(active, reply);
  Vq storage = c.extract(c.ldv(st, s), reply);
  // pull_ponded tanks.py:78-100 -> QueueTank.pull_storage tanks.py:541-564
  const double ponded_v = pymax(storage.x[0] - *c.sewer(WSI_SEWER_CAPACITY, w), 0.0);
  Vq ponded = c.vchange(active, pymin(ponded_v, active.x[0]));
  active = c.extract(active, ponded);
  storage = c.extract(storage, ponded);
  if (ponded.x[0] > c.FA) {
    // push_distributed(ponded, of_type=["Land"]) has no eligible arc (the flattener refuses
    // Sewer->Land arcs), so everything is forced back (sewer.py:236-237, tanks.py:526-530)
    storage = c.sum(storage, ponded);
    active = c.sum(active, ponded);
  }
  c.stv(act, s, active);
  c.stv(st, s, storage);
}

// QueueTank._end_timestep tanks.py:631-635 + AltQueueArc.end_timestep arcs.py:640-660; with decays:
// DecayQueueTank._end_timestep tanks.py:688-696 + DecayArcAlt._end_timestep arcs.py:778-805
template <class C>
WSI_HD void sewer_end_timestep(C& c, int w) {
  using Vq = typename C::Vq;
  constexpr int V = C::V;
  const size_t s = c.s_sewer();
  const int nq = c.D().nq;
  const int n_decay = (int)*c.sewer(WSI_SEWER_N_DECAY, w);
  double* stp = c.sewer(WSI_SEWER_STORAGE(V), w);
  if (n_decay > 0) {  // what decayed in the delay line this step leaves the storage, then the tally restarts
    Vq st = c.ldv(stp, s);
    for (int i = 0; i < WSI_MAX_DECAY; i++) {
      if (i >= n_decay) break;
      const int pol = (int)*c.sewer(WSI_SEWER_DECAY_POL + i, w);
      if (bit(c.addm(), pol)) vsub<C::P>(st, pol, *c.sewer(WSI_SEWER_TOTAL_DECAYED + i, w));
      *c.sewer(WSI_SEWER_TOTAL_DECAYED + i, w) = 0.0;
    }
    c.stv(stp, s, st);
  }
  *c.sewer(WSI_SEWER_IA_FLOW_IN, w) = 0.0;
  unsigned keys = (unsigned)*c.sewer(WSI_SEWER_QMASK, w);
  int max_travel = 1;
  for (int k = 2; k < nq; k++)
    if (keys & (1u << k)) max_travel = k;
  double* q = c.sewer(WSI_SEWER_QUEUE(V), w);
  const Vq q0 = c.ldv(q, s);
  Vq q1 = c.ldv(q + (size_t)V * s, s);
  queue_make_decay(c, w, q1);  // i = 0: queue[0] = make_decay(queue_[1])
  // arcs.py:653-658: queue[i] = queue_[i+1]; queue[i+1] = empty, for the keys that exist
  unsigned newkeys = keys | 1u;
  for (int i = 1; i < nq; i++) {
    const bool take = (i < max_travel) && ((keys >> (i + 1)) & 1u);
    if (take) {
      Vq moved = c.ldv(q + (size_t)(i + 1) * V * s, s);
      queue_make_decay(c, w, moved);
      c.stv(q + (size_t)i * V * s, s, moved);
      newkeys |= 1u << i;
    } else if ((keys >> i) & 1u) {
      c.stv(q + (size_t)i * V * s, s, vq_zero<C::P>());
    }
  }
  *c.sewer(WSI_SEWER_QMASK, w) = (double)newkeys;
  if (n_decay > 0) {
    // queue[0] = sum_vqip(queue[0], make_decay(queue_[0])); it waits there for the next update_queue
    Vq d0 = q0;
    queue_make_decay(c, w, d0);
    c.stv(q, s, c.sum(q1, d0));
  } else {
    // queue[0] = sum_vqip(queue_[0], queue_[1]), then update_queue() empties it into active
    const Vq merged = c.sum(q0, q1);
    double* act = c.sewer(WSI_SEWER_ACTIVE_STORAGE(V), w);
    c.stv(act, s, c.sum(c.ldv(act, s), merged));
    c.stv(q, s, vq_zero<C::P>());
  }
  // storage_ = copy(storage)
  c.stv(c.sewer(WSI_SEWER_STORAGE_LAG(V), w), s, c.ldv(stp, s));
}

// ---- QueueGroundwater.distribute storage.py:345-363 ------------------------------------------
template <class C>
WSI_HD void qgw_distribute_pre(C& c, int node, typename C::Vq& v) {
  const int w = c.D().node_gidx[node];
  sewer_update_queue(c, w);
  v = c.ldv(c.sewer(WSI_SEWER_ACTIVE_STORAGE(C::V), w), c.s_sewer());
}
template <class C>
WSI_HD void qgw_distribute_post(C& c, int node, const typename C::Vq& remaining) {
  using Vq = typename C::Vq;
  constexpr int V = C::V;
  const int w = c.D().node_gidx[node];
  const size_t s = c.s_sewer();
  double* act = c.sewer(WSI_SEWER_ACTIVE_STORAGE(V), w);
  double* st = c.sewer(WSI_SEWER_STORAGE(V), w);
  Vq active = c.ldv(act, s);
  const Vq sent = c.vchange(active, active.x[0] - remaining.x[0]);
  // QueueTank.pull_storage tanks.py:541-564
  const Vq reply = c.vchange(active, pymin(sent.x[0], active.x[0]));
  c.stv(act, s, c.extract(active, reply));
  c.stv(st, s, c.extract(c.ldv(st, s), reply));
}

// ---- Catchment ------------------------------------------------------------------------------
template <class C>
WSI_HD typename C::Vq catch_get_flow(const C& c, int g) {  // catchment.py:60-75
  const WsiModelDesc& d = c.D();
  const int cid = (int)d.catchm[(size_t)g * c.R + c.r];
  const double* row = d.catch_inputs + ((size_t)c.t * d.n_cseries + cid) * C::V;
  typename C::Vq v;
#pragma unroll
  for (int p = 0; p <= C::P; p++) v.x[p] = row[p];
#pragma unroll
  for (int p = 1; p <= C::P; p++)
    if (bit(c.addm(), p)) v.x[p] *= v.x[0];
  return v;
}
// Boundary node of a partitioned network (backend/partition.py): the VQIP the cut-off upstream
// river parked on the cut arc at this timestep, as totals; its series id is the node's group index.
template <class C>
WSI_HD typename C::Vq boundary_flow(const C& c, int node) {
  const WsiModelDesc& d = c.D();
  const double* row = d.catch_inputs + ((size_t)c.t * d.n_cseries + d.node_gidx[node]) * C::V;
  typename C::Vq v;
#pragma unroll
  for (int p = 0; p <= C::P; p++) v.x[p] = row[p];
  return v;
}
// Put the series value on the node's out-arc exactly as a river whose release is admitted in
// full books it (wsi_net.cuh, pending_mode 0); `deliver`: also hand it to the receiving river
// (the sequential engine; the network engine's receivers pick `pending` up themselves).
template <class C>
WSI_HD void boundary_release(C& c, int node, bool deliver) {
  constexpr int V = C::V;
  const WsiModelDesc& d = c.D();
  const int a0 = d.out_arcs[d.out_ptr[node]];
  const Obj A = c.arc_obj(a0);
  const typename C::Vq v = boundary_flow(c, node);
  c.stv(A(WSI_ARC_PENDING(V)), A.s, v);
  *A(WSI_ARC_PENDING_MODE) = 0.0;
  arc_book(c, a0, v);
  if (deliver) {
    const int recv = d.arc_out[a0];
    if (d.node_type[recv] == WSI_NT_RIVER) {
      const int tk = (int)*c.river(WSI_RIVER_TANK, d.node_gidx[recv]);
      tank_push_force(c, c.tank(WSI_TANK_STORAGE(V), tk), c.s_tank(), v);
    }
  }
}

// Catchment.route catchment.py:77-88: what is still to be routed, and -- after the push -- the
// mass-balance terms, which need unrouted_water (the reply)
}  // namespace wsi
#include "wsi_works.cuh"  // pull side, treatment works, demand (needs catch_get_flow above)
namespace wsi {

template <class C>
WSI_HD void catchment_route_pre(C& c, int node, typename C::Vq& v) {
  v = catchment_avail(c, node);  // get_avail catchment.py:90-104: what abstractions have left
}
template <class C>
WSI_HD void catchment_route_post(C& c, int node, const typename C::Vq& reply) {
  using Vq = typename C::Vq;
  if (c.run_flags & WSI_RUN_MASS_BALANCE) {
    // in = total_in + get_flow(); out = total_out + unrouted_water (catchment.py:55-56)
    Vq in_ = total_in(c, node);
    ds_add(c, in_, catch_get_flow(c, c.D().node_gidx[node]));
    Vq out_ = total_out(c, node);
    ds_add(c, out_, reply);
    store_mb_terms(c, node, in_, vq_zero<C::P>(), out_);
  }
}

// ---- mass balance of the remaining node classes (core.py:370-429) --------------------------
template <class C>
WSI_HD typename C::Vq tank_ds(const C& c, int tk) {  // tanks.py:70-76 / 449-457
  constexpr int V = C::V;
  using Vq = typename C::Vq;
  const Vq st = c.ldv_add(c.tank(WSI_TANK_STORAGE(V), tk), c.s_tank());
  const Vq st_ = c.ldv_add(c.tank(WSI_TANK_STORAGE_LAG(V), tk), c.s_tank());
  Vq dsv;
#pragma unroll
  for (int p = 0; p <= C::P; p++) dsv.x[p] = (p == 0 || bit(c.addm(), p)) ? st.x[p] - st_.x[p] : 0.0;
  if (*c.tank(WSI_TANK_IS_DECAY, tk) != 0.0) {
    const Obj Tk = c.tank_obj(tk);
    dsv = c.sum(dsv, decay_total<C>(c, Tk, WSI_TANK_N_DECAY));
  }
  return dsv;
}

// arcs are trivially balanced (vqip_in is vqip_out, arcs.py:83-85, 155-156); they only feed the
// system totals of model.py:1016-1023
template <class C>
WSI_HD void arc_system_totals(C& c, int a) {
  const typename C::Vq v = arc_vqip_add(c, a);
#pragma unroll
  for (int p = 0; p <= C::P; p++)
    if (p == 0 || bit(c.addm(), p)) {
      c.mb_acc[(0 * C::V + p) * c.mb_stride] += v.x[p];
      c.mb_acc[(2 * C::V + p) * c.mb_stride] += v.x[p];
    }
}

template <class C>
WSI_HD void node_mass_balance(C& c, int node, typename C::Vq& in_, typename C::Vq& ds, typename C::Vq& out_) {
  using Vq = typename C::Vq;
  constexpr int V = C::V;
  const WsiModelDesc& d = c.D();
  const int g = d.node_gidx[node];
  const int type = d.node_type[node];
  if (type == WSI_NT_LAND) {
    land_mass_balance_terms(c, node, in_, ds, out_);
  } else if (type == WSI_NT_CATCHMENT) {  // formed inside its op
    load_mb_terms(c, node, in_, ds, out_);
  } else if (type == WSI_NT_BOUNDARY) {  // what it releases is what its arc carries
    in_ = vq_zero<C::P>();
    ds_add(c, in_, boundary_flow(c, node));
    out_ = total_out(c, node);
    ds = vq_zero<C::P>();
  } else {
    const Vq tin = total_in(c, node);
    in_ = tin;
    out_ = total_out(c, node);
    ds = vq_zero<C::P>();
    switch (type) {
      case WSI_NT_WASTE:  // waste.py:38
        ds_add(c, out_, tin);
        break;
      case WSI_NT_QUEUEGROUNDWATER:
      case WSI_NT_SEWER: {
        const Vq st = c.ldv_add(c.sewer(WSI_SEWER_STORAGE(V), g), c.s_sewer());
        const Vq st_ = c.ldv_add(c.sewer(WSI_SEWER_STORAGE_LAG(V), g), c.s_sewer());
#pragma unroll
        for (int p = 0; p <= C::P; p++)
          if (p == 0 || bit(c.addm(), p)) ds.x[p] += st.x[p] - st_.x[p];
        break;
      }
      case WSI_NT_GROUNDWATER:
      case WSI_NT_STORAGE:
      case WSI_NT_RESERVOIR:
      case WSI_NT_RIVERRESERVOIR:
        ds_add(c, ds, tank_ds(c, g));
        break;
      case WSI_NT_WWTW:
      case WSI_NT_FWTW:
      case WSI_NT_DEMAND:
      case WSI_NT_UNLIMITEDDISTRIBUTION:
#pragma unroll
        for (int p = 0; p <= C::P; p++)
          if (p == 0 || bit(c.addm(), p)) (void)works_mb_lane(c, type, g, p, in_.x[p], ds.x[p], out_.x[p]);
        break;
      case WSI_NT_RIVER: {
        ds_add(c, in_, c.ldv_add(c.river(WSI_RIVER_BIO_IN(V), g), c.s_river()));
        ds_add(c, out_, c.ldv_add(c.river(WSI_RIVER_BIO_OUT(V), g), c.s_river()));
        ds_add(c, ds, tank_ds(c, (int)*c.river(WSI_RIVER_TANK, g)));
        break;
      }
      default:
        break;
    }
  }
  mb_check(c, node, in_, ds, out_);
}

// ---- the same mass balance, one VQIP lane at a time --------------------------------------------
// The schedule kernel checks every node of its replica after the schedule (fused tail): formed lane
// by lane, a node's (in, ds, out) are three scalars, so the whole pass needs a handful of
// registers.  Every lane value is built with the additions of node_mass_balance in the same order
// (bit-identical terms).  LANE is a compile-time lane (0 volume, else an additive pollutant).
template <int LANE, class C>
WSI_HD double arcs_lane_sum(const C& c, const int* ptr, const int* list, int node) {
  double v = 0.0;
  for (int i = ptr[node]; i < ptr[node + 1]; i++) v += arc_lane(c, list[i], LANE);
  return v;
}
template <int LANE, class C>
WSI_HD double tank_ds_lane(const C& c, int tk) {  // tank_ds
  constexpr int V = C::V;
  double dsv = *c.tank(WSI_TANK_STORAGE(V) + LANE, tk) - *c.tank(WSI_TANK_STORAGE_LAG(V) + LANE, tk);
  if (*c.tank(WSI_TANK_IS_DECAY, tk) != 0.0) dsv = dsv + decay_total_of(c.tank_obj(tk), WSI_TANK_N_DECAY, LANE);
  return dsv;
}
template <int LANE, class C>
WSI_HD void node_mb_lane(const C& c, int node, double& in_, double& ds, double& out_) {
  constexpr int V = C::V;
  const WsiModelDesc& d = c.D();
  const int g = d.node_gidx[node];
  const int type = d.node_type[node];
  if (type == WSI_NT_LAND) {  // land_mass_balance_terms
    const size_t ss = c.s_surf();
    in_ = arcs_lane_sum<LANE>(c, d.in_ptr, d.in_arcs, node);
    out_ = arcs_lane_sum<LANE>(c, d.out_ptr, d.out_arcs, node);
    ds = 0.0;
    for (int i = d.land_surf_ptr[g]; i < d.land_surf_ptr[g + 1]; i++) {
      const Obj S = c.surf_obj(d.land_surf[i]);
      in_ += *S(WSI_SURF_MB_IN(V) + LANE);
      out_ += *S(WSI_SURF_MB_OUT(V) + LANE);
    }
    for (int i = d.land_surf_ptr[g]; i < d.land_surf_ptr[g + 1]; i++) {
      const Obj S = c.surf_obj(d.land_surf[i]);
      if ((int)*S(WSI_SURF_TYPE) == WSI_ST_IMPERVIOUSSURFACE) {
        double dsv = *S(WSI_SURF_STORAGE(V) + LANE) - *S(WSI_SURF_MB_DS(V) + LANE);
        dsv = dsv + decay_total_of(S, WSI_SURF_N_DECAY, LANE);
        ds += dsv;
      } else {
        ds += *S(WSI_SURF_MB_DS(V) + LANE);
      }
    }
    const double* parked = d.mb_terms + (size_t)d.mb_term_slot[node] * 3 * V * c.R + c.r;
    ds += *c.land(WSI_LAND_SURFACE_RUNOFF(V) + LANE, g) - parked[(size_t)(0 * V + LANE) * c.R];
    ds += *c.land(WSI_LAND_SUBSURFACE_RUNOFF(V) + LANE, g) - parked[(size_t)(1 * V + LANE) * c.R];
    ds += *c.land(WSI_LAND_PERCOLATION(V) + LANE, g) - parked[(size_t)(2 * V + LANE) * c.R];
    (void)ss;
  } else if (type == WSI_NT_CATCHMENT) {  // formed inside its op
    const double* p = d.mb_terms + (size_t)d.mb_term_slot[node] * 3 * V * c.R + c.r;
    in_ = p[(size_t)(0 * V + LANE) * c.R];
    ds = p[(size_t)(1 * V + LANE) * c.R];
    out_ = p[(size_t)(2 * V + LANE) * c.R];
  } else if (type == WSI_NT_BOUNDARY) {
    in_ = 0.0 + d.catch_inputs[((size_t)c.t * d.n_cseries + g) * V + LANE];
    out_ = arcs_lane_sum<LANE>(c, d.out_ptr, d.out_arcs, node);
    ds = 0.0;
  } else {
    const double tin = arcs_lane_sum<LANE>(c, d.in_ptr, d.in_arcs, node);
    in_ = tin;
    out_ = arcs_lane_sum<LANE>(c, d.out_ptr, d.out_arcs, node);
    ds = 0.0;
    switch (type) {
      case WSI_NT_WASTE:
        out_ += tin;
        break;
      case WSI_NT_QUEUEGROUNDWATER:
      case WSI_NT_SEWER:
        ds += *c.sewer(WSI_SEWER_STORAGE(V) + LANE, g) - *c.sewer(WSI_SEWER_STORAGE_LAG(V) + LANE, g);
        break;
      case WSI_NT_GROUNDWATER:
      case WSI_NT_STORAGE:
      case WSI_NT_RESERVOIR:
      case WSI_NT_RIVERRESERVOIR:
        ds += tank_ds_lane<LANE>(c, g);
        break;
      case WSI_NT_WWTW:
      case WSI_NT_FWTW:
      case WSI_NT_DEMAND:
      case WSI_NT_UNLIMITEDDISTRIBUTION:
        (void)works_mb_lane(c, type, g, LANE, in_, ds, out_);
        break;
      case WSI_NT_RIVER:
        in_ += *c.river(WSI_RIVER_BIO_IN(V) + LANE, g);
        out_ += *c.river(WSI_RIVER_BIO_OUT(V) + LANE, g);
        ds += tank_ds_lane<LANE>(c, (int)*c.river(WSI_RIVER_TANK, g));
        break;
      default:
        break;
    }
  }
}
// One lane of every node of the replica: violations (flagged on the node) and this replica's
// contribution to the system totals of model.py:1016-1036 (nodes and -- `with_arcs` -- arcs).
template <int LANE, class C>
WSI_HD int replica_mb_lane(C& c, bool with_arcs, double& sys_in, double& sys_ds, double& sys_out) {
  constexpr int V = C::V;
  const WsiModelDesc& d = c.D();
  const int row = c.t - d.rec_t0;
  const bool rec = d.record_mb && row >= 0 && row < d.rec_T;
  int viol = 0;
  sys_in = 0.0;
  sys_ds = 0.0;
  sys_out = 0.0;
  if (with_arcs)
    for (int a = 0; a < d.n_arcs; a++) {
      const double v = arc_lane(c, a, LANE);
      sys_in += v;
      sys_out += v;
    }
  for (int node = 0; node < d.n_nodes; node++) {
    double in_, ds, out_;
    node_mb_lane<LANE>(c, node, in_, ds, out_);
    if (mb_lane_bad(c, in_, ds, out_)) {
      viol++;
      c.flag(node, WSI_FLAG_MASS_BALANCE);
    }
    sys_in += in_;
    sys_ds += ds;
    sys_out += out_;
    if (rec) {
      double* o = d.out_mb + ((size_t)row * d.n_nodes + node) * 3 * V * c.R + c.r;
      o[(size_t)(0 * V + LANE) * c.R] = in_;
      o[(size_t)(1 * V + LANE) * c.R] = ds;
      o[(size_t)(2 * V + LANE) * c.R] = out_;
    }
  }
  return viol;
}

// all mass-balance lanes (volume + additive pollutants), each handed to `sink.take(c, lane, violations,
// sys_in, sys_ds, sys_out)`; `active` is false for the padding threads of a block, which still take
// part in the sink's reductions
template <int LANE, class C, class SINK>
WSI_HD void replica_mb_lanes(C& c, bool active, bool with_arcs, SINK& sink) {
  if constexpr (LANE <= C::P) {
    if (LANE == 0 || bit(c.addm(), LANE)) {
      double a = 0.0, b = 0.0, o = 0.0;
      int viol = 0;
      if (active) viol = replica_mb_lane<LANE>(c, with_arcs, a, b, o);
      sink.take(c, LANE, viol, a, b, o);
    }
    replica_mb_lanes<LANE + 1>(c, active, with_arcs, sink);
  }
}
// host build: into the caller's accumulators
struct MbHostSink {
  template <class C>
  WSI_HD void take(C& c, int lane, int viol, double a, double b, double o) {
    c.viol += viol;
    if (c.mb_acc) {
      c.mb_acc[(0 * C::V + lane) * c.mb_stride] += a;
      c.mb_acc[(1 * C::V + lane) * c.mb_stride] += b;
      c.mb_acc[(2 * C::V + lane) * c.mb_stride] += o;
    }
  }
};

// ---- end_timestep (model.py:1172-1176) ------------------------------------------------------
template <class C>
WSI_HD void store_end_timestep(C& c, int tk) {  // tanks.py:348-351 / 442-447
  constexpr int V = C::V;
  double* stp = c.tank(WSI_TANK_STORAGE(V), tk);
  typename C::Vq st = c.ldv(stp, c.s_tank());
  c.stv(c.tank(WSI_TANK_STORAGE_LAG(V), tk), c.s_tank(), st);
  if (*c.tank(WSI_TANK_IS_DECAY, tk) != 0.0 && (int)*c.tank(WSI_TANK_N_DECAY, tk) > 0) {
    const Obj Tk = c.tank_obj(tk);
    decay_end_timestep(c, st, Tk, WSI_TANK_N_DECAY, c.forcing((int)*c.tank(WSI_TANK_FORCING_ID, tk), 2));
    c.stv(stp, c.s_tank(), st);
  }
}

// end_timestep of one object; `kind` indexes the group (0 arc, 2 surface, 3 tank, 4 river,
// 5 sewer).  Objects are independent here, so the CUDA library runs one thread per
// (object, replica) (wsi_end_kernel); the fused per-replica step loops over the same list.
template <class C>
WSI_HD void end_item(C& c, int kind, int obj) {
  constexpr int V = C::V;
  switch (kind) {
    case 0:  // Arc.end_timestep arcs.py:259-265 (network engine; the replica engine's arcs expire by stamp)
      if (c.D().engine == 1) {
        *c.arc(WSI_ARC_FLOW_IN, obj) = 0.0;
        c.stv(c.arc(WSI_ARC_VQIP_IN(V), obj), c.s_arc(), vq_zero<C::P>());
      }
      break;
    case 2: {  // Surface (DecayTank) end_timestep tanks.py:442-447, land.py:266-275
      const Obj S = c.surf_obj(obj);
      if ((int)*S(WSI_SURF_N_DECAY) > 0) {
        const int L = (int)*S(WSI_SURF_LAND);
        const double temperature = c.forcing((int)*c.land(WSI_LAND_FORCING_ID, L), 2);
        double* stp = S(WSI_SURF_STORAGE(V));
        typename C::Vq st = c.ldv(stp, c.s_surf());
        decay_end_timestep(c, st, S, WSI_SURF_N_DECAY, temperature);
        c.stv(stp, c.s_surf(), st);
      }
      break;
    }
    case 3:
      store_end_timestep(c, obj);
      break;
    case 4:  // River.end_timestep_ storage.py:951-955 (its tank is a kind-3 item)
      c.stv(c.river(WSI_RIVER_BIO_IN(V), obj), c.s_river(), vq_zero<C::P>());
      c.stv(c.river(WSI_RIVER_BIO_OUT(V), obj), c.s_river(), vq_zero<C::P>());
      break;
    case 5:
      sewer_end_timestep(c, obj);
      break;
    case 6:  // WWTW / FWTW (their tank is a kind-3 item)
      works_end_timestep(c, obj);
      break;
    case 7:
      demand_end_timestep(c, obj);
      break;
    case 9:  // RiverReservoir.end_timestep_ storage.py:1129-1133 (its tank is a kind-3 item)
      *c.tank(WSI_TANK_ENV_SATISFIED, obj) = 0.0;
      break;
    case 8:  // UnlimitedDistribution.end_timestep distribution.py:191-200
      c.stv(c.distrib(WSI_DISTRIB_SUPPLIED(V), obj), c.s_distrib(), vq_zero<C::P>());
      break;
    default:
      break;
  }
}

// end_timestep of everything a node owns (the replica engine's settle pass: one thread per
// (node, replica) checks the node's mass balance and then ends its timestep)
template <class C>
WSI_HD void end_node(C& c, int node) {
  const WsiModelDesc& d = c.D();
  const int g = d.node_gidx[node];
  switch (d.node_type[node]) {
    case WSI_NT_LAND:
      for (int i = d.land_surf_ptr[g]; i < d.land_surf_ptr[g + 1]; i++) end_item(c, 2, d.land_surf[i]);
      break;
    case WSI_NT_GROUNDWATER:
    case WSI_NT_STORAGE:
    case WSI_NT_RESERVOIR:
      end_item(c, 3, g);
      break;
    case WSI_NT_RIVERRESERVOIR:
      end_item(c, 3, g);
      end_item(c, 9, g);
      break;
    case WSI_NT_RIVER:
      end_item(c, 4, g);
      end_item(c, 3, (int)*c.river(WSI_RIVER_TANK, g));
      break;
    case WSI_NT_QUEUEGROUNDWATER:
    case WSI_NT_SEWER:
      end_item(c, 5, g);
      break;
    case WSI_NT_WWTW:
    case WSI_NT_FWTW:
      end_item(c, 6, g);
      end_item(c, 3, (int)*c.works(WSI_WORKS_TANK, g));
      break;
    case WSI_NT_DEMAND:
      end_item(c, 7, g);
      break;
    case WSI_NT_UNLIMITEDDISTRIBUTION:
      end_item(c, 8, g);
      break;
    default:
      break;
  }
}

template <class C>
WSI_HDN void end_timestep(C& c) {
  const WsiModelDesc& d = c.D();
  for (int i = 0; i < d.n_end; i++) end_item(c, d.end_kind[i], d.end_obj[i]);
}

// ---- recording (model.py:1063-1170): copy probe planes into out_probe[t, k, r] ---------------
template <class C>
WSI_HD const double* probe_src(const C& c, int group, int plane, int obj) {
  switch (group) {
    case 0: return c.arc(plane, obj);
    case 1: return c.land(plane, obj);
    case 2: return c.surf(plane, obj);
    case 3: return c.tank(plane, obj);
    case 4: return c.river(plane, obj);
    case 7: return c.works(plane, obj);
    case 8: return c.demand(plane, obj);
    case 9: return c.distrib(plane, obj);
    default: return c.sewer(plane, obj);
  }
}
// series probe k: only the members of the record window are kept (rec_R == 0: all of them)
// value of a probed plane; the transients of an arc nothing was booked on this step read as zero
template <class C>
WSI_HD double probe_value(const C& c, int group, int plane, int obj) {
  if (group == 0 && (plane == WSI_ARC_FLOW_IN || (plane >= WSI_ARC_VQIP_IN(C::V) && plane < WSI_ARC_VQIP_IN(C::V) + C::V)) &&
      !arc_live(c, obj))
    return 0.0;
  return *probe_src(c, group, plane, obj);
}
template <class C>
WSI_HD void record_probe(const C& c, int row, int k) {
  const WsiModelDesc& d = c.D();
  const int nr = d.rec_R > 0 ? d.rec_R : d.R;
  const int rr = c.r - (d.rec_R > 0 ? d.rec_r0 : 0);
  if (rr < 0 || rr >= nr) return;
  d.out_probe[((size_t)row * d.n_probe + k) * (size_t)nr + rr] = probe_value(c, d.probe_group[k], d.probe_plane[k], d.probe_obj[k]);
}
// reduced probe k: running sum and maximum over the timesteps (what the objectives of
// model.py:1177-1197 need of a member's series, without keeping the series)
template <class C>
WSI_HD void reduce_probe(const C& c, int k) {
  const WsiModelDesc& d = c.D();
  const double v = probe_value(c, d.red_group[k], d.red_plane[k], d.red_obj[k]);
  double* sum = d.red_out + (size_t)k * c.R + c.r;
  double* mx = sum + (size_t)d.n_red * c.R;
  *sum = *sum + v;
  if (v > *mx) *mx = v;
}
// item k of the recording pass: series probes first, then the reduced ones
template <class C>
WSI_HD void record_item(const C& c, int row, int k) {
  const WsiModelDesc& d = c.D();
  if (k < d.n_probe) {
    if (row >= 0 && row < d.rec_T) record_probe(c, row, k);
  } else {
    reduce_probe(c, k - d.n_probe);
  }
}
template <class C>
WSI_HD void record(const C& c) {
  const WsiModelDesc& d = c.D();
  const int row = c.t - d.rec_t0;
  for (int k = 0; k < d.n_probe + d.n_red; k++) record_item(c, row, k);
}

// Out-of-line ops receive a COPY of the context: a context whose address is handed to a function
// that is not inlined would have to live in local memory for the whole kernel.
template <class C>
WSI_HDN void river_calc_call(C c, int node) {
  river_calc(c, node);
}
template <class C>
WSI_HDN void reservoir_abstract_call(C c, int node) {
  reservoir_abstract(c, node);
}
// Reservoir.make_abstractions without in-arcs: pull_distributed returns an empty VQIP (nodes.py:376-407),
// which is then pushed (twice) into the tank (storage.py:983-989)
template <class C>
WSI_HDN void reservoir_settle_call(C c, int node) {
  constexpr int V = C::V;
  const int g = c.D().node_gidx[node];
  double* st = c.tank(WSI_TANK_STORAGE(V), g);
  const typename C::Vq spill = tank_push(c, st, c.s_tank(), *c.tank(WSI_TANK_CAPACITY, g), vq_zero<C::P>());
  tank_push_force(c, st, c.s_tank(), spill);
}

// ---- one schedule entry -----------------------------------------------------------------------
// An entry runs as one or more sub-steps (Land.run: one per surface, then its tanks); each sub-step
// prepares at most one push_distributed call, and the push itself exists ONCE in the kernel (all
// entries share it), which keeps the schedule kernel's code and its live registers small: between
// a sub-step's two halves only the VQIP being pushed (then the reply) and three scalars are alive.
template <class C>
WSI_HD void run_op(C& c, int op, int node) {
  using Vq = typename C::Vq;
  constexpr int V = C::V;
  const int n_sub = op == WSI_OP_LAND_RUN ? land_substeps(c, node) : 1;
#pragma unroll 1
  for (int sub = 0; sub < n_sub; sub++) {
    Vq v;
    OpAux aux = {0.0, 0.0, 0.0, 0};
    int filter = WSI_FILTER_ALL;
    bool push = true;
    switch (op) {
      case WSI_OP_LAND_RUN: push = land_pre(c, node, sub, v, filter, aux); break;
      case WSI_OP_GW_INFILTRATE: gw_infiltrate_pre(c, node, v); filter = WSI_FILTER_SEWER; break;
      case WSI_OP_SEWER_DISCHARGE: sewer_discharge_pre(c, node, v); filter = WSI_FILTER_ALL; break;
      case WSI_OP_GW_DISTRIBUTE: gw_distribute_pre(c, node, v); filter = WSI_FILTER_NRW; break;
      case WSI_OP_RIVER_CALC: river_calc_call(c, node); push = false; break;
      case WSI_OP_RESERVOIR_ABSTRACT:
        if constexpr (C::kLean) reservoir_settle_call(c, node);  // no in-arcs: nothing to pull
        else reservoir_abstract_call(c, node);
        push = false;
        break;
      case WSI_OP_CATCHMENT_ROUTE: catchment_route_pre(c, node, v); filter = WSI_FILTER_NRW; break;
      case WSI_OP_RIVER_DISTRIBUTE: river_distribute_pre(c, node, v); filter = WSI_FILTER_NRW; break;
      case WSI_OP_STORAGE_DISTRIBUTE: storage_distribute_pre(c, node, v); filter = WSI_FILTER_ALL; break;
      case WSI_OP_BOUNDARY_RELEASE: boundary_release(c, node, true); push = false; break;
      case WSI_OP_FWTW_TREAT:
        if constexpr (!C::kLean) fwtw_treat_water(c, node);
        push = false;
        break;
      case WSI_OP_DEMAND_CREATE:
        if constexpr (!C::kLean) demand_create(c, node);
        push = false;
        break;
      case WSI_OP_WWTW_CALC:
        if constexpr (!C::kLean) wwtw_calculate_discharge(c, node);
        push = false;
        break;
      case WSI_OP_WWTW_DISCHARGE:
        if constexpr (!C::kLean) wwtw_make_discharge(c, node);
        push = false;
        break;
      case WSI_OP_LAND_IRRIGATE:
        if constexpr (!C::kLean) land_irrigate(c, node);
        push = false;
        break;
      case WSI_OP_RESERVOIR_ENVIRONMENTAL:
        if constexpr (!C::kLean) rres_satisfy_environmental(c, node);
        push = false;
        break;
      case WSI_OP_QGW_DISTRIBUTE:
        if constexpr (!C::kLean) {
          qgw_distribute_pre(c, node, v);
          filter = WSI_FILTER_ALL;
        } else {
          push = false;
        }
        break;
      default: push = false; break;
    }
    WSI_BLOCK_ALIGN(c);
    Vq reply;
    if (push) (void)push_out(c, node, v, filter, reply);
    else reply = vq_zero<C::P>();
    WSI_BLOCK_ALIGN(c);
    if (!push) continue;
    switch (op) {
      case WSI_OP_LAND_RUN: land_post(c, node, sub, reply, aux); break;
      case WSI_OP_SEWER_DISCHARGE: sewer_discharge_post(c, node, reply); break;
      case WSI_OP_QGW_DISTRIBUTE:
        if constexpr (!C::kLean) qgw_distribute_post(c, node, reply);
        break;
      case WSI_OP_CATCHMENT_ROUTE: catchment_route_post(c, node, reply); break;
      case WSI_OP_GW_INFILTRATE:
      case WSI_OP_GW_DISTRIBUTE:
      case WSI_OP_STORAGE_DISTRIBUTE:  // retained water goes back into the tank (storage.py:127,229,245)
        tank_push_force(c, c.tank(WSI_TANK_STORAGE(V), c.D().node_gidx[node]), c.s_tank(), reply);
        break;
      case WSI_OP_RIVER_DISTRIBUTE:  // storage.py:934-936
        tank_push_force(c, c.tank(WSI_TANK_STORAGE(V), (int)*c.river(WSI_RIVER_TANK, c.D().node_gidx[node])),
                        c.s_tank(), reply);
        break;
      default: break;
    }
  }
}
// out-of-line entry points of single ops for the network engine's scalar fall-backs (wsi_net.cuh);
// written out rather than routed through run_op, which would instantiate every other op for the
// network kernels' context as well
template <class C>
WSI_HDN void catchment_route(C& c, int node) {
  typename C::Vq v, reply;
  catchment_route_pre(c, node, v);
  (void)push_out(c, node, v, WSI_FILTER_NRW, reply);
  catchment_route_post(c, node, reply);
}
template <class C>
WSI_HDN void river_distribute(C& c, int node) {
  typename C::Vq v, reply;
  river_distribute_pre(c, node, v);
  (void)push_out(c, node, v, WSI_FILTER_NRW, reply);
  tank_push_force(c, c.tank(WSI_TANK_STORAGE(C::V), (int)*c.river(WSI_RIVER_TANK, c.D().node_gidx[node])), c.s_tank(),
                  reply);
}

template <class C>
WSI_HD void net_schedule(C& c);  // wsi_net.cuh

// ---- the whole timestep of one replica --------------------------------------------------------
// FUSED_SURFACES: this thread also runs the replica's surfaces (host build / tiny models);
// the CUDA library runs them in wsi_surface_kernel instead and instantiates <false>.
template <bool FUSED_SURFACES, class C>
WSI_HD void replica_step(C& c, int sched_begin, int sched_end) {
  const WsiModelDesc& d = c.D();
  // entries that reach into a Land's surfaces before Land.run (garden irrigation) come first
  const int split = (d.engine == 0 && d.sched_split > sched_begin && d.sched_split < sched_end) ? d.sched_split : sched_begin;
  if (c.run_flags & WSI_RUN_SCHEDULE)
    for (int k = sched_begin; k < split; k++) run_op(c, d.sched_op[k], d.sched_node[k]);
  if constexpr (FUSED_SURFACES) {
    if (c.run_flags & WSI_RUN_SURFACES)
      for (int u = 0; u < d.n_surf; u++) surface_phase<-1>(c, u);
  }
  if (c.run_flags & WSI_RUN_SCHEDULE) {
    if (d.engine == 1) net_schedule(c);
    else
      for (int k = split; k < sched_end; k++) run_op(c, d.sched_op[k], d.sched_node[k]);
  }
  if (c.run_flags & WSI_RUN_MASS_BALANCE) {
    MbHostSink sink;
    replica_mb_lanes<0>(c, true, c.mb_acc != nullptr, sink);
  }
  if (c.run_flags & WSI_RUN_RECORD) record(c);
  if (c.run_flags & WSI_RUN_END_TIMESTEP) end_timestep(c);
}

template <bool SMEM = false, bool LEAN = false, class CFG>
WSI_HD Ctx<CFG, SMEM, LEAN> make_ctx(const WsiModelDesc* host_d, int slot, const CFG& cfg, int r, int t, uint32_t flags) {
  Ctx<CFG, SMEM, LEAN> c;
  c.host_d = host_d;
  c.slot = slot;
  c.cfg = cfg;
  c.R = (size_t)c.D().R;
  c.r = r;
  c.t = t;
  c.FA = c.D().float_accuracy;
  c.run_flags = flags;
  c.viol = 0;
  c.tag = 0;
  c.block_sync = false;
  c.mb_acc = nullptr;
  c.mb_stride = 1;
  return c;
}

}  // namespace wsi
