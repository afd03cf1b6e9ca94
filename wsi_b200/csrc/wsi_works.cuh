// Pull side of the protocol and the nodes that use it: Node.pull_distributed (nodes.py:345-407),
// Arc.send_pull_request / send_pull_check (arcs.py:160-257), the pull handlers of junction Node,
// Catchment (catchment.py:106-138), Storage / Groundwater / Reservoir (storage.py:96-99), FWTW and
// WWTW (wtw.py), Reservoir.make_abstractions (storage.py:983-989), and the treatment works and
// residential demand themselves (wtw.py:14-657, demand.py:14-324).  Included by wsi_device.cuh.
//
// These nodes are few per model (one works, one demand node, a handful of abstraction points), so
// they are written in the general recursive form -- one call per junction level, VQIPs by value --
// and kept out of line; the schedule kernel's hot path (Land, tanks, the chain push) does not pay
// for them.  Pull requests and pull checks only ever carry a volume (arcs.py:176-186, nodes.py:389).
#pragma once

namespace wsi {

// ---- what a Catchment still has to give: get_avail catchment.py:90-104 -------------------------
template <class C>
WSI_HD typename C::Vq catchment_avail(C& c, int node) {
  const WsiModelDesc& d = c.D();
  typename C::Vq v = catch_get_flow(c, d.node_gidx[node]);
  for (int i = d.out_ptr[node]; i < d.out_ptr[node + 1]; i++)  // abstractions already made
    v = c.vchange(v, v.x[0] - arc_lane(c, d.out_arcs[i], 0));
  return v;
}

template <int D, class C>
WSI_HDN double arc_pull_check_vol(C& c, int a, bool has_v, double v_vol);
template <int D, class C>
WSI_HD typename C::Vq pull_distributed(C& c, int node, double volume, int of_type = 0);

// in-arcs a pull may use: every one (of_type None), or -- of_type ["River", "Node"] -- those that
// come from a River, then those from a junction Node (get_direction_arcs nodes.py:218-256).
// Returns the count; arcs[] holds at most kMaxCand of them.
constexpr int kPullAll = 0, kPullRiverNode = 1, kPullIrrigation = 2;
template <class C>
WSI_HD int pull_arcs(const C& c, int node, int of_type, int* arcs) {
  const WsiModelDesc& d = c.D();
  int n = 0;
  if (of_type == kPullAll) {
    for (int i = d.in_ptr[node]; i < d.in_ptr[node + 1]; i++, n++)
      if (n < kMaxCand) arcs[n] = d.in_arcs[i];
    return n;
  }
  if (of_type == kPullIrrigation) {  // ["River", "Node", "Groundwater", "Reservoir"] land.py:2364-2367, by class NAME
    for (int pass = 0; pass < 4; pass++)
      for (int i = d.in_ptr[node]; i < d.in_ptr[node + 1]; i++) {
        const int ty = d.node_type[d.arc_in[d.in_arcs[i]]];
        const bool hit = pass == 0   ? ty == WSI_NT_RIVER
                         : pass == 1 ? ty == WSI_NT_NODE
                         : pass == 2 ? (ty == WSI_NT_GROUNDWATER || ty == WSI_NT_QUEUEGROUNDWATER)
                                     : (ty == WSI_NT_RESERVOIR || ty == WSI_NT_RIVERRESERVOIR);
        if (hit) {
          if (n < kMaxCand) arcs[n] = d.in_arcs[i];
          n++;
        }
      }
    return n;
  }
  for (int pass = 0; pass < 2; pass++)
    for (int i = d.in_ptr[node]; i < d.in_ptr[node + 1]; i++)
      if (d.node_type[d.arc_in[d.in_arcs[i]]] == (pass == 0 ? WSI_NT_RIVER : WSI_NT_NODE)) {
        if (n < kMaxCand) arcs[n] = d.in_arcs[i];
        n++;
      }
  return n;
}

// River.pull_check_river storage.py:660-690: the tank plus what could be drawn through it from
// upstream, less the minimum required flow (a volume per timestep, hence / riverrc)
template <int D, class C>
WSI_HD double river_pull_check_vol(C& c, int node, bool has_v, double v_vol) {
  constexpr int V = C::V;
  const int g = c.D().node_gidx[node];
  double avail = *c.tank(WSI_TANK_STORAGE(V), (int)*c.river(WSI_RIVER_TANK, g));
  if constexpr (D > 0) {
    int arcs[kMaxCand];
    const int n = pull_arcs(c, node, kPullRiverNode, arcs);
    double upstream = 0.0;
#pragma unroll 1
    for (int i = 0; i < n && i < kMaxCand; i++) {
      double a = arc_pull_check_vol<D - 1>(c, arcs[i], false, 0.0);
      if (a < c.FA) a = 0.0;
      upstream += a;
    }
    avail += upstream;
  } else {
    c.flag(node, WSI_FLAG_DEPTH);
  }
  const double mrf = *c.river(WSI_RIVER_MRF, g) / *c.river(WSI_RIVER_RIVERRC, g);
  const double avail_vol = pymax(avail - mrf, 0.0);
  return vol_vchange(avail, has_v ? pymin(avail_vol, v_vol) : avail_vol);
}

template <class C>
WSI_HD typename C::Vq push_from(C& c, int node, const typename C::Vq& v, int filter, int tag);

// in_port.pull_check(vqip)["volume"]
template <int D, class C>
WSI_HD double node_pull_check_vol(C& c, int node, bool has_v, double v_vol) {
  constexpr int V = C::V;
  const WsiModelDesc& d = c.D();
  const int g = d.node_gidx[node];
  switch (d.node_type[node]) {
    case WSI_NT_NODE: {  // pull_check_basic nodes.py:510-527 -> check_basic :480-508
      if constexpr (D == 0) {
        c.flag(node, WSI_FLAG_DEPTH);
        return 0.0;
      } else {
        double avail = 0.0;
#pragma unroll 1
        for (int i = d.in_ptr[node]; i < d.in_ptr[node + 1]; i++)
          avail = avail + arc_pull_check_vol<D - 1>(c, d.in_arcs[i], false, 0.0);
        if (has_v) avail = vol_vchange(avail, pymin(avail, v_vol));
        return avail;
      }
    }
    case WSI_NT_DISTRIBUTION: {  // pull_check_basic under decorate_leakage_check distribution.py:63-98
      if constexpr (D == 0) {
        c.flag(node, WSI_FLAG_DEPTH);
        return 0.0;
      } else {
        const double leakage = *c.distrib(WSI_DISTRIB_LEAKAGE, g);
        const double want = (leakage > 0 && has_v) ? v_vol / (1 - leakage) : v_vol;
        double avail = 0.0;
#pragma unroll 1
        for (int i = d.in_ptr[node]; i < d.in_ptr[node + 1]; i++)
          avail = avail + arc_pull_check_vol<D - 1>(c, d.in_arcs[i], false, 0.0);
        if (has_v) avail = vol_vchange(avail, pymin(avail, want));
        if (leakage > 0) avail = avail - vol_vchange(avail, avail * leakage);
        return avail;
      }
    }
    case WSI_NT_CATCHMENT: {  // pull_check_abstraction catchment.py:106-121
      const double avail = catchment_avail(c, node).x[0];
      return has_v ? vol_vchange(avail, pymin(avail, v_vol)) : avail;
    }
    case WSI_NT_GROUNDWATER:
    case WSI_NT_STORAGE:
    case WSI_NT_RESERVOIR:
    case WSI_NT_RIVERRESERVOIR: {  // tank.get_avail tanks.py:102-130
      const double vol = *c.tank(WSI_TANK_STORAGE(V), g);
      return has_v ? vol_vchange(vol, pymin(vol, v_vol)) : vol;
    }
    case WSI_NT_FWTW: {  // pull_check_fwtw wtw.py:612-623: the service reservoirs
      const double vol = *c.tank(WSI_TANK_STORAGE(V), (int)*c.works(WSI_WORKS_TANK, g));
      return has_v ? vol_vchange(vol, pymin(vol, v_vol)) : vol;
    }
    case WSI_NT_WWTW:  // pull_check_reuse wtw.py:392-403: whatever has been treated
      return *c.works(WSI_WORKS_TREATED(V), g);
    case WSI_NT_UNLIMITEDDISTRIBUTION:  // distribution.py:166-168: always everything
      return kUnbounded;
    case WSI_NT_RIVER:
      return river_pull_check_vol<D>(c, node, has_v, v_vol);
    default:  // Land / Demand deny pulls (nodes.py:563-593); pulling through a Sewer is not on this path
      c.flag(node, WSI_FLAG_DENY);
      return 0.0;
  }
}

// ---- RiverReservoir (storage.py:992-1133): a Reservoir on a river -- what it cannot hold spills on
// downstream, and the spill counts towards its environmental flow ----------------------------------
// push_check_river_reservoir storage.py:1084-1107
template <int D, class C>
WSI_HDN double rres_push_check_vol(C c, int node, bool has_v, double v_vol) {
  constexpr int V = C::V;
  const WsiModelDesc& d = c.D();
  const int g = d.node_gidx[node];
  double downstream = 0.0;  // get_connected(direction="push", of_type=[Node, River, Waste])["avail"]
  if constexpr (D > 0) {
    const int b = d.cand_ptr[node * WSI_NF + WSI_FILTER_NRW], e = d.cand_ptr[node * WSI_NF + WSI_FILTER_NRW + 1];
#pragma unroll 1
    for (int i = b; i < e; i++) {
      double a = arc_excess_vol<D - 1>(c, d.cand_arcs[i], false, 0.0);
      if (a < c.FA) a = 0.0;
      downstream += a;
    }
  } else {
    c.flag(node, WSI_FLAG_DEPTH);
  }
  const double vol = *c.tank(WSI_TANK_STORAGE(V), g);
  const double excess = tank_excess_vol(*c.tank(WSI_TANK_CAPACITY, g), vol, false, 0.0);
  double new_v = excess + downstream;
  if (has_v) new_v = pymin(v_vol, new_v);
  return vol_vchange(excess, new_v);
}
// push_set_river_reservoir storage.py:1062-1082
template <int D, class C>
WSI_HDN typename C::Vq rres_push_set(C c, int node, typename C::Vq vqip) {
  constexpr int V = C::V;
  using Vq = typename C::Vq;
  const int g = c.D().node_gidx[node];
  double* st = c.tank(WSI_TANK_STORAGE(V), g);
  tank_push_force(c, st, c.s_tank(), vqip);
  const Vq spill = tank_pull(c, st, c.s_tank(), pymax(*st - *c.tank(WSI_TANK_CAPACITY, g), 0.0));  // pull_ponded tanks.py:78-100
  Vq reply = spill;
  if constexpr (D > 0) {
    const int tag = c.tag;
    c.tag = 0;
    reply = push_distributed<D - 1>(c, node, spill, WSI_FILTER_NRW);
    c.tag = tag;
  } else {
    c.flag(node, WSI_FLAG_DEPTH);
  }
  double* sat = c.tank(WSI_TANK_ENV_SATISFIED, g);
  *sat = *sat + (spill.x[0] - reply.x[0]);
  return reply;
}
// satisfy_environmental storage.py:1109-1127
template <class C>
WSI_HDN void rres_satisfy_environmental(C c, int node) {
  constexpr int V = C::V;
  using Vq = typename C::Vq;
  const int g = c.D().node_gidx[node];
  double* st = c.tank(WSI_TANK_STORAGE(V), g);
  double* sat = c.tank(WSI_TANK_ENV_SATISFIED, g);
  const double to_satisfy = pymax(*c.tank(WSI_TANK_ENVIRONMENTAL_FLOW, g) - *sat, 0.0);
  const Vq environmental = tank_pull(c, st, c.s_tank(), to_satisfy);
  const Vq reply = push_distributed<kJunctionDepth>(c, node, environmental, WSI_FILTER_ALL);
  tank_push_force(c, st, c.s_tank(), reply);
  if (reply.x[0] > c.FA) c.flag(node, WSI_FLAG_DEFICIT);  // "warning: environmental not able to push"
  *sat = *sat + environmental.x[0];
}

// Arc.get_excess(direction="pull")["volume"] arcs.py:230-257
template <int D, class C>
WSI_HDN double arc_pull_check_vol(C& c, int a, bool has_v, double v_vol) {
  if (c.D().arc_flags[a] & WSI_ARC_DENY_PULL) return 0.0;  // PushArc.send_pull_check_deny arcs.py:881-893
  const double pipe_excess = *c.arc(WSI_ARC_CAPACITY, a) - arc_flow_in(c, a);
  const double ne = node_pull_check_vol<D>(c, c.D().arc_in[a], has_v, v_vol);
  return vol_vchange(ne, pymin(pipe_excess, ne));
}

// in_port.pull_set({"volume": volume})
template <int D, class C>
WSI_HD typename C::Vq node_pull_set(C& c, int node, double volume) {
  using Vq = typename C::Vq;
  constexpr int V = C::V;
  const WsiModelDesc& d = c.D();
  const int g = d.node_gidx[node];
  switch (d.node_type[node]) {
    case WSI_NT_NODE:  // a junction passes the request upstream (nodes.py:61)
      if constexpr (D == 0) {
        c.flag(node, WSI_FLAG_DEPTH);
        return vq_zero<C::P>();
      } else {
        return pull_distributed<D - 1>(c, node, volume);
      }
    case WSI_NT_DISTRIBUTION:  // Node.pull_distributed under decorate_leakage_set distribution.py:14-60
      if constexpr (D == 0) {
        c.flag(node, WSI_FLAG_DEPTH);
        return vq_zero<C::P>();
      } else {
        const double leakage = *c.distrib(WSI_DISTRIB_LEAKAGE, g);
        if (!(leakage > 0)) return pull_distributed<D - 1>(c, node, volume);
        Vq reply = pull_distributed<D - 1>(c, node, volume / (1 - leakage));
        const Vq leaked = c.vchange(reply, reply.x[0] * leakage);
        reply = c.extract(reply, leaked);
        const Vq unsuccessful = push_from(c, node, leaked, WSI_FILTER_GROUNDWATER, 0);
        if (unsuccessful.x[0] > c.FA) {  // "warning, distribution leakage not going to GW"
          c.flag(node, WSI_FLAG_DEFICIT);
          reply = c.sum(reply, unsuccessful);
        }
        return reply;
      }
    case WSI_NT_CATCHMENT: {  // pull_set_abstraction catchment.py:123-138 (the arc keeps the books)
      const Vq avail = catchment_avail(c, node);
      return c.vchange(avail, pymin(avail.x[0], volume));
    }
    case WSI_NT_GROUNDWATER:
    case WSI_NT_STORAGE:
    case WSI_NT_RESERVOIR:
    case WSI_NT_RIVERRESERVOIR:  // tank.pull_storage storage.py:98
      return tank_pull(c, c.tank(WSI_TANK_STORAGE(V), g), c.s_tank(), volume);
    case WSI_NT_FWTW: {  // pull_set_fwtw wtw.py:625-639
      const Vq pulled = tank_pull(c, c.tank(WSI_TANK_STORAGE(V), (int)*c.works(WSI_WORKS_TANK, g)), c.s_tank(), volume);
      double* tp = c.works(WSI_WORKS_TOTAL_PULLED(V), g);
      c.stv(tp, c.s_works(), c.sum(c.ldv(tp, c.s_works()), pulled));
      return pulled;
    }
    case WSI_NT_RIVER: {  // pull_set_river storage.py:692-709: first the tank, the rest from upstream
      const double avail = river_pull_check_vol<D>(c, node, true, volume);
      const Vq pulled = tank_pull(c, c.tank(WSI_TANK_STORAGE(V), (int)*c.river(WSI_RIVER_TANK, g)), c.s_tank(), avail);
      if constexpr (D == 0) {
        return pulled;
      } else {
        const Vq more = pull_distributed<D - 1>(c, node, avail - pulled.x[0], kPullRiverNode);
        return c.sum(pulled, more);
      }
    }
    case WSI_NT_UNLIMITEDDISTRIBUTION: {  // pull_set_unlimited distribution.py:173-189
      const Vq vq = vq_volume<C::P>(volume);
      double* sp = c.distrib(WSI_DISTRIB_SUPPLIED(V), g);
      c.stv(sp, c.s_distrib(), c.sum(c.ldv(sp, c.s_distrib()), vq));
      return vq;
    }
    case WSI_NT_WWTW: {  // pull_set_reuse wtw.py:376-390
      double* tr = c.works(WSI_WORKS_TREATED(V), g);
      const Vq treated = c.ldv(tr, c.s_works());
      const double reply_vol = pymin(volume, treated.x[0]);
      const Vq reply = c.vchange(treated, reply_vol);
      c.stv(tr, c.s_works(), c.vchange(treated, treated.x[0] - reply_vol));
      return reply;
    }
    default:
      c.flag(node, WSI_FLAG_DENY);
      return vq_zero<C::P>();
  }
}

// Arc.send_pull_request({"volume": volume}) arcs.py:160-197
template <int D, class C>
WSI_HDN typename C::Vq arc_send_pull_request(C& c, int a, double volume) {
  if (c.D().arc_flags[a] & WSI_ARC_DENY_PULL) return vq_zero<C::P>();  // PushArc.send_pull_deny arcs.py:865-879
  const double excess_in = arc_pull_check_vol<D>(c, a, true, volume);
  const double not_pulled = pymax(volume - excess_in, 0.0);
  volume -= not_pulled;
  const typename C::Vq vq = node_pull_set<D>(c, c.D().arc_in[a], volume);
  arc_book(c, a, vq);
  return vq;
}

// Node.pull_distributed({"volume": volume}, of_type) nodes.py:345-407 (default tag)
template <int D, class C>
WSI_HDN typename C::Vq pull_distributed_call(C& c, int node, double volume, int of_type) {
  using Vq = typename C::Vq;
  const WsiModelDesc& d = c.D();
  const int ib = d.in_ptr[node];
  int arcs[kMaxCand];
  const int n_in = pull_arcs(c, node, of_type, arcs);
  if (d.in_ptr[node + 1] - ib == 1)  // one in-arc: pull from it if its type is wanted (nodes.py:360-375)
    return n_in == 1 ? arc_send_pull_request<D>(c, d.in_arcs[ib], volume) : vq_zero<C::P>();
  Vq pulled = vq_zero<C::P>();
  if (n_in > kMaxCand) {  // refused by the flattener; never silently truncated
    c.flag(node, WSI_FLAG_DEPTH);
    return pulled;
  }
  double alloc[kMaxCand];
  double deficit = volume;
  int iter = 0;
#pragma unroll 1
  for (;;) {
    // get_connected(direction="pull") nodes.py:271-317
    double avail = 0.0, priority = 0.0;
#pragma unroll 1
    for (int i = 0; i < n_in; i++) {
      double a = arc_pull_check_vol<D>(c, arcs[i], false, 0.0);
      if (a < c.FA) a = 0.0;
      const double pref = *c.arc(WSI_ARC_PREFERENCE, arcs[i]);
      avail += a;
      priority += a * pref;
      alloc[i] = a * pref;
    }
    if (!(deficit > c.FA && avail > c.FA && iter < kMaxIter)) break;
#pragma unroll 1
    for (int i = 0; i < n_in; i++) {
      const Vq received = arc_send_pull_request<D>(c, arcs[i], deficit * alloc[i] / priority);
      pulled = c.sum(pulled, received);
    }
    deficit = volume - pulled.x[0];
    iter++;
  }
  if (iter == kMaxIter) c.flag(node, WSI_FLAG_MAXITER);
  return pulled;
}
template <int D, class C>
WSI_HD typename C::Vq pull_distributed(C& c, int node, double volume, int of_type) {
  return pull_distributed_call<D>(c, node, volume, of_type);
}

// a push made from inside one of the out-of-line ops below
template <class C>
WSI_HD typename C::Vq push_from(C& c, int node, const typename C::Vq& v, int filter, int tag) {
  c.tag = tag;
  const typename C::Vq reply = push_distributed<kJunctionDepth>(c, node, v, filter);
  c.tag = 0;
  return reply;
}

// ---- Reservoir.make_abstractions storage.py:983-989 ------------------------------------------
template <class C>
WSI_HD void reservoir_abstract(C& c, int node) {
  constexpr int V = C::V;
  const int g = c.D().node_gidx[node];
  double* st = c.tank(WSI_TANK_STORAGE(V), g);
  const double cap = *c.tank(WSI_TANK_CAPACITY, g);
  const typename C::Vq reply = pull_distributed<kJunctionDepth>(c, node, tank_excess_vol(cap, st[0], false, 0.0));
  const typename C::Vq spill = tank_push(c, st, c.s_tank(), cap, reply);
  tank_push_force(c, st, c.s_tank(), spill);
}

// ---- Sewer.push_set_land sewer.py:178-199: a push spread over the pipe time-area diagram -----
template <class C>
WSI_HDN typename C::Vq sewer_push_timearea(C c, int w, typename C::Vq vqip) {
  typename C::Vq reply = vq_zero<C::P>();
  const int n = (int)*c.sewer(WSI_SEWER_N_TIMEAREA, w);
#pragma unroll 1
  for (int i = 0; i < n; i++) {
    const double frac = *c.sewer(WSI_SEWER_TA_FRAC + i, w);
    const typename C::Vq part = c.vchange(vqip, vqip.x[0] * frac);
    reply = c.sum(reply, sewer_push_storage(c, w, part, (int)*c.sewer(WSI_SEWER_TA_TIME + i, w)));
  }
  return reply;
}

// ---- a Land's GardenSurface answering the ("Demand", "Garden") tag (land.py:2402-2437) ----------
template <class C>
WSI_HDN double garden_irrigation_demand(C c, int land) {  // calculate_irrigation_demand
  const int u = (int)*c.land(WSI_LAND_GARDEN_SURF, land);
  if (u < 0) return 0.0;  // the Land's default handler: an empty VQIP (land.py:93)
  double demand = pymax(*c.surf(WSI_SURF_OUT_EVAPORATION, u) - *c.surf(WSI_SURF_OUT_PRECIPITATION, u), 0.0);
  const double cmd = tank_excess_vol(*c.surf(WSI_SURF_CAPACITY, u), *c.surf(WSI_SURF_STORAGE(C::V), u), false, 0.0) /
                     *c.surf(WSI_SURF_AREA, u);  // get_cmd land.py:886
  if (cmd <= c.FA) demand = 0.0;
  return demand;
}
template <class C>
WSI_HDN void garden_receive_irrigation(C c, int land, typename C::Vq vqip) {  // push_storage(force=True)
  const int u = (int)*c.land(WSI_LAND_GARDEN_SURF, land);
  if (u < 0) return;
  double* st = c.surf(WSI_SURF_STORAGE(C::V), u);
  // the water arrives before the surface runs this step; what the tank held until now is its
  // storage_ for this step's ds() (tanks.py:70-76): keep it for surface_phase
  if (*c.surf(WSI_SURF_IRRIGATED_AT, u) != (double)c.t) {
    c.stv_add(c.surf(WSI_SURF_PRE_IRRIGATION(C::V), u), c.s_surf(), c.ldv_add(st, c.s_surf()));
    *c.surf(WSI_SURF_IRRIGATED_AT, u) = (double)c.t;
  }
  tank_push_force(c, st, c.s_surf(), vqip);
}

// Land.apply_irrigation land.py:192-198 -> IrrigationSurface.irrigate land.py:2347-2381: what the crop
// evaporated beyond the rain, pulled through the Land's in-arcs, into the soil tank and its nutrient pools
template <class C>
WSI_HDN void land_irrigate(C c, int node) {
  using Vq = typename C::Vq;
  constexpr int P = C::P;
  constexpr int V = C::V;
  const WsiModelDesc& d = c.D();
  const int L = d.node_gidx[node];
  const auto& cfg = c.cfg;
#pragma unroll 1
  for (int i = d.land_surf_ptr[L]; i < d.land_surf_ptr[L + 1]; i++) {
    const int u = d.land_surf[i];
    const double coef = *c.surf(WSI_SURF_IRRIGATION_COEFFICIENT, u);
    if (coef < 0.0) continue;
    if (!(*c.surf(WSI_SURF_DAYS_AFTER_SOW, u) > 0.0)) continue;  // `if self.days_after_sow:` (None and 0 are falsy)
    double demand = pymax(*c.surf(WSI_SURF_OUT_EVAPORATION, u) - *c.surf(WSI_SURF_OUT_PRECIPITATION, u), 0.0) * coef;
    if (!(demand > c.FA)) continue;
    double* st = c.surf(WSI_SURF_STORAGE(V), u);
    const double cmd = tank_excess_vol(*c.surf(WSI_SURF_CAPACITY, u), *st, false, 0.0) / *c.surf(WSI_SURF_AREA, u);
    if (cmd <= c.FA) demand = 0.0;
    const Vq supplied = pull_distributed_call<kJunctionDepth>(c, node, demand, kPullIrrigation);
    tank_push_force(c, st, c.s_surf(), supplied);
    // allocate_organic_irrigation / allocate_inorganic_irrigation nutrient_pool.py:192-224
    if (cfg.orgn() >= 0) {
      double* org = c.surf(WSI_SURF_POOL_DIS_ORG, u);
      double* ino = c.surf(WSI_SURF_POOL_DIS_INORG, u);
      const size_t ss = c.s_surf();
      org[0] = org[0] + vget<P>(supplied, cfg.orgn());
      org[ss] = org[ss] + vget<P>(supplied, cfg.orgp());
      ino[0] = ino[0] + (vget<P>(supplied, cfg.nh4()) + vget<P>(supplied, cfg.no3()));
      ino[ss] = ino[ss] + vget<P>(supplied, cfg.po4());
    }
    if (c.run_flags & WSI_RUN_MASS_BALANCE) {  // decay_ds tanks.py:449-457 again, on the watered tank
      const Obj S = c.surf_obj(u);
      const Vq now = c.ldv_add(st, c.s_surf());
      const Vq old = c.ldv_add(S(WSI_SURF_PRE_IRRIGATION(V)), c.s_surf());
      Vq dsv;
#pragma unroll
      for (int p = 0; p <= P; p++) dsv.x[p] = (p == 0 || bit(c.addm(), p)) ? now.x[p] - old.x[p] : 0.0;
      dsv = c.sum(dsv, decay_total<C>(c, S, WSI_SURF_N_DECAY));
      c.stv_add(S(WSI_SURF_MB_DS(V)), c.s_surf(), dsv);
    }
  }
}

// ---- WTW.treat_current_input wtw.py:118-160 ---------------------------------------------------
template <class C>
WSI_HD void works_treat_current_input(C& c, int node, int g) {
  using Vq = typename C::Vq;
  constexpr int P = C::P;
  constexpr int V = C::V;
  const size_t s = c.s_works();
  const Vq influent = c.ldv(c.works(WSI_WORKS_CURRENT_INPUT(V), g), s);
  Vq liquor = c.ldv(c.works(WSI_WORKS_LIQUOR(V), g), s);
  Vq solids = c.ldv(c.works(WSI_WORKS_SOLIDS(V), g), s);
  const double* lm = c.works(WSI_WORKS_LIQUOR_MULT(V), g);
  const double* pc = c.works(WSI_WORKS_PROC_CONST(V), g);
  const double* pe = c.works(WSI_WORKS_PROC_EXP(V), g);
  const double lm_vol = lm[0];
  Vq out = vq_zero<P>();  // discharge_holder
  const double den = liquor.x[0] + influent.x[0] * lm_vol;
  if (den == 0) c.flag(node, WSI_FLAG_NAN);  // the reference raises ZeroDivisionError (wtw.py:127-133)
#pragma unroll
  for (int p = 1; p <= P; p++)
    if (!bit(c.addm(), p)) {
      out.x[p] = influent.x[p];
      liquor.x[p] = (liquor.x[p] * liquor.x[0] + influent.x[p] * influent.x[0] * lm_vol) / den;
    }
  const double T = vget<P>(influent, c.cfg.temp());
#pragma unroll
  for (int p = 1; p <= P; p++)
    if (bit(c.addm(), p)) {
      const double temp_factor = pow(pe[(size_t)p * s], 20.0 - T);  // constants.DECAY_REFERENCE_TEMPERATURE
      out.x[p] = influent.x[p] * pc[(size_t)p * s] * temp_factor;
      liquor.x[p] = influent.x[p] * lm[(size_t)p * s];
    }
  out.x[0] = influent.x[0] * pc[0] * 1.0;
  liquor.x[0] = influent.x[0] * lm_vol;
  solids.x[0] = influent.x[0] * *c.works(WSI_WORKS_PERCENT_SOLIDS, g);
#pragma unroll
  for (int p = 1; p <= P; p++)
    if (bit(c.addm(), p)) solids.x[p] = influent.x[p] - out.x[p] - liquor.x[p];
  double* tr = c.works(WSI_WORKS_TREATED(V), g);
  c.stv(tr, s, c.sum(c.ldv(tr, s), out));
  c.stv(c.works(WSI_WORKS_LIQUOR(V), g), s, liquor);
  c.stv(c.works(WSI_WORKS_SOLIDS(V), g), s, solids);
}

// ---- WWTW as a push receiver: push_check_sewer / push_set_sewer wtw.py:319-374 ------------------
template <class C>
WSI_HD double works_excess_throughput(C& c, int g) {  // wtw.py:109-116
  return pymax(*c.works(WSI_WORKS_THROUGHPUT_CAPACITY, g) - *c.works(WSI_WORKS_CURRENT_INPUT(C::V), g), 0.0);
}
template <class C>
WSI_HDN double wwtw_push_check_vol(C c, int g, bool has_v, double v_vol) {
  const int tk = (int)*c.works(WSI_WORKS_TANK, g);
  const double excess_tank =
      tank_excess_vol(*c.tank(WSI_TANK_CAPACITY, tk), *c.tank(WSI_TANK_STORAGE(C::V), tk), false, 0.0);
  double vol = excess_tank + works_excess_throughput(c, g);
  if (!has_v) return vol;  // v_change_vqip(empty_vqip(), vol)
  vol = pymin(vol, v_vol);
  return vol_vchange(v_vol, vol);
}
template <class C>
WSI_HDN typename C::Vq wwtw_push_set(C c, int g, typename C::Vq vqip0) {
  using Vq = typename C::Vq;
  constexpr int V = C::V;
  const int tk = (int)*c.works(WSI_WORKS_TANK, g);
  const Vq sent_direct = c.vchange(vqip0, pymin(works_excess_throughput(c, g), vqip0.x[0]));
  double* ci = c.works(WSI_WORKS_CURRENT_INPUT(V), g);
  c.stv(ci, c.s_works(), c.sum(c.ldv(ci, c.s_works()), sent_direct));
  if (sent_direct.x[0] == vqip0.x[0]) return vq_zero<C::P>();
  const Vq rest = c.vchange(vqip0, vqip0.x[0] - sent_direct.x[0]);
  const Vq back = tank_push(c, c.tank(WSI_TANK_STORAGE(V), tk), c.s_tank(), *c.tank(WSI_TANK_CAPACITY, tk), rest);
  if (back.x[0] < c.FA) return vq_zero<C::P>();
  return back;
}

// ---- WWTW.calculate_discharge wtw.py:278-304, make_discharge :306-317 ---------------------------
template <class C>
WSI_HDN void wwtw_calculate_discharge(C c, int node) {
  using Vq = typename C::Vq;
  constexpr int V = C::V;
  const int g = c.D().node_gidx[node];
  const int tk = (int)*c.works(WSI_WORKS_TANK, g);
  const size_t s = c.s_works();
  double* ci = c.works(WSI_WORKS_CURRENT_INPUT(V), g);
  double* st = c.tank(WSI_TANK_STORAGE(V), tk);
  const double excess = works_excess_throughput(c, g);
  if (st[0] > c.FA && excess > c.FA) {
    // to_pull = v_change_vqip(storage, min(excess, avail)); pull_storage reads its volume (wtw.py:290-296)
    const Vq cleared = tank_pull(c, st, c.s_tank(), vol_vchange(st[0], pymin(excess, st[0])));
    c.stv(ci, s, c.sum(c.ldv(ci, s), cleared));
  }
  c.stv(ci, s, c.sum(c.ldv(ci, s), c.ldv(c.works(WSI_WORKS_LIQUOR(V), g), s)));
  works_treat_current_input(c, node, g);
}
template <class C>
WSI_HDN void wwtw_make_discharge(C c, int node) {
  using Vq = typename C::Vq;
  constexpr int V = C::V;
  const int g = c.D().node_gidx[node];
  double* tr = c.works(WSI_WORKS_TREATED(V), g);
  const Vq reply = push_from(c, node, c.ldv(tr, c.s_works()), WSI_FILTER_ALL, 0);
  c.stv(tr, c.s_works(), vq_zero<C::P>());
  if (reply.x[0] > c.FA)
    tank_push_force(c, c.tank(WSI_TANK_STORAGE(V), (int)*c.works(WSI_WORKS_TANK, g)), c.s_tank(), reply);
}

// ---- FWTW.treat_water wtw.py:559-610 --------------------------------------------------------------
template <class C>
WSI_HDN void fwtw_treat_water(C c, int node) {
  using Vq = typename C::Vq;
  constexpr int V = C::V;
  const int g = c.D().node_gidx[node];
  const int tk = (int)*c.works(WSI_WORKS_TANK, g);
  const size_t s = c.s_works();
  double* st = c.tank(WSI_TANK_STORAGE(V), tk);
  const double cap = *c.tank(WSI_TANK_CAPACITY, tk);
  const double target =
      pymin(tank_excess_vol(cap, st[0], false, 0.0), *c.works(WSI_WORKS_THROUGHPUT_CAPACITY, g));
  const Vq throughput = pull_distributed<kJunctionDepth>(c, node, target);
  const Vq deficit = c.vchange(c.ldv(c.works(WSI_WORKS_PREVIOUS_PULLED(V), g), s), pymax(target - throughput.x[0], 0.0));
  c.stv(c.works(WSI_WORKS_CURRENT_INPUT(V), g), s, c.sum(throughput, deficit));
  double* td = c.works(WSI_WORKS_TOTAL_DEFICIT(V), g);
  const Vq total_deficit = c.sum(c.ldv(td, s), deficit);
  c.stv(td, s, total_deficit);
  if (total_deficit.x[0] > c.FA) c.flag(node, WSI_FLAG_DEFICIT);  // "Service reservoirs not filled"
  works_treat_current_input(c, node, g);
  // liquor and solids go back to the sewers
  const Vq push_back = c.sum(c.ldv(c.works(WSI_WORKS_LIQUOR(V), g), s), c.ldv(c.works(WSI_WORKS_SOLIDS(V), g), s));
  const Vq rejected = push_from(c, node, push_back, WSI_FILTER_SEWER, 0);
  double* us = c.works(WSI_WORKS_UNPUSHED_SLUDGE(V), g);
  c.stv(us, s, c.sum(c.ldv(us, s), rejected));
  // treated water into the service reservoirs
  const Vq excess = tank_push(c, st, c.s_tank(), cap, c.ldv(c.works(WSI_WORKS_TREATED(V), g), s));
  tank_push_force(c, st, c.s_tank(), excess);
}

// ---- Demand.create_demand demand.py:79-119 with ResidentialDemand.get_demand :236-324 ----------
template <class C>
WSI_HDN void demand_create(C c, int node) {
  using Vq = typename C::Vq;
  constexpr int P = C::P;
  constexpr int V = C::V;
  const int g = c.D().node_gidx[node];
  const size_t s = c.s_demand();
  // garden (get_garden_demand demand.py:250-272): what the attached Lands' gardens ask for through the
  // ("Demand", "Garden") push check, times the gardening efficiency
  const WsiModelDesc& d = c.D();
  double excess = 0.0;
  c.tag = 2;
  for (int i = d.cand_ptr[node * WSI_NF + WSI_FILTER_LAND]; i < d.cand_ptr[node * WSI_NF + WSI_FILTER_LAND + 1]; i++) {
    double a = arc_excess_vol<kJunctionDepth>(c, d.cand_arcs[i], false, 0.0);
    if (a < c.FA) a = 0.0;
    excess += a;
  }
  c.tag = 0;
  const double population = *c.demand(WSI_DEMAND_POPULATION, g);
  if (population < 0) {  // base Demand (demand.py:14-131): {"default": constant_demand carrying pollutant_load}
    Vq pol = c.ldv(c.demand(WSI_DEMAND_POLLUTANT_LOAD(V), g), s);
    pol.x[0] = *c.demand(WSI_DEMAND_PER_CAPITA, g);
    const double requested = 0.0 + pol.x[0];
    const Vq got = pull_distributed<kJunctionDepth>(c, node, requested);
    c.stv(c.demand(WSI_DEMAND_TOTAL_RECEIVED(V), g), s, got);
    if (requested - got.x[0] > c.FA) c.flag(node, WSI_FLAG_DEFICIT);  // "demand deficit"
    double* tb0 = c.demand(WSI_DEMAND_TOTAL_BACKUP(V), g);
    c.stv(tb0, s, c.sum(c.ldv(tb0, s), push_from(c, node, pol, WSI_FILTER_ALL, 0)));
    double* td0 = c.demand(WSI_DEMAND_TOTAL_DEMAND(V), g);
    c.stv(td0, s, c.sum(c.ldv(td0, s), pol));
    return;
  }
  Vq garden = vq_zero<P>();
  garden.x[0] = excess * *c.demand(WSI_DEMAND_GARDENING_EFFICIENCY, g);
  // house: per-capita consumption carrying the per-capita loads
  Vq foul = c.ldv(c.demand(WSI_DEMAND_POLLUTANT_LOAD(V), g), s);
#pragma unroll
  for (int p = 1; p <= P; p++)
    if (bit(c.addm(), p)) foul.x[p] *= population;
  foul.x[0] = population * *c.demand(WSI_DEMAND_PER_CAPITA, g);
  const double w = *c.demand(WSI_DEMAND_CONSTANT_WEIGHTING, g);
  vset<P>(foul, c.cfg.temp(),
          c.forcing((int)*c.demand(WSI_DEMAND_FORCING_ID, g), 2) * (1 - w) + *c.demand(WSI_DEMAND_CONSTANT_TEMP, g) * w);
  const double total_requested = 0.0 + garden.x[0] + foul.x[0];
  const Vq received = pull_distributed<kJunctionDepth>(c, node, total_requested);
  c.stv(c.demand(WSI_DEMAND_TOTAL_RECEIVED(V), g), s, received);
  if (total_requested - received.x[0] > c.FA) c.flag(node, WSI_FLAG_DEFICIT);  // "demand deficit"
  double* tb = c.demand(WSI_DEMAND_TOTAL_BACKUP(V), g);
  Vq backup = c.ldv(tb, s);
  backup = c.sum(backup, push_from(c, node, garden, WSI_FILTER_LAND, 2));
  backup = c.sum(backup, push_from(c, node, foul, WSI_FILTER_SEWER, 1));
  c.stv(tb, s, backup);
  double* tdm = c.demand(WSI_DEMAND_TOTAL_DEMAND(V), g);
  c.stv(tdm, s, c.sum(c.sum(c.ldv(tdm, s), garden), foul));
}

// ---- end_timestep of works and demand nodes (wtw.py:404-420, 641-651; demand.py:133-137) -------
template <class C>
WSI_HD void works_end_timestep(C& c, int g) {
  constexpr int V = C::V;
  const size_t s = c.s_works();
  const typename C::Vq zero = vq_zero<C::P>();
  if ((int)*c.works(WSI_WORKS_KIND, g) == 0) {  // WWTW (treated is reset by make_discharge)
    c.stv(c.works(WSI_WORKS_LIQUOR_LAG(V), g), s, c.ldv(c.works(WSI_WORKS_LIQUOR(V), g), s));
    c.stv(c.works(WSI_WORKS_CURRENT_INPUT(V), g), s, zero);
    c.stv(c.works(WSI_WORKS_SOLIDS(V), g), s, zero);
  } else {  // FWTW
    c.stv(c.works(WSI_WORKS_TOTAL_DEFICIT(V), g), s, zero);
    c.stv(c.works(WSI_WORKS_PREVIOUS_PULLED(V), g), s, c.ldv(c.works(WSI_WORKS_TOTAL_PULLED(V), g), s));
    c.stv(c.works(WSI_WORKS_TOTAL_PULLED(V), g), s, zero);
    c.stv(c.works(WSI_WORKS_TREATED(V), g), s, zero);
    c.stv(c.works(WSI_WORKS_UNPUSHED_SLUDGE(V), g), s, zero);
  }
}
template <class C>
WSI_HD void demand_end_timestep(C& c, int g) {
  constexpr int V = C::V;
  const typename C::Vq zero = vq_zero<C::P>();
  c.stv(c.demand(WSI_DEMAND_TOTAL_DEMAND(V), g), c.s_demand(), zero);
  c.stv(c.demand(WSI_DEMAND_TOTAL_BACKUP(V), g), c.s_demand(), zero);
  c.stv(c.demand(WSI_DEMAND_TOTAL_RECEIVED(V), g), c.s_demand(), zero);
}

// ---- mass-balance terms beyond the arcs (wtw.py:203-208, 516-518; demand.py:62-64), one lane ----
// returns false for node types that have none
template <class C>
WSI_HD bool works_mb_lane(const C& c, int type, int g, int lane, double& in_, double& ds, double& out_) {
  constexpr int V = C::V;
  if (type == WSI_NT_WWTW) {
    const int tk = (int)*c.works(WSI_WORKS_TANK, g);
    out_ += *c.works(WSI_WORKS_SOLIDS(V) + lane, g);
    ds += *c.tank(WSI_TANK_STORAGE(V) + lane, tk) - *c.tank(WSI_TANK_STORAGE_LAG(V) + lane, tk);
    ds += *c.works(WSI_WORKS_LIQUOR(V) + lane, g) - *c.works(WSI_WORKS_LIQUOR_LAG(V) + lane, g);
    return true;
  }
  if (type == WSI_NT_FWTW) {
    const int tk = (int)*c.works(WSI_WORKS_TANK, g);
    in_ += *c.works(WSI_WORKS_TOTAL_DEFICIT(V) + lane, g);
    ds += *c.tank(WSI_TANK_STORAGE(V) + lane, tk) - *c.tank(WSI_TANK_STORAGE_LAG(V) + lane, tk);
    out_ += *c.works(WSI_WORKS_UNPUSHED_SLUDGE(V) + lane, g);
    return true;
  }
  if (type == WSI_NT_UNLIMITEDDISTRIBUTION) {  // distribution.py:171
    in_ += *c.distrib(WSI_DISTRIB_SUPPLIED(V) + lane, g);
    return true;
  }
  if (type == WSI_NT_DEMAND) {
    in_ += *c.demand(WSI_DEMAND_TOTAL_DEMAND(V) + lane, g);
    out_ += *c.demand(WSI_DEMAND_TOTAL_BACKUP(V) + lane, g);
    out_ += *c.demand(WSI_DEMAND_TOTAL_RECEIVED(V) + lane, g);
    return true;
  }
  return false;
}

}  // namespace wsi
