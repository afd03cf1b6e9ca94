// Level-synchronous routing over a dendritic river network (one thread per node, launches per
// level) -- the execution strategy for a single large graph, where parallelism has to come from
// the nodes rather than from replicas.
//
// Reference semantics being reproduced: Catchment.route (catchment.py:77-88) and River.distribute
// (storage.py:926-936) executed in Model.river_discharge_order (model.py:720-735, 1006-1007),
// i.e. upstream before downstream; each push goes Arc.send_push_request (arcs.py:112-158) ->
// junction Nodes forward (nodes.py:62-66) -> River.push_set_river force-adds to the tank
// (storage.py:711-721) / Waste swallows (waste.py:40-49).
//
// Parallel formulation: a river's distribute is split in two.  At its own level the river pulls
// its outflow from the tank and parks it on its out-arc (`pending`).  The node at the end of the
// junction chain then *delivers* its incoming paths in a fixed order (catchments in orchestration
// order, then rivers in discharge order -- the order the reference pushes them) by running the
// very same push code as the sequential engine (push_distributed / catchment_route), and hands
// whatever was refused back to the sender's tank.  Every object on a path (the chain's arcs, the
// receiving tank) and the sender's tank at that point in time have exactly one writer thread, so
// there are no atomics, the result does not depend on scheduling and is bit-identical to the
// sequential order.  Levels are "upstream depth": a level only depends on lower ones.
//
// Restrictions checked on the host (backend/network.py): nodes are River / junction Node / Waste /
// Catchment, every sender and junction has exactly one out-arc (a push reaches one receiver),
// junction chains are at most kJunctionDepth long.
#pragma once

namespace wsi {

// Second half of River.distribute (storage.py:926-936) for the river `sender`, run by the thread
// of the node that receives it: the general case, through the sequential engine's push code.
template <class C>
WSI_HDN void net_deliver_river(C& c, int sender) {
  constexpr int V = C::V;
  const WsiModelDesc& d = c.D();
  const int a0 = d.out_arcs[d.out_ptr[sender]];
  const typename C::Vq outflow = c.ldv(c.arc(WSI_ARC_PENDING(V), a0), c.s_arc());
  const typename C::Vq reply = push_distributed<kJunctionDepth>(c, sender, outflow, WSI_FILTER_NRW);
  const int tk = (int)*c.river(WSI_RIVER_TANK, d.node_gidx[sender]);
  tank_push_force(c, c.tank(WSI_TANK_STORAGE(V), tk), c.s_tank(), reply);
}

// Before the levels, for every river: River.calculate_discharge, then the Catchment.route of the
// catchments that drain into it.  Both only involve this river's own objects, and the reference
// runs them in this order before any River.distribute (model.py:1000-1007), so they need not
// wait for upstream rivers.
template <class C>
WSI_HD void net_pre(C& c, int node) {
  const WsiModelDesc& d = c.D();
  river_calc(c, node);
#pragma unroll 1
  for (int p = d.net_path_ptr[node]; p < d.net_path_split[node]; p++) catchment_route(c, d.net_path_sender[p]);
}

// One river of a level: take in the upstream rivers' releases in discharge order, then release.
//
// River.distribute (storage.py:926-936) is  outflow = pull(tank); reply = push(outflow);
// tank.push(reply).  When the out-arc goes straight to a River / Waste and admits the whole
// release (pipe_excess >= volume > 0 -- the arc has no other sender, so the river can tell by
// itself), Arc.send_push_request (arcs.py:112-158) reduces to not_pushed == 0, sent == outflow,
// reply == empty (the identities in arc_send_push_request).  The river then books the release on
// its arc and re-blends the empty reply into its own tank right away (tank_push_force's empty
// case), marks the arc pending_mode 0, and the receiver only has to add `pending` to its tank.
// Otherwise (junction chain, binding capacity, dry release) pending_mode is 1 and the receiver's
// thread runs the whole push through the sequential engine's code (net_deliver_river).
//
// NetItem is the part of the work that does not depend on upstream rivers (indices, the tank as
// calculate_discharge and the catchments left it, the routing coefficient, the out-arc): the
// single-block tail kernel loads it one level ahead so that only  pending -> sum -> pull ->
// pending  is on the critical path of a level.
template <class C>
struct NetRec {  // stage 1 of the look-ahead: indices only
  int node, tk, g, out_arc, pb, pe, direct, first_arc;
};
template <class C>
struct NetItem {  // stage 2: everything that does not depend on upstream rivers
  NetRec<C> rec;
  typename C::Vq T, booked_in;  // the tank; the out-arc's vqip_in so far
  double rc, cap, flow_in;
};
template <class C>
WSI_HD NetRec<C> net_rec_load(const C& c, int i) {
  const WsiModelDesc& d = c.D();
  const int32_t* rec = d.net_item + (size_t)i * 8;
  NetRec<C> r;
  r.node = rec[0];
  r.tk = rec[1];
  r.g = rec[2];
  r.out_arc = rec[3];
  r.pb = rec[4];
  r.pe = rec[5];
  r.direct = rec[6];
  r.first_arc = rec[7];
  return r;
}
template <class C>
WSI_HD NetItem<C> net_item_load(const C& c, const NetRec<C>& rec) {
  constexpr int V = C::V;
  NetItem<C> it;
  it.rec = rec;
  it.T = c.ldv(c.tank(WSI_TANK_STORAGE(V), rec.tk), c.s_tank());
  it.rc = *c.river(WSI_RIVER_RIVERRC, rec.g);
  if (rec.out_arc >= 0) {
    const Obj A = c.arc_obj(rec.out_arc);
    it.cap = *A(WSI_ARC_CAPACITY);
    it.flow_in = *A(WSI_ARC_FLOW_IN);
    it.booked_in = c.ldv(A(WSI_ARC_VQIP_IN(V)), c.s_arc());
  } else {
    it.cap = it.flow_in = 0.0;
    it.booked_in = vq_zero<C::P>();
  }
  return it;
}
template <class C>
WSI_HD void net_level(C& c, NetItem<C>& it) {
  using Vq = typename C::Vq;
  constexpr int V = C::V;
  const WsiModelDesc& d = c.D();
  const NetRec<C>& rec = it.rec;
  double* st = c.tank(WSI_TANK_STORAGE(V), rec.tk);
  const size_t s_t = c.s_tank(), s_a = c.s_arc();
  Vq& T = it.T;
  if (rec.pb < rec.pe) {
    // the loads of path p + 1 are issued before path p is folded into the tank
    Obj A = c.arc_obj(rec.first_arc);
    Vq v = c.ldv(A(WSI_ARC_PENDING(V)), s_a);
    double mode = *A(WSI_ARC_PENDING_MODE);
    int sender = 0;
#pragma unroll 1
    for (int p = rec.pb; p < rec.pe; p++) {
      Vq v_next = v;
      double mode_next = 0.0;
      int sender_next = 0;
      if (mode != 0.0) sender = d.net_rpath[(size_t)p * 4 + 3];
      if (p + 1 < rec.pe) {
        const int32_t* rp = d.net_rpath + (size_t)(p + 1) * 4;
        const Obj An = c.arc_obj(rp[0]);
        v_next = c.ldv(An(WSI_ARC_PENDING(V)), s_a);
        mode_next = *An(WSI_ARC_PENDING_MODE);
      }
      if (mode == 0.0) {
        T = c.sum(T, v);
      } else {
        c.stv(st, s_t, T);
        net_deliver_river(c, sender);
        T = c.ldv(st, s_t);
      }
      v = v_next;
      mode = mode_next;
      (void)sender_next;
    }
  }
  if (rec.out_arc < 0) {  // no single receiver: the sequential code is already local
    c.stv(st, s_t, T);
    river_distribute(c, rec.node);
    return;
  }
  const Vq outflow = vq_pull(c, T, T.x[0] * it.rc);
  const Obj A = c.arc_obj(rec.out_arc);
  const bool booked = rec.direct && outflow.x[0] > 0 && it.cap - it.flow_in >= outflow.x[0];
  c.stv(A(WSI_ARC_PENDING(V)), s_a, outflow);  // what the next level waits for goes out first
  *A(WSI_ARC_PENDING_MODE) = booked ? 0.0 : 1.0;
  if (booked) {
    if (T.x[0] > 0) {  // tank.push_storage(empty reply, force=True)
#pragma unroll
      for (int q = 1; q <= C::P; q++)
        if (!bit(c.addm(), q)) {
          const double num = 0.0 * 0.0 + T.x[q] * T.x[0];
          T.x[q] = div_z(num, T.x[0]);
        }
    }
    *A(WSI_ARC_FLOW_IN) = it.flow_in + outflow.x[0];
    c.stv(A(WSI_ARC_VQIP_IN(V)), s_a, c.sum(it.booked_in, outflow));
  }
  c.stv(st, s_t, T);
}

// Nodes that only receive (Waste; rivers outside the discharge order): everything still
// undelivered, in the reference's order.
template <class C>
WSI_HD void net_final(C& c, int node) {
  const WsiModelDesc& d = c.D();
  const bool pre_done = d.node_type[node] == WSI_NT_RIVER;  // its catchments went in with net_pre
#pragma unroll 1
  for (int p = pre_done ? d.net_path_split[node] : d.net_path_ptr[node]; p < d.net_path_ptr[node + 1]; p++) {
    const int sender = d.net_path_sender[p];
    if (d.node_type[sender] == WSI_NT_CATCHMENT) {
      catchment_route(c, sender);
      continue;
    }
    const int a0 = d.out_arcs[d.out_ptr[sender]];
    if (*c.arc(WSI_ARC_PENDING_MODE, a0) != 0.0) {
      net_deliver_river(c, sender);
    } else if (pre_done) {  // already booked by the sender: only the tank is left (a Waste keeps nothing)
      const int tk = (int)*c.river(WSI_RIVER_TANK, d.node_gidx[node]);
      tank_push_force(c, c.tank(WSI_TANK_STORAGE(C::V), tk), c.s_tank(),
                      c.ldv(c.arc(WSI_ARC_PENDING(C::V), a0), c.s_arc()));
    }
  }
}

// item kinds of the network engine's item-parallel kernel
constexpr int kNetPre = 0;    // every river: calculate_discharge + its catchments
constexpr int kNetLevel = 1;  // the rivers of one level
constexpr int kNetFinal = 2;  // receive-only nodes
constexpr int kNetBoundary = 3;  // boundary nodes of a partitioned network: park the series value

// `i` indexes net_calc_node (pre), net_level_node / net_item (level) or net_final_node (final)
template <class C>
WSI_HD void net_item(C& c, int kind, int i) {
  const WsiModelDesc& d = c.D();
  switch (kind) {
    case kNetPre: net_pre(c, d.net_calc_node[i]); break;
    case kNetLevel: {
      NetItem<C> it = net_item_load(c, net_rec_load(c, i));
      net_level(c, it);
      break;
    }
    case kNetBoundary: boundary_release(c, d.net_bnd_node[i], false); break;
    default: net_final(c, d.net_final_node[i]); break;
  }
}

// ---- lane-parallel form of net_level (device only) ---------------------------------------------
// The depth of the network, not its size, bounds a timestep: the main stem is a chain of
// thousands of levels with one river each, and a single thread spends ~2000 instructions per
// river.  A VQIP is V independent lanes once the two volumes are known, so the critical levels
// give each river a group of kLanes threads: lane q holds element q of every VQIP, volumes are
// broadcast with a shuffle, and every lane performs exactly the arithmetic the scalar code
// performs for its element (same operations, same order: results are bit-identical).  Rivers
// that need the general push (pending_mode 1, no single receiver) fall back to the scalar code
// on the group's lane 0.
#if defined(__CUDACC__)
constexpr int kLanes = 16;

constexpr int kLanePaths = 4;  // river paths whose arcs are resolved during the look-ahead
template <class C>
struct LaneItem {
  NetRec<C> rec;
  double Tq, bkq, rc, cap, flow_in;
  int a0[kLanePaths];
};
template <class C>
__device__ __forceinline__ LaneItem<C> lane_item_load(const C& c, const NetRec<C>& rec, int q) {
  constexpr int V = C::V;
  const WsiModelDesc& d = c.D();
  LaneItem<C> it;
  it.rec = rec;
  const bool live = q < V;
  it.Tq = live ? c.tank(WSI_TANK_STORAGE(V), rec.tk)[(size_t)q * c.s_tank()] : 0.0;
  it.rc = *c.river(WSI_RIVER_RIVERRC, rec.g);
  it.cap = it.flow_in = it.bkq = 0.0;
  if (rec.out_arc >= 0) {
    const Obj A = c.arc_obj(rec.out_arc);
    it.cap = *A(WSI_ARC_CAPACITY);
    it.flow_in = *A(WSI_ARC_FLOW_IN);
    if (live) it.bkq = A(WSI_ARC_VQIP_IN(V))[(size_t)q * c.s_arc()];
  }
#pragma unroll
  for (int j = 0; j < kLanePaths; j++)
    it.a0[j] = rec.pb + j < rec.pe ? d.net_rpath[(size_t)(rec.pb + j) * 4] : rec.first_arc;
  return it;
}
// element q of sum_vqip(a, b) (vq_sum): a0, b0 are the volumes
// (lanes beyond the VQIP pass blend = false and idle along the additive branch)
__device__ __forceinline__ double lane_sum(bool blend, int q, double aq, double a0, double bq, double b0) {
  const double tot = a0 + b0;
  if (q == 0) return tot;
  if (!blend) return aq + bq;
  const bool wet = tot > 0;
  const double den = wet ? tot : 1.0;
  const double num = bq * b0 + aq * a0;
  return wet ? div_z(num, den) : aq;
}
template <class C>
WSI_HDN void net_level_scalar(C& c, int i) {
  NetItem<C> it = net_item_load(c, net_rec_load(c, i));
  net_level(c, it);
}
// Pull everything net_level_lanes will touch for item `i` towards L2 (and its pages into the
// TLB): a river's VQIPs are plane-major, so its V elements live in V different pages, and the
// main stem visits rivers scattered over gigabytes.  Run by otherwise idle groups of the tail
// block a few levels ahead of the consumer, off its critical path.
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
template <class C>
__device__ __forceinline__ void net_prefetch_item(const C& c, int i, int q) {
  constexpr int V = C::V;
  const WsiModelDesc& d = c.D();
  const NetRec<C> rec = net_rec_load(c, i);
  const size_t s_t = c.s_tank(), s_a = c.s_arc();
  if (q < V) prefetch_l2(c.tank(WSI_TANK_STORAGE(V), rec.tk) + (size_t)q * s_t);
  if (q == V) prefetch_l2(c.river(WSI_RIVER_RIVERRC, rec.g));
  if (rec.out_arc >= 0) {
    const Obj A = c.arc_obj(rec.out_arc);
    if (q < V) {
      prefetch_l2(A(WSI_ARC_VQIP_IN(V)) + (size_t)q * s_a);
      prefetch_l2(A(WSI_ARC_PENDING(V)) + (size_t)q * s_a);
    }
    if (q == V) prefetch_l2(A(WSI_ARC_CAPACITY));
    if (q == V + 1) prefetch_l2(A(WSI_ARC_FLOW_IN));
    if (q == V + 2) prefetch_l2(A(WSI_ARC_PENDING_MODE));
  }
#pragma unroll 1
  for (int p = rec.pb; p < rec.pe; p++) {
    const Obj A = c.arc_obj(d.net_rpath[(size_t)p * 4]);
    if (q < V) prefetch_l2(A(WSI_ARC_PENDING(V)) + (size_t)q * s_a);
    if (q == V) prefetch_l2(A(WSI_ARC_PENDING_MODE));
  }
}

// `gmask`: the warp lanes of this group; q: lane inside the group.  Two parts, so that a caller
// can signal "level done" between them:
//   net_lanes_release   what the next level waits for: fold the upstream releases into the tank,
//                       pull this river's release, park it on the out-arc.  Returns true -- having
//                       stored nothing -- when the river needs the general push code
//                       (net_level_scalar).
//   net_lanes_settle    the rest: the empty reply re-blended into the tank, the release booked on
//                       the arc, the tank stored.
template <class C>
struct LaneOut {
  double Tq, oq, o0;
  bool booked;
};
template <class C>
__device__ __forceinline__ bool net_lanes_release(C& c, const LaneItem<C>& it, int q, unsigned gmask, LaneOut<C>& out) {
  constexpr int V = C::V;
  static_assert(V <= kLanes, "a VQIP must fit a lane group");
  const WsiModelDesc& d = c.D();
  const NetRec<C>& rec = it.rec;
  const size_t s_a = c.s_arc();
  const bool live = q < V;
  const bool blend = q != 0 && live && !bit(c.addm(), q);  // non-additive lane
  const int np = rec.pe - rec.pb;
  // all loads first, then the folds (which are a dependent chain through the tank)
  double vq[kLanePaths], md[kLanePaths];
#pragma unroll
  for (int j = 0; j < kLanePaths; j++) {
    vq[j] = md[j] = 0.0;
    if (j < np) {
      const Obj A = c.arc_obj(it.a0[j]);
      md[j] = *A(WSI_ARC_PENDING_MODE);
      if (live) vq[j] = A(WSI_ARC_PENDING(V))[(size_t)q * s_a];
    }
  }
  bool general = rec.out_arc < 0;
  double Tq = it.Tq;
#pragma unroll
  for (int j = 0; j < kLanePaths; j++)
    if (j < np) {
      general |= md[j] != 0.0;
      const double v0 = __shfl_sync(gmask, vq[j], 0, kLanes), T0 = __shfl_sync(gmask, Tq, 0, kLanes);
      Tq = lane_sum(blend, q, Tq, T0, vq[j], v0);
    }
#pragma unroll 1
  for (int p = rec.pb + kLanePaths; p < rec.pe; p++) {
    const Obj A = c.arc_obj(d.net_rpath[(size_t)p * 4]);
    const double mode = *A(WSI_ARC_PENDING_MODE);
    const double v = live ? A(WSI_ARC_PENDING(V))[(size_t)q * s_a] : 0.0;
    general |= mode != 0.0;
    const double v0 = __shfl_sync(gmask, v, 0, kLanes), T0 = __shfl_sync(gmask, Tq, 0, kLanes);
    Tq = lane_sum(blend, q, Tq, T0, v, v0);
  }
  if (general) return true;  // uniform in the group; nothing stored: the caller runs the scalar code
  // vq_pull: reply = v_change_vqip(T, min(T0 * rc, T0)); T = extract_vqip(T, reply)
  const double T0 = __shfl_sync(gmask, Tq, 0, kLanes);
  double oq = 0.0;
  if (T0 != 0) {
    const double nv = pymin(T0 * it.rc, T0);
    if (T0 > 0) {
      const double ratio = div_z(nv, T0);
      oq = blend ? Tq : Tq * ratio;
    } else {
      oq = q == 0 ? nv : Tq;
    }
    if (!blend) Tq = Tq - oq;
  }
  const double o0 = __shfl_sync(gmask, oq, 0, kLanes);
  const Obj A = c.arc_obj(rec.out_arc);
  const bool booked = rec.direct && o0 > 0 && it.cap - it.flow_in >= o0;
  if (live) A(WSI_ARC_PENDING(V))[(size_t)q * s_a] = oq;
  if (q == 0) *A(WSI_ARC_PENDING_MODE) = booked ? 0.0 : 1.0;
  out.Tq = Tq;
  out.oq = oq;
  out.o0 = o0;
  out.booked = booked;
  return false;
}
template <class C>
__device__ __forceinline__ void net_lanes_settle(C& c, const LaneItem<C>& it, int q, unsigned gmask,
                                                 const LaneOut<C>& out) {
  constexpr int V = C::V;
  const NetRec<C>& rec = it.rec;
  const bool live = q < V;
  const bool blend = q != 0 && live && !bit(c.addm(), q);
  double Tq = out.Tq;
  if (out.booked) {
    const Obj A = c.arc_obj(rec.out_arc);
    const double T0b = __shfl_sync(gmask, Tq, 0, kLanes);
    if (blend && T0b > 0) {  // tank.push_storage(empty reply, force=True)
      const double num = 0.0 * 0.0 + Tq * T0b;
      Tq = div_z(num, T0b);
    }
    const double bk0 = __shfl_sync(gmask, it.bkq, 0, kLanes);
    if (q == 0) *A(WSI_ARC_FLOW_IN) = it.flow_in + out.o0;
    if (live) A(WSI_ARC_VQIP_IN(V))[(size_t)q * c.s_arc()] = lane_sum(blend, q, it.bkq, bk0, out.oq, out.o0);
  }
  if (live) c.tank(WSI_TANK_STORAGE(V), rec.tk)[(size_t)q * c.s_tank()] = Tq;
}

// split-phase block barrier (mbarrier): arrive as soon as the next level's inputs are stored,
// wait after the work nobody is waiting for
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int n) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(n));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(smem_u32(b)),
      "r"(parity)
      : "memory");
}
#endif  // __CUDACC__

// The whole routing phase of one replica in level order (host build and tiny models; the CUDA
// library launches net_item per (node, replica) instead, one launch per level).
template <class C>
WSI_HD void net_schedule(C& c) {
  const WsiModelDesc& d = c.D();
  for (int i = 0; i < d.n_net_bnd; i++) net_item(c, kNetBoundary, i);
  for (int i = 0; i < d.n_net_calc; i++) net_item(c, kNetPre, i);
  for (int l = 0; l < d.n_net_levels; l++)
    for (int i = d.net_level_ptr[l]; i < d.net_level_ptr[l + 1]; i++) net_item(c, kNetLevel, i);
  for (int i = 0; i < d.n_net_final; i++) net_item(c, kNetFinal, i);
}

}  // namespace wsi
