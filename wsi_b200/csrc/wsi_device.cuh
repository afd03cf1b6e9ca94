// Per-replica timestep logic of the wsi_b200 kernels.
//
// One thread advances one replica (ensemble member / independent catchment) of the
// template graph through one model timestep: the orchestration schedule, river routing,
// per-node mass balance, recording and end_timestep -- the body of the loop at
// wsimod/orchestration/model.py:993-1176.  State lives in the SoA group tensors
// (include/wsi_layout.h); plane `f` of object `o` for replica `r` is at
// base[(f * n_obj + o) * R + r], so a warp of 32 consecutive replicas reads 256
// contiguous bytes per plane.
//
// Everything here is __host__ __device__: the shipped library only instantiates it inside
// __global__ kernels (wsi_api.cu); tests/hostsim compiles the same code for the CPU so the
// CPU-only CI can exercise the kernel logic (it is never loaded by the product).
#pragma once

#include "../../include/wsi_b200.h"
#include "../../include/wsi_layout.h"
#include "wsi_vq.cuh"

namespace wsi {

constexpr double kUnbounded = 1e15;  // constants.UNBOUNDED_CAPACITY
constexpr int kMaxIter = 5;          // constants.MAXITER
constexpr int kMaxCand = WSI_MAX_CAND;  // arcs one push_distributed call can split over (the flattener refuses more)
constexpr int kJunctionDepth = 3;    // nested junction Nodes a push may traverse (the Oxford demo chains 3)

// The model descriptor lives in constant memory on the device (direct constant-bank loads with
// immediate offsets from every function, inlined or not); wsi_api.cu copies it into a slot
// before launching.  The host build (tests/hostsim) reads it through a pointer.
#ifndef WSI_DESC_SLOTS
#define WSI_DESC_SLOTS 32
#endif
#if defined(__CUDACC__)
static __constant__ WsiModelDesc g_wsi_desc[WSI_DESC_SLOTS];  // one copy per configuration unit
// The schedule kernel stages its descriptor at the start of its dynamic shared memory
// (wsi_stage_desc): its device functions then read it with immediate-offset shared loads, whereas
// a constant-bank access indexed by the run-time slot needs a uniform-register move per access
// (9 % of that kernel's instructions).  Small kernels with many blocks keep the constant slot --
// there the 640-byte copy per block costs more than it saves (measured).
extern __shared__ __align__(16) unsigned char wsi_smem[];
__device__ __forceinline__ void wsi_stage_desc(int slot) {
  const uint32_t* src = reinterpret_cast<const uint32_t*>(&g_wsi_desc[slot]);
  uint32_t* dst = reinterpret_cast<uint32_t*>(wsi_smem);
  for (int i = threadIdx.x; i < (int)(sizeof(WsiModelDesc) / 4); i += blockDim.x) dst[i] = src[i];
  __syncthreads();
}
#endif

// Re-align the warps of a block inside a long schedule entry (instruction-cache locality, see
// wsi_step_kernel); only at points every thread of the block reaches.
#if defined(__CUDA_ARCH__) && !defined(WSI_NO_ALIGN)
#define WSI_BLOCK_ALIGN(c)                \
  do {                                    \
    if ((c).block_sync) __syncthreads();  \
  } while (0)
#else
#define WSI_BLOCK_ALIGN(c)
#endif

// One object's column in a group tensor: plane k lives at p[k * s]
struct Obj {
  double* p;
  size_t s;
  WSI_HD double* operator()(int plane) const { return p + (size_t)plane * s; }
};

// SMEM: the kernel staged its descriptor in shared memory (wsi_stage_desc) and device functions
// read it from there; otherwise from the constant-memory slot.
// LEAN: the schedule kernel's context for models without treatment works, demand, distribution,
// queue groundwater or abstractions (WsiModelDesc.features == 0) -- their handlers are compiled out,
// which keeps the hot kernel's code and register pressure where they were before those nodes existed.
template <class CFG, bool SMEM = false, bool LEAN = false>
struct Ctx {
  static constexpr bool kInlinePush = SMEM;  // the schedule kernel's context (see push_distributed)
  static constexpr bool kLean = LEAN;
  static constexpr int P = CFG::P;
  static constexpr int V = P + 1;
  using Vq = VQ<P>;

  const WsiModelDesc* host_d;  // host build only
  int slot;                    // constant-memory slot of the descriptor (device)
  CFG cfg;
  size_t R;
  int r;
  int t;
  double FA;
  uint32_t run_flags;
  bool block_sync;  // device: every thread of the block walks the same schedule -> WSI_BLOCK_ALIGN may sync
  int tag;         // tag of the push being made: 0 "default", 1 "Demand" (Sewer.push_set_land, sewer.py:95-98),
                   // 2 ("Demand", "Garden") (a Land's GardenSurface, land.py:120-128)
  int viol;        // mass-balance violations found for this replica this step
  double* mb_acc;  // [3*V] per-thread system-total accumulators (stride mb_stride) or null
  int mb_stride;

  // additive-pollutant mask: a compile-time constant for the default pollutant set -- also inside
  // the noinline functions, where a copy kept in the context would be a run-time value (twelve
  // run-time bit tests in every VQIP operation)
  WSI_HD uint32_t addm() const { return cfg.add_mask(); }

  WSI_HD const WsiModelDesc& D() const {
#if defined(__CUDA_ARCH__)
    if constexpr (SMEM) return *reinterpret_cast<const WsiModelDesc*>(wsi_smem);
    else return g_wsi_desc[slot];
#else
    return *host_d;
#endif
  }

  // ---- plane addressing ----
  WSI_HD double* pl(double* base, int n_obj, int plane, int obj) const {
    return base + ((size_t)plane * (size_t)n_obj + (size_t)obj) * R + (size_t)r;
  }
  WSI_HD double* arc(int plane, int a) const { return pl(D().arc, D().n_arcs, plane, a); }
  WSI_HD double* land(int plane, int i) const { return pl(D().land, D().n_land, plane, i); }
  WSI_HD double* surf(int plane, int i) const { return pl(D().surf, D().n_surf, plane, i); }
  WSI_HD double* tank(int plane, int i) const { return pl(D().tank, D().n_tank, plane, i); }
  WSI_HD double* river(int plane, int i) const { return pl(D().river, D().n_river, plane, i); }
  WSI_HD double* sewer(int plane, int i) const { return pl(D().sewer, D().n_sewer, plane, i); }
  WSI_HD double* works(int plane, int i) const { return pl(D().works, D().n_works, plane, i); }
  WSI_HD double* demand(int plane, int i) const { return pl(D().demand, D().n_demand, plane, i); }
  WSI_HD size_t s_works() const { return (size_t)D().n_works * R; }
  WSI_HD size_t s_demand() const { return (size_t)D().n_demand * R; }
  WSI_HD double* distrib(int plane, int i) const { return pl(D().distrib, D().n_distrib, plane, i); }
  WSI_HD size_t s_distrib() const { return (size_t)D().n_distrib * R; }
  WSI_HD Obj arc_obj(int a) const { return Obj{D().arc + (size_t)a * R + r, (size_t)D().n_arcs * R}; }
  WSI_HD Obj land_obj(int i) const { return Obj{D().land + (size_t)i * R + r, (size_t)D().n_land * R}; }
  WSI_HD Obj surf_obj(int i) const { return Obj{D().surf + (size_t)i * R + r, (size_t)D().n_surf * R}; }
  WSI_HD Obj tank_obj(int i) const { return Obj{D().tank + (size_t)i * R + r, (size_t)D().n_tank * R}; }
  WSI_HD Obj river_obj(int i) const { return Obj{D().river + (size_t)i * R + r, (size_t)D().n_river * R}; }
  WSI_HD Obj sewer_obj(int i) const { return Obj{D().sewer + (size_t)i * R + r, (size_t)D().n_sewer * R}; }
  WSI_HD size_t s_arc() const { return (size_t)D().n_arcs * R; }
  WSI_HD size_t s_land() const { return (size_t)D().n_land * R; }
  WSI_HD size_t s_surf() const { return (size_t)D().n_surf * R; }
  WSI_HD size_t s_tank() const { return (size_t)D().n_tank * R; }
  WSI_HD size_t s_river() const { return (size_t)D().n_river * R; }
  WSI_HD size_t s_sewer() const { return (size_t)D().n_sewer * R; }

  WSI_HD Vq ldv(const double* p, size_t stride) const {
    Vq v;
#pragma unroll
    for (int i = 0; i < V; i++) v.x[i] = p[(size_t)i * stride];
    return v;
  }
  WSI_HD void stv(double* p, size_t stride, const Vq& v) const {
#pragma unroll
    for (int i = 0; i < V; i++) p[(size_t)i * stride] = v.x[i];
  }
  // volume + additive lanes only (everything a mass balance reads, core.py:400-404)
  WSI_HD Vq ldv_add(const double* p, size_t stride) const {
    Vq v;
#pragma unroll
    for (int i = 0; i < V; i++) v.x[i] = (i == 0 || bit(addm(), i)) ? p[(size_t)i * stride] : 0.0;
    return v;
  }
  WSI_HD void stv_add(double* p, size_t stride, const Vq& v) const {
#pragma unroll
    for (int i = 0; i < V; i++)
      if (i == 0 || bit(addm(), i)) p[(size_t)i * stride] = v.x[i];
  }
  WSI_HD void flag(int node, uint32_t f) const { D().node_flags[(size_t)node * R + r] |= f; }

  // ---- forcing ----
  WSI_HD double forcing(int fid, int col) const {
    return ldp(D().node_forcing + ((size_t)t * D().n_fseries + fid) * 3 + col);
  }
  WSI_HD double dep(int dep_id, int col) const {
    return ldp(D().surf_inputs + ((size_t)D().month_idx[t] * D().n_dep_series + dep_id) * WSI_N_DEP + col);
  }

  // ---- algebra shorthands ----
  WSI_HD Vq sum(const Vq& a, const Vq& b) const { return vq_sum<P>(addm(), a, b); }
  WSI_HD Vq extract(const Vq& a, const Vq& b) const { return vq_extract<P>(addm(), a, b); }
  WSI_HD Vq vchange(const Vq& a, double v) const { return vq_vchange<P>(addm(), a, v); }
};

// =========================================================================================
// Arc transients.  An arc's flow_in / vqip_in (arcs.py:64-66, 152-156) live for one timestep and
// Arc.end_timestep resets them (arcs.py:259-265).  The replica engine never writes those zeros:
// the arc's `stamp` plane names the timestep its transients belong to, and anything else reads as
// the empty arc -- the first booking of a step simply overwrites.  (The network engine keeps its
// own bookkeeping and zeroes the arcs in its end_timestep pass: there every arc is live.)
// =========================================================================================
template <class C>
WSI_HD bool arc_live(const C& c, int a) {
  return c.D().engine == 1 || *c.arc(WSI_ARC_STAMP, a) == (double)c.t;
}
template <class C>
WSI_HD double arc_flow_in(const C& c, int a) {
  return arc_live(c, a) ? *c.arc(WSI_ARC_FLOW_IN, a) : 0.0;
}
template <class C>
WSI_HD double arc_lane(const C& c, int a, int lane) {  // one lane of vqip_in
  return arc_live(c, a) ? *c.arc(WSI_ARC_VQIP_IN(C::V) + lane, a) : 0.0;
}
template <class C>
WSI_HD typename C::Vq arc_vqip_add(const C& c, int a) {  // volume + additive lanes of vqip_in
  return arc_live(c, a) ? c.ldv_add(c.arc(WSI_ARC_VQIP_IN(C::V), a), c.s_arc()) : vq_zero<C::P>();
}
// book `sent` on the arc: flow_in += volume, vqip_in = sum_vqip(vqip_in, sent) (arcs.py:152-156)
template <class C>
WSI_HD void arc_book(C& c, int a, const typename C::Vq& sent) {
  double* fi = c.arc(WSI_ARC_FLOW_IN, a);
  double* vi = c.arc(WSI_ARC_VQIP_IN(C::V), a);
  if (arc_live(c, a)) {
    *fi = *fi + sent.x[0];
    c.stv(vi, c.s_arc(), c.sum(c.ldv(vi, c.s_arc()), sent));
  } else {  // first booking of the step: the arc is empty
    *fi = 0.0 + sent.x[0];
    c.stv(vi, c.s_arc(), c.sum(vq_zero<C::P>(), sent));
    *c.arc(WSI_ARC_STAMP, a) = (double)c.t;
  }
}

// =========================================================================================
// Tanks addressed in place (global memory read-modify-write)
// =========================================================================================
template <class C>
struct TankRef {
  double* st;     // plane 0 of the storage VQIP
  size_t stride;  // distance between planes
  double capacity;
};

// volume and additive lanes all zero (non-additive lanes of a zero-volume VQIP never reach a
// result: they are always multiplied by the volume)
template <class C>
WSI_HD bool vq_is_empty(const C& c, const typename C::Vq& v) {
  bool z = v.x[0] == 0.0;
#pragma unroll
  for (int p = 1; p <= C::P; p++)
    if (bit(c.addm(), p)) z = z && v.x[p] == 0.0;
  return z;
}

// compare_vqip core.py:354-368: every lane within FLOAT_ACCURACY
template <class C>
WSI_HD bool vq_compare(const C& c, const typename C::Vq& a, const typename C::Vq& b) {
  bool same = true;
#pragma unroll
  for (int p = 0; p <= C::P; p++)
    if (fabs(a.x[p] - b.x[p]) > c.FA) same = false;
  return same;
}

// Tank.push_storage(force=True) tanks.py:189-192: storage = sum_vqip(storage, v).  When v is
// empty the volume and additive lanes are unchanged (x + 0) and only the non-additive lanes are
// re-blended, (v_p * 0 + s_p * vol) / vol, so only those are read and written.
template <class C>
WSI_HD void tank_push_force(const C& c, double* st, size_t stride, const typename C::Vq& v) {
  if (vq_is_empty(c, v)) {
    const double vol = st[0];
    if (vol > 0) {
#pragma unroll
      for (int p = 1; p <= C::P; p++)
        if (!bit(c.addm(), p)) {
          const double num = v.x[p] * v.x[0] + st[(size_t)p * stride] * vol;
          st[(size_t)p * stride] = div_z(num, vol);
        }
    }
    return;
  }
  c.stv(st, stride, c.sum(c.ldv(st, stride), v));
}
// Tank.get_excess()["volume"] tanks.py:132-158
WSI_HD double tank_excess_vol(double capacity, double vol, bool has_v, double v_vol) {
  double e = pymax(capacity - vol, 0.0);
  if (has_v) e = pymin(v_vol, e);
  return vol_vchange(vol, e);
}
// Tank.push_storage tanks.py:194-205 (capacity limited); returns the reply
template <class C>
WSI_HD typename C::Vq tank_push(const C& c, double* st, size_t stride, double capacity,
                                const typename C::Vq& v) {
  typename C::Vq s = c.ldv(st, stride);
  const double excess = tank_excess_vol(capacity, s.x[0], false, 0.0);
  typename C::Vq reply = c.vchange(v, pymax(v.x[0] - excess, 0.0));
  typename C::Vq entered = c.vchange(v, v.x[0] - reply.x[0]);
  c.stv(st, stride, c.sum(s, entered));
  return reply;
}
// Tank.pull_storage tanks.py:207-239 on a register-resident storage
template <class C>
WSI_HD typename C::Vq vq_pull(const C& c, typename C::Vq& s, double vol) {
  if (s.x[0] == 0) return vq_zero<C::P>();
  typename C::Vq reply = c.vchange(s, pymin(vol, s.x[0]));
  s = c.extract(s, reply);
  return reply;
}
template <class C>
WSI_HD typename C::Vq tank_pull(const C& c, double* st, size_t stride, double vol) {
  typename C::Vq s = c.ldv(st, stride);
  if (s.x[0] == 0) return vq_zero<C::P>();
  typename C::Vq reply = vq_pull(c, s, vol);
  c.stv(st, stride, s);
  return reply;
}
// Tank.pull_pollutants tanks.py:241-266: clips the request (in place) and extracts it
template <class C>
WSI_HD void vq_pull_pollutants(const C& c, typename C::Vq& s, typename C::Vq& req) {
#pragma unroll
  for (int p = 1; p <= C::P; p++)
    if (bit(c.addm(), p)) req.x[p] = pymin(s.x[p], req.x[p]);
  req.x[0] = pymin(s.x[0], req.x[0]);
  s = c.extract(s, req);
}

// =========================================================================================
// Push protocol: Arc.send_push_check / send_push_request (arcs.py:112-257),
// Node.push_distributed / check_basic (nodes.py:409-508) and the push handlers of every
// supported node class.  D is the number of junction `Node`s the call may still traverse.
// =========================================================================================
template <int D, class C>
WSI_HD double arc_excess_vol(C& c, int a, bool has_v, double v_vol);
template <int D, class C>
WSI_HD typename C::Vq arc_send_push_request(C& c, int a, const typename C::Vq& vqip);
template <int D, class C>
WSI_HD typename C::Vq node_push_set(C& c, int node, const typename C::Vq& vqip);
template <int D, class C>
WSI_HD typename C::Vq push_distributed(C& c, int node, const typename C::Vq& vqip, int filter);

// treatment works as push receivers (wsi_works.cuh)
// (out of line, and handed a COPY of the context: rare receivers must neither grow every
// instantiation of the push code nor force the caller's context into memory)
template <int D, class C>
WSI_HDN double rres_push_check_vol(C c, int node, bool has_v, double v_vol);
template <int D, class C>
WSI_HDN typename C::Vq rres_push_set(C c, int node, typename C::Vq vqip);
template <class C>
WSI_HDN double wwtw_push_check_vol(C c, int g, bool has_v, double v_vol);
template <class C>
WSI_HDN typename C::Vq wwtw_push_set(C c, int g, typename C::Vq vqip);
template <class C>
WSI_HDN typename C::Vq sewer_push_timearea(C c, int w, typename C::Vq vqip);
template <class C>
WSI_HDN double garden_irrigation_demand(C c, int land);
template <class C>
WSI_HDN void garden_receive_irrigation(C c, int land, typename C::Vq vqip);

// sewer QueueTank helpers -------------------------------------------------------------------
template <class C>
WSI_HD void sewer_update_queue(C& c, int w) {  // AltQueueArc.update_queue arcs.py:603-638
  constexpr int V = C::V;
  const size_t s = c.s_sewer();
  double* q0 = c.sewer(WSI_SEWER_QUEUE(V), w);
  double* act = c.sewer(WSI_SEWER_ACTIVE_STORAGE(V), w);
  const typename C::Vq q = c.ldv(q0, s);
  tank_push_force(c, act, s, q);  // QueueTank.push_set tanks.py:610-629: active = sum(active, q)
  if (!vq_is_empty(c, q) || q.x[0] != 0.0) c.stv(q0, s, vq_zero<C::P>());
  else {
#pragma unroll
    for (int p = 1; p <= C::P; p++)
      if (!bit(c.addm(), p)) q0[(size_t)p * s] = 0.0;
  }
}
// DecayObj.make_decay (core.py:474-521) of a VQIP entering or waiting in a DecayQueueTank's delay
// line (DecayArcAlt, arcs.py:731-805): diff = x * min(k * theta^(T - 20), 1) leaves the VQIP and
// accumulates in the arc's total_decayed.  A store without decays is left alone.
template <class C>
WSI_HD void queue_make_decay(C& c, int w, typename C::Vq& v) {
  const int n = (int)*c.sewer(WSI_SEWER_N_DECAY, w);
  if (n <= 0) return;
  const double temperature = c.forcing((int)*c.sewer(WSI_SEWER_FORCING_ID, w), 2);
  for (int i = 0; i < WSI_MAX_DECAY; i++) {
    if (i >= n) break;
    const int pol = (int)*c.sewer(WSI_SEWER_DECAY_POL + i, w);
    const double f = *c.sewer(WSI_SEWER_DECAY_CONST + i, w) * pow(*c.sewer(WSI_SEWER_DECAY_EXP + i, w), temperature - 20.0);
    const double x = vget<C::P>(v, pol);
    const double diff = x * pymin(f, 1.0);
    vset<C::P>(v, pol, x - diff);
    // total_decayed = sum_vqip(total_decayed, diff): a zero-volume VQIP, so only additive lanes accumulate
    double* td = c.sewer(WSI_SEWER_TOTAL_DECAYED + i, w);
    if (bit(c.addm(), pol)) *td = *td + diff;
  }
}

// QueueTank.push_storage(vqip, time) tanks.py:513-539 through AltQueueArc (arcs.py:360-407).
template <class C>
WSI_HD typename C::Vq sewer_push_storage(C& c, int w, const typename C::Vq& v0, int time) {
  constexpr int V = C::V;
  using Vq = typename C::Vq;
  const size_t s = c.s_sewer();
  double* st = c.sewer(WSI_SEWER_STORAGE(V), w);
  Vq reply;
  if (v0.x[0] < c.FA) {
    reply = vq_zero<C::P>();
  } else {
    const double cap = *c.sewer(WSI_SEWER_CAPACITY, w);
    double* ia_flow_in = c.sewer(WSI_SEWER_IA_FLOW_IN, w);
    const double pipe_excess = kUnbounded - *ia_flow_in;
    const double svol = st[0];
    double ne = vol_vchange(svol, pymax(cap - svol, 0.0));  // get_excess()
    ne = pymin(v0.x[0], ne);                                // QueueTank.push_check tanks.py:605-608
    const double excess_in = vol_vchange(ne, pymin(pipe_excess, ne));
    Vq not_pushed = c.vchange(v0, pymax(v0.x[0] - excess_in, 0.0));
    Vq v = c.extract(v0, not_pushed);
    if constexpr (!C::kLean) time += (int)*c.sewer(WSI_SEWER_QUEUE_DELAY, w);  // + number_of_timesteps (arcs.py:385)
    *ia_flow_in = *ia_flow_in + v.x[0] / (double)(time + 1);  // enter_arc arcs.py:423-427
    if constexpr (!C::kLean) queue_make_decay(c, w, v);  // DecayArcAlt.enter_queue arcs.py:750-776: decay on entry
    double* qm = c.sewer(WSI_SEWER_QMASK, w);
    unsigned qmask = (unsigned)*qm;
    double* q = c.sewer(WSI_SEWER_QUEUE(V) + time * V, w);
    if (qmask & (1u << time)) {  // AltQueueArc.enter_queue arcs.py:595-601
      c.stv(q, s, c.sum(c.ldv(q, s), v));
    } else {
      c.stv(q, s, v);
      *qm = (double)(qmask | (1u << time));
    }
    sewer_update_queue(c, w);
    reply = c.sum(not_pushed, vq_zero<C::P>());  // + (empty) backflow
  }
  Vq entered = c.vchange(v0, v0.x[0] - reply.x[0]);
  c.stv(st, s, c.sum(c.ldv(st, s), entered));
  return reply;
}

template <int D, class C>
WSI_HD double node_push_check_vol(C& c, int node, bool has_v, double v_vol) {
  const WsiModelDesc& d = c.D();
  const int g = d.node_gidx[node];
  constexpr int V = C::V;
  switch (d.node_type[node]) {
    case WSI_NT_NODE: {  // push_check_basic nodes.py:529-545 -> check_basic :480-508
      if constexpr (D == 0) {
        c.flag(node, WSI_FLAG_DEPTH);
        return 0.0;
      } else {
        const int b = d.cand_ptr[node * WSI_NF + WSI_FILTER_NRWR];
        const int e = d.cand_ptr[node * WSI_NF + WSI_FILTER_NRWR + 1];
        double avail = 0.0;
#pragma unroll 1
        for (int i = b; i < e; i++) avail = avail + arc_excess_vol<D - 1>(c, d.cand_arcs[i], false, 0.0);
        if (has_v) avail = vol_vchange(avail, pymin(avail, v_vol));
        return avail;
      }
    }
    case WSI_NT_WASTE:  // push_check_accept nodes.py:611-623
    case WSI_NT_RIVER:
      return has_v ? v_vol : kUnbounded;
    case WSI_NT_GROUNDWATER:
    case WSI_NT_STORAGE:
    case WSI_NT_RESERVOIR:  // tank.get_excess storage.py:71
      return tank_excess_vol(*c.tank(WSI_TANK_CAPACITY, g), *c.tank(WSI_TANK_STORAGE(V), g), has_v,
                             v_vol);
    case WSI_NT_SEWER: {  // push_check_sewer sewer.py:145-161
      const double vol = *c.sewer(WSI_SEWER_STORAGE(V), g);
      const double ex = vol_vchange(vol, pymax(*c.sewer(WSI_SEWER_CAPACITY, g) - vol, 0.0));
      if (!has_v) return ex;
      return vol_vchange(ex, pymin(ex, v_vol));
    }
    case WSI_NT_RIVERRESERVOIR:  // push_check_river_reservoir storage.py:1084-1107
      if constexpr (!C::kLean) return rres_push_check_vol<D>(c, node, has_v, v_vol);
      break;
    case WSI_NT_QUEUEGROUNDWATER:  // the stale Tank's get_excess (storage.py:60, 301-315)
      if constexpr (!C::kLean)
        return tank_excess_vol(*c.sewer(WSI_SEWER_STALE_CAPACITY, g), *c.sewer(WSI_SEWER_STALE_VOLUME, g), has_v, v_vol);
      break;
    case WSI_NT_WWTW:  // push_check_sewer wtw.py:319-342 (its "Sewer" and "default" handlers)
      if constexpr (!C::kLean) return wwtw_push_check_vol(c, g, has_v, v_vol);
      break;
    case WSI_NT_LAND:  // only the ("Demand", "Garden") tag is answered (land.py:93, 120-128): irrigation demand
      if constexpr (!C::kLean)
        if (c.tag == 2) return garden_irrigation_demand(c, g);
      break;
    default:  // Land / Catchment / FWTW / Demand deny pushes (nodes.py:595-609)
      break;
  }
  c.flag(node, WSI_FLAG_DENY);
  return 0.0;
}

template <int D, class C>
WSI_HD typename C::Vq node_push_set(C& c, int node, const typename C::Vq& vqip) {
  const WsiModelDesc& d = c.D();
  const int g = d.node_gidx[node];
  constexpr int V = C::V;
  switch (d.node_type[node]) {
    case WSI_NT_NODE:  // junction: forward (nodes.py:62-66)
      if constexpr (D == 0) {
        c.flag(node, WSI_FLAG_DEPTH);
        return vqip;
      } else {  // a junction forwards with the default tag (nodes.py:62-66)
        const int tag = c.tag;
        c.tag = 0;
        const typename C::Vq r = push_distributed<D - 1>(c, node, vqip, WSI_FILTER_NRWR);
        c.tag = tag;
        return r;
      }
    case WSI_NT_WASTE:  // waste.py:40-49
      return vq_zero<C::P>();
    case WSI_NT_RIVER: {  // push_set_river storage.py:711-721
      const int tk = (int)*c.river(WSI_RIVER_TANK, g);
      tank_push_force(c, c.tank(WSI_TANK_STORAGE(V), tk), c.s_tank(), vqip);
      return vq_zero<C::P>();
    }
    case WSI_NT_GROUNDWATER:
    case WSI_NT_STORAGE:
    case WSI_NT_RESERVOIR:  // push_set_storage storage.py:106-118
      return tank_push(c, c.tank(WSI_TANK_STORAGE(V), g), c.s_tank(), *c.tank(WSI_TANK_CAPACITY, g),
                       vqip);
    case WSI_NT_SEWER:  // push_set_sewer sewer.py:163-176; tagged "Demand": push_set_land sewer.py:178-199
      if constexpr (!C::kLean)
        if (c.tag == 1) return sewer_push_timearea(c, g, vqip);
      return sewer_push_storage(c, g, vqip, (int)*c.sewer(WSI_SEWER_PIPE_TIME, g));
    case WSI_NT_RIVERRESERVOIR:  // push_set_river_reservoir storage.py:1062-1082
      if constexpr (!C::kLean) return rres_push_set<D>(c, node, vqip);
      break;
    case WSI_NT_QUEUEGROUNDWATER:  // push_set_timearea storage.py:329-343
      if constexpr (!C::kLean) return sewer_push_timearea(c, g, vqip);
      break;
    case WSI_NT_WWTW:  // push_set_sewer wtw.py:344-374
      if constexpr (!C::kLean) return wwtw_push_set(c, g, vqip);
      break;
    case WSI_NT_LAND:  // ("Demand", "Garden"): GardenSurface.receive_irrigation_demand land.py:2425-2437
      if constexpr (!C::kLean)
        if (c.tag == 2) {
          garden_receive_irrigation(c, g, vqip);
          return vq_zero<C::P>();
        }
      break;
    default:
      break;
  }
  c.flag(node, WSI_FLAG_DENY);
  return vqip;
}

// Arc.get_excess(direction="push")["volume"] arcs.py:230-257
template <int D, class C>
WSI_HD double arc_excess_vol(C& c, int a, bool has_v, double v_vol) {
  const double pipe_excess = *c.arc(WSI_ARC_CAPACITY, a) - arc_flow_in(c, a);
  const double ne = node_push_check_vol<D>(c, c.D().arc_out[a], has_v, v_vol);
  return vol_vchange(ne, pymin(pipe_excess, ne));
}

// Arc.send_push_request arcs.py:112-158.  When the whole volume is admitted
// (max(vol - excess, 0) == 0) not_pushed = v_change_vqip(vqip, 0) has zero volume and zero
// additive lanes and extract_vqip(vqip, not_pushed) is vqip itself; when the receiver then hands
// nothing back, sent == vqip and sum_vqip(reply, not_pushed) leaves the reply unchanged.  Those
// identities skip four VQIP operations on the common path without changing any value.
template <int D, class C>
WSI_HD typename C::Vq arc_send_push_request(C& c, int a, const typename C::Vq& vqip) {
  using Vq = typename C::Vq;
  const double excess_in = arc_excess_vol<D>(c, a, true, vqip.x[0]);
  const double np = pymax(vqip.x[0] - excess_in, 0.0);
  const bool admitted = np == 0.0 && vqip.x[0] > 0;
  Vq not_pushed, v;
  if (admitted) {
    v = vqip;
  } else {
    not_pushed = c.vchange(vqip, np);
    v = c.extract(vqip, not_pushed);
  }
  Vq reply = node_push_set<D>(c, c.D().arc_out[a], v);
  Vq sent;
  if (admitted && vq_is_empty(c, reply)) {
    sent = v;
  } else {
    if (admitted) not_pushed = c.vchange(vqip, np);
    sent = c.extract(v, reply);
    reply = c.sum(reply, not_pushed);
  }
  arc_book(c, a, sent);
  return reply;
}

// Node.push_distributed nodes.py:409-478 -- general form (any number of candidate arcs, junctions
// that split): recursive through the junction Nodes, one call per level.  The schedule kernel only
// comes here when a push is NOT a simple chain (push_out below); it is kept out of line so that
// its VQIP temporaries and candidate tables cost stack only when it actually runs.
template <int D, class C>
WSI_HD typename C::Vq push_distributed_body(C& c, int node, const typename C::Vq& vqip, int filter);
template <int D, class C>
WSI_HDN typename C::Vq push_distributed_call(C& c, int node, const typename C::Vq& vqip, int filter) {
  return push_distributed_body<D>(c, node, vqip, filter);
}
template <int D, class C>
WSI_HD typename C::Vq push_distributed(C& c, int node, const typename C::Vq& vqip, int filter) {
  return push_distributed_call<D>(c, node, vqip, filter);
}
template <int D, class C>
WSI_HD typename C::Vq push_distributed_body(C& c, int node, const typename C::Vq& vqip, int filter) {
  using Vq = typename C::Vq;
  const WsiModelDesc& d = c.D();
  const int cb = d.cand_ptr[node * WSI_NF + filter];
  const int nc = d.cand_ptr[node * WSI_NF + filter + 1] - cb;  // <= kMaxCand (refused by the flattener otherwise)
  const int* cand = d.cand_arcs + cb;
  const int n_out = d.out_ptr[node + 1] - d.out_ptr[node];
  const bool single = n_out == 1;  // nodes.py:424-438: one out-arc -> send everything down it
  if (single && nc <= 0) return vqip;
  double not_pushed = vqip.x[0];
  Vq not_pushed_ = vqip;
  double alloc[kMaxCand], cap_[kMaxCand];
  double avail = 0.0, priority = 0.0;
  if (!single) {
    // get_connected nodes.py:271-317
#pragma unroll 1
    for (int i = 0; i < nc; i++) {
      double a = arc_excess_vol<D>(c, cand[i], false, 0.0);
      if (a < c.FA) a = 0.0;
      const double pref = *c.arc(WSI_ARC_PREFERENCE, cand[i]);
      avail += a;
      priority += a * pref;
      alloc[i] = a * pref;
      cap_[i] = a;
    }
    if (not_pushed > avail) {
      priority = avail;
#pragma unroll 1
      for (int i = 0; i < nc; i++) alloc[i] = cap_[i];
    }
  }
  int iter = 0;
#pragma unroll 1
  while (single || (not_pushed > c.FA && avail > c.FA && iter < kMaxIter)) {
    const double amount = pymin(avail, not_pushed);
    const int n_send = single ? 1 : nc;
#pragma unroll 1
    for (int i = 0; i < n_send; i++) {
      const int arc = single ? d.out_arcs[d.out_ptr[node]] : cand[i];
      const Vq to_send = single ? vqip : c.vchange(not_pushed_, amount * alloc[i] / priority);
      const Vq reply = arc_send_push_request<D>(c, arc, to_send);
      if (single) return reply;
      const Vq sent = c.extract(to_send, reply);
      not_pushed_ = c.extract(not_pushed_, sent);
    }
    not_pushed = not_pushed_.x[0];
    iter++;
    if (!(not_pushed > c.FA)) break;  // the loop ends here: get_connected's result would go unused
    avail = 0.0;
    priority = 0.0;
#pragma unroll 1
    for (int i = 0; i < nc; i++) {
      double a = arc_excess_vol<D>(c, cand[i], false, 0.0);
      if (a < c.FA) a = 0.0;
      const double pref = *c.arc(WSI_ARC_PREFERENCE, cand[i]);
      avail += a;
      priority += a * pref;
      alloc[i] = a * pref;
    }
  }
  if (iter == kMaxIter) c.flag(node, WSI_FLAG_MAXITER);
  return not_pushed_;
}

// =========================================================================================
// Simple chains: the push protocol without recursion
//
// Almost every push of a real model travels down a *simple chain*: an arc, then zero or more
// junction `Node`s that each forward to exactly one arc, then a node that stores or swallows
// the water.  The host lists the arcs of the chain that starts at every arc (arc_chain_ptr /
// arc_chain; empty when a junction on the way splits or the chain is longer than kMaxChain).
// For such a chain the recursion arc.send_push_request -> node.push_set -> push_distributed ->
// arc.send_push_request ... (arcs.py:112-158, nodes.py:62-66, 409-438) unrolls into three
// passes over at most kMaxChain arcs -- the excess the chain offers (volumes only, bottom up),
// the request each arc admits (top down), the receiver's push_set and the arcs' bookkeeping
// (bottom up) -- with exactly the arithmetic of the recursive form, no calls and nothing
// indexed at run time, so every VQIP stays in registers.  When an arc on the way does not admit
// the whole request the general recursive code runs instead (nothing has been changed yet).
// =========================================================================================
constexpr int kMaxChain = kJunctionDepth + 1;

#if !defined(__CUDA_ARCH__) && defined(WSI_HOST_STATS)  // development aid (host build): which path a send takes
static long long g_send_stats[8];
#define WSI_STAT(i) (g_send_stats[i]++)
#else
#define WSI_STAT(i)
#endif

// vol_vchange(x, min(p, x)): the request x clamped by a pipe's remaining room p.  When p >= x
// this is x * (x / x) == x exactly, so the division is only evaluated when the pipe binds.
WSI_HD double clamp_to(double x, double p) { return p < x ? vol_vchange(x, p) : x; }

template <class C>
WSI_HD int chain_of(const C& c, int a, const int*& arcs) {
  const WsiModelDesc& d = c.D();
  if (!d.arc_chain_ptr) return 0;
  const int b = d.arc_chain_ptr[a];
  arcs = d.arc_chain + b;
  return d.arc_chain_ptr[a + 1] - b;
}

// What a chain can take: pipe[k] = capacity - flow_in of its k-th arc and E[k] =
// Arc.get_excess("push") of that arc without a request (arcs.py:230-257 with vqip None; a
// junction's push_check without a request is the excess of its only arc, nodes.py:480-508).
template <class C>
WSI_HD void chain_excess(C& c, const int* arcs, int n, double (&pipe)[kMaxChain], double (&E)[kMaxChain]) {
  double u = node_push_check_vol<0>(c, c.D().arc_out[arcs[n - 1]], false, 0.0);
#pragma unroll
  for (int k = kMaxChain - 1; k >= 0; k--) {
    pipe[k] = 0.0;
    E[k] = 0.0;
    if (k < n) {
      pipe[k] = *c.arc(WSI_ARC_CAPACITY, arcs[k]) - arc_flow_in(c, arcs[k]);
      u = clamp_to(u, pipe[k]);
      E[k] = u;
      u = 0.0 + u;  // the junction above sums its candidates' excesses from zero
    }
  }
}

template <class C>
WSI_HDN typename C::Vq arc_send_general_call(C c, int a, typename C::Vq vqip, int* viol) {
  const typename C::Vq r = arc_send_push_request<kJunctionDepth>(c, a, vqip);
  *viol = c.viol;
  return r;
}
template <class C>
WSI_HDN typename C::Vq push_general_call(C c, int node, typename C::Vq vqip, int filter, int* viol) {
  const typename C::Vq r = push_distributed_body<kJunctionDepth>(c, node, vqip, filter);
  *viol = c.viol;
  return r;
}

// Arc.send_push_request(vqip) on arc `a` (== arc_send_push_request<kJunctionDepth>).  Returns true
// when the reply is empty (zero volume and zero additive lanes; `reply` is then all zeros -- the
// non-additive lanes of an empty VQIP never reach a result, they are only ever multiplied by its
// zero volume).
// An empty request (zero volume, zero additive lanes) travelling down arcs[k0 .. n) to node X: on
// every arc not_pushed = v_change_vqip(vqip, 0) is the request itself, what travels on is empty
// again, and so is every reply.  What is left of sum_vqip with an empty operand is the re-blend
// of the non-additive lanes, (x_p * 0 + s_p * vol) / vol: of the receiver's store (push_set_river
// storage.py:711-721, Tank.push_storage tanks.py:160-205, QueueTank.push_storage below
// FLOAT_ACCURACY tanks.py:513-539) and of every arc's vqip_in (arcs.py:152-156); Waste takes
// nothing.
// push_set of the chain's receiver; a tag only reaches a receiver right behind the first arc --
// junctions forward with the default one (nodes.py:62-66)
template <class C>
WSI_HD typename C::Vq chain_push_set(C& c, int X, const typename C::Vq& v, int n) {
  const int tag = c.tag;
  if (n > 1) c.tag = 0;
  const typename C::Vq r = node_push_set<0>(c, X, v);
  c.tag = tag;
  return r;
}

template <class C>
WSI_HD void chain_push_empty(C& c, const int* arcs, int k0, int n, int X) {
  constexpr int V = C::V;
  const WsiModelDesc& d = c.D();
  const typename C::Vq none = vq_zero<C::P>();
  const int g = d.node_gidx[X];
  switch (d.node_type[X]) {
    case WSI_NT_RIVER:
      tank_push_force(c, c.tank(WSI_TANK_STORAGE(V), (int)*c.river(WSI_RIVER_TANK, g)), c.s_tank(), none);
      break;
    case WSI_NT_GROUNDWATER:
    case WSI_NT_STORAGE:
    case WSI_NT_RESERVOIR:
      tank_push_force(c, c.tank(WSI_TANK_STORAGE(V), g), c.s_tank(), none);
      break;
    case WSI_NT_QUEUEGROUNDWATER:
    case WSI_NT_SEWER: {  // once per push_storage call: one, or one per time-area entry (push_set_land / _timearea)
      int reps = 1;
      if constexpr (!C::kLean)
        if ((c.tag == 1 && n == 1) || d.node_type[X] == WSI_NT_QUEUEGROUNDWATER) reps = (int)*c.sewer(WSI_SEWER_N_TIMEAREA, g);
      for (int i = 0; i < reps; i++) tank_push_force(c, c.sewer(WSI_SEWER_STORAGE(V), g), c.s_sewer(), none);
      break;
    }
    case WSI_NT_WWTW:  // current_input = sum_vqip(current_input, empty) wtw.py:359
      if constexpr (!C::kLean) tank_push_force(c, c.works(WSI_WORKS_CURRENT_INPUT(V), g), c.s_works(), none);
      break;
    case WSI_NT_WASTE:
      break;
    default:
      c.flag(X, WSI_FLAG_DENY);
      break;
  }
#pragma unroll 1
  for (int k = k0; k < n; k++)  // an arc nothing was booked on this step is empty: nothing to re-blend
    if (arc_live(c, arcs[k])) tank_push_force(c, c.arc(WSI_ARC_VQIP_IN(V), arcs[k]), c.s_arc(), none);
}

template <class C>
WSI_HD bool arc_send(C& c, int a, const typename C::Vq& vqip, typename C::Vq& reply) {
  using Vq = typename C::Vq;
  const int* arcs = nullptr;
  const int n = chain_of(c, a, arcs);
  const double v0 = vqip.x[0];
  if (n > 0) {
    double pipe[kMaxChain], E[kMaxChain];
    chain_excess(c, arcs, n, pipe, E);
    const WsiModelDesc& d = c.D();
    const int X = d.arc_out[arcs[n - 1]];
    // Top-down: what every arc admits of the request (the receiver itself, or a junction handing
    // the question down -- check_basic nodes.py:480-508: min(excess of its arc, request)).  The
    // request is unchanged on its way down while all of it is admitted; `j` is the first arc that
    // does not admit everything (arcs.py:126-133), below it travels w = vqip - not_pushed.
    int j = -1;
    bool ok = true, room = true, below_empty = false;
    double cv = v0, ratio = 0.0;
#pragma unroll
    for (int k = 0; k < kMaxChain; k++) {
      if (k < n) {
        double ne;
        if (k == n - 1) {
          ne = node_push_check_vol<0>(c, X, true, cv);
        } else {
          const double below = E[k + 1 < kMaxChain ? k + 1 : 0];
          ne = clamp_to(below, cv);
        }
        const double excess_in = clamp_to(ne, pipe[k]);
        room = room && excess_in >= 0.0;
        if (!below_empty && pymax(cv - excess_in, 0.0) != 0.0) {
          if (j >= 0 || !(cv > 0)) {
            ok = false;  // a second arc binds: general code
          } else {
            j = k;
            ratio = div_z(pymax(cv - excess_in, 0.0), cv);  // v_change_vqip(vqip, not_pushed volume)
            cv = cv - cv * ratio;
            if (!(cv > 0)) {  // nothing admitted: what travels on is empty iff the ratio is exactly 1
              below_empty = true;
              ok = ok && ratio == 1.0 && cv == 0.0;
            }
          }
        }
      }
    }
    if (j < 0 && v0 > 0) {
      WSI_STAT(0);
      reply = chain_push_set(c, X, vqip, n);
      if (vq_is_empty(c, reply)) {
#pragma unroll 1
        for (int k = 0; k < n; k++) arc_book(c, arcs[k], vqip);
        reply = vq_zero<C::P>();
        return true;
      }
      // the receiver handed something back: every arc on the way books what stayed below it
      // and adds its own (empty) not_pushed to the reply, deepest arc first (arcs.py:140-150)
      const Vq not_pushed = c.vchange(vqip, 0.0);
#pragma unroll 1
      for (int k = n - 1; k >= 0; k--) {
        const Vq sent = c.extract(vqip, reply);
        reply = c.sum(reply, not_pushed);
        arc_book(c, arcs[k], sent);
      }
      return false;
    }
    if (j < 0 && room && vq_is_empty(c, vqip)) {  // a dry store's release
      WSI_STAT(1);
      chain_push_empty(c, arcs, 0, n, X);
      reply = vq_zero<C::P>();
      return true;
    }
    if (j >= 0 && ok && room && v0 > 0) {
      // Arc j admits only part of the request: not_pushed_j = v_change_vqip(vqip, np_j), and
      // w = vqip - not_pushed_j travels on (fully admitted below, or empty when nothing fits).
      WSI_STAT(7);
      Vq w;
      w.x[0] = cv;
#pragma unroll
      for (int p = 1; p <= C::P; p++) w.x[p] = bit(c.addm(), p) ? vqip.x[p] - vqip.x[p] * ratio : vqip.x[p];
      if (below_empty) {
        chain_push_empty(c, arcs, j + 1, n, X);
        reply = vq_zero<C::P>();
      } else {
        reply = chain_push_set(c, X, w, n);
        const bool rempty = vq_is_empty(c, reply);
        const Vq np_w = c.vchange(w, 0.0);
#pragma unroll 1
        for (int k = n - 1; k > j; k--) {
          if (rempty) {
            arc_book(c, arcs[k], w);
          } else {
            const Vq sent = c.extract(w, reply);
            reply = c.sum(reply, np_w);
            arc_book(c, arcs[k], sent);
          }
        }
      }
      {
        Vq np_j;  // v_change_vqip(vqip, np_j): the same ratio as above
        np_j.x[0] = v0 * ratio;
#pragma unroll
        for (int p = 1; p <= C::P; p++) np_j.x[p] = bit(c.addm(), p) ? vqip.x[p] * ratio : vqip.x[p];
        const Vq sent = c.extract(w, reply);
        reply = c.sum(reply, np_j);
        arc_book(c, arcs[j], sent);
      }
      const Vq np_0 = c.vchange(vqip, 0.0);
#pragma unroll 1
      for (int k = j - 1; k >= 0; k--) {
        const Vq sent = c.extract(vqip, reply);
        reply = c.sum(reply, np_0);
        arc_book(c, arcs[k], sent);
      }
      return false;
    }
  }
  WSI_STAT(2);
  if (n <= 0) WSI_STAT(3);
  else if (v0 > 0) WSI_STAT(4);       // several arcs bind
  else if (v0 == 0) WSI_STAT(5);      // zero volume, not empty (or no room)
  else WSI_STAT(6);
  int viol = 0;
  reply = arc_send_general_call<C>(c, a, vqip, &viol);
  c.viol = viol;
  return false;
}

// Arc.get_excess("push")["volume"] without a request, arcs.py:230-257
template <class C>
WSI_HD double arc_excess_free(C& c, int a) {
  const int* arcs = nullptr;
  const int n = chain_of(c, a, arcs);
  if (n > 0) {
    double pipe[kMaxChain], E[kMaxChain];
    chain_excess(c, arcs, n, pipe, E);
    return E[0];
  }
  return arc_excess_vol<kJunctionDepth>(c, a, false, 0.0);
}

// Node.push_distributed nodes.py:409-478 as the schedule entries call it; returns true when the
// reply is empty (`reply` all zeros).  One candidate arc (or one out-arc) is handled here with
// scalars; several candidates go to the general form.
template <class C>
WSI_HD bool push_out(C& c, int node, const typename C::Vq& vqip, int filter, typename C::Vq& reply) {
  using Vq = typename C::Vq;
  const WsiModelDesc& d = c.D();
  const int cb = d.cand_ptr[node * WSI_NF + filter];
  const int nc = d.cand_ptr[node * WSI_NF + filter + 1] - cb;
  const int n_out = d.out_ptr[node + 1] - d.out_ptr[node];
  const bool single = n_out == 1;  // nodes.py:424-438
  if (nc <= 0) {  // nobody to push to (single: nodes.py:436-438; else the loop never runs)
    reply = vqip;
    return false;
  }
  if (!single && nc > 1) {
    int viol = 0;
    reply = push_general_call<C>(c, node, vqip, filter, &viol);
    c.viol = viol;
    return false;
  }
  const int arc = single ? d.out_arcs[d.out_ptr[node]] : d.cand_arcs[cb];
  double avail = 0.0, priority = 1.0, alloc = 1.0, pref = 1.0;
  double not_pushed = vqip.x[0];
  if (!single) {  // get_connected nodes.py:271-317 for the one candidate
    double a = arc_excess_free(c, arc);
    if (a < c.FA) a = 0.0;
    pref = *c.arc(WSI_ARC_PREFERENCE, arc);
    avail = 0.0 + a;
    priority = 0.0 + a * pref;
    alloc = a * pref;
    if (not_pushed > avail) {
      priority = avail;
      alloc = a;
    }
  }
  reply = vqip;  // not_pushed_
  int iter = 0;
#pragma unroll 1
  while (single || (not_pushed > c.FA && avail > c.FA && iter < kMaxIter)) {
    const double amount = pymin(avail, not_pushed);
    const Vq to_send = single ? vqip : c.vchange(reply, amount * alloc / priority);
    Vq back;
    const bool empty = arc_send(c, arc, to_send, back);
    if (single) {
      reply = back;
      return empty;
    }
    if (empty) {
      reply = c.extract(reply, to_send);  // sent == to_send
    } else {
      const Vq sent = c.extract(to_send, back);
      reply = c.extract(reply, sent);
    }
    not_pushed = reply.x[0];
    iter++;
    if (!(not_pushed > c.FA)) break;
    double a = arc_excess_free(c, arc);
    if (a < c.FA) a = 0.0;
    avail = 0.0 + a;
    priority = 0.0 + a * pref;
    alloc = a * pref;
  }
  if (iter == kMaxIter) c.flag(node, WSI_FLAG_MAXITER);
  return false;
}

// =========================================================================================
// Mass-balance bookkeeping
// =========================================================================================
// (in, ds, out) of one object checked as WSIObj.mass_balance does (core.py:403-427);
// optionally recorded and accumulated into the per-thread system totals.
// 10 ** int(log10(x)) of core.py:410 without transcendental calls or table look-ups (a table
// indexed by a per-thread exponent serialises in the constant cache).  10^k for 0 <= k <= 22 is
// exact in binary64 and is built from exact products; int() truncates towards zero, so the
// magnitude is 10^floor(log10 x) for x >= 1 and 1/10^j with the smallest 10^-j >= x below 1
// (Python evaluates 10 ** -j as the correctly rounded float, which 1.0 / 10^j reproduces).
WSI_HD double pow10i(int k) {  // 0 <= k <= 22
  double p = 1.0;
  if (k & 1) p *= 1e1;
  if (k & 2) p *= 1e2;
  if (k & 4) p *= 1e4;
  if (k & 8) p *= 1e8;
  if (k & 16) p *= 1e16;
  return p;
}
inline WSI_HDN double magnitude10(double x) {
  if (!(x > 1e-22 && x < 1e22)) return pow(10.0, (double)(int)log10(x));
  int e2;
  (void)frexp(x, &e2);  // x = m * 2^e2, m in [0.5, 1)
  if (x >= 1.0) {
    int k = (int)((double)(e2 - 1) * 0.30102999566398120);  // <= floor(log10 x), off by at most 1
    if (k < 0) k = 0;
    if (k < 22 && pow10i(k + 1) <= x) k++;
    return pow10i(k);
  }
  // x < 1: j = number of leading decimal zeros, magnitude 10^-j with 10^-j >= x > 10^-(j+1)
  int j = (int)((double)(-e2) * 0.30102999566398120);  // <= true j, off by at most 1
  if (j < 0) j = 0;
  if (j < 22 && 1.0 / pow10i(j + 1) >= x) j++;
  return 1.0 / pow10i(j);
}

// one lane of WSIObj.mass_balance's test, core.py:404-419: |in/m - ds/m - out/m| > FLOAT_ACCURACY
// with m = 10**int(log10(largest)) (m = 1 when largest <= FLOAT_ACCURACY).  m lies in (x/10, x] for
// x >= 1 and in [x, 10x) below 1, so most lanes are decided from those bounds; the exact power of
// ten (and the three divisions) are only evaluated when the residual is within 1 % of a possible
// threshold.
template <class C>
WSI_HD bool mb_lane_bad(const C& c, double a, double dd, double o) {
  const double largest = pymax(pymax(a, o), dd);
  const double resid = fabs(a - dd - o);
  if (!(largest > c.FA)) return resid > c.FA;
  const double m_lo = largest >= 1.0 ? largest * 0.1 : largest;
  const double m_hi = largest >= 1.0 ? largest : largest * 10.0;
  if (c.FA >= 1e-12 && resid < 0.99 * c.FA * m_lo) return false;
  if (c.FA >= 1e-12 && resid > 1.01 * c.FA * m_hi) return true;
  const double magnitude = magnitude10(largest);
  return fabs(a / magnitude - dd / magnitude - o / magnitude) > c.FA;
}

template <class C>
WSI_HD void mb_check(C& c, int node, const typename C::Vq& in_, const typename C::Vq& ds_,
                     const typename C::Vq& out_) {
  constexpr int V = C::V;
  int viol = 0;
#pragma unroll
  for (int p = 0; p <= C::P; p++) {
    if (p != 0 && !bit(c.addm(), p)) continue;
    if (mb_lane_bad(c, in_.x[p], ds_.x[p], out_.x[p])) viol++;
  }
  if (viol) {
    c.flag(node, WSI_FLAG_MASS_BALANCE);
    c.viol += viol;
  }
  if (c.mb_acc) {
#pragma unroll
    for (int p = 0; p <= C::P; p++) {
      if (p != 0 && !bit(c.addm(), p)) continue;
      c.mb_acc[(0 * V + p) * c.mb_stride] += in_.x[p];
      c.mb_acc[(1 * V + p) * c.mb_stride] += ds_.x[p];
      c.mb_acc[(2 * V + p) * c.mb_stride] += out_.x[p];
    }
  }
  const WsiModelDesc& d = c.D();
  const int row = c.t - d.rec_t0;
  if (d.record_mb && row >= 0 && row < d.rec_T) {
    double* o = d.out_mb + ((size_t)row * d.n_nodes + node) * 3 * V * c.R + c.r;
#pragma unroll
    for (int p = 0; p < V; p++) {
      o[(size_t)(0 * V + p) * c.R] = in_.x[p];
      o[(size_t)(1 * V + p) * c.R] = ds_.x[p];
      o[(size_t)(2 * V + p) * c.R] = out_.x[p];
    }
  }
}

// (in, ds, out) of a node whose terms are formed inside its schedule op (Land, Catchment) are
// handed to the mass-balance pass through mb_terms[slot][3][V][R]
template <class C>
WSI_HD void store_mb_terms(const C& c, int node, const typename C::Vq& in_, const typename C::Vq& ds_,
                           const typename C::Vq& out_) {
  const WsiModelDesc& d = c.D();
  double* p = d.mb_terms + (size_t)d.mb_term_slot[node] * 3 * C::V * c.R + c.r;
  c.stv_add(p, c.R, in_);
  c.stv_add(p + (size_t)C::V * c.R, c.R, ds_);
  c.stv_add(p + (size_t)2 * C::V * c.R, c.R, out_);
}
template <class C>
WSI_HD void load_mb_terms(const C& c, int node, typename C::Vq& in_, typename C::Vq& ds_, typename C::Vq& out_) {
  const WsiModelDesc& d = c.D();
  const double* p = d.mb_terms + (size_t)d.mb_term_slot[node] * 3 * C::V * c.R + c.r;
  in_ = c.ldv_add(p, c.R);
  ds_ = c.ldv_add(p + (size_t)C::V * c.R, c.R);
  out_ = c.ldv_add(p + (size_t)2 * C::V * c.R, c.R);
}

// Node.total_in / total_out nodes.py:104-132, volume + additive lanes (the non-additive lanes
// of a mass-balance term are never compared, core.py:400-404, so they are not blended here)
template <class C>
WSI_HD typename C::Vq total_in(const C& c, int node) {
  const WsiModelDesc& d = c.D();
  typename C::Vq v = vq_zero<C::P>();
  for (int i = d.in_ptr[node]; i < d.in_ptr[node + 1]; i++) {
    const typename C::Vq a = arc_vqip_add(c, d.in_arcs[i]);
#pragma unroll
    for (int p = 0; p <= C::P; p++) v.x[p] += a.x[p];
  }
  return v;
}
template <class C>
WSI_HD typename C::Vq total_out(const C& c, int node) {
  const WsiModelDesc& d = c.D();
  typename C::Vq v = vq_zero<C::P>();
  for (int i = d.out_ptr[node]; i < d.out_ptr[node + 1]; i++) {
    const typename C::Vq a = arc_vqip_add(c, d.out_arcs[i]);
#pragma unroll
    for (int p = 0; p <= C::P; p++) v.x[p] += a.x[p];
  }
  return v;
}
// ds_ accumulation of mass_balance (core.py:397-401): volume + additive only
template <class C>
WSI_HD void ds_add(const C& c, typename C::Vq& acc, const typename C::Vq& v) {
#pragma unroll
  for (int p = 0; p <= C::P; p++)
    if (p == 0 || bit(c.addm(), p)) acc.x[p] += v.x[p];
}

// =========================================================================================
// Decay (core.py:474-521) on a register-resident storage; decay slots of a group start at
// plane `f_n` (n_decay, pol[4], const[4], exp[4], total[4], pre[4])
// =========================================================================================
template <class C, class PlaneFn>
WSI_HD void decay_end_timestep(const C& c, typename C::Vq& st, PlaneFn plane, int f_n,
                               double temperature) {
  const int n = (int)*plane(f_n);
  for (int i = 0; i < WSI_MAX_DECAY; i++) {
    if (i >= n) break;
    const int pol = (int)*plane(f_n + 1 + i);
    const double k = *plane(f_n + 1 + WSI_MAX_DECAY + i);
    const double e = *plane(f_n + 1 + 2 * WSI_MAX_DECAY + i);
    const double x = vget<C::P>(st, pol);
    const double f = k * pow(e, temperature - 20.0);  // constants.DECAY_REFERENCE_TEMPERATURE
    const double diff = x * pymin(f, 1.0);
    *plane(f_n + 1 + 4 * WSI_MAX_DECAY + i) = x;  // storage_[pol] (tanks.py:445)
    vset<C::P>(st, pol, x - diff);
    // total_decayed = sum_vqip(empty, diff): only additive lanes accumulate (core.py:491)
    *plane(f_n + 1 + 3 * WSI_MAX_DECAY + i) = bit(c.addm(), pol) ? 0.0 + diff : 0.0;
  }
}
// total_decayed as a VQIP (zero volume)
template <class C, class PlaneFn>
WSI_HD typename C::Vq decay_total(const C& c, PlaneFn plane, int f_n) {
  typename C::Vq td = vq_zero<C::P>();
  const int n = (int)*plane(f_n);
  for (int i = 0; i < WSI_MAX_DECAY; i++) {
    if (i >= n) break;
    vset<C::P>(td, (int)*plane(f_n + 1 + i), *plane(f_n + 1 + 3 * WSI_MAX_DECAY + i));
  }
  return td;
}
template <class PlaneFn>
WSI_HD double decay_total_of(PlaneFn plane, int f_n, int pol) {
  const int n = (int)*plane(f_n);
  double r = 0.0;
  for (int i = 0; i < WSI_MAX_DECAY; i++) {
    if (i >= n) break;
    if ((int)*plane(f_n + 1 + i) == pol) r = *plane(f_n + 1 + 3 * WSI_MAX_DECAY + i);
  }
  return r;
}

}  // namespace wsi

#include "wsi_land.cuh"
#include "wsi_nodes.cuh"
#include "wsi_net.cuh"
