// The sm_100a kernels of libwsi_b200.so and their launch sequence, as templates over the pollutant
// configuration; instantiated once per configuration by wsi_cfg_*.cu (WSI_DEFINE_CONFIG_UNIT).
//
// Kernels
//   wsi_surface_kernel<CFG,TYPE>  one thread per (surface, replica): Surface.run of one
//                         surface class (wsi_land.cuh surface_phase); one launch per class.
//   wsi_step_kernel<CFG>  one thread per replica: the orchestration schedule (routing) with the
//                         Land / Catchment mass balances fused in.
//   wsi_mb_kernel<CFG>    one thread per (node | arc, replica): remaining mass balances; per-block
//                         partials (violation count, system totals) via shared memory / shuffles.
//   wsi_mb_finalize       fixed-order reduction of the per-block partials -> mb_sums[t].
//   wsi_record_kernel     one thread per (probe, replica): recording is a strided copy.
//   wsi_end_kernel        one thread per (end_timestep item, replica).
//   wsi_net_kernel / wsi_net_tail_kernel   network engine (desc.engine == 1): one thread per (river
//                         node, replica), one launch per upstream-depth level; the narrow main-stem
//                         levels in a single block.  Replaces wsi_step_kernel for one large
//                         dendritic network; mb / record / end then use a flat (item, replica) index.
//
// The path is HBM-bound fp64 element-wise work: no tensor cores.  Grid sizing,
// coalescing and the byte accounting are described in DESIGN.md.
#pragma once
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "wsi_host.h"
#include "wsi_device.cuh"
#include "wsi_warp.cuh"

using namespace wsi;
using namespace wsi_host;

namespace {

// Minimum resident blocks per SM requested from ptxas (caps registers per thread): the kernels
// are latency-bound, so occupancy matters more than avoiding a few spills.
#ifndef WSI_SURF_MINB
#define WSI_SURF_MINB 6  // 80 registers: the merged launch prefers threads in flight (gpurun sweep r2j/r2k: 3 -> 0.134, 4 -> 0.128, 5 -> 0.138, 6 -> 0.119, 7 -> 0.123, 8 -> 0.133 ms)
#endif
#ifndef WSI_STEP_MINB
#define WSI_STEP_MINB 3
#endif
#ifndef WSI_OBJ_MINB
#define WSI_OBJ_MINB 6
#endif
#ifndef WSI_MB_MINB  // mass balance / record / end_timestep: streaming kernels
#define WSI_MB_MINB 5  // 96 registers: 6 (80) spills more, 4 (128) starves the end_timestep copy (gpurun sweep, round 1)
#endif
constexpr size_t kDescSmem = sizeof(WsiModelDesc);  // dynamic shared memory: the staged descriptor

// Work-item decode of the object-parallel kernels.  Replica engine (FLAT = false): blockIdx.y is
// the item and blockIdx.x tiles the replicas, so every block is homogeneous in item.  Network
// engine (FLAT = true): one flat index over (item, replica) with the replica fastest, so a graph
// with few replicas and a million nodes still fills the machine (and gridDim.y's 65535 limit
// does not apply).
// The timestep a kernel works on: its argument, or -- inside a captured CUDA graph, which is the
// same for every step -- the device-side step counter (net_scratch[2], advanced by
// wsi_advance_kernel at the end of each step).
__device__ __forceinline__ int step_of(const WsiModelDesc& d, int t_arg) {
  return t_arg >= 0 ? t_arg : d.net_scratch[2];
}

template <bool FLAT>
__device__ __forceinline__ bool decode_item(const WsiModelDesc& d, int n_items, int& item, int& r) {
  if constexpr (FLAT) {
    const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    item = (int)(gid / (size_t)d.R);
    r = (int)(gid % (size_t)d.R);
    return item < n_items;
  } else {
    item = blockIdx.y;
    r = blockIdx.x * blockDim.x + threadIdx.x;
    return r < d.R;
  }
}

// One thread per (surface slot, replica).  blockIdx.y selects the slot inside the surface
// class being launched, so every block (and warp) is homogeneous in surface class and reads
// 256 contiguous bytes per plane.
template <class CFG, int TYPE>
__global__ void __launch_bounds__(kSurfBlock, WSI_SURF_MINB)
    wsi_surface_kernel(const int slot, const CFG cfg, const int t_arg, const uint32_t flags) {
  const WsiModelDesc& d = g_wsi_desc[slot];
  const int t = step_of(d, t_arg);
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= d.R) return;
  const int u = d.surf_order[d.surf_class_ptr[TYPE] + blockIdx.y];
  Ctx<CFG> c = make_ctx(nullptr, slot, cfg, r, t, flags);
  surface_phase<TYPE>(c, u);
}

// All surface classes in ONE launch: blockIdx.y walks surf_order backwards, so the blocks of the
// heaviest class (GrowingSurface, last in the order) are scheduled first and the light classes fill
// the tail of the launch; every block is still homogeneous in class.
template <class CFG>
__global__ void __launch_bounds__(kSurfBlock, WSI_SURF_MINB)
    wsi_surfaces_kernel(const int slot, const CFG cfg, const int t_arg, const uint32_t flags) {
  const WsiModelDesc& d = g_wsi_desc[slot];
  const int t = step_of(d, t_arg);
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= d.R) return;
  const int k = (int)gridDim.y - 1 - (int)blockIdx.y;
  const int u = d.surf_order[k];
  Ctx<CFG> c = make_ctx(nullptr, slot, cfg, r, t, flags);
  if (k >= d.surf_class_ptr[WSI_ST_GROWINGSURFACE]) surface_phase<WSI_ST_GROWINGSURFACE>(c, u);
  else if (k >= d.surf_class_ptr[WSI_ST_PERVIOUSSURFACE]) surface_phase<WSI_ST_PERVIOUSSURFACE>(c, u);
  else if (k >= d.surf_class_ptr[WSI_ST_IMPERVIOUSSURFACE]) surface_phase<WSI_ST_IMPERVIOUSSURFACE>(c, u);
  else surface_phase<WSI_ST_SURFACE>(c, u);
}

// One thread per replica: the orchestration schedule (Land.run after its surfaces, Groundwater,
// Sewer, River, Catchment ops and river routing, in the reference's order).  Land / Catchment
// form their mass-balance terms here (everything is still at hand) and leave them in mb_terms.
// Routing is sequential inside a replica by construction.
// Block-level sink of the schedule kernel's fused mass balance: violations via one atomic per warp,
// the replica's system-total contributions via a butterfly over the warp (every lane ends with the
// same sum, fixed order) into one shared-memory row per warp.
template <int V>
struct MbBlockSink {
  double (*s_w)[3 * V];
  int viol;
  template <class C>
  __device__ __forceinline__ void take(C& c, int lane, int v, double a, double b, double o) {
    viol += v;
#pragma unroll
    for (int k = 16; k > 0; k >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, k);
      b += __shfl_xor_sync(0xffffffffu, b, k);
      o += __shfl_xor_sync(0xffffffffu, o, k);
    }
    if ((threadIdx.x & 31) == 0) {
      double* row = s_w[threadIdx.x >> 5];
      row[0 * V + lane] = a;
      row[1 * V + lane] = b;
      row[2 * V + lane] = o;
    }
  }
};

// One thread per replica: the orchestration schedule (Land.run after its surfaces, Groundwater,
// Sewer, River, Catchment ops and river routing, in the reference's order), then -- fused tail --
// the replica's mass balances (lane by lane, wsi_nodes.cuh), recording and end_timestep: the
// thread reads back what it has just written (L1 / L2) instead of three more kernels re-streaming
// the state from HBM.  Routing is sequential inside a replica by construction.
template <class CFG, bool LEAN>
__global__ void __launch_bounds__(kBlock, WSI_STEP_MINB)
    wsi_step_kernel(const int slot, const CFG cfg, const int t_arg, const int sched_begin, const int sched_end,
                    const uint32_t flags, const int row_arg, const int fused) {
  constexpr int V = CFG::P + 1;
  constexpr int kWarps = kBlock / 32;
  __shared__ double s_w[kWarps][3 * V];
  wsi_stage_desc(slot);
  const WsiModelDesc& d = g_wsi_desc[slot];
  const int t = step_of(d, t_arg);
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = r < d.R;
  Ctx<CFG, true, LEAN> c = make_ctx<true, LEAN>(nullptr, slot, cfg, active ? r : 0, t, flags);
  c.block_sync = (blockIdx.x + 1) * blockDim.x <= (unsigned)d.R;  // full block: all threads walk the schedule
  if (flags & WSI_RUN_SCHEDULE) {
    for (int k = sched_begin; k < sched_end; k++) {
      if (active) run_op(c, d.sched_op[k], d.sched_node[k]);
      // The kernel stalls on instruction fetch more than on anything else (ncu: no_instruction):
      // a warp walks ~14k instructions.  Re-aligning the block's 8 warps after every schedule
      // entry keeps them in the same stretch of code.
#ifndef WSI_NO_OPSYNC
      __syncthreads();
#endif
    }
  }
  // recording: the thread copies what it has just written (recording precedes end_timestep,
  // model.py:1063-1176, and the settle pass below must not race with it)
  if (active && (flags & WSI_RUN_RECORD)) {
    const int row = t_arg >= 0 ? row_arg : t - d.rec_t0;
    for (int k = 0; k < d.n_probe + d.n_red; k++) record_item(c, row, k);
  }
  if (!fused) return;  // wsi_settle_kernel runs mass balance and end_timestep one thread per (node, replica)
  if (flags & WSI_RUN_MASS_BALANCE) {
    const bool want_sums = d.mb_sums != nullptr;
    MbBlockSink<V> sink = {s_w, 0};
    replica_mb_lanes<0>(c, active, want_sums, sink);
    int v = sink.viol;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0 && v != 0 && d.mb_count) atomicAdd(&d.mb_count[t], v);
    if (want_sums) {
      __syncthreads();
      if (threadIdx.x < 3 * V) {
        const int p = threadIdx.x % V;
        double acc = 0.0;
        if (p == 0 || bit(c.addm(), p))
          for (int w = 0; w < kWarps; w++) acc += s_w[w][threadIdx.x];
        d.mb_partial[(size_t)blockIdx.x * 3 * V + threadIdx.x] = acc;
      }
    }
  }
  if (!active) return;
  if (flags & WSI_RUN_END_TIMESTEP)
    for (int i = 0; i < d.n_end; i++) end_item(c, d.end_kind[i], d.end_obj[i]);
}

// One thread per (node or arc, replica): WSIObj.mass_balance (core.py:370-429) of every node and
// the system totals of model.py:1016-1036.  Block-reduced: violations via warp shuffle + one
// atomic per warp; the (in, ds, out) lanes via warp shuffles, a fixed cross-warp sum, and one
// row of mb_partial per block (summed in block order by wsi_mb_finalize: deterministic).
template <class CFG, bool FLAT>
__global__ void __launch_bounds__(kObjBlock, WSI_MB_MINB)
    wsi_mb_kernel(const int slot, const CFG cfg, const int t_arg, const uint32_t flags, const int n_items) {
  constexpr int V = CFG::P + 1;
  constexpr int kWarps = kObjBlock / 32;
  __shared__ double s_w[kWarps][3 * V];
  const WsiModelDesc& d = g_wsi_desc[slot];
  const int t = step_of(d, t_arg);
  int item, r;
  const bool active = decode_item<FLAT>(d, n_items, item, r);
  const bool want_sums = d.mb_sums != nullptr;
  Ctx<CFG> c = make_ctx(nullptr, slot, cfg, active ? r : 0, t, flags);
  using Vq = typename Ctx<CFG>::Vq;
  Vq term[3];
  term[0] = vq_zero<CFG::P>();
  term[1] = vq_zero<CFG::P>();
  term[2] = vq_zero<CFG::P>();
  if (active) {
    if (item < d.n_nodes) {
      node_mass_balance(c, item, term[0], term[1], term[2]);
    } else {  // arcs are trivially balanced; they only feed the system totals
      term[0] = arc_vqip_add(c, item - d.n_nodes);
      term[2] = term[0];
    }
  }
  int v = active ? c.viol : 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0 && v != 0 && d.mb_count) atomicAdd(&d.mb_count[t], v);
  if (want_sums) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int k = 0; k < 3; k++) {
      double a[16];
#pragma unroll
      for (int p = 0; p < 16; p++) a[p] = (p < V && active && (p == 0 || bit(c.addm(), p))) ? term[k].x[p < V ? p : 0] : 0.0;
      int p_mine;
      const double x = warp_sum_16(a, lane, p_mine);
      if ((lane & 1) == 0 && p_mine < V && (p_mine == 0 || bit(c.addm(), p_mine))) s_w[warp][k * V + p_mine] = x;
    }
    __syncthreads();
    if (threadIdx.x < 3 * V) {
      const int p = threadIdx.x % V;
      double acc = 0.0;
      if (p == 0 || bit(c.addm(), p))
        for (int w = 0; w < kWarps; w++) acc += s_w[w][threadIdx.x];
      d.mb_partial[((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 3 * V + threadIdx.x] = acc;
    }
  }
}

// Replica engine, large ensembles: one thread per (node, replica) checks the node's mass balance
// (core.py:370-429; its in-arcs also stand for the arcs' own -- trivially balanced -- terms in the
// system totals of model.py:1016-1036: every arc is the in-arc of exactly one node) and then ends
// the timestep of what the node owns (model.py:1172-1176).  Arcs need no pass of their own: their
// transients expire by stamp (wsi_device.cuh).  Block-reduced like wsi_mb_kernel.
template <class CFG>
__global__ void __launch_bounds__(kObjBlock, WSI_MB_MINB)
    wsi_settle_kernel(const int slot, const CFG cfg, const int t_arg, const uint32_t flags) {
  constexpr int V = CFG::P + 1;
  constexpr int kWarps = kObjBlock / 32;
  __shared__ double s_w[kWarps][3 * V];
  const WsiModelDesc& d = g_wsi_desc[slot];
  const int t = step_of(d, t_arg);
  int node, r;
  const bool active = decode_item<false>(d, d.n_nodes, node, r);
  Ctx<CFG> c = make_ctx(nullptr, slot, cfg, active ? r : 0, t, flags);
  if (flags & WSI_RUN_MASS_BALANCE) {
    using Vq = typename Ctx<CFG>::Vq;
    const bool want_sums = d.mb_sums != nullptr;
    Vq term[3];
    term[0] = vq_zero<CFG::P>();
    term[1] = vq_zero<CFG::P>();
    term[2] = vq_zero<CFG::P>();
    if (active) {
      node_mass_balance(c, node, term[0], term[1], term[2]);
      if (want_sums) {
        const Vq tin = total_in(c, node);
        ds_add(c, term[0], tin);
        ds_add(c, term[2], tin);
      }
    }
    int v = active ? c.viol : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0 && v != 0 && d.mb_count) atomicAdd(&d.mb_count[t], v);
    if (want_sums) {
      const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
      for (int k = 0; k < 3; k++) {
        double a[16];
#pragma unroll
        for (int p = 0; p < 16; p++) a[p] = (p < V && active && (p == 0 || bit(c.addm(), p))) ? term[k].x[p < V ? p : 0] : 0.0;
        int p_mine;
        const double x = warp_sum_16(a, lane, p_mine);
        if ((lane & 1) == 0 && p_mine < V && (p_mine == 0 || bit(c.addm(), p_mine))) s_w[warp][k * V + p_mine] = x;
      }
      __syncthreads();
      if (threadIdx.x < 3 * V) {
        const int p = threadIdx.x % V;
        double acc = 0.0;
        if (p == 0 || bit(c.addm(), p))
          for (int w = 0; w < kWarps; w++) acc += s_w[w][threadIdx.x];
        d.mb_partial[((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 3 * V + threadIdx.x] = acc;
      }
    }
  }
  if (active && (flags & WSI_RUN_END_TIMESTEP)) end_node(c, node);
}

// One thread per (probe, replica): Model.run's recording (model.py:1063-1170) is a strided copy.
template <class CFG, bool FLAT>
__global__ void __launch_bounds__(kObjBlock, WSI_MB_MINB)
    wsi_record_kernel(const int slot, const CFG cfg, const int t_arg, const int row_arg) {
  const WsiModelDesc& d = g_wsi_desc[slot];
  const int t = step_of(d, t_arg);
  const int row = t_arg >= 0 ? row_arg : t - d.rec_t0;
  int item, r;
  if (!decode_item<FLAT>(d, d.n_probe + d.n_red, item, r)) return;
  Ctx<CFG> c = make_ctx(nullptr, slot, cfg, r, t, 0u);
  record_item(c, row, item);
}

// One thread per (end_timestep item, replica): arcs reset, DecayTank decay, QueueTank shift.
template <class CFG, bool FLAT>
__global__ void __launch_bounds__(kObjBlock, WSI_MB_MINB)
    wsi_end_kernel(const int slot, const CFG cfg, const int t_arg) {
  const WsiModelDesc& d = g_wsi_desc[slot];
  const int t = step_of(d, t_arg);
  int item, r;
  if (!decode_item<FLAT>(d, d.n_end, item, r)) return;
  Ctx<CFG> c = make_ctx(nullptr, slot, cfg, r, t, 0u);
  end_item(c, d.end_kind[item], d.end_obj[item]);
}

// Network engine: one thread per (node of a work list, replica) -- River.calculate_discharge of
// every river, one upstream-depth level of the routing, or the final deliveries (wsi_net.cuh).
template <class CFG>
__global__ void __launch_bounds__(kObjBlock, WSI_OBJ_MINB)
    wsi_net_kernel(const int slot, const CFG cfg, const int t_arg, const uint32_t flags, const int kind, const int begin,
                   const int count) {
  const WsiModelDesc& d = g_wsi_desc[slot];
  const int t = step_of(d, t_arg);
  int item, r;
  if (!decode_item<true>(d, count, item, r)) return;
  Ctx<CFG> c = make_ctx(nullptr, slot, cfg, r, t, flags);
  net_item(c, kind, begin + item);
}
// Mid-width levels: a group of kLanes threads per (river, replica) (wsi_net.cuh, lane-parallel
// form) -- fewer instructions on the critical path of the level than one thread per river.
constexpr int kLaneBlock = 256;
constexpr size_t kLaneLevelMax = 32768;  // (river, replica) items per level up to which lane groups are used
template <class CFG>
__global__ void __launch_bounds__(kLaneBlock)
    wsi_net_lane_kernel(const int slot, const CFG cfg, const int t_arg, const uint32_t flags, const int begin,
                        const int count) {
  const WsiModelDesc& d = g_wsi_desc[slot];
  const int t = step_of(d, t_arg);
  const size_t gid = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) / kLanes;
  const int q = threadIdx.x % kLanes;
  const unsigned gmask = 0xFFFFu << (threadIdx.x & 16);
  const int item = (int)(gid / (size_t)d.R);
  if (item >= count) return;  // whole groups leave together
  using Cx = Ctx<CFG>;
  Cx c = make_ctx(nullptr, slot, cfg, (int)(gid % (size_t)d.R), t, flags);
  LaneItem<Cx> it = lane_item_load(c, net_rec_load(c, begin + item), q);
  LaneOut<Cx> out;
  const bool general = net_lanes_release(c, it, q, gmask, out);
  if (!general) net_lanes_settle(c, it, q, gmask, out);
  else if (q == 0) net_level_scalar(c, begin + item);
}

// The narrow downstream end of the network (levels with at most kTailGroups work items: the
// main stem) in ONE block, levels separated by __syncthreads() instead of kernel boundaries -- a
// launch per level would cost more than the level's arithmetic.  Half of the block's lane groups
// are consumers, one per (river, replica); what does not depend on upstream rivers is loaded one
// level (state) and two levels (indices) ahead, so that only pending -> sum -> pull -> pending
// is on a level's critical path.  The other half run kTailAhead levels ahead and prefetch.
// The kernel contains no calls: when a river needs the general push code it stops at that level
// and wsi_net_tail_general (launched right behind it) takes over from there.
#ifdef WSI_NET_DEBUG  // development aid: per-phase clock64() totals of the tail kernel's thread 0
__device__ unsigned long long g_net_dbg[16];
#define NET_DBG(i)                                        \
  do {                                                    \
    if (threadIdx.x == 0) {                               \
      const long long now_ = clock64();                   \
      g_net_dbg[i] += (unsigned long long)(now_ - dbg_t); \
      dbg_t = now_;                                       \
    }                                                     \
  } while (0)
#else
#define NET_DBG(i)
#endif
constexpr int kNetTail = 512;
constexpr int kTailGroups = kNetTail / kLanes / 2;
constexpr int kTailAhead = 12;
constexpr int kTailRounds = 3;     // (lane kernel, general kernel) pairs launched per step and range
constexpr int kClusterCtas = 8;    // CTAs of the cluster variant (portable maximum)
constexpr int kClusterMax = kClusterCtas * kTailGroups;  // rivers per level it takes in one round

__device__ __forceinline__ unsigned cluster_ctarank() {
  unsigned r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire;" ::: "memory"); }

// CLUSTER = false: one block, levels [begin, end) separated by a split-phase mbarrier (the main
// stem: <= kTailGroups rivers per level).  CLUSTER = true: a thread-block cluster of kClusterCtas
// such blocks, levels separated by the hardware cluster barrier (release / acquire at cluster
// scope) -- the mid-width levels (<= kClusterMax rivers), which would otherwise be one launch
// each.  Wider levels than a round takes are handled in further rounds (correct, just slower).
template <class CFG, bool CLUSTER>
__global__ void __launch_bounds__(kNetTail)
    wsi_net_tail_kernel(const int slot, const CFG cfg, const int t_arg, const uint32_t flags, const int level_begin_,
                        const int level_end_) {
  __shared__ int s_general[2];  // "a river of this level needs the general code", by level parity (see below)
  __shared__ uint64_t s_bar;
  const WsiModelDesc& d = g_wsi_desc[slot];
  const int t = step_of(d, t_arg);
  const int L = level_end_ >= 0 ? level_end_ : d.n_net_levels;
  const int level_begin = level_begin_ >= 0 ? level_begin_ : min(d.net_scratch[0], L);  // < 0: resume
  const int R = d.R;
  const int cta = CLUSTER ? (int)cluster_ctarank() : 0;
  const int stride = (CLUSTER ? kClusterCtas : 1) * kTailGroups;  // (river, replica) items per round
  const int grp = threadIdx.x / kLanes, q = threadIdx.x % kLanes;
  const unsigned gmask = 0xFFFFu << (threadIdx.x & 16);
  const bool prefetcher = grp >= kTailGroups;
  const int w0 = cta * kTailGroups + (prefetcher ? grp - kTailGroups : grp);  // this group's item in round 0
  const int r = w0 % R, k = w0 / R;
  using Cx = Ctx<CFG>;
  Cx c = make_ctx(nullptr, slot, cfg, r, t, flags);
  const int32_t* __restrict__ lptr = d.net_level_ptr;
  // The flag of level l is written before a group arrives at the level's barrier and read right
  // after the wait.  A group may already be in level l + 1 (and set ITS flag) while a lagging group
  // has not read level l's flag yet, so the two levels must not share a word: the flag is indexed
  // by level parity.  Two words suffice -- nobody enters level l + 2 before everybody has arrived
  // at level l + 1, i.e. has read level l's flag -- and a set flag ends the sweep, so no word is
  // ever cleared inside the loop.
  volatile int* g_general = d.net_scratch + 3;  // [2] the cluster's flags (net_scratch[3], [4])
  if (threadIdx.x == 0) {
    s_general[0] = 0;
    s_general[1] = 0;
    if (!CLUSTER) mbar_init(&s_bar, kNetTail);
    if (CLUSTER && cta == 0) {
      g_general[0] = 0;
      g_general[1] = 0;
    }
  }
  // consumer groups: unconditional look-ahead loads -- absent items read item 0, a select
  // would have to wait for the loaded values
  NetRec<Cx> rec2 = {};
  LaneItem<Cx> item1 = {};
  bool has1 = false, has2 = false;
  int i1 = 0, i2 = 0;
  if (!prefetcher) {
    has1 = level_begin < L && k < lptr[level_begin + 1] - lptr[level_begin];
    has2 = level_begin + 1 < L && k < lptr[level_begin + 2] - lptr[level_begin + 1];
    i1 = has1 ? lptr[level_begin] + k : 0;
    i2 = has2 ? lptr[level_begin + 1] + k : 0;
    item1 = lane_item_load(c, net_rec_load(c, i1), q);
    rec2 = net_rec_load(c, i2);
  } else {
    for (int l = level_begin; l < level_begin + kTailAhead && l < L; l++)
      if (k < lptr[l + 1] - lptr[l]) net_prefetch_item(c, lptr[l] + k, q);
  }
  __syncthreads();
  if (CLUSTER) {  // the flag reset above is visible to the whole cluster before anybody may set it
    cluster_arrive();
    cluster_wait();
  }
#ifdef WSI_NET_DEBUG
  long long dbg_t = clock64();
#endif
  int l = level_begin;
  bool stop = false;
  for (; l < L && !stop; l++) {
    if (prefetcher) {
      if (CLUSTER) cluster_arrive();
      else mbar_arrive(&s_bar);  // produces nothing anybody waits for: only paced by the wait below
      const int lp = l + kTailAhead;
      if (lp < L && k < lptr[lp + 1] - lptr[lp]) net_prefetch_item(c, lptr[lp] + k, q);
    } else {
      LaneItem<Cx> cur = item1;
      const bool has = has1;
      const int i0 = i1;
      has1 = has2;
      i1 = i2;
      item1 = lane_item_load(c, rec2, q);
      has2 = l + 2 < L && k < lptr[l + 3] - lptr[l + 2];
      i2 = has2 ? lptr[l + 2] + k : 0;
      rec2 = net_rec_load(c, i2);
      NET_DBG(0);
      LaneOut<Cx> out;
      const bool general0 = has && net_lanes_release(c, cur, q, gmask, out);
      bool general = general0;
      // further rounds of a level wider than one round takes (rare; no look-ahead for them)
      const int count = (lptr[l + 1] - lptr[l]) * R;
      for (int w = w0 + stride; w < count; w += stride) {
        Cx cw = make_ctx(nullptr, slot, cfg, w % R, t, flags);
        const int iw = lptr[l] + w / R;
        LaneItem<Cx> itw = lane_item_load(cw, net_rec_load(cw, iw), q);
        LaneOut<Cx> ow;
        if (net_lanes_release(cw, itw, q, gmask, ow)) {
          general = true;
        } else {
          net_lanes_settle(cw, itw, q, gmask, ow);
          if (q == 0) d.net_done[(size_t)iw * R + (w % R)] = t + 1;
        }
      }
      if (general && q == 0) {
        if (CLUSTER) g_general[l & 1] = 1;
        else s_general[l & 1] = 1;
      }
      NET_DBG(1);
      if (CLUSTER) cluster_arrive();
      else mbar_arrive(&s_bar);  // level l + 1 may start as soon as everybody's releases are parked
      if (has && !general0) {
        net_lanes_settle(c, cur, q, gmask, out);
        if (q == 0) d.net_done[(size_t)i0 * R + r] = t + 1;
      }
      NET_DBG(3);
    }
    if (CLUSTER) cluster_wait();
    else mbar_wait(&s_bar, (uint32_t)(l - level_begin) & 1u);
    NET_DBG(2);
    stop = CLUSTER ? (g_general[l & 1] != 0) : (s_general[l & 1] != 0);
    // rivers of level l that took the lane path are marked in net_done; a general river has
    // stored nothing and is done by wsi_net_tail_general
  }
  if (stop) l--;  // level l (now l - 1 after the loop's increment) is the one to resume at
  __syncthreads();  // everybody has read net_scratch[0] / s_general
  if (threadIdx.x == 0 && cta == 0) {
    d.net_scratch[0] = l;  // first level still to do (the range's end: none)
    d.net_scratch[1] = stop ? 1 : 0;
  }
}

// General continuation of the tail with the scalar code (any push protocol case): `max_levels`
// levels from the one the lane kernel stopped at (rivers it completed there, net_done == t + 1,
// are skipped), then -- `finish` -- the final deliveries.  The launcher alternates the two kernels
// a few times, so that a handful of general rivers per step do not push the whole main stem onto
// the slow path; the last general launch runs to the end.
template <class CFG>
__global__ void __launch_bounds__(kObjBlock)
    wsi_net_tail_general(const int slot, const CFG cfg, const int t_arg, const uint32_t flags, const int max_levels,
                         const int finish) {
  const WsiModelDesc& d = g_wsi_desc[slot];
  const int t = step_of(d, t_arg);
  const int R = d.R;
  const int L = d.n_net_levels;
  if (!finish && d.net_scratch[1] == 0) return;  // the lane kernel got through its range
  const int resume = d.net_scratch[0];
  const int stop = (L - resume > max_levels) ? resume + max_levels : L;
  for (int l = resume; l < stop; l++) {
    const int begin = d.net_level_ptr[l], count = d.net_level_ptr[l + 1] - begin;
    for (size_t w = threadIdx.x; w < (size_t)count * R; w += kObjBlock) {
      const int i = begin + (int)(w / (size_t)R), r = (int)(w % (size_t)R);
      if (l == resume && d.net_done[(size_t)i * R + r] == t + 1) continue;
      Ctx<CFG> cl = make_ctx(nullptr, slot, cfg, r, t, flags);
      net_item(cl, kNetLevel, i);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    d.net_scratch[0] = stop;
    d.net_scratch[1] = 0;
  }
  if (!finish) return;
  for (size_t w = threadIdx.x; w < (size_t)d.n_net_final * R; w += kObjBlock) {
    Ctx<CFG> cf = make_ctx(nullptr, slot, cfg, (int)(w % (size_t)R), t, flags);
    net_item(cf, kNetFinal, (int)(w / (size_t)R));
  }
}

// Device-side step counter of the graph mode.
__global__ void wsi_set_step_kernel(const int slot, const int t) { g_wsi_desc[slot].net_scratch[2] = t; }
__global__ void wsi_advance_kernel(const int slot) { g_wsi_desc[slot].net_scratch[2] += 1; }

// Sum the mass-balance partial rows of one timestep: block j reduces column j of
// partial[n_rows][n] with a fixed strided pass + fixed shared-memory tree (deterministic).
constexpr int kFinBlock = 256;
__global__ void __launch_bounds__(kFinBlock)
    wsi_mb_finalize(const int slot, const int t_arg, int n_rows, int n) {
  __shared__ double s[kFinBlock];
  const WsiModelDesc& d = g_wsi_desc[slot];
  const double* __restrict__ partial = d.mb_partial;
  double* __restrict__ sums = d.mb_sums + (size_t)step_of(d, t_arg) * n;
  const int j = blockIdx.x;
  double acc = 0.0;
  for (int b = threadIdx.x; b < n_rows; b += kFinBlock) acc += partial[(size_t)b * n + j];
  s[threadIdx.x] = acc;
  __syncthreads();
  for (int o = kFinBlock / 2; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) sums[j] = s[0];
}

template <class CFGR>
CFGR runtime_cfg_as(const WsiModelDesc* d) {
  CFGR c;
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

#define WSI_CHECK_LAUNCH(name)                                                                         \
  do {                                                                                                 \
    g_launches++;                                                                                      \
    cudaError_t err_ = cudaGetLastError();                                                             \
    if (err_ != cudaSuccess) return fail(WSI_ERR_LAUNCH, name " launch failed: %s", cudaGetErrorString(err_)); \
  } while (0)

template <class CFG, int TYPE>
int launch_surfaces(const WsiModelDesc* d, int slot, const CFG& cfg, int64_t t, uint32_t flags, cudaStream_t s) {
  const int n = d->surf_class_ptr[TYPE + 1] - d->surf_class_ptr[TYPE];
  if (n <= 0) return WSI_OK;
  const dim3 grid((d->R + kSurfBlock - 1) / kSurfBlock, n);
  wsi_surface_kernel<CFG, TYPE><<<grid, kSurfBlock, 0, s>>>(slot, cfg, (int)t, flags);
  WSI_CHECK_LAUNCH("wsi_surface_kernel");
  return WSI_OK;
}

// One timestep = surface kernels (one per class) -> schedule kernel -> mass-balance kernel ->
// finalize -> record kernel -> end_timestep kernel, all on one stream.
template <class CFG>
int launch(const WsiModelDesc* d, int slot, const CFG& cfg, int64_t t, int b, int e, uint32_t flags, cudaStream_t s) {
  constexpr int V = CFG::P + 1;
  const bool want_mb = (flags & WSI_RUN_MASS_BALANCE) != 0;
  const bool want_sums = want_mb && d->mb_sums != nullptr;
  // schedule entries that reach into the surfaces before Land.run (garden irrigation) run first
  if (d->engine == 0 && (flags & WSI_RUN_SCHEDULE) && (flags & WSI_RUN_SURFACES) && d->sched_split > b && d->sched_split < e) {
    wsi_step_kernel<CFG, false><<<n_blocks(d), kBlock, kDescSmem, s>>>(slot, cfg, (int)t, b, d->sched_split, WSI_RUN_SCHEDULE, 0, 0);
    WSI_CHECK_LAUNCH("wsi_step_kernel");
    b = d->sched_split;
  }
#ifndef WSI_SURF_SPLIT
  if ((flags & WSI_RUN_SURFACES) && d->n_surf > 0 && d->n_surf <= 65535) {
    const dim3 grid((d->R + kSurfBlock - 1) / kSurfBlock, d->n_surf);
    wsi_surfaces_kernel<CFG><<<grid, kSurfBlock, 0, s>>>(slot, cfg, (int)t, flags);
    WSI_CHECK_LAUNCH("wsi_surfaces_kernel");
  } else
#endif
  if (flags & WSI_RUN_SURFACES) {
    int rc = launch_surfaces<CFG, WSI_ST_SURFACE>(d, slot, cfg, t, flags, s);
    if (!rc) rc = launch_surfaces<CFG, WSI_ST_IMPERVIOUSSURFACE>(d, slot, cfg, t, flags, s);
    if (!rc) rc = launch_surfaces<CFG, WSI_ST_PERVIOUSSURFACE>(d, slot, cfg, t, flags, s);
    if (!rc) rc = launch_surfaces<CFG, WSI_ST_GROWINGSURFACE>(d, slot, cfg, t, flags, s);
    if (rc) return rc;
  }
  const bool net = d->engine == 1;
  if ((flags & WSI_RUN_SCHEDULE) && e > b) {
    if (net) {
      if (d->n_net_bnd > 0) {
        wsi_net_kernel<CFG><<<n_flat_blocks(d, d->n_net_bnd), kObjBlock, 0, s>>>(slot, cfg, (int)t, flags,
                                                                                 kNetBoundary, 0, d->n_net_bnd);
        WSI_CHECK_LAUNCH("wsi_net_kernel");
      }
      if (d->n_net_calc > 0) {
        wsi_net_kernel<CFG><<<n_flat_blocks(d, d->n_net_calc), kObjBlock, 0, s>>>(slot, cfg, (int)t, flags,
                                                                                  kNetPre, 0, d->n_net_calc);
        WSI_CHECK_LAUNCH("wsi_net_kernel");
      }
      // wide levels: one launch each -- a thread per (river, replica) while the level fills the
      // machine that way, lane groups below that; then ONE cluster launch for the levels that fit a
      // round of the cluster kernel, then ONE block for the levels that fit a round of the tail
      // kernel (the main stem).  Each of the two alternates with the general kernel a few times.
      const int32_t* lp = d->net_level_ptr_host;
      const int L = d->n_net_levels;
      auto first_after_which_all_fit = [&](size_t limit) {
        int l = L;
        while (l > 0 && (size_t)(lp[l] - lp[l - 1]) * (size_t)d->R <= limit) l--;
        return l;
      };
      const int tail = first_after_which_all_fit((size_t)kTailGroups);
      const int mid = cluster_enabled() ? first_after_which_all_fit((size_t)kClusterMax) : tail;
      for (int l = 0; l < mid; l++) {
        const int count = lp[l + 1] - lp[l];
        if (count <= 0) continue;
        const size_t work = (size_t)count * (size_t)d->R;
        if (work > kLaneLevelMax) {
          wsi_net_kernel<CFG><<<n_flat_blocks(d, count), kObjBlock, 0, s>>>(slot, cfg, (int)t, flags, kNetLevel,
                                                                            lp[l], count);
          WSI_CHECK_LAUNCH("wsi_net_kernel");
        } else {
          const int blocks = (int)((work * kLanes + kLaneBlock - 1) / kLaneBlock);
          wsi_net_lane_kernel<CFG><<<blocks, kLaneBlock, 0, s>>>(slot, cfg, (int)t, flags, lp[l], count);
          WSI_CHECK_LAUNCH("wsi_net_lane_kernel");
        }
      }
      bool started = false;  // has a lane kernel of this step already written net_scratch[0]?
      if (mid < tail) {
        cudaLaunchConfig_t lc = {};
        lc.gridDim = dim3(kClusterCtas);
        lc.blockDim = dim3(kNetTail);
        lc.stream = s;
        cudaLaunchAttribute attr;
        attr.id = cudaLaunchAttributeClusterDimension;
        attr.val.clusterDim.x = kClusterCtas;
        attr.val.clusterDim.y = 1;
        attr.val.clusterDim.z = 1;
        lc.attrs = &attr;
        lc.numAttrs = 1;
        for (int k = 0; k < kTailRounds; k++) {
          cudaError_t err = cudaLaunchKernelEx(&lc, wsi_net_tail_kernel<CFG, true>, slot, cfg, (int)t, flags,
                                               started ? -1 : mid, tail);
          if (err != cudaSuccess) return fail(WSI_ERR_LAUNCH, "cluster launch failed: %s", cudaGetErrorString(err));
          WSI_CHECK_LAUNCH("wsi_net_tail_kernel<cluster>");
          started = true;
          wsi_net_tail_general<CFG><<<1, kObjBlock, 0, s>>>(slot, cfg, (int)t, flags, 1, 0);
          WSI_CHECK_LAUNCH("wsi_net_tail_general");
        }
      }
      for (int k = 0; k < kTailRounds; k++) {
        const bool last = k == kTailRounds - 1;
        wsi_net_tail_kernel<CFG, false><<<1, kNetTail, 0, s>>>(slot, cfg, (int)t, flags, started ? -1 : tail, -1);
        WSI_CHECK_LAUNCH("wsi_net_tail_kernel");
        started = true;
        wsi_net_tail_general<CFG><<<1, kObjBlock, 0, s>>>(slot, cfg, (int)t, flags, last ? 0x7fffffff : 1, last ? 1 : 0);
        WSI_CHECK_LAUNCH("wsi_net_tail_general");
      }
    } else {
      // Small ensembles are launch-latency bound: the schedule kernel then also runs the other
      // phases asked for (fused tail, one thread per replica).  Large ones have the threads to hide
      // the tail's dependent loads only when it runs one thread per (object, replica): measured on
      // config 2 at R = 1e5, fused 0.228 ms vs 0.13 ms for the three object-parallel kernels.
      const bool fused = d->R <= fused_tail_max_r();
      // the lean instantiation for models without treatment works / demand / distribution / queue groundwater /
      // abstractions (desc.features == 0): those handlers are compiled out of the hot kernel
      bool lean = false;
      if constexpr (CFG::kLeanKernel) {
        if (d->features == 0) {
          lean = true;
          wsi_step_kernel<CFG, true><<<n_blocks(d), kBlock, kDescSmem, s>>>(slot, cfg, (int)t, b, e, flags,
                                                                              (int)t - d->rec_t0, fused ? 1 : 0);
        }
      }
      if (!lean)
        wsi_step_kernel<CFG, false><<<n_blocks(d), kBlock, kDescSmem, s>>>(slot, cfg, (int)t, b, e, flags,
                                                                             (int)t - d->rec_t0, fused ? 1 : 0);
      WSI_CHECK_LAUNCH("wsi_step_kernel");
      int rows = n_blocks(d);
      if (!fused && (flags & (WSI_RUN_MASS_BALANCE | WSI_RUN_END_TIMESTEP)) && d->n_nodes > 0) {
        const dim3 grid(n_obj_blocks(d), d->n_nodes);
        wsi_settle_kernel<CFG><<<grid, kObjBlock, 0, s>>>(slot, cfg, (int)t, flags);
        WSI_CHECK_LAUNCH("wsi_settle_kernel");
        rows = d->n_nodes * n_obj_blocks(d);
      }
      if (want_sums) {
        const int n = 3 * V;
        wsi_mb_finalize<<<n, kFinBlock, 0, s>>>(slot, (int)t, rows, n);
        WSI_CHECK_LAUNCH("wsi_mb_finalize");
      }
      return WSI_OK;
    }
  }
  // What follows runs when the schedule kernel did not (phase-by-phase callers) and for the network
  // engine.  Replica engine: the settle pass in two halves around the recording, network engine: the
  // flat (item, replica) kernels -- in the order of model.py:1016-1176.
  auto settle = [&](uint32_t f) {
    const dim3 grid(n_obj_blocks(d), d->n_nodes);
    wsi_settle_kernel<CFG><<<grid, kObjBlock, 0, s>>>(slot, cfg, (int)t, f);
  };
  if (want_mb) {
    const int items = d->n_nodes + (want_sums && net ? d->n_arcs : 0);
    const int rows = net ? n_flat_blocks(d, items) : d->n_nodes * n_obj_blocks(d);
    if (items > 0) {
      if (net) {
        wsi_mb_kernel<CFG, true><<<rows, kObjBlock, 0, s>>>(slot, cfg, (int)t, flags, items);
        WSI_CHECK_LAUNCH("wsi_mb_kernel");
      } else {
        settle(flags & ~WSI_RUN_END_TIMESTEP);
        WSI_CHECK_LAUNCH("wsi_settle_kernel");
      }
    }
    if (want_sums) {
      const int n = 3 * V;
      wsi_mb_finalize<<<n, kFinBlock, 0, s>>>(slot, (int)t, rows, n);
      WSI_CHECK_LAUNCH("wsi_mb_finalize");
    }
  }
  if (flags & WSI_RUN_RECORD) {
    // series probes only inside the record window; reduced probes every step
    const int row = (int)t - d->rec_t0;
    const bool in_window = d->n_probe > 0 && (t < 0 || (row >= 0 && row < d->rec_T));
    const int items = d->n_probe + d->n_red;
    if (d->n_red > 0 || in_window) {
      if (net) {
        wsi_record_kernel<CFG, true><<<n_flat_blocks(d, items), kObjBlock, 0, s>>>(slot, cfg, (int)t, row);
      } else {
        const dim3 grid(n_obj_blocks(d), items);
        wsi_record_kernel<CFG, false><<<grid, kObjBlock, 0, s>>>(slot, cfg, (int)t, row);
      }
      WSI_CHECK_LAUNCH("wsi_record_kernel");
    }
  }
  if (flags & WSI_RUN_END_TIMESTEP) {
    if (net && d->n_end > 0) {
      wsi_end_kernel<CFG, true><<<n_flat_blocks(d, d->n_end), kObjBlock, 0, s>>>(slot, cfg, (int)t);
      WSI_CHECK_LAUNCH("wsi_end_kernel");
    } else if (!net && d->n_nodes > 0) {
      settle(flags & ~WSI_RUN_MASS_BALANCE);
      WSI_CHECK_LAUNCH("wsi_settle_kernel");
    }
  }
  return WSI_OK;
}


}  // namespace

// The host functions a configuration unit exports (wsi_host.h ConfigUnit).  MAKE_CFG(d) builds the
// configuration object from the descriptor.
#define WSI_DEFINE_CONFIG_UNIT(NAME, MAKE_CFG)                                                              \
  namespace {                                                                                               \
  int dispatch_##NAME(const WsiModelDesc* d, int slot, int64_t t, int b, int e, uint32_t flags, cudaStream_t s) { \
    return launch(d, slot, MAKE_CFG, t, b, e, flags, s);                                                     \
  }                                                                                                         \
  cudaError_t upload_##NAME(const WsiModelDesc* d, int slot, cudaStream_t s) {                               \
    return cudaMemcpyToSymbolAsync(g_wsi_desc, d, sizeof(WsiModelDesc), (size_t)slot * sizeof(WsiModelDesc), \
                                   cudaMemcpyHostToDevice, s);                                              \
  }                                                                                                         \
  void set_step_##NAME(int slot, int t, cudaStream_t s) { wsi_set_step_kernel<<<1, 1, 0, s>>>(slot, t); }   \
  void advance_##NAME(int slot, cudaStream_t s) { wsi_advance_kernel<<<1, 1, 0, s>>>(slot); }              \
  }                                                                                                         \
  namespace wsi_host {                                                                                      \
  const ConfigUnit wsi_unit_##NAME = {#NAME, dispatch_##NAME, upload_##NAME, set_step_##NAME, advance_##NAME}; \
  }
