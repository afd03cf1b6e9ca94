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

// Deliver the incoming paths of `node` (a River or Waste at the end of junction chains).
template <class C>
WSI_HD void net_gather(C& c, int node) {
  using Vq = typename C::Vq;
  constexpr int V = C::V;
  const WsiModelDesc& d = c.D();
#pragma unroll 1
  for (int p = d.net_path_ptr[node]; p < d.net_path_ptr[node + 1]; p++) {
    const int sender = d.net_path_sender[p];
    if (d.node_type[sender] == WSI_NT_CATCHMENT) {
      catchment_route(c, sender);  // the flow is forcing data: nothing to wait for
    } else {
      // second half of River.distribute storage.py:926-936
      const int a0 = d.out_arcs[d.out_ptr[sender]];
      const Vq outflow = c.ldv(c.arc(WSI_ARC_PENDING(V), a0), c.s_arc());
      const Vq reply = push_distributed<kJunctionDepth>(c, sender, outflow, WSI_FILTER_NRW);
      const int tk = (int)*c.river(WSI_RIVER_TANK, d.node_gidx[sender]);
      tank_push_force(c, c.tank(WSI_TANK_STORAGE(V), tk), c.s_tank(), reply);
    }
  }
}

// First half of River.distribute storage.py:926-936 for a river whose inflows have been delivered
template <class C>
WSI_HD void net_river_release(C& c, int node) {
  constexpr int V = C::V;
  const WsiModelDesc& d = c.D();
  if (d.out_ptr[node + 1] - d.out_ptr[node] != 1) {  // no receiver: the sequential code is already local
    river_distribute(c, node);
    return;
  }
  const int g = d.node_gidx[node];
  const int tk = (int)*c.river(WSI_RIVER_TANK, g);
  double* st = c.tank(WSI_TANK_STORAGE(V), tk);
  const typename C::Vq outflow = tank_pull(c, st, c.s_tank(), st[0] * *c.river(WSI_RIVER_RIVERRC, g));
  c.stv(c.arc(WSI_ARC_PENDING(V), d.out_arcs[d.out_ptr[node]]), c.s_arc(), outflow);
}

// item kinds of the network engine's item-parallel kernel
constexpr int kNetRiverCalc = 0;   // River.calculate_discharge of every river
constexpr int kNetLevel = 1;       // deliver + release of the rivers of one level
constexpr int kNetGatherOnly = 2;  // Waste nodes and rivers outside the discharge order

template <class C>
WSI_HD void net_item(C& c, int kind, int node) {
  switch (kind) {
    case kNetRiverCalc:
      river_calc(c, node);
      break;
    case kNetLevel:
      net_gather(c, node);
      net_river_release(c, node);
      break;
    default:
      net_gather(c, node);
      break;
  }
}

// The whole routing phase of one replica in level order (host build and tiny models; the CUDA
// library launches net_item per (node, replica) instead, one launch per level).
template <class C>
WSI_HD void net_schedule(C& c) {
  const WsiModelDesc& d = c.D();
  for (int i = 0; i < d.n_net_calc; i++) net_item(c, kNetRiverCalc, d.net_calc_node[i]);
  for (int l = 0; l < d.n_net_levels; l++)
    for (int i = d.net_level_ptr[l]; i < d.net_level_ptr[l + 1]; i++) net_item(c, kNetLevel, d.net_level_node[i]);
  for (int i = 0; i < d.n_net_final; i++) net_item(c, kNetGatherOnly, d.net_final_node[i]);
}

}  // namespace wsi
