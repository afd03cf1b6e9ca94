/*
 * wsi_b200 -- C ABI of the B200-native WSIMOD land/tank/arc backend.
 *
 * The reference (ImperialCollegeLondon/wsi) is pure Python and has no FFI; the
 * boundary it exposes for this path is Model.load / Model.run
 * (wsimod/orchestration/model.py:216, :893).  This header declares what a ctypes
 * binding inside Model.run calls instead of the per-node Python loop of
 * model.py:993-1176.  All buffers are caller-owned (torch tensors): the library
 * never allocates, frees or retains them.  Pointers are DEVICE pointers for
 * libwsi_b200.so; the CPU oracle (oracle/, test infrastructure only) accepts the
 * same struct with HOST pointers.
 *
 * Data layout: see include/wsi_layout.h (generated) and DESIGN.md section 3.
 * Group tensors are fp64 [n_fields][n_obj * R]; object `slot` of replica `r` is
 * column slot * R + r.
 */
#ifndef WSI_B200_H
#define WSI_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WSI_ABI_VERSION 2

/* error classes returned by every entry point (0 == success) */
#define WSI_OK 0
#define WSI_ERR_BAD_DESC 1     /* inconsistent descriptor (sizes, null pointers, P unsupported) */
#define WSI_ERR_LAUNCH 2       /* cudaGetLastError() after a launch */
#define WSI_ERR_UNSUPPORTED 3  /* feature outside the supported path */

/* sticky per-node status bits written to WsiModelDesc.node_flags.  The reference
 * prints advisories for these (core.py:423, nodes.py:476, land.py:1808) instead
 * of raising; they map to bits, not exceptions. */
#define WSI_FLAG_MASS_BALANCE 1u   /* core.py:419-427 */
#define WSI_FLAG_MAXITER 2u        /* nodes.py:475-476 */
#define WSI_FLAG_POOL_MISMATCH 4u  /* "nutrient pool not tracking soil tank" land.py:1808 */
#define WSI_FLAG_NAN 8u            /* reference would raise ZeroDivisionError (land.py:1462) */
#define WSI_FLAG_DEPTH 16u         /* junction chain deeper than the kernel supports */
#define WSI_FLAG_DENY 32u          /* push to / pull from a node whose handler denies (nodes.py:563-609) */
#define WSI_FLAG_DEFICIT 64u       /* "Service reservoirs not filled" / "demand deficit" (wtw.py:583, demand.py:91) */

/* arc_flags bits */
#define WSI_ARC_DENY_PUSH 1 /* PullArc: pushes and push checks answer "nothing" (arcs.py:808-850) */
#define WSI_ARC_DENY_PULL 2 /* PushArc: pulls and pull checks answer "nothing" (arcs.py:852-893) */

/* run flags for wsi_sched_range */
#define WSI_RUN_SCHEDULE 1u      /* execute schedule entries [begin, end) */
#define WSI_RUN_MASS_BALANCE 2u  /* per-node mass balance (core.py:370) */
#define WSI_RUN_RECORD 4u        /* copy probe planes into out_probe (model.py:1063-1170) */
#define WSI_RUN_END_TIMESTEP 8u  /* node/arc end_timestep (model.py:1172-1176) */
#define WSI_RUN_SURFACES 16u     /* Surface.run of every Land surface (land.py:425-449) */
#define WSI_RUN_ALL 31u

typedef struct WsiModelDesc {
  /* ---- sizes ---- */
  int32_t abi_version;
  int32_t P;  /* number of pollutants, constants.POLLUTANTS (constants.py:21) */
  int32_t V;  /* P + 1 */
  int32_t R;  /* replicas (ensemble members / independent catchments) */
  int32_t T;  /* timesteps covered by the forcing tables */
  int32_t n_nodes, n_arcs, n_land, n_surf, n_tank, n_river, n_sewer, n_catch;
  int32_t n_sched;
  int32_t nq; /* queue slots per sewer QueueTank (max travel time + 1, >= 2) */
  int32_t n_fseries, n_dep_series, n_months, n_cseries;
  int32_t n_probe;    /* planes recorded per step */
  int32_t rec_t0;     /* out_probe row 0 corresponds to timestep rec_t0 */
  int32_t rec_T;      /* rows available in out_probe / out_mb */
  int32_t record_mb;  /* 1: write per-node (in, ds, out) into out_mb */
  uint32_t additive_mask; /* bit i (1..P): vqip plane i is an ADDITIVE pollutant */
  int32_t idx[8];     /* vqip plane of nitrate, nitrite, ammonia, phosphate,
                         org-phosphorus, org-nitrogen, solids, temperature; -1 absent */
  double float_accuracy; /* constants.FLOAT_ACCURACY */

  /* ---- template graph (int32) ---- */
  const int32_t* node_type;  /* [n_nodes] WSI_NT_* */
  const int32_t* node_gidx;  /* [n_nodes] index inside the node's group */
  const int32_t* arc_in;     /* [n_arcs] in_port node */
  const int32_t* arc_out;    /* [n_arcs] out_port node */
  const int32_t* out_ptr;    /* [n_nodes+1] CSR, insertion order of out_arcs */
  const int32_t* out_arcs;
  const int32_t* in_ptr;     /* [n_nodes+1] CSR, insertion order of in_arcs */
  const int32_t* in_arcs;
  const int32_t* cand_ptr;   /* [n_nodes*WSI_NF+1] CSR of arcs per (node, filter) */
  const int32_t* cand_arcs;
  const int32_t* land_surf_ptr; /* [n_land+1] CSR land -> surface slots */
  const int32_t* land_surf;
  const int32_t* surf_order; /* [n_surf] surface slots sorted by surface class */
  int32_t surf_class_ptr[5]; /* CSR over surf_order: class k = [ptr[k], ptr[k+1]) */
  int32_t features;          /* bit 0: the model has treatment works, demand, distribution, queue groundwater, river
                                reservoirs, irrigation or abstractions (pulls); 0 lets the library use the schedule
                                kernel without them.  bit 1: IrrigationSurface / VariableAreaSurface planes are in use */
  const int32_t* end_kind;   /* [n_end] end_timestep items: group id (0 arc, 2 surf, 3 tank, 4 river, 5 sewer) */
  const int32_t* end_obj;    /* [n_end] object slot */
  int32_t n_end;
  int32_t reserved1;
  const int32_t* sched_op;   /* [n_sched] WSI_OP_* */
  const int32_t* sched_node; /* [n_sched] node slot */

  /* ---- network engine (engine == 1): level-synchronous routing of a dendritic river network,
   *      one thread per (node, replica); see wsi_b200/csrc/wsi_net.cuh ---- */
  int32_t engine;            /* 0: replica engine (one thread per replica runs the schedule), 1: network */
  int32_t n_net_levels;
  int32_t n_net_calc;
  int32_t n_net_final;
  int32_t n_net_bnd;
  int32_t reserved2;
  const int32_t* net_level_ptr;   /* [n_net_levels+1] CSR over net_level_node (upstream depth ascending) */
  const int32_t* net_level_ptr_host; /* HOST copy of net_level_ptr (the launcher sizes one grid per level) */
  const int32_t* net_level_node;  /* river node slots */
  const int32_t* net_calc_node;   /* [n_net_calc] the rivers (calculate_discharge + their catchments' route) */
  const int32_t* net_final_node;  /* [n_net_final] gather-only nodes (Waste, rivers outside the order) */
  const int32_t* net_bnd_node;    /* [n_net_bnd] Boundary nodes (partitioned networks) */
  const int32_t* net_path_ptr;    /* [n_nodes+1] incoming paths per receiving node */
  const int32_t* net_path_split;  /* [n_nodes] first path of the node whose sender is a River: catchments
                                     come first (they are routed before any river discharges) */
  const int32_t* net_path_sender; /* sender node of each path */
  const int32_t* net_path_arc_ptr; /* CSR over net_path_arcs: arcs along each path */
  const int32_t* net_path_arcs;
  const int32_t* net_item;   /* [n_level_items][8] per entry of net_level_node: node, tank slot, river slot,
                                out-arc (-1: none), [begin, end) into net_rpath, 1 if the out-arc goes
                                straight to a River / Waste, arc of the first river path */
  int32_t* net_scratch;      /* [8] device scratch: [0] level at which the fast tail kernel handed over, [1] stop
                                flag, [2] device-side step counter, [3..4] cluster "general" flags by level parity */
  int32_t* net_done;         /* [n_level_items] device scratch: t + 1 of the last step in which the fast tail
                                kernel completed the item (the general kernel skips those when it resumes) */
  const int32_t* net_rpath;  /* [.][4] river paths in delivery order: sender's arc, sender's tank slot,
                                1 if the arc goes straight to the receiver, sender node */

  /* ---- SoA groups, fp64 [n_fields][n_obj*R] ---- */
  double* arc;
  double* land;
  double* surf;
  double* tank;
  double* river;
  double* sewer;
  double* catchm; /* [1][n_catch*R]: series id per catchment */

  /* ---- forcing ---- */
  const double* node_forcing; /* [T][n_fseries][3] precipitation, et0, temperature */
  const double* surf_inputs;  /* [n_months][n_dep_series][WSI_N_DEP] */
  const double* catch_inputs; /* [T][n_cseries][V] flow + pollutant concentrations */
  const int32_t* month_idx;   /* [T] row of surf_inputs */
  const int32_t* dayofyear;   /* [T] t.dayofyear */
  const int32_t* doy_adj;     /* [T] leap-shifted day of year (land.py:1524-1529) */

  /* ---- outputs ---- */
  const int32_t* probe_group; /* [n_probe] 0 arc,1 land,2 surf,3 tank,4 river,5 sewer */
  const int32_t* probe_plane; /* [n_probe] plane (field) index inside the group */
  const int32_t* probe_obj;   /* [n_probe] object slot */
  double* out_probe;          /* [rec_T][n_probe][R] */
  double* out_mb;             /* [rec_T][n_nodes][3][V][R] or NULL */
  double* mb_terms;           /* [n_fused][3][V][R] scratch: (in, ds, out) of nodes whose terms are formed
                                 inside their schedule op (Land, Catchment) */
  const int32_t* mb_term_slot; /* [n_nodes] row of mb_terms, -1 for other nodes */
  double* mb_partial;         /* [wsi_num_blocks()][3*V] scratch for the block reduction */
  double* mb_sums;            /* [T][3][V] system totals sys_in, sys_ds, sys_out */
  int32_t* mb_count;          /* [T] number of (node, component) violations */
  uint32_t* node_flags;       /* [n_nodes*R] sticky WSI_FLAG_* */

  /* ---- simple chains (ABI 2): arc_chain[arc_chain_ptr[a] .. arc_chain_ptr[a+1]) lists arc `a` followed by
   *      the arcs a push down `a` travels through junction Nodes that forward to exactly one arc
   *      (nodes.py:62-66, 424-438); empty when a junction on the way splits, dead-ends, or the chain is
   *      longer than the kernels unroll.  NULL: every push takes the general recursive code. ---- */
  const int32_t* arc_chain_ptr; /* [n_arcs+1] */
  const int32_t* arc_chain;

  /* ---- reduced outputs and member window (ABI 2): what an ensemble run brings home without streaming
   *      every member's series (objectives of model.py:1177-1197 evaluated on the device) ---- */
  int32_t n_red;      /* planes reduced over time instead of recorded as series */
  int32_t rec_r0;     /* out_probe holds the members [rec_r0, rec_r0 + rec_R) only ... */
  int32_t rec_R;      /* ... rec_R == 0: every member (out_probe is [rec_T][n_probe][R]) */
  int32_t reserved3;
  const int32_t* red_group; /* [n_red] like probe_group / probe_plane / probe_obj */
  const int32_t* red_plane;
  const int32_t* red_obj;
  double* red_out;    /* [2][n_red][R]: running sum and running maximum over the timesteps stepped so far
                         (caller-initialised: 0 and -inf) */

  /* ---- treatment works and demand (ABI 2): SoA groups like the ones above ---- */
  int32_t n_works, n_demand;
  double* works;      /* WWTW / FWTW (wtw.py) */
  double* demand;     /* ResidentialDemand (demand.py) */
  double* distrib;    /* UnlimitedDistribution (distribution.py) */
  int32_t n_distrib;
  int32_t sched_split; /* > 0: schedule entries [0, sched_split) run BEFORE the surfaces of the timestep (they reach into
                          a Land's surfaces: garden irrigation by a Demand node, demand.py:250 / land.py:2402);
                          0: surfaces first, then the whole schedule (nothing before Land.run touches them) */
  const int32_t* arc_flags; /* [n_arcs] WSI_ARC_DENY_PUSH: PullArc (arcs.py:808-850; the caller also leaves such arcs
                               out of cand_arcs), WSI_ARC_DENY_PULL: PushArc (arcs.py:852-893) */
} WsiModelDesc;

/* ABI version of the loaded library (== WSI_ABI_VERSION). */
int wsi_version(void);

/* sizeof(WsiModelDesc) the library was compiled with: a binding compares it with its own
 * mirror of the struct before passing descriptors (catches a stale build). */
int wsi_desc_size(void);

/* Thread-local message for the last non-zero return. */
const char* wsi_last_error(void);

/* Structural checks on a descriptor (no device access). */
int wsi_validate(const WsiModelDesc* d);

/* Rows of mb_partial a timestep may write (blocks that reduce mass-balance partials). */
int wsi_num_blocks(const WsiModelDesc* d);

/* One model timestep `t` for every replica: orchestration schedule, river
 * routing, mass balance, recording, end_timestep -- replaces one iteration of
 * the loop at model.py:993-1176.  Asynchronous on `stream` (cudaStream_t). */
int wsi_step(const WsiModelDesc* d, int64_t t, void* stream);

/* Timesteps [t0, t1), launched back to back from C (no per-step ctypes cost). */
int wsi_run_steps(const WsiModelDesc* d, int64_t t0, int64_t t1, void* stream);

/* Part of a timestep: schedule entries [begin, end) and the phases in `flags`
 * (WSI_RUN_*).  wsi_step == wsi_sched_range(d, t, 0, n_sched, WSI_RUN_ALL).
 * WSI_RUN_SURFACES launches the surface kernels (one thread per surface, one launch per
 * surface class) before the per-replica schedule kernel; a LAND_RUN schedule entry then
 * only runs the Land's own tanks and pushes.  Land / Catchment mass balances are fused into
 * their ops, so WSI_RUN_MASS_BALANCE must accompany WSI_RUN_SURFACES | WSI_RUN_SCHEDULE. */
int wsi_sched_range(const WsiModelDesc* d, int64_t t, int32_t begin, int32_t end,
                    uint32_t flags, void* stream);

/* Kernels launched by this library since load (bench.py's gpu_launches). */
int64_t wsi_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif
