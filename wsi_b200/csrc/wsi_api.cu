// libwsi_b200.so: the C ABI of include/wsi_b200.h.  Host code only: descriptor validation, the
// constant-memory slot table, CUDA-graph capture / replay of a timestep and the choice of the
// configuration unit (wsi_cfg_*.cu) whose kernels run the model.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <mutex>

#include "../../include/wsi_layout.h"
#include "wsi_host.h"

namespace wsi_host {

thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};

int fail(int code, const char* fmt, const char* detail) {
  snprintf(g_err, sizeof g_err, fmt, detail);
  return code;
}

// The cluster variant of the tail kernel (mid-width levels in one launch of a thread-block cluster)
// is correct and tested but measured no faster than one lane-kernel launch per level inside the
// CUDA graph (N = 1e6: 10.2 vs 10.0 ms per step), so it is opt-in: WSI_B200_CLUSTER=1.
bool cluster_enabled() {
  static const bool on = [] {
    const char* e = getenv("WSI_B200_CLUSTER");
    return e && e[0] && e[0] != '0';
  }();
  return on;
}

// Replicas up to which the schedule kernel runs mass balance and end_timestep itself (two launches per step)
// instead of handing them to wsi_settle_kernel.  Default 0 = never: one thread per (node,
// replica) beats one thread walking a replica's nodes at every ensemble size measured, a single model included
// (profiles/r02_fused_tail_threshold.md: 35-45 % of the step at R = 1 ... 4096).  WSI_B200_FUSED_TAIL_MAX_R
// overrides it -- the tests use it to put the fused sequence under the same parity.
int fused_tail_max_r() {
  const char* e = getenv("WSI_B200_FUSED_TAIL_MAX_R");
  return e && e[0] ? atoi(e) : 0;
}

}  // namespace wsi_host

using namespace wsi_host;

namespace {

#define WSI_CHECK_LAUNCH(name)                                                                         \
  do {                                                                                                 \
    g_launches++;                                                                                      \
    cudaError_t err_ = cudaGetLastError();                                                             \
    if (err_ != cudaSuccess) return fail(WSI_ERR_LAUNCH, name " launch failed: %s", cudaGetErrorString(err_)); \
  } while (0)

bool is_default12(const WsiModelDesc* d) {
  static const int def_idx[8] = {10, 11, 4, 3, 2, 12, 5, 9};
  if (d->P != 12 || d->additive_mask != 0x1CFCu) return false;
  for (int i = 0; i < 8; i++)
    if (d->idx[i] != def_idx[i]) return false;
  return true;
}

// The configuration unit whose kernels run this model (nullptr: none instantiated).
const ConfigUnit* unit_of(const WsiModelDesc* d) {
  if (is_default12(d)) return &wsi_unit_default12;
#ifndef WSI_ONLY_DEFAULT12  // development builds link a single configuration
  if (d->P == 12 && d->additive_mask == kOxfordMask) return &wsi_unit_oxford12;
  if (d->P == 12) return &wsi_unit_runtime12;
  if (d->P == 2) return &wsi_unit_runtime2;
  if (d->P == 15) return &wsi_unit_runtime15;
#endif
  return nullptr;
}

// Copy the descriptor into a constant-memory slot, ordered on the launching stream.  A descriptor
// (identified by its address: one per model) keeps its slot from call to call, so models stepped
// concurrently on different streams -- the shards of a partitioned network -- never overwrite each
// other's descriptor; calls for one model are ordered by its stream.  When more than
// WSI_DESC_SLOTS models are alive the least recently used slot is recycled after a device
// synchronisation.
struct SlotEntry {
  const void* key;
  unsigned long long last_use;
  bool uploaded;
};
struct GraphEntry {
  const void* key = nullptr;
  unsigned long long hash = 0;
  cudaGraphExec_t exec = nullptr;
  long long launches = 0;
};
// Constant memory, streams and graphs belong to a device: everything the library remembers between
// calls is kept per device (one process may drive several GPUs, one host thread each).
constexpr int kMaxDevices = 16;
struct DeviceState {
  SlotEntry slots[WSI_DESC_SLOTS] = {};
  WsiModelDesc image[WSI_DESC_SLOTS] = {};  // what each slot holds (uploads of an unchanged descriptor are skipped)
  unsigned long long clock = 0;
  GraphEntry graphs[WSI_DESC_SLOTS];
  cudaStream_t capture_stream = nullptr;
};
DeviceState g_dev[kMaxDevices];
std::mutex g_slot_mutex;
DeviceState* device_state() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return nullptr;
  return &g_dev[dev];
}
int upload_desc(const WsiModelDesc* d, cudaStream_t s, int* slot) {
  const ConfigUnit* u = unit_of(d);
  if (!u) return fail(WSI_ERR_UNSUPPORTED, "no kernel instantiated for this number of pollutants%s");
  DeviceState* ds = device_state();
  if (!ds) return fail(WSI_ERR_LAUNCH, "no current CUDA device (or more than 16 devices)%s");
  std::lock_guard<std::mutex> lock(g_slot_mutex);
  int found = -1, lru = 0;
  for (int i = 0; i < WSI_DESC_SLOTS; i++) {
    if (ds->slots[i].key == (const void*)d) found = i;
    if (ds->slots[i].last_use < ds->slots[lru].last_use) lru = i;
  }
  if (found < 0) {
    found = lru;
    if (ds->slots[lru].key != nullptr) cudaDeviceSynchronize();  // its last user may still be running
    ds->slots[lru].key = (const void*)d;
    ds->slots[lru].uploaded = false;
  }
  ds->slots[found].last_use = ++ds->clock;
  *slot = found;
  // A copy to a __constant__ symbol is ordered against every kernel that may read the bank; a
  // model that is stepped in chunks re-sends the same bytes each time, so compare first.  (The
  // image is only trusted for the stream-ordered case: same model, same stream, which is the rule.)
  if (ds->slots[found].uploaded && memcmp(&ds->image[found], d, sizeof(WsiModelDesc)) == 0) return WSI_OK;
  ds->slots[found].uploaded = false;
  cudaError_t err = u->upload(d, *slot, s);
  if (err != cudaSuccess) return fail(WSI_ERR_LAUNCH, "descriptor upload failed: %s", cudaGetErrorString(err));
  memcpy(&ds->image[found], d, sizeof(WsiModelDesc));  // committed only once the copy is enqueued
  ds->slots[found].uploaded = true;
  return WSI_OK;
}

int dispatch(const WsiModelDesc* d, int slot, int64_t t, int b, int e, uint32_t flags, cudaStream_t s) {
  const ConfigUnit* u = unit_of(d);
  if (!u) return fail(WSI_ERR_UNSUPPORTED, "no kernel instantiated for this number of pollutants%s");
  return u->dispatch(d, slot, t, b, e, flags, s);
}

// ---- CUDA graph of one timestep ------------------------------------------------------------------
// A timestep of the network engine is hundreds of small dependent launches (one per level); issued
// one by one they cost more host and launch latency than they compute.  wsi_run_steps therefore
// captures ONE timestep -- launched with t = -1, i.e. "read the step from net_scratch[2]" -- into a
// graph, instantiates it once per model and replays it for every step; wsi_advance_kernel at the
// end of the graph moves the device-side step counter.  The graph is re-captured when anything in
// the descriptor (or the level sizes) changes.
bool graphs_enabled() {
  static const bool on = [] {
    const char* e = getenv("WSI_B200_NO_GRAPH");
    return !(e && e[0] && e[0] != '0');
  }();
  return on;
}
std::mutex g_graph_mutex;

unsigned long long fnv(unsigned long long h, const void* p, size_t n) {
  const unsigned char* b = (const unsigned char*)p;
  for (size_t i = 0; i < n; i++) h = (h ^ b[i]) * 1099511628211ull;
  return h;
}

int run_steps_graph(const WsiModelDesc* d, int slot, int64_t t0, int64_t t1, cudaStream_t stream) {
  std::lock_guard<std::mutex> lock(g_graph_mutex);
  // everything a captured launch bakes in (grid sizes, kernel arguments); pointers and the record
  // window are read from the descriptor in constant memory when the graph runs
  const int32_t baked[] = {slot, d->P, (int32_t)d->additive_mask, d->idx[0], d->idx[1], d->idx[2], d->idx[3], d->idx[4],
                           d->idx[5], d->idx[6], d->idx[7], d->R, d->n_nodes, d->n_arcs, d->n_surf, d->surf_class_ptr[0],
                           d->surf_class_ptr[1], d->surf_class_ptr[2], d->surf_class_ptr[3], d->surf_class_ptr[4],
                           d->n_probe, d->n_red, d->n_end, d->n_sched, d->sched_split, d->engine, d->n_net_levels, d->n_net_calc, d->n_net_final,
                           d->n_net_bnd, d->mb_sums ? 1 : 0, fused_tail_max_r()};
  unsigned long long h = fnv(1469598103934665603ull, baked, sizeof baked);
  if (d->engine == 1) h = fnv(h, d->net_level_ptr_host, sizeof(int32_t) * (size_t)(d->n_net_levels + 1));
  DeviceState* ds = device_state();
  if (!ds) return fail(WSI_ERR_LAUNCH, "no current CUDA device (or more than 16 devices)%s");
  GraphEntry& g = ds->graphs[slot];
  cudaStream_t& g_capture_stream = ds->capture_stream;
  if (g.key != (const void*)d || g.hash != h || !g.exec) {
    if (g.exec) cudaGraphExecDestroy(g.exec);
    g = GraphEntry();
    if (!g_capture_stream && cudaStreamCreateWithFlags(&g_capture_stream, cudaStreamNonBlocking) != cudaSuccess)
      return fail(WSI_ERR_LAUNCH, "cannot create the capture stream%s");
    const long long before = g_launches.load();
    cudaError_t err = cudaStreamBeginCapture(g_capture_stream, cudaStreamCaptureModeThreadLocal);
    if (err != cudaSuccess) return fail(WSI_ERR_LAUNCH, "graph capture failed: %s", cudaGetErrorString(err));
    int rc = dispatch(d, slot, -1, 0, d->n_sched, WSI_RUN_ALL, g_capture_stream);
    unit_of(d)->advance(slot, g_capture_stream);
    g_launches++;
    if (!rc && cudaGetLastError() != cudaSuccess) rc = fail(WSI_ERR_LAUNCH, "wsi_advance_kernel launch failed%s");
    cudaGraph_t graph = nullptr;
    err = cudaStreamEndCapture(g_capture_stream, &graph);
    g.launches = g_launches.load() - before;
    g_launches -= g.launches;  // captured, not executed
    if (rc) {
      if (graph) cudaGraphDestroy(graph);
      return rc;
    }
    if (err != cudaSuccess || !graph) return fail(WSI_ERR_LAUNCH, "graph capture failed: %s", cudaGetErrorString(err));
    err = cudaGraphInstantiate(&g.exec, graph, 0);
    cudaGraphDestroy(graph);
    if (err != cudaSuccess) return fail(WSI_ERR_LAUNCH, "graph instantiation failed: %s", cudaGetErrorString(err));
    g.key = (const void*)d;
    g.hash = h;
  }
  unit_of(d)->set_step(slot, (int)t0, stream);
  WSI_CHECK_LAUNCH("wsi_set_step_kernel");
  for (int64_t t = t0; t < t1; t++) {
    cudaError_t err = cudaGraphLaunch(g.exec, stream);
    if (err != cudaSuccess) return fail(WSI_ERR_LAUNCH, "graph launch failed: %s", cudaGetErrorString(err));
    g_launches += g.launches;
  }
  return WSI_OK;
}

}  // namespace

extern "C" {

int wsi_version(void) { return WSI_ABI_VERSION; }

int wsi_desc_size(void) { return (int)sizeof(WsiModelDesc); }

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
  if (d->n_end > 0 && (!d->end_kind || !d->end_obj)) return fail(WSI_ERR_BAD_DESC, "end_timestep items missing%s");
  if (d->engine != 0 && d->engine != 1) return fail(WSI_ERR_BAD_DESC, "unknown engine%s");
  if (d->engine == 0 && (d->n_probe + d->n_red > 65535 || d->n_end > 65535 || d->n_nodes + d->n_arcs > 65535))
    return fail(WSI_ERR_UNSUPPORTED,
                "more than 65535 probes / nodes+arcs / end items per template need the network engine%s");
  if (d->engine == 1 && (d->n_net_levels < 0 || !d->net_level_ptr || !d->net_level_ptr_host || !d->net_level_node ||
                         !d->net_calc_node || !d->net_final_node || !d->net_bnd_node || !d->net_path_ptr || !d->net_path_sender ||
                         !d->net_path_split || !d->net_path_arc_ptr || !d->net_path_arcs || !d->net_item || !d->net_rpath || !d->net_scratch || !d->net_done))
    return fail(WSI_ERR_BAD_DESC, "network plan missing%s");
  if (d->engine == 1 && d->n_surf > 0) return fail(WSI_ERR_UNSUPPORTED, "the network engine has no Land nodes%s");
  if (d->n_surf > 0 && (!d->surf_order || d->surf_class_ptr[0] != 0 || d->surf_class_ptr[4] != d->n_surf))
    return fail(WSI_ERR_BAD_DESC, "surface class order missing%s");
  if ((d->n_arcs > 0 && !d->arc) || (d->n_land > 0 && !d->land) || (d->n_surf > 0 && !d->surf) ||
      (d->n_tank > 0 && !d->tank) || (d->n_river > 0 && !d->river) || (d->n_sewer > 0 && !d->sewer) ||
      (d->n_catch > 0 && !d->catchm))
    return fail(WSI_ERR_BAD_DESC, "group tensor missing%s");
  if (!d->node_forcing || !d->month_idx || !d->dayofyear || !d->doy_adj)
    return fail(WSI_ERR_BAD_DESC, "forcing arrays missing%s");
  if (d->n_probe > 0 && (!d->probe_group || !d->probe_plane || !d->probe_obj || !d->out_probe))
    return fail(WSI_ERR_BAD_DESC, "probe arrays missing%s");
  if (d->n_red < 0 || (d->n_red > 0 && (!d->red_group || !d->red_plane || !d->red_obj || !d->red_out)))
    return fail(WSI_ERR_BAD_DESC, "reduced-probe arrays missing%s");
  if (d->rec_R < 0 || d->rec_r0 < 0 || d->rec_r0 + d->rec_R > d->R) return fail(WSI_ERR_BAD_DESC, "record member window outside [0, R)%s");
  if (d->record_mb && !d->out_mb) return fail(WSI_ERR_BAD_DESC, "record_mb without out_mb%s");
  if (d->mb_sums && !d->mb_partial) return fail(WSI_ERR_BAD_DESC, "mb_sums without mb_partial%s");
  if (d->n_nodes > 0 && (!d->mb_terms || !d->mb_term_slot)) return fail(WSI_ERR_BAD_DESC, "mb_terms missing%s");
  if (!d->node_flags) return fail(WSI_ERR_BAD_DESC, "node_flags missing%s");
  if (!unit_of(d))  // other counts: the caller pads with zero additive pollutants (wsi_b200/backend/flatten.py)
    return fail(WSI_ERR_UNSUPPORTED, "kernels are instantiated for P in {2, 12, 15}%s");
  return WSI_OK;
}

int wsi_num_blocks(const WsiModelDesc* d) { return d ? n_partial_rows(d) : 0; }

int wsi_sched_range(const WsiModelDesc* d, int64_t t, int32_t begin, int32_t end, uint32_t flags, void* stream) {
  int rc = wsi_validate(d);
  if (rc) return rc;
  if (t < 0 || t >= d->T) return fail(WSI_ERR_BAD_DESC, "timestep outside the forcing tables%s");
  if (begin < 0 || end > d->n_sched || begin > end) return fail(WSI_ERR_BAD_DESC, "bad schedule range%s");
  int slot;
  rc = upload_desc(d, (cudaStream_t)stream, &slot);
  if (rc) return rc;
  return dispatch(d, slot, t, begin, end, flags, (cudaStream_t)stream);
}

int wsi_step(const WsiModelDesc* d, int64_t t, void* stream) {
  if (!d) return fail(WSI_ERR_BAD_DESC, "null descriptor%s");
  return wsi_sched_range(d, t, 0, d->n_sched, WSI_RUN_ALL, stream);
}

int wsi_run_steps(const WsiModelDesc* d, int64_t t0, int64_t t1, void* stream) {
  int rc = wsi_validate(d);
  if (rc) return rc;
  if (t0 < 0 || t1 > d->T || t0 > t1) return fail(WSI_ERR_BAD_DESC, "timestep range outside the forcing tables%s");
  int slot;
  rc = upload_desc(d, (cudaStream_t)stream, &slot);
  if (rc) return rc;
  if (t1 - t0 >= 2 && d->net_scratch && graphs_enabled()) return run_steps_graph(d, slot, t0, t1, (cudaStream_t)stream);
  for (int64_t t = t0; t < t1; t++) {
    rc = dispatch(d, slot, t, 0, d->n_sched, WSI_RUN_ALL, (cudaStream_t)stream);
    if (rc) return rc;
  }
  return WSI_OK;
}

int64_t wsi_launch_count(void) { return (int64_t)g_launches.load(); }


}  // extern "C"
