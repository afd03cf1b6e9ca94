// Host-side glue shared by the C ABI (wsi_api.cu) and the per-configuration translation units
// (wsi_cfg_*.cu).  Every pollutant configuration is compiled in its own unit (they build in
// parallel) with its own copy of the kernels and of the __constant__ descriptor slots, and exports
// a small table of host functions.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

#include "../../include/wsi_b200.h"

#ifndef WSI_DESC_SLOTS
#define WSI_DESC_SLOTS 32
#endif
#ifndef WSI_STEP_BLOCK
#define WSI_STEP_BLOCK 256
#endif

namespace wsi_host {

extern thread_local char g_err[512];
extern std::atomic<long long> g_launches;
int fail(int code, const char* fmt, const char* detail = "");
bool cluster_enabled();
int fused_tail_max_r();

constexpr int kBlock = WSI_STEP_BLOCK;  // replicas per block of the per-replica schedule kernel
#ifndef WSI_SURF_BLOCK
#define WSI_SURF_BLOCK 128
#endif
constexpr int kSurfBlock = WSI_SURF_BLOCK;         // surfaces per block of the surface kernels
constexpr int kObjBlock = 128;          // replicas per block of the object-parallel kernels

inline int n_blocks(const WsiModelDesc* d) { return (d->R + kBlock - 1) / kBlock; }
inline int n_obj_blocks(const WsiModelDesc* d) { return (d->R + kObjBlock - 1) / kObjBlock; }
inline int n_flat_blocks(const WsiModelDesc* d, int n_items) {
  return (int)(((size_t)n_items * (size_t)d->R + kObjBlock - 1) / kObjBlock);
}
inline int n_partial_rows(const WsiModelDesc* d) {
  return d->engine == 1 ? n_flat_blocks(d, d->n_nodes + d->n_arcs) : (d->n_nodes + d->n_arcs) * n_obj_blocks(d);
}

// what one configuration unit exports
struct ConfigUnit {
  const char* name;
  // one timestep (or part of it): the launch sequence of DESIGN.md 2a / 2b; t < 0: graph mode
  int (*dispatch)(const WsiModelDesc* d, int slot, int64_t t, int b, int e, uint32_t flags, cudaStream_t s);
  cudaError_t (*upload)(const WsiModelDesc* d, int slot, cudaStream_t s);  // into the unit's constant slot
  void (*set_step)(int slot, int t, cudaStream_t s);                        // device-side step counter
  void (*advance)(int slot, cudaStream_t s);
};
extern const ConfigUnit wsi_unit_default12, wsi_unit_oxford12, wsi_unit_runtime12, wsi_unit_runtime2, wsi_unit_runtime15;

constexpr uint32_t kOxfordMask = 0x1FFCu;  // 12 pollutants, every one additive but plane 1 (temperature)

}  // namespace wsi_host
