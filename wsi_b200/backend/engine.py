"""Drive the CUDA kernels over a flattened model (product path; CUDA only)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi


class GpuRunner:
    """Holds a flattened model on one GPU and steps it through ``libwsi_b200.so``.

    Parameters mirror what ``Model.run`` needs: which planes to record every step
    (``probes``), how many steps the record buffer covers (``rec_T``) and whether the
    per-node mass-balance terms are kept (``record_mb``).  ``engine``: "replica" (one thread per
    replica walks the orchestration schedule), "network" (level-synchronous routing of one large
    dendritic river network, one thread per node) or "auto" (network when the model is eligible
    and has few replicas per node).
    """

    def __init__(self, fm, probes=(), rec_t0=0, rec_T=None, record_mb=False, device="cuda:0",
                 system_sums=True, engine="auto", reduce_probes=(), rec_members=None):
        import torch

        self.torch = torch
        self.lib = capi.load_library()
        self.fm = fm
        self.device = torch.device(device)
        self.buffers = capi.ModelBuffers(fm, device=device, probes=probes, rec_t0=rec_t0, rec_T=rec_T,
                                         record_mb=record_mb, engine=engine, reduce_probes=reduce_probes,
                                         rec_members=rec_members)
        self.desc = self.buffers.desc
        if not system_sums:
            self.desc.mb_sums = C.cast(None, C.POINTER(C.c_double))
        if self.lib.wsi_num_blocks(C.byref(self.desc)) > self.buffers.partial_rows():
            raise RuntimeError("mb_partial scratch smaller than the library needs (block sizes out of sync)")
        capi.check(self.lib, self.lib.wsi_validate(C.byref(self.desc)))
        self._last_t = -1

    def _fresh_arcs(self, t0, stream):
        """Arc transients are valid for the timestep their stamp names (replica engine): stepping a
        timestep again, or going back in time on the same buffers, must not find old stamps."""
        if self.buffers.engine == 0 and int(t0) <= self._last_t and self.fm.n_obj["arc"]:
            plane = self.buffers.group("arc")[self.fm.plane("arc", "stamp")]
            with self.torch.cuda.stream(stream if stream is not None else self.torch.cuda.current_stream(self.device)):
                plane.fill_(-1.0)

    def _stream(self, stream):
        if stream is None:
            stream = self.torch.cuda.current_stream(self.device)
        return C.c_void_p(stream.cuda_stream)

    def run(self, t0, t1, stream=None):
        """Timesteps [t0, t1) back to back (asynchronous on the stream)."""
        with self.torch.cuda.device(self.device):
            self._fresh_arcs(t0, stream)
            self._last_t = max(self._last_t, int(t1) - 1)
            capi.check(self.lib, self.lib.wsi_run_steps(C.byref(self.desc), int(t0), int(t1), self._stream(stream)))

    def step(self, t, stream=None):
        with self.torch.cuda.device(self.device):
            self._fresh_arcs(t, stream)
            self._last_t = max(self._last_t, int(t))
            capi.check(self.lib, self.lib.wsi_step(C.byref(self.desc), int(t), self._stream(stream)))

    def sched_range(self, t, begin, end, flags, stream=None):
        with self.torch.cuda.device(self.device):
            capi.check(
                self.lib,
                self.lib.wsi_sched_range(C.byref(self.desc), int(t), int(begin), int(end), int(flags), self._stream(stream)),
            )

    def synchronize(self):
        self.torch.cuda.synchronize(self.device)

    def set_catch_series(self, ids, values, t0=0):
        """Overwrite the inflow series ``ids`` for timesteps ``t0 ..`` (catch_inputs[t0:t0+n, ids, :])
        with ``values[n, len(ids), V]`` on the device (current stream) -- the boundary series of a
        partitioned network's trunk."""
        ci = self.buffers.array("catch_inputs")
        key = tuple(ids)
        if getattr(self, "_series_key", None) != key:
            self._series_key = key
            self._series_idx = self.torch.as_tensor(list(ids), dtype=self.torch.int64, device=ci.device)
        values = self.torch.as_tensor(values, dtype=ci.dtype).to(ci.device)
        ci[t0:t0 + values.shape[0]].index_copy_(1, self._series_idx, values)

    def private_stream(self):
        """A CUDA stream of this runner's own (models stepped concurrently, parallel.run_partitioned)."""
        if getattr(self, "_stream_own", None) is None:
            self._stream_own = self.torch.cuda.Stream(self.device)
        return self._stream_own

    # ---- results ----
    def out_probe(self):
        return self.buffers.to_numpy("out_probe")

    def reduced(self):
        """``(sum, max)`` over the timesteps stepped so far of every reduced probe, ``[n_red, R]`` each."""
        a = self.buffers.to_numpy("red_out")
        return a[0], a[1]

    def out_mb(self):
        return self.buffers.to_numpy("out_mb") if self.buffers.record_mb else None

    def mb_sums(self):
        return self.buffers.to_numpy("mb_sums")

    def mb_count(self):
        return self.buffers.to_numpy("mb_count")

    def node_flags(self):
        return self.buffers.to_numpy("node_flags").view(np.uint32)

    def groups(self):
        return self.buffers.groups_numpy()

    def launch_count(self):
        return int(self.lib.wsi_launch_count())
