"""Per-timestep ensemble API with HOST buffers (the end-to-end path of bench.py).

``EnsembleSession.step(t, forcing_row)`` takes the step's forcing from pinned host memory,
advances every member one timestep on the GPU and brings the recorded arc VQIPs and the
mass-balance violation count back to pinned host memory.  Copies run on a second stream and
are double-buffered, so the D2H of step t overlaps the kernel of step t+1.
"""
from __future__ import annotations

from .backend.engine import GpuRunner
from .backend.probes import ProbeSet


class EnsembleSession:
    def __init__(self, fm, record_arcs, device="cuda:0"):
        import torch

        self.torch = torch
        self.fm = fm
        self.ps = ProbeSet(fm).add_flows(record_arcs)
        self.runner = GpuRunner(fm, probes=self.ps.probes, rec_t0=0, rec_T=2, record_mb=False, device=device)
        dev = self.runner.device
        self.stream = torch.cuda.Stream(device=dev)
        self.copy_stream = torch.cuda.Stream(device=dev)
        self.d_forcing = self.runner.buffers.array("node_forcing")
        self.d_out = self.runner.buffers.array("out_probe")  # [2, n_probe, R]
        self.d_cnt = self.runner.buffers.array("mb_count")
        self.h_out = [torch.empty(self.d_out.shape[1:], dtype=torch.float64).pin_memory() for _ in range(2)]
        self.h_cnt = [torch.empty(1, dtype=torch.int32).pin_memory() for _ in range(2)]
        self.k_done = [torch.cuda.Event() for _ in range(2)]
        self.c_done = [torch.cuda.Event() for _ in range(2)]
        self._pending = [False, False]
        self.h2d_bytes_per_step = self.d_forcing[0].numel() * 8
        self.d2h_bytes_per_step = self.h_out[0].numel() * 8 + 4
        self.last = None

    def step(self, t, forcing_row_host):
        """Advance timestep ``t``; returns the host result of the step whose buffer slot is
        being recycled (step t-2) or None."""
        torch = self.torch
        slot = t % 2
        done = None
        if self._pending[slot]:
            self.c_done[slot].synchronize()  # result of step t-2 is now in h_out[slot]
            done = (self.h_out[slot], int(self.h_cnt[slot][0]))
            self.last = done
        with torch.cuda.stream(self.stream):
            self.d_forcing[t].copy_(forcing_row_host, non_blocking=True)
            self.stream.wait_event(self.c_done[slot]) if self._pending[slot] else None
            self.runner.buffers.set_rec_window(t - slot)
            self.runner.step(t, stream=self.stream)
            self.k_done[slot].record(self.stream)
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self.k_done[slot])
            self.h_out[slot].copy_(self.d_out[slot], non_blocking=True)
            self.h_cnt[slot].copy_(self.d_cnt[t : t + 1], non_blocking=True)
            self.c_done[slot].record(self.copy_stream)
        self._pending[slot] = True
        return done

    def drain(self):
        """Wait for everything in flight; returns the outstanding host results (numpy views)."""
        out = []
        for slot in (0, 1):
            if self._pending[slot]:
                self.c_done[slot].synchronize()
                out.append(self.h_out[slot].numpy())
                self._pending[slot] = False
        self.stream.synchronize()
        return out
