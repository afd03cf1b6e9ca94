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
        # three host slots for two device rows: the buffer handed out by step() (result of step t-2)
        # is not written again before the NEXT call, so it stays valid until then
        self.NH = 3
        self.h_out = [torch.empty(self.d_out.shape[1:], dtype=torch.float64).pin_memory() for _ in range(self.NH)]
        self.h_cnt = [torch.empty(1, dtype=torch.int32).pin_memory() for _ in range(self.NH)]
        self.k_done = [torch.cuda.Event() for _ in range(2)]
        self.c_done = [torch.cuda.Event() for _ in range(self.NH)]
        self._pending = {}  # timestep -> host slot of results still in flight / not yet handed out
        self.h2d_bytes_per_step = self.d_forcing[0].numel() * 8
        self.d2h_bytes_per_step = self.h_out[0].numel() * 8 + 4
        self.last = None

    def step(self, t, forcing_row_host):
        """Advance timestep ``t``; returns ``(outputs, violations)`` of step ``t-2`` (pinned host
        tensor ``[n_probe, R]`` -- valid until the next call of ``step``/``drain``) or None."""
        torch = self.torch
        row = t % 2  # device record row, reused every second step
        hs = t % self.NH
        done = None
        if (t - 2) in self._pending:
            h = self._pending.pop(t - 2)
            self.c_done[h].synchronize()  # result of step t-2 is in h_out[h]; its device row is free
            done = (self.h_out[h], int(self.h_cnt[h][0]))
            self.last = done
        with torch.cuda.stream(self.stream):
            self.d_forcing[t].copy_(forcing_row_host, non_blocking=True)
            self.runner.buffers.set_rec_window(t - row)
            self.runner.step(t, stream=self.stream)
            self.k_done[row].record(self.stream)
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self.k_done[row])
            self.h_out[hs].copy_(self.d_out[row], non_blocking=True)
            self.h_cnt[hs].copy_(self.d_cnt[t : t + 1], non_blocking=True)
            self.c_done[hs].record(self.copy_stream)
        self._pending[t] = hs
        return done

    def drain(self):
        """Wait for everything in flight; returns the outstanding host results in timestep order
        (numpy views of the pinned buffers, valid until the next call of ``step``)."""
        out = []
        for t in sorted(self._pending):
            h = self._pending[t]
            self.c_done[h].synchronize()
            out.append(self.h_out[h].numpy())
        self._pending = {}
        self.stream.synchronize()
        return out
