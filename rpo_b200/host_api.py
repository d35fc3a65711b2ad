"""Host-buffer entry point of the projection path: the call a user of the reference makes with CPU data.

The reference's callers hold numpy / CPU-torch batches (replay samples, rollout states — rpo_ddpg.py:93,165-174)
and receive CPU arrays back (rpo_ddpg.py:108-109).  `HostProjector` takes pinned host buffers, streams them to the
GPU in chunks on several CUDA streams (H2D copy, kernels and D2H copy of different chunks overlap on the two DMA
engines and the SMs), and writes the results into pinned host buffers.  State rows cross PCIe whole (contiguous
DMA): uploading only the 33 columns the path reads with a strided copy (rpo_copy_rows_h2d, 132 of every 228 bytes)
was measured 3x SLOWER end to end (3.1e7 vs 9.9e7 projected actions/s) — the DMA engine handles it row by row.
Pass `state_cols=33` only when the caller already holds a compact (B,33) [pd|qd|soc] matrix.
Device staging buffers are allocated once per stream slot.  Used by bench.py for the end-to-end (`e2e`) number.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import ops


def bind_host_thread_to_device(device) -> bool:
    """Pin the calling thread to the CPU cores NVML reports as closest to `device` (same NUMA node / PCIe root), so that
    pinned buffers allocated afterwards are first-touched in that node's memory and the H2D / D2H DMA does not cross
    the socket interconnect.  Matters when several ranks drive their own GPU from one host; a no-op (False) if NVML or
    the affinity call is unavailable.  Call it before allocating the pinned buffers."""
    try:
        import pynvml
        pynvml.nvmlInit()
        p = torch.cuda.get_device_properties(torch.device(device))
        bus_id = "%08x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        try:
            handle = pynvml.nvmlDeviceGetHandleByPciBusId(bus_id)
        except TypeError:
            handle = pynvml.nvmlDeviceGetHandleByPciBusId(bus_id.encode())
        pynvml.nvmlDeviceSetCpuAffinity(handle)
        return True
    except Exception:
        return False


class HostProjector(object):
    def __init__(self, device, chunk: int = 1 << 16, n_streams: int = 4, corr_lr: float = 1e-4, momentum: float = 0.0,
                 corr_eps: float = 1e-5, n_steps: int = 10, tol: float = 1e-5, max_iters: int = 50):
        self.device = torch.device(device)
        self.chunk = int(chunk)
        self.streams = [torch.cuda.Stream(device=self.device) for _ in range(n_streams)]
        self.p = dict(lr=corr_lr, momentum=momentum, corr_eps=corr_eps, n_steps=n_steps, tol=tol, max_iters=max_iters)
        f32 = dict(dtype=torch.float32, device=self.device)
        self.slots = [dict(x=torch.empty(self.chunk * 57, **f32), z=torch.empty(self.chunk, 14, **f32),
                           g=torch.empty(self.chunk, 43, **f32)) for _ in range(n_streams)]
        self.h2d_bytes = 0
        self.d2h_bytes = 0
        self.launches = 0

    def _schedule(self, B: int):
        """Row ranges of the chunks.  The pipeline's fill (first chunk's upload, nothing to overlap with) and drain (last
        chunk's kernels + download) are pure latency, so the batch is cut small at both ends — chunk/4, chunk/2, chunk, ...,
        chunk, chunk/2, chunk/4 — when it is large enough for that to matter."""
        c = self.chunk
        if B < 6 * c or c < 4096:
            return [(b0, min(B, b0 + c)) for b0 in range(0, B, c)]
        head = [c // 4, c // 2]
        out, b0 = [], 0
        for n in head:
            out.append((b0, b0 + n))
            b0 += n
        tail = sum(head)
        while B - tail - b0 > 0:
            n = min(c, B - tail - b0)
            out.append((b0, b0 + n))
            b0 += n
        for n in reversed(head):
            out.append((b0, b0 + n))
            b0 += n
        return out

    @staticmethod
    def pinned(*shape, dtype=torch.float32):
        return torch.empty(*shape, dtype=dtype, pin_memory=True)

    def project_backward(self, state: torch.Tensor, z: torch.Tensor, dl_dy: Optional[torch.Tensor],
                         y_corr_out: torch.Tensor, dl_dz_out: Optional[torch.Tensor]):
        """state (B, >=33) rows [pd|qd|soc|...], z (B,14), dl_dy (B,43): pinned CPU float32 ->
        y_corr_out (B,43), dl_dz_out (B,14) pinned CPU.  Completion + correction (+ GRG backward of the completion
        when dl_dy is given).  Blocks until the results have landed in host memory."""
        B = z.shape[0]
        p = self.p
        cols = state.shape[1]
        assert state.dtype == torch.float32 and state.is_contiguous() and 33 <= cols <= 57
        cur = torch.cuda.current_stream(self.device)
        for s in self.streams:
            s.wait_stream(cur)
        self.h2d_bytes = self.d2h_bytes = self.launches = 0
        with torch.cuda.device(self.device):
            for i, (b0, b1) in enumerate(self._schedule(B)):
                n = b1 - b0
                s = self.streams[i % len(self.streams)]
                slot = self.slots[i % len(self.slots)]
                with torch.cuda.stream(s):
                    d_x, d_z = slot["x"][:n * cols].view(n, cols), slot["z"][:n]
                    d_x.copy_(state[b0:b1], non_blocking=True)
                    d_z.copy_(z[b0:b1], non_blocking=True)
                    self.h2d_bytes += n * (cols + 14) * 4
                    yc, y, _, _ = ops.evopf_project(d_x, d_z, p["lr"], p["momentum"], p["corr_eps"], p["n_steps"],
                                                    tol=p["tol"], max_iters=p["max_iters"], want_y=dl_dy is not None,
                                                    want_flags=False)
                    self.launches += 1
                    y_corr_out[b0:b1].copy_(yc, non_blocking=True)
                    self.d2h_bytes += yc.numel() * 4
                    if dl_dy is not None:
                        d_g = slot["g"][:n]
                        d_g.copy_(dl_dy[b0:b1], non_blocking=True)
                        self.h2d_bytes += d_g.numel() * 4
                        dz, _ = ops.evopf_complete_bwd(y, d_g, need_dx=False)
                        self.launches += 1
                        dl_dz_out[b0:b1].copy_(dz, non_blocking=True)
                        self.d2h_bytes += dz.numel() * 4
        for s in self.streams:
            cur.wait_stream(s)
        cur.synchronize()
        return y_corr_out, dl_dz_out
