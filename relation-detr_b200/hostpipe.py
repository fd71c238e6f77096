"""Host-buffer front end of the MSDA operator: H2D copy -> forward+backward -> D2H copy, pipelined.

The hot path is 2.5 ms of GPU work per batch but 640 MB in and 640 MB out over PCIe, so a caller whose
tensors live in host memory is transfer-bound.  ``MsdaHostPipeline`` keeps three CUDA streams busy
(copy-in of step i+1, compute of step i, copy-out of step i-1) with ``depth`` device input buffers and
pinned host output buffers, at most ``depth`` steps in flight (default 3: with two, the copy-in of step i waits for the
copy-out of step i-2, which makes the period (copy-in + compute + copy-out) / 2 = 14.7 ms at configs[1] instead of the
slower copy's 14.0 ms -- profiles/r02ah_exp_e2e_depth.txt); every step still copies all of its inputs in and all of its results out.  The
compute goes through the public operator (``MultiScaleDeformableAttnFunction.apply`` + autograd).
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from .ops import MultiScaleDeformableAttnFunction

_IN_KEYS = ("value", "sampling_locations", "attention_weights", "grad_output")
_OUT_KEYS = ("out", "grad_value", "grad_loc", "grad_attn")


class MsdaHostPipeline:
    def __init__(self, spatial_shapes: torch.Tensor, level_start_index: torch.Tensor, device, depth: int = 3):
        self.device = torch.device(device)
        self.ss = spatial_shapes.to(self.device)
        self.lsi = level_start_index.to(self.device)
        self.depth = depth
        self.s_in = torch.cuda.Stream(self.device)
        self.s_compute = torch.cuda.Stream(self.device)
        self.s_out = torch.cuda.Stream(self.device)
        self.dev_in: List[Optional[Dict[str, torch.Tensor]]] = [None] * depth
        self.host_out: List[Optional[Dict[str, torch.Tensor]]] = [None] * depth
        self.ev_in_ready = [torch.cuda.Event() for _ in range(depth)]
        self.ev_compute_done = [torch.cuda.Event() for _ in range(depth)]
        self.ev_out_done = [torch.cuda.Event() for _ in range(depth)]
        self.step_index = 0
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    def submit(self, host: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        """Enqueue one step on pinned host tensors (keys: value, sampling_locations, attention_weights,
        grad_output).  Returns the pinned host tensors the results will land in.  They are valid after ``wait()``
        and are OVERWRITTEN by the ``depth``-th submit that follows: consume (or copy) them before resubmitting
        that slot."""
        i = self.step_index
        b = i % self.depth
        self.step_index += 1
        if i >= self.depth:
            # Bound the run-ahead to `depth` steps: the operator allocates its outputs (640 MB per step at configs[1]) from
            # torch's caching allocator, and a block handed to the copy-out stream is not reusable before that copy has
            # finished -- a host that submits ten steps at once makes the allocator cudaMalloc (and thereby synchronise)
            # gigabytes inside the pipeline (measured: 14 ms per step with a warm cache, 27-47 ms without).
            self.ev_out_done[b].synchronize()
        with torch.cuda.stream(self.s_in):
            if self.dev_in[b] is None:
                self.dev_in[b] = {k: torch.empty(host[k].shape, dtype=host[k].dtype, device=self.device) for k in _IN_KEYS}
                for t in self.dev_in[b].values():  # allocated under s_in, read by the kernels on s_compute: keep the
                    t.record_stream(self.s_compute)  # caching allocator from recycling the block while compute is pending
            else:
                self.s_in.wait_event(self.ev_compute_done[b])  # step i-depth has finished reading this buffer
            for k in _IN_KEYS:
                self.dev_in[b][k].copy_(host[k], non_blocking=True)
            self.ev_in_ready[b].record(self.s_in)
        self.h2d_bytes = sum(host[k].numel() * host[k].element_size() for k in _IN_KEYS)
        with torch.cuda.stream(self.s_compute):
            self.s_compute.wait_event(self.ev_in_ready[b])
            d = self.dev_in[b]
            v = d["value"].detach().requires_grad_(True)
            loc = d["sampling_locations"].detach().requires_grad_(True)
            attn = d["attention_weights"].detach().requires_grad_(True)
            out = MultiScaleDeformableAttnFunction.apply(v, self.ss, self.lsi, loc, attn, 64)
            out.backward(d["grad_output"])  # autograd runs the backward on the forward's stream
            results = dict(out=out.detach(), grad_value=v.grad, grad_loc=loc.grad, grad_attn=attn.grad)
            self.ev_compute_done[b].record(self.s_compute)
        with torch.cuda.stream(self.s_out):
            self.s_out.wait_event(self.ev_compute_done[b])
            if self.host_out[b] is None:
                self.host_out[b] = {k: torch.empty(t.shape, dtype=t.dtype).pin_memory() for k, t in results.items()}
            # (the previous use of this pinned buffer was copied on this same stream => ordered)
            for k in _OUT_KEYS:
                results[k].record_stream(self.s_out)
                self.host_out[b][k].copy_(results[k], non_blocking=True)
            self.ev_out_done[b].record(self.s_out)
        self.d2h_bytes = sum(t.numel() * t.element_size() for t in results.values())
        return self.host_out[b]

    def wait(self) -> None:
        self.s_out.synchronize()
        self.s_compute.synchronize()
        self.s_in.synchronize()
