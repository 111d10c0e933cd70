"""Host-side inference loop for a model whose token / channel mixers run in the sm_100a kernels: pinned host batches in,
pinned host logits out, with the host->device copy of batch i+1 and the device->host copy of result i-1 overlapped with
the compute of batch i (two device staging buffers, one copy stream, CUDA events — no host synchronisation inside the loop).

This is the plumbing a serving process or an evaluation loop puts around ``model(x)`` (the reference's own loops —
speed_gpu.py:11-27, engine.py evaluate — copy and compute back to back on one stream); it is what ``bench.py`` times as ``e2e``.
"""
from __future__ import annotations

from typing import Iterable, Iterator, Optional

import torch


class PipelinedInference:
    """``for logits_host in PipelinedInference(model, autocast_dtype).run(host_batches): ...``

    ``host_batches`` yields pinned CPU tensors of one fixed shape/dtype; every result is written to one of THREE pinned
    host buffers used in rotation, so a yielded tensor stays valid until the next-but-one ``next()`` (while result i is in
    the consumer's hands, result i+1 is being copied out and result i+2 goes to the third buffer)."""

    def __init__(self, model: torch.nn.Module, autocast_dtype: Optional[torch.dtype] = torch.bfloat16, device: Optional[torch.device] = None,
                 cuda_graph: bool = False):
        """``cuda_graph=True``: the forward of each of the two staging buffers is captured once in a CUDA graph (after an eager warm-up run
        that fills the packed-weight caches) and replayed afterwards — one graph launch per batch instead of ~50 kernel launches plus the
        Python between them.  Every kernel of the eval path is stream-ordered and allocation-free, so the capture is legal
        (tests/test_model.py::test_whole_model_is_cuda_graph_capturable); the graphs are rebuilt when the batch shape changes."""
        self.model = model.eval()
        self.dtype = autocast_dtype
        self.device = device or next(model.parameters()).device
        self.copy_stream = torch.cuda.Stream(self.device)
        self._xin = [None, None]
        self._yout = [None, None, None]
        self.cuda_graph = cuda_graph
        self._graphs = [None, None]      # (graph, static output) per staging buffer

    def _forward(self, cur: int) -> torch.Tensor:
        x = self._xin[cur]
        if not self.cuda_graph:
            with torch.autocast("cuda", dtype=self.dtype, enabled=self.dtype is not None):
                return self.model(x)
        g = self._graphs[cur]
        if g is None or g[2] is not x:
            with torch.autocast("cuda", dtype=self.dtype, enabled=self.dtype is not None):
                self.model(x)                                     # eager warm-up on this very buffer (weight caches, kernel attributes)
            main = torch.cuda.current_stream(self.device)
            side = torch.cuda.Stream(self.device)
            side.wait_stream(main)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.stream(side):
                with torch.cuda.graph(graph, stream=side):
                    with torch.autocast("cuda", dtype=self.dtype, enabled=self.dtype is not None):
                        y = self.model(x)
            main.wait_stream(side)
            g = (graph, y, x)
            self._graphs[cur] = g
        g[0].replay()
        return g[1]

    def _stage(self, i: int, xh: torch.Tensor) -> torch.cuda.Event:
        """copy host batch -> device staging buffer i on the copy stream; returns the event that marks its arrival"""
        if self._xin[i] is None or self._xin[i].shape != xh.shape or self._xin[i].dtype != xh.dtype:
            # a fresh block from the caching allocator may have just been freed by a kernel that is still queued on the main stream:
            # the copy stream must not write it before the main stream has drained to this point
            self._xin[i] = torch.empty(xh.shape, dtype=xh.dtype, device=self.device)
            self.copy_stream.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(self.copy_stream):
            self._xin[i].copy_(xh, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.copy_stream)
        return ev

    @torch.no_grad()
    def run(self, host_batches: Iterable[torch.Tensor]) -> Iterator[torch.Tensor]:
        main = torch.cuda.current_stream(self.device)
        it = iter(host_batches)
        try:
            nxt = next(it)
        except StopIteration:
            return
        arrived = self._stage(0, nxt)
        consumed = [None, None]          # compute-done events: a staging buffer may be refilled only after its batch was consumed
        i = 0
        pending = None                   # (host result buffer, its D2H-done event) of the previous step
        while nxt is not None:
            cur = i & 1
            main.wait_event(arrived)
            y = self._forward(cur)
            consumed[cur] = torch.cuda.Event()
            consumed[cur].record(main)
            try:
                nxt = next(it)
            except StopIteration:
                nxt = None
            if nxt is not None:          # prefetch batch i+1 while batch i computes
                if consumed[cur ^ 1] is not None:
                    self.copy_stream.wait_event(consumed[cur ^ 1])
                arrived = self._stage(cur ^ 1, nxt)
            yo = i % 3
            if self._yout[yo] is None or self._yout[yo].shape != y.shape or self._yout[yo].dtype != y.dtype:
                self._yout[yo] = torch.empty(y.shape, dtype=y.dtype).pin_memory()
            self._yout[yo].copy_(y, non_blocking=True)       # D2H on the main stream, behind the compute of this batch
            done = torch.cuda.Event()
            done.record(main)
            if pending is not None:
                pending[1].synchronize()
                yield pending[0]
            pending = (self._yout[yo], done)
            i += 1
        if pending is not None:
            pending[1].synchronize()
            yield pending[0]
