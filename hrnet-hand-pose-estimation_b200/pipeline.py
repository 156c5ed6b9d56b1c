"""Streaming inference over host batches: the public end-to-end entry point.

`StreamingPredictor(model).run(batches)` takes an iterable of pinned host tensors [B,3,H,W] float32 and yields the
decoded joints [B,J,2] (host tensors), overlapping the host->device copy of batch i+1 and the device->host read of
batch i-1 with the network of batch i (two device input slots, one copy stream).  Every batch still pays its own
H2D copy, forward, decode and D2H read; they are just not serialised.
The per-batch work is what tools/evaluate_2D.py:223-238 of the reference does (model -> get_final_preds -> .cpu()).
"""
import torch

from .utils.heatmap_decoding import get_final_preds


class StreamingPredictor:
    def __init__(self, model, use_softmax=None, depth=2):
        self.model = model
        self.use_softmax = (model.variant == "softmax") if use_softmax is None else use_softmax
        self.depth = depth
        self.device = next(model.parameters()).device
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self._slots = None

    def _ensure(self, shape):
        if self._slots is None or self._slots[0]["x"].shape != shape:
            J = self.model.arch.num_joints
            self._slots = [dict(x=torch.empty(shape, dtype=torch.float32, device=self.device),
                                out=torch.empty((shape[0], J, 2), dtype=torch.float32).pin_memory(),
                                copied=torch.cuda.Event(), done=torch.cuda.Event(), used=False)
                           for _ in range(self.depth)]

    def run(self, batches):
        main = torch.cuda.current_stream(self.device)
        pending = []                      # slots whose results have not been handed out yet (FIFO)
        i = 0
        for host in batches:
            self._ensure(tuple(host.shape))
            s = self._slots[i % self.depth]
            if s["used"]:
                if pending and pending[0] is s:      # hand out the oldest result before its slot is reused
                    pending.pop(0)
                    s["done"].synchronize()
                    yield s["out"]
                self.copy_stream.wait_event(s["done"])
            with torch.cuda.stream(self.copy_stream):
                s["x"].copy_(host, non_blocking=True)
                s["copied"].record(self.copy_stream)
            main.wait_event(s["copied"])
            out = self.model(s["x"])
            coords = get_final_preds(out[0], self.use_softmax)
            s["out"].copy_(coords, non_blocking=True)
            s["done"].record(main)
            s["used"] = True
            pending.append(s)
            i += 1
        for s in pending:
            s["done"].synchronize()
            yield s["out"]


class StreamingTrainer:
    """Training over host batches: `StreamingTrainer(engine).run(batches)` takes an iterable of pinned host tuples
    (images [B,3,H,W], gt heat maps [B,J,h,w], gt joints [B,J,2], visibility [B,J]) and yields, per step and in order,
    the losses [total, heat-map, pose2d] as a host tensor.  What lib/core/function.py:24-162 (train_helper) does per
    iteration - .cuda() the batch, forward, losses, backward, optimizer step, .item() the losses - with the copies taken
    off the critical path: two device input slots filled by a copy stream (the H2D of batch i+1 runs under the network of
    batch i) and the losses of step i read back while step i+1 is already queued (the reference's .item() per term per
    step, function.py:1375, stalls the GPU instead).  Every step still pays its own H2D copy and its own D2H read."""

    def __init__(self, engine, allreduce=None, depth=2):
        self.engine, self.allreduce, self.depth = engine, allreduce, depth
        self.device = engine.device
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self._slots = None

    def _ensure(self, host):
        shapes = tuple(tuple(t.shape) for t in host)
        if self._slots is None or self._slots[0]["shapes"] != shapes:
            self._slots = [dict(shapes=shapes, dev=[torch.empty(t.shape, dtype=t.dtype, device=self.device) for t in host],
                                out=torch.empty(3, dtype=torch.float32).pin_memory(), copied=torch.cuda.Event(),
                                done=torch.cuda.Event(), used=False) for _ in range(self.depth)]

    def run(self, batches):
        main = torch.cuda.current_stream(self.device)
        pending = []
        i = 0
        for host in batches:
            self._ensure(host)
            s = self._slots[i % self.depth]
            if s["used"]:
                if pending and pending[0] is s:          # results are handed out in order, before their slot is reused
                    pending.pop(0)
                    s["done"].synchronize()
                    yield s["out"]
                self.copy_stream.wait_event(s["done"])
            with torch.cuda.stream(self.copy_stream):
                for d, h in zip(s["dev"], host):
                    d.copy_(h, non_blocking=True)
                s["copied"].record(self.copy_stream)
            main.wait_event(s["copied"])
            p = self.engine.train_step(*s["dev"], allreduce=self.allreduce)
            s["out"].copy_(p.losses, non_blocking=True)
            s["done"].record(main)
            s["used"] = True
            pending.append(s)
            if len(pending) >= self.depth:               # keep at most depth-1 steps in flight behind the consumer
                o = pending.pop(0)
                o["done"].synchronize()
                yield o["out"]
            i += 1
        for s in pending:
            s["done"].synchronize()
            yield s["out"]
