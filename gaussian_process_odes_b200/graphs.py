"""CUDA-graph capture of a whole ELBO forward+backward step.

The reference's real training shapes are tiny (1-6 trajectories, 125-3000 shooting segments): after the integrator is
one kernel, a step is ~100 small launches whose CPU dispatch cost (PyTorch op overhead, ctypes calls) dominates.
Everything in the step is stream-ordered and shape-static, so it is captured once and replayed: the only per-step host
work is drawing the GP cache's random numbers with numpy (same generator, same order as the reference,
``src/core/dsvgp.py:100-103,83``) and copying them into static device buffers.

    step = GraphedStep(model, lambda: compute_loss_shooting(model, ys, ts, num_samples=5)[0])
    for it in range(n):
        loss = step()          # parameters' .grad hold the new gradients; call optimizer.step() yourself

Both solvers can be captured. For dopri5 the number of accepted steps stays on the device (the adjoint and the
gradient contraction read it there, ``gpode_dopri5_bwd_dev`` / ``gpode_param_grad_dev``) and the checkpoint capacity is
fixed at capture time (8 x grid length, at least 64 steps); ``GraphedStep.check()`` reads the solver status back and
raises if an integration ran out of capacity or failed.
"""
import numpy as np
import torch

from .core import dsvgp as _dsvgp
from .core import kernels as _kernels


class _StaticDraws:
    """Replaces the three host samplers by static device buffers that ``refresh()`` refills in call order."""

    RING = 4   # pinned staging buffers per slot: the host may run at most RING refreshes ahead of the stream

    def __init__(self, device):
        self.device = device
        self.slots = []      # (kind, shape, pinned host ring [{"buf", "ev"}], device_buffer)
        self.turn = []       # next ring entry of every slot
        self.cursor = 0
        self.recording = True
        self._saved = None

    def _provide(self, kind, shape):
        shape = tuple(shape)
        if self.recording:
            host = [{"buf": torch.empty(shape, dtype=torch.float32).pin_memory(), "ev": None}
                    for _ in range(self.RING)]
            dev = torch.empty(shape, dtype=torch.float32, device=self.device)
            self.slots.append((kind, shape, host, dev))
            self.turn.append(0)
            self._fill(len(self.slots) - 1)
            return dev
        kind_, shape_, _, dev = self.slots[self.cursor]
        assert (kind_, shape_) == (kind, shape), "sampler call sequence changed between capture and replay"
        self.cursor += 1
        return dev

    def _fill(self, i):
        kind, shape, ring, dev = self.slots[i]
        if kind == "uniform":
            arr = np.random.uniform(low=0.0, high=1.0, size=shape)
        else:
            arr = np.random.normal(size=shape)
        # A replay takes milliseconds, a refresh microseconds: the host runs ahead of the stream. A pinned buffer is
        # therefore re-used only after the event recorded behind ITS last copy (same protocol as
        # misc.torch_utils._PinnedRing); overwriting it earlier would tear the draw of a step still in flight.
        ent = ring[self.turn[i]]
        self.turn[i] = (self.turn[i] + 1) % self.RING
        if ent["ev"] is not None:
            ent["ev"].synchronize()
        ent["buf"].copy_(torch.from_numpy(arr.astype(np.float32)))
        dev.copy_(ent["buf"], non_blocking=True)
        if ent["ev"] is None:
            ent["ev"] = torch.cuda.Event()
        ent["ev"].record()

    def refresh(self):
        for i in range(len(self.slots)):
            self._fill(i)

    def install(self):
        self._saved = (_dsvgp.sample_normal, _dsvgp.sample_uniform, _kernels.sample_normal)
        _dsvgp.sample_normal = lambda shape, seed=None: self._provide("normal", shape)
        _dsvgp.sample_uniform = lambda shape, seed=None: self._provide("uniform", shape)
        _kernels.sample_normal = lambda shape, seed=None: self._provide("normal", shape)

    def uninstall(self):
        if self._saved is not None:
            _dsvgp.sample_normal, _dsvgp.sample_uniform, _kernels.sample_normal = self._saved
            self._saved = None


class GraphedStep:
    def __init__(self, model, loss_fn, warmup=3):
        params = [p for p in model.parameters() if p.requires_grad]
        if not params or not params[0].is_cuda:
            raise RuntimeError("GraphedStep needs a model on a CUDA device")
        self.model, self.loss_fn, self.params = model, loss_fn, params
        # a cache left over from an eager step keeps that step's autograd graph -- and the parameters' AccumulateGrad
        # nodes, bound to the eager stream -- alive, which would break capture on the graph's side stream
        for mod in model.modules():
            if isinstance(mod, _dsvgp.DSVGP_Layer):
                for attr in ("nu", "rff_omega", "rff_phase", "rff_weights", "_ell_dimwise", "_var_dimwise"):
                    if isinstance(getattr(mod, attr, None), torch.Tensor):
                        setattr(mod, attr, getattr(mod, attr).detach())
        self.draws = _StaticDraws(params[0].device)
        self.draws.install()
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for i in range(warmup):  # warm-up on a side stream (allocator, lazy inits, grid caches)
                    if i == 1:
                        self.draws.recording = False
                    self.draws.cursor = 0
                    for p in params:
                        p.grad = None
                    loss_fn().backward()
            torch.cuda.current_stream().wait_stream(side)
            self.draws.recording = False
            for p in params:
                p.grad = None
            self.graph = torch.cuda.CUDAGraph()
            self.draws.cursor = 0
            from . import ops as _ops
            n_before = len(_ops.DEVICE_COUNT_STATS)
            with torch.cuda.graph(self.graph):
                self.loss = loss_fn()
                self.loss.backward()
            self.dopri5_stats = _ops.DEVICE_COUNT_STATS[n_before:]   # static tensors of the captured integrations
            del _ops.DEVICE_COUNT_STATS[n_before:]
        finally:
            self.draws.uninstall()

    def check(self):
        """Synchronises and raises if a captured dopri5 integration of the LAST replay failed (status 1 attempt limit,
        2 step-size underflow, 3 checkpoint capacity exceeded). Returns the list of [nfe, accepted, rejected, status]."""
        out = [[int(v) for v in s.cpu()] for s in self.dopri5_stats]
        for st in out:
            if st[3] != 0:
                from ._lib import GpodeError
                raise GpodeError("captured dopri5 integration failed: status %d (accepted %d steps)" % (st[3], st[1]))
        return out

    def __call__(self):
        """New GP draw (host numpy -> static buffers), replay fwd+bwd; returns the static loss tensor."""
        self.draws.refresh()
        self.graph.replay()
        return self.loss


class GraphedPrediction:
    """CUDA-graph replay of one posterior-predictive sample: new GP function draw (``build_cache``) + one integration
    of ``x0_fn()`` over ``ts`` under ``torch.no_grad()`` -- the body of the reference's ``compute_predictions`` loop
    (``src/gpode/model_builder.py:60-78``), which rebuilds the cache for each of its 128 samples.

        pred = GraphedPrediction(model, ts)            # x0 ~ q(x0) each sample
        samples = pred.sample_many(128)                # (128, N, T, D)
    """

    def __init__(self, model, ts, x0_fn=None, warmup=2):
        from .misc.torch_utils import insert_zero_t0
        dev = next(model.parameters()).device
        self.model = model
        dist = model.x0_distribution if hasattr(model, "x0_distribution") else model.state_distribution.x0
        self.x0_fn = x0_fn if x0_fn is not None else (lambda: dist.sample().squeeze(0))
        self.ts0 = insert_zero_t0(ts)
        self.draws = _StaticDraws(dev)
        self.draws.install()
        try:
            with torch.no_grad():
                side = torch.cuda.Stream()
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    for i in range(warmup):
                        if i == 1:
                            self.draws.recording = False
                        self.draws.cursor = 0
                        model(self.x0_fn(), self.ts0)
                torch.cuda.current_stream().wait_stream(side)
                self.draws.recording = False
                self.draws.cursor = 0
                self.graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self.graph):
                    self.out = model(self.x0_fn(), self.ts0)[:, 1:]
        finally:
            self.draws.uninstall()

    def sample(self):
        """One predictive trajectory set ``(N, T, D)`` (a view of the static output buffer: clone to keep it)."""
        self.draws.refresh()
        self.graph.replay()
        return self.out

    def sample_many(self, n):
        return torch.stack([self.sample().clone() for _ in range(n)], 0)
