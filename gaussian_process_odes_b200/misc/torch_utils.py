"""Step-grid helpers and small utilities (mirror of reference ``src/misc/torch_utils.py``).

``insert_zero_t0`` / ``compute_ts_dense`` define the RK4 step grid and must match the reference to the bit, so they
run the very same float32 CPU ops (``torch.linspace`` per interval); the result is cached per time tensor so that a
training loop pays the (tiny) host work and the device->host read of ``ts`` once, not every ELBO evaluation."""
import os
import random
import weakref

import numpy as np
import torch

from .settings import settings

dtype = settings.torch_float


def numpy2torch(x):
    dev = settings.device
    return torch.tensor(x, dtype=dtype).to(dev) if type(x) is np.ndarray else x.to(dev)


def torch2numpy(x):
    return x if type(x) is np.ndarray else x.detach().cpu().numpy()


def restore_model(model, filename):
    checkpt = torch.load(filename, map_location=lambda storage, loc: storage)
    model.load_state_dict(checkpt['state_dict'])
    return model


def save_model(model, filename):
    torch.save({'state_dict': model.state_dict()}, filename)


def save_model_optimizer(model, optimizer, filename):
    torch.save({'state_dict': model.state_dict(), 'optimizer_state_dict': optimizer.state_dict()}, filename)


_GRID_CACHE = {}


def _cached(kind, ts, extra, build):
    # valid only while the very same tensor object is alive and unmodified (weakref + version counter)
    key = (kind, id(ts), extra)
    hit = _GRID_CACHE.get(key)
    if hit is not None and hit[0]() is ts and hit[1] == ts._version:
        return hit[2]
    if len(_GRID_CACHE) > 64:
        _GRID_CACHE.clear()
    out = build(ts.detach().to("cpu", torch.float32)).to(ts.device)
    _GRID_CACHE[key] = (weakref.ref(ts), ts._version, out)
    return out


def insert_zero_t0(ts):
    """Prepend t=0 and shift the rest by one sampling interval (reference ``torch_utils.py:36-38``)."""
    return _cached("zero", ts, None, lambda t: torch.cat([torch.tensor([0.0]), t + t[1] - t[0]]))


def compute_ts_dense(ts, ts_dense_scale):
    """(ts_dense_scale - 1) equal sub-steps per interval (reference ``torch_utils.py:41-48``)."""
    if ts_dense_scale <= 1:
        return ts

    def build(t):
        return torch.cat([torch.linspace(t1, t2, ts_dense_scale)[:-1] for (t1, t2) in zip(t[:-1], t[1:])] + [t[-1:]])

    return _cached("dense", ts, int(ts_dense_scale), build)


def seed_everything(seed):
    random.seed(seed)
    os.environ['PYTHONHASHSEED'] = str(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed(seed)


class _PinnedRing:
    """Pinned staging buffers for ``host_to_device``: per (device, shape, dtype) a small ring of page-locked tensors,
    each with the CUDA event of its last copy. Re-using a slot waits for that event (normally long complete: a slot
    comes round again ``depth`` draws of the same shape later), so the host can run at most ``depth`` steps ahead of
    the stream. Costs a host memcpy, one async copy and one event record per draw -- ``Tensor.pin_memory()`` per draw
    was measured at ~0.1 ms each on the launch-bound training shapes."""

    depth = 4

    def __init__(self):
        self.slots = {}

    def stage(self, t, device):
        key = (device, tuple(t.shape), t.dtype)
        ring = self.slots.get(key)
        if ring is None:
            ring = self.slots[key] = {"i": 0, "buf": [torch.empty(t.shape, dtype=t.dtype).pin_memory()
                                                      for _ in range(self.depth)],
                                      "ev": [None] * self.depth}
        i = ring["i"]
        ring["i"] = (i + 1) % self.depth
        if ring["ev"][i] is not None:
            ring["ev"][i].synchronize()
        buf = ring["buf"][i]
        buf.copy_(t)
        with torch.cuda.device(device):
            out = torch.empty(t.shape, dtype=t.dtype, device=device)
            out.copy_(buf, non_blocking=True)
            ev = ring["ev"][i] or torch.cuda.Event()
            ev.record()
            ring["ev"][i] = ev
        return out


_pinned_ring = _PinnedRing()


def host_to_device(t, device):
    """Move a small host-side random draw to ``device`` WITHOUT synchronising the host with the stream: a pageable
    ``tensor.to('cuda')`` blocks the caller until every kernel already queued has drained (once per draw, i.e. four
    times per ``build_cache``), which leaves the GPU idle while the next step's launches are being issued. The draw is
    staged in a pinned ring buffer and copied asynchronously. Tensors already on ``device`` (CUDA-graph static
    buffers) pass through; ``GPODE_SYNC_DRAWS=1`` restores the blocking copy."""
    device = torch.device(device)
    if t.device.type != 'cpu' or device.type != 'cuda' or os.environ.get('GPODE_SYNC_DRAWS'):
        return t.to(device)
    if torch.cuda.is_current_stream_capturing():
        return t.to(device)  # never reached by GraphedStep (its samplers return device buffers); fail loudly if it is
    return _pinned_ring.stage(t, device)
