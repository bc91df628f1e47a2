"""Host->device input staging for the step loops (reference train.py:86 / predict.py:171 do a blocking `.cuda()`
per step): the next batch's pinned-host tensors are copied on a side stream while the current step computes."""
import torch

from . import _lib


def label_discs(centers, height=288, width=512, sigma=2.5, out=None):
    """The training labels of the reference's dataset (dataset.py:400-410 `_get_heatmap`, called with integer centres at
    :632) built ON the device from the coordinates: ``centers`` int (N, L, 2) = (cx, cy) per frame -> float32
    (N, L, height, width) binary discs of radius ``sigma``, all-zero maps where cx == cy == 0. 8 bytes per map cross
    PCIe instead of a 590 KB fp32 heatmap. ``out``: optional preallocated float32 (N, L, height, width) CUDA tensor."""
    lib = _lib.load()
    _lib.require_cuda(centers)
    if centers.dim() != 3 or centers.shape[2] != 2:
        raise RuntimeError(f"label_discs expects integer centres (N, L, 2), got {tuple(centers.shape)}")
    c = centers.to(torch.int32).contiguous()
    n, l = c.shape[0], c.shape[1]
    if out is None:
        out = torch.empty((n, l, height, width), dtype=torch.float32, device=c.device)
    elif tuple(out.shape) != (n, l, height, width) or out.dtype != torch.float32 or not out.is_contiguous() or out.device != c.device:
        raise RuntimeError(f"label_discs: out must be a contiguous float32 {(n, l, height, width)} tensor on {c.device}")
    _lib.check(lib.tnb_label_discs(c.data_ptr(), n * l, height, width, float(sigma), out.data_ptr(), _lib.stream_ptr()))
    return out


class DevicePrefetcher:
    """Iterate over an iterable of tuples of (pinned) host tensors, yielding the same tuples on the GPU.

    The copy of batch k+1 is enqueued on a dedicated CUDA stream as soon as batch k is handed out, so it overlaps
    with the compute of batch k; the consumer's stream waits on the copy's event, no host synchronisation.
    Two fixed sets of device buffers are cycled (no allocator traffic in the loop): a yielded batch is valid until
    the next-but-one ``next()`` call, which is what a train / predict loop needs. Only COPIES run on the side stream:
    running the staging kernels (FramePreprocessor, label discs) there too, next to the compute of the previous batch, was
    measured and dropped - small grids interleaved with persistent one-CTA-per-SM kernels made the step time erratic
    (22.8 ... 37 ms per step over four runs, profiles/r2_summary.md).
    """

    def __init__(self, host_batches, device=None):
        self.it = iter(host_batches)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else device
        self.stream = torch.cuda.Stream(device=self.device)
        self.slots = [None, None]     # device buffers of each slot
        self.release = [None, None]   # event on the consumer stream after which the slot may be overwritten
        self.count = 0
        self.next = None
        self.event = None
        self._preload()

    def _buffers(self, slot, batch):
        bufs = self.slots[slot]
        ok = bufs is not None and len(bufs) == len(batch) and all(
            (not torch.is_tensor(t)) or (b.shape == t.shape and b.dtype == t.dtype) for b, t in zip(bufs, batch))
        if not ok:
            cur = torch.cuda.current_stream(self.device)
            bufs = tuple(torch.empty(t.shape, dtype=t.dtype, device=self.device) if torch.is_tensor(t) else t
                         for t in batch)
            for b in bufs:
                if torch.is_tensor(b):
                    b.record_stream(self.stream)  # freed memory must also wait for the copy stream
            ev = torch.cuda.Event()
            ev.record(cur)                        # allocation-order safety: copy only after `cur` got here
            self.release[slot] = ev
            self.slots[slot] = bufs
        return bufs

    def _preload(self):
        try:
            batch = next(self.it)
        except StopIteration:
            self.next = None
            return
        slot = self.count & 1
        self.count += 1
        bufs = self._buffers(slot, batch)
        with torch.cuda.stream(self.stream):
            if self.release[slot] is not None:
                self.stream.wait_event(self.release[slot])
            for b, t in zip(bufs, batch):
                if torch.is_tensor(t):
                    b.copy_(t, non_blocking=True)
            self.event = torch.cuda.Event()
            self.event.record(self.stream)
        self.next = (slot, bufs)

    def __iter__(self):
        return self

    def __next__(self):
        if self.next is None:
            raise StopIteration
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(self.event)
        slot, batch = self.next
        # everything that used the other slot has been enqueued on `cur` by now: it may be refilled after this point
        ev = torch.cuda.Event()
        ev.record(cur)
        self.release[1 - slot] = ev
        self._preload()
        return batch


class ScalarReader:
    """Read a device scalar (the step's loss) on the host WITHOUT draining the compute stream.

    ``loss.item()`` (reference train.py:94) is a cudaMemcpy + synchronize of the current stream: placed after
    ``loss.backward()`` it waits for the whole backward pass, and the GPU then idles while the host enqueues the next step
    (0.3 ms of a 23 ms step, ``tools/diag_e2e.py``). ``read(t)`` records an event where ``t`` is ready, copies it to a
    pinned host slot on a side stream behind that event and returns at once; ``value()`` waits for that copy only -
    everything enqueued on the compute stream after ``read`` (the backward pass) keeps running.

        loss = WBCELoss(model(x), y); reader.read(loss); loss.backward(); v = reader.value()
    """

    def __init__(self, device=None):
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else device
        self.stream = torch.cuda.Stream(device=self.device)
        self.host = torch.zeros(1, dtype=torch.float32).pin_memory()
        self.ready = torch.cuda.Event()
        self.done = torch.cuda.Event()
        self.pending = False

    def read(self, t):
        _lib.require_cuda(t)
        if t.numel() != 1:
            raise RuntimeError(f"ScalarReader.read expects a one-element tensor, got shape {tuple(t.shape)}")
        if self.pending:
            raise RuntimeError("ScalarReader.read() while the previous value has not been taken (one host slot: call value())")
        src = t.detach().reshape(1).float()
        self.ready.record(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(self.ready)
            self.host.copy_(src, non_blocking=True)
            self.done.record(self.stream)
        src.record_stream(self.stream)
        self.pending = True

    def value(self):
        if not self.pending:
            raise RuntimeError("ScalarReader.value() without a preceding read()")
        self.done.synchronize()
        self.pending = False
        return float(self.host[0])
