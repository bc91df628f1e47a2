"""Host->device input staging for the step loops (reference train.py:86 / predict.py:171 do a blocking `.cuda()`
per step): the next batch's pinned-host tensors are copied on a side stream while the current step computes."""
import torch


class DevicePrefetcher:
    """Iterate over an iterable of tuples of (pinned) host tensors, yielding the same tuples on the GPU.

    The copy of batch k+1 is enqueued on a dedicated CUDA stream as soon as batch k is handed out, so it overlaps
    with the compute of batch k; the consumer's stream waits on the copy's event, no host synchronisation.
    """

    def __init__(self, host_batches, device=None):
        self.it = iter(host_batches)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else device
        self.stream = torch.cuda.Stream(device=self.device)
        self.next = None
        self.event = None
        self._preload()

    def _preload(self):
        try:
            batch = next(self.it)
        except StopIteration:
            self.next = None
            return
        with torch.cuda.stream(self.stream):
            self.next = tuple(t.to(self.device, non_blocking=True) if torch.is_tensor(t) else t for t in batch)
            self.event = torch.cuda.Event()
            self.event.record(self.stream)

    def __iter__(self):
        return self

    def __next__(self):
        if self.next is None:
            raise StopIteration
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(self.event)
        batch = self.next
        for t in batch:
            if torch.is_tensor(t):
                t.record_stream(cur)  # the caching allocator must not recycle it while `cur` still uses it
        self._preload()
        return batch
