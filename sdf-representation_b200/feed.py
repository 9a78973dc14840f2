"""Device-resident data feed.

The reference feeds the training loop from a host ``DataLoader`` (``dataloader/load_data.py:74-82``: ``TensorDataset`` of
CPU tensors, ``shuffle=True``; ``executor/executordistributed.py:37``: ``num_workers=30``) and copies every batch with
``x_batch.to(device)`` on the compute stream (``executor/executor.py:174``).  At >= 30 M points/s that host path is the
bottleneck.  Two replacements with the same iteration protocol (``for x_batch, y_batch in feed``):

* ``DeviceFeed``: X / Y live in HBM; each epoch draws ``torch.randperm`` on the device and yields row-gathered batches --
  no host work and no H2D copy per step (the reference's seeded CPU shuffle order is available with ``generator=``).
* ``PinnedPrefetcher``: batches stay in pinned host memory (data sets larger than HBM); a copy stream moves batch i+1
  while batch i computes, so the per-step H2D copy is still paid every step but overlaps the kernels.
"""
import torch


class DeviceFeed:
    """Shuffled mini-batches of a device-resident (X [n,3], Y [n,4]) data set (``DataLoader(shuffle=True)`` semantics:
    a fresh permutation per epoch, ragged last batch kept unless ``drop_last``)."""

    def __init__(self, X, Y, batch_size, shuffle=True, drop_last=False, device=None, generator=None):
        device = torch.device(device) if device is not None else (X.device if X.is_cuda else torch.device("cuda"))
        if device.type != "cuda":
            raise RuntimeError("DeviceFeed keeps the data set in HBM: it needs a CUDA device (no CPU fallback)")
        self.X = X.to(device=device, dtype=torch.float32).contiguous()
        self.Y = None if Y is None else Y.to(device=device, dtype=torch.float32).contiguous()
        if self.Y is not None and self.Y.shape[0] != self.X.shape[0]:
            raise ValueError("X and Y must have the same number of rows")
        self.batch_size, self.shuffle, self.drop_last = int(batch_size), bool(shuffle), bool(drop_last)
        self.generator = generator  # a CPU torch.Generator reproduces the reference's host-side shuffle order

    def __len__(self):
        n = self.X.shape[0]
        return n // self.batch_size if self.drop_last else (n + self.batch_size - 1) // self.batch_size

    def __iter__(self):
        n = self.X.shape[0]
        if self.shuffle:
            if self.generator is not None and self.generator.device.type == "cpu":
                perm = torch.randperm(n, generator=self.generator).to(self.X.device, non_blocking=True)
            else:
                perm = torch.randperm(n, device=self.X.device, generator=self.generator)
        for i in range(len(self)):
            lo, hi = i * self.batch_size, min((i + 1) * self.batch_size, n)
            if self.shuffle:
                idx = perm[lo:hi]
                yield (self.X.index_select(0, idx), None if self.Y is None else self.Y.index_select(0, idx))
            else:
                yield (self.X[lo:hi], None if self.Y is None else self.Y[lo:hi])


class PinnedPrefetcher:
    """Iterate over host batches (tuples of pinned tensors) with the H2D copy of batch i+1 overlapped with the compute
    of batch i on a side stream.  ``depth`` device buffers are recycled (no allocation per step)."""

    def __init__(self, host_batches, device, depth=2):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("PinnedPrefetcher copies to a CUDA device (no CPU fallback)")
        self.host = host_batches
        self.depth = max(2, int(depth))
        self.stream = torch.cuda.Stream(self.device)
        self._slots = [None] * self.depth
        self._free = [None] * self.depth  # event: the consumer is done with the slot

    def __len__(self):
        return len(self.host)

    def _issue(self, i, slot):
        batch = self.host[i]
        if self._slots[slot] is None or any(d.shape != h.shape for d, h in zip(self._slots[slot], batch)):
            self._slots[slot] = tuple(torch.empty(h.shape, dtype=h.dtype, device=self.device) for h in batch)
        with torch.cuda.stream(self.stream):
            if self._free[slot] is not None:
                self.stream.wait_event(self._free[slot])
            for d, h in zip(self._slots[slot], batch):
                d.copy_(h, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        return ev

    def __iter__(self):
        n = len(self.host)
        if n == 0:
            return
        ready = [None] * self.depth
        for j in range(min(self.depth - 1, n)):
            ready[j % self.depth] = self._issue(j, j % self.depth)
        cur = torch.cuda.current_stream(self.device)
        for i in range(n):
            slot = i % self.depth
            nxt = i + self.depth - 1
            if nxt < n:
                ready[nxt % self.depth] = self._issue(nxt, nxt % self.depth)
            cur.wait_event(ready[slot])
            yield self._slots[slot]
            done = torch.cuda.Event()
            done.record(cur)
            self._free[slot] = done
