"""Input pipeline and device-side metric logging for the training steps.

What the reference does per iteration (src/data.py:65-81, src/utils.py:55-57,259-262, src/logger.py:34-50,
src/train_gan.py:134-137,177-181):

    DataLoader(num_workers=0, collate_fn=input_collate)   list of sample dicts -> dict of LISTS   (main thread)
    collate(input)                                        torch.stack per key                      (main thread)
    to_device(input, device)                              synchronous pageable H2D copies
    ... step ...
    metric.evaluate(...) -> .item() x 3                   three host synchronisations per iteration
    logger.append(evaluation, 'train', n)                 running means on the host

At the B200 step rate (>= 17 k img/s per GPU) that serial chain starves the device: the loader, the stack
and the copy all sit on the critical path, and the ``.item()`` calls drain the launch queue every iteration.
Same semantics, restructured:

    DeviceFeeder    a producer thread pulls from the loader, stacks each key straight INTO pinned staging
                    buffers (ring of `depth`), and issues the H2D copies on a dedicated copy stream; the
                    training stream only waits on the copy's event.  Batches arrive as device tensors while
                    the previous step is still running.
    DeviceMeter     ``logger.append`` without the host round trip: the weighted sums live in a device tensor
                    (one tiny add per step), and are read back — one asynchronous copy — every ``flush_every``
                    steps or when ``mean`` is asked for.  The means it reports equal Logger.append's.

``input_collate`` / ``collate`` / ``to_device`` keep the reference's names and meaning so that a training script
written against the reference reads the same.
"""
import queue
import threading

import torch


def input_collate(batch):
    """list of sample dicts -> dict of lists (src/data.py:65-73); anything else goes to torch's default collate."""
    if isinstance(batch[0], dict):
        output = {key: [] for key in batch[0].keys()}
        for b in batch:
            for key in b:
                output[key].append(b[key])
        return output
    return torch.utils.data.default_collate(batch)


def collate(input):
    """dict of lists -> dict of stacked tensors (src/utils.py:259-262)."""
    for k in input:
        input[k] = torch.stack(input[k], 0)
    return input


def to_device(input, device):
    """src/utils.py:55-57 for the dict / list / tensor nests the loaders produce."""
    if isinstance(input, dict):
        return {k: to_device(v, device) for k, v in input.items()}
    if isinstance(input, (list, tuple)):
        return type(input)(to_device(v, device) for v in input)
    return input.to(device)


class SyntheticImages(torch.utils.data.Dataset):
    """A host-resident uint8 image set with labels, served like the reference's datasets after their transform
    (ToTensor + Normalize(0.5, 0.5): float32 CHW in [-1, 1]) — ``{'img', 'label'}`` per sample."""

    def __init__(self, n, data_shape, classes, seed=0):
        g = torch.Generator().manual_seed(seed)
        self.img = torch.randint(0, 256, (n, *data_shape), dtype=torch.uint8, generator=g)
        self.label = torch.randint(0, classes, (n,), generator=g)

    def __len__(self):
        return self.img.size(0)

    def __getitem__(self, i):
        return {'img': self.img[i].float().div_(127.5).sub_(1.0), 'label': self.label[i]}

    def batch(self, idx):
        """Vectorised equivalent of ``collate(input_collate([self[i] for i in idx]))``."""
        return {'img': self.img[idx].float().div_(127.5).sub_(1.0), 'label': self.label[idx]}


class DeviceFeeder:
    """Iterate a host loader as device batches, copies overlapped with compute.

    loader: iterable of batches — dicts of tensors (already collated) or dicts of lists (``input_collate``'s
    output; stacked here, directly into pinned memory).  depth: staging ring size (2 = double buffering).
    Each yielded dict holds device tensors valid until `depth - 1` further batches have been requested; the
    current stream has already been made to wait for their copy.
    """

    _END = object()

    def __init__(self, loader, device, depth=2):
        if depth < 2:
            raise ValueError('Not valid depth (>= 2)')
        self.loader, self.device, self.depth = loader, torch.device(device), depth
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self.host = [None] * depth           # pinned staging dicts
        self.dev = [None] * depth            # device dicts
        self.copied = [None] * depth         # event: H2D of slot i finished
        self.consumed = [None] * depth       # event: the step that used slot i was enqueued-and-passed
        self.h2d_bytes = 0

    def _stage(self, slot, batch):
        """Stack / copy `batch` into the pinned buffers of `slot` (allocated on first use or shape change)."""
        out = self.host[slot] or {}
        for k, v in batch.items():
            if isinstance(v, (list, tuple)):
                shape, dtype = (len(v), *v[0].shape), v[0].dtype
            else:
                shape, dtype = tuple(v.shape), v.dtype
            buf = out.get(k)
            if buf is None or tuple(buf.shape) != shape or buf.dtype != dtype:
                buf = torch.empty(shape, dtype=dtype).pin_memory()
                out[k] = buf
            if isinstance(v, (list, tuple)):
                torch.stack(list(v), 0, out=buf)
            else:
                buf.copy_(v)
        self.host[slot] = out
        return out

    def _producer(self, q, free):
        try:
            for batch in self.loader:
                slot = free.get()
                if slot is None:
                    return
                q.put((slot, self._stage(slot, batch)))
            q.put(self._END)
        except BaseException as e:          # surface loader errors in the consumer
            q.put(e)

    def iterate(self, loader):
        """Iterate another loader through the same staging ring (pinned and device buffers are kept: no
        cudaHostAlloc / cudaMalloc — both synchronise the device — after the first pass)."""
        self.loader = loader
        return iter(self)

    def __iter__(self):
        q, free = queue.Queue(maxsize=self.depth), queue.Queue()
        for s in range(self.depth):
            free.put(s)
        worker = threading.Thread(target=self._producer, args=(q, free), daemon=True)
        worker.start()
        in_flight = []
        try:
            while True:
                item = q.get()
                if item is self._END:
                    break
                if isinstance(item, BaseException):
                    raise item
                slot, host = item
                cur = torch.cuda.current_stream(self.device)
                with torch.cuda.stream(self.copy_stream):
                    if self.consumed[slot] is not None:
                        self.copy_stream.wait_event(self.consumed[slot])     # the step that read this slot is done
                    dev = self.dev[slot] or {}
                    nbytes = 0
                    for k, h in host.items():
                        d = dev.get(k)
                        if d is None or d.shape != h.shape or d.dtype != h.dtype:
                            d = torch.empty(h.shape, dtype=h.dtype, device=self.device)
                            dev[k] = d
                        d.copy_(h, non_blocking=True)
                        nbytes += h.numel() * h.element_size()
                    self.dev[slot] = dev
                    ev = torch.cuda.Event()
                    ev.record(self.copy_stream)
                    self.copied[slot] = ev
                self.h2d_bytes = nbytes
                cur.wait_event(ev)
                yield dev
                done = torch.cuda.Event()
                done.record(cur)                   # everything enqueued for this batch so far
                self.consumed[slot] = done
                in_flight.append((slot, ev))
                # a staging slot may be overwritten by the producer once its H2D copy has completed
                while len(in_flight) >= self.depth:
                    s, e = in_flight.pop(0)
                    e.synchronize()
                    free.put(s)
        finally:
            free.put(None)
            while in_flight:
                s, e = in_flight.pop(0)
                e.synchronize()


class DeviceMeter:
    """Running means of per-step scalars without per-step host synchronisation (replaces the
    ``.item()`` + ``Logger.append`` pair, src/train_gan.py:177-181 + src/logger.py:34-50).

    ``append(result, tag, n)``: result maps names to 0-d device tensors (or numbers); the weighted sums
    ``sum(n * value)`` and counts accumulate on the device.  ``mean`` triggers (or, after ``flush_every`` steps,
    picks up) one small device->host copy and returns ``{f'{tag}/{name}': mean}`` — the values
    ``Logger.mean`` would hold after the same appends.
    """

    def __init__(self, device, flush_every=50, capacity=32):
        self.device = torch.device(device)
        self.flush_every = flush_every
        self.names = {}
        self.acc = torch.zeros(capacity, dtype=torch.float64, device=self.device)
        self.count = [0.0] * capacity
        self.host = torch.zeros(capacity, dtype=torch.float64).pin_memory() if self.device.type == 'cuda' \
            else torch.zeros(capacity, dtype=torch.float64)
        self.host_count = [0.0] * capacity
        self.flushed = None
        self.steps = 0
        self.d2h_bytes_per_flush = capacity * 8

    def _index(self, name):
        i = self.names.get(name)
        if i is None:
            i = len(self.names)
            if i >= self.acc.numel():
                raise ValueError('Not valid metric count: raise DeviceMeter(capacity=...)')
            self.names[name] = i
        return i

    def append(self, result, tag, n=1):
        idx, vals = [], []
        for k, v in result.items():
            if v is None:
                continue
            i = self._index(f'{tag}/{k}')
            self.count[i] += n
            idx.append(i)
            vals.append(v.detach().to(torch.float64).reshape(()) if torch.is_tensor(v)
                        else torch.tensor(float(v), dtype=torch.float64, device=self.device))
        if idx:
            contiguous = idx == list(range(idx[0], idx[0] + len(idx)))
            stacked = torch.stack(vals) * float(n)
            if contiguous:
                self.acc[idx[0]:idx[0] + len(idx)].add_(stacked)
            else:
                self.acc.index_add_(0, torch.tensor(idx, device=self.device), stacked)
        self.steps += 1
        if self.flush_every and self.steps % self.flush_every == 0:
            self.flush()

    def flush(self):
        """Start one asynchronous device->host copy of the sums (no synchronisation)."""
        self.host.copy_(self.acc, non_blocking=True)
        self.host_count = list(self.count)
        if self.device.type == 'cuda':
            self.flushed = torch.cuda.Event()
            self.flushed.record(torch.cuda.current_stream(self.device))

    @property
    def mean(self):
        """Means over everything appended so far (synchronises on one small copy)."""
        self.flush()
        if self.flushed is not None:
            self.flushed.synchronize()
        return {name: float(self.host[i]) / self.host_count[i] for name, i in self.names.items() if self.host_count[i]}

    def reset(self):
        self.acc.zero_()
        self.count = [0.0] * len(self.count)
        self.steps = 0
