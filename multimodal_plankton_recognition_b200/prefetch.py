"""Host -> HBM staging for the loss path: a prefetcher for (image_emb, profile_emb) pairs.

The reference trains through a Lightning loop whose DataLoader hands host batches to the device one
step at a time (reference scripts/train_multi.py:78-85, src/model.py `training_step`); the copy of
step i+1 does not overlap the compute of step i.  At B=4096, d=256 the fused loss step is ~85 us of
GPU time while the 8 MiB of fp32 embeddings take ~170 us over PCIe, so the serial order leaves the
GPU idle two thirds of the time.  `HostPairPrefetcher` keeps `depth` device slots filled by the native
stager (`plk_stager_*` in libplk.so: a copy stream plus per-slot events), so the copies of the next
batches run under the current step; `read_async` brings a scalar result back through a pinned slot
and an event of its own, so reading a loss never drains the queue.
"""
from __future__ import annotations

from typing import Callable, Iterable, Iterator, Tuple

import torch

from . import _lib


class HostPairPrefetcher:
    """Iterate device copies of host (x, y) pairs, `depth` batches ahead of the consumer.

    The yielded tensors are views of a ring of device slots: a slot is rewritten once the consumer
    has asked for `depth` further batches, and the rewrite is ordered (by an event recorded on the
    consumer's stream at that moment) after everything the consumer had queued by then.  Use the
    tensors within the step they were yielded for -- the usual contract of a training loop.
    Host batches should be pinned (`tensor.pin_memory()`); pageable ones are pinned on the fly,
    which costs a host copy.
    """

    def __init__(self, batches: Iterable[Tuple[torch.Tensor, torch.Tensor]], device, depth: int = 2) -> None:
        if depth < 2:
            raise ValueError("depth must be >= 2 (one slot in use, one in flight)")
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise ValueError("HostPairPrefetcher stages into CUDA memory")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self._lib = _lib.load()
        self._src = iter(batches)
        self._depth = depth
        self._n_read = 2 * depth
        with torch.cuda.device(self.device):
            self._h = self._lib.plk_stager_create(depth, self._n_read)
        if not self._h:
            raise _lib.PlkError(f"plk_stager_create: {self._lib.plk_last_error().decode()}")
        self._slots = [None] * depth            # (dx, dy) device buffers
        self._hold = [None] * depth             # host tensors kept alive while their copy is in flight
        self._issued = 0
        self._taken = 0
        self._exhausted = False
        self._host = torch.zeros(self._n_read, dtype=torch.float32).pin_memory()
        self._host_ptr = self._host.data_ptr()
        self._reads = 0

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                torch.cuda.synchronize(self.device)
                self._lib.plk_stager_destroy(h)
            except Exception:
                pass

    # -- host -> device -------------------------------------------------------------------------
    def _issue(self) -> bool:
        if self._exhausted:
            return False
        try:
            hx, hy = next(self._src)
        except StopIteration:
            self._exhausted = True
            return False
        if hx.is_cuda or hy.is_cuda:
            raise ValueError("HostPairPrefetcher expects host tensors")
        if not hx.is_contiguous() or not hy.is_contiguous():
            hx, hy = hx.contiguous(), hy.contiguous()
        if not (hx.is_pinned() and hy.is_pinned()):
            hx, hy = hx.pin_memory(), hy.pin_memory()
        s = self._issued % self._depth
        slot = self._slots[s]
        if slot is None or slot[0].shape != hx.shape or slot[1].shape != hy.shape or \
                slot[0].dtype != hx.dtype or slot[1].dtype != hy.dtype:
            if slot is not None:                 # the old buffers may still be in use by queued work
                torch.cuda.synchronize(self.device)
            slot = (torch.empty(hx.shape, device=self.device, dtype=hx.dtype),
                    torch.empty(hy.shape, device=self.device, dtype=hy.dtype))
            self._slots[s] = slot
        # plk_stager_issue first waits (on the host) for the slot's previous copy: only then may the host
        # tensors of that copy be dropped -- a tensor pinned on the fly or a DataLoader pin_memory batch goes
        # back to the pinned pool when its last reference dies and could be rewritten under a queued DMA
        self._lib.check(self._lib.plk_stager_issue(self._h, s, slot[0].data_ptr(), hx.data_ptr(),
                                                   hx.numel() * hx.element_size(), slot[1].data_ptr(),
                                                   hy.data_ptr(), hy.numel() * hy.element_size()),
                        "plk_stager_issue")
        self._hold[s] = (hx, hy)
        self._issued += 1
        return True

    def __iter__(self) -> Iterator[Tuple[torch.Tensor, torch.Tensor]]:
        return self

    def __next__(self) -> Tuple[torch.Tensor, torch.Tensor]:
        if self._taken >= self._issued and not self._issue():
            raise StopIteration
        s = self._taken % self._depth
        release = (self._taken - 1) % self._depth if self._taken > 0 else -1
        stream = torch._C._cuda_getCurrentRawStream(self.device.index)
        # the previous batch's slot may be rewritten after what the consumer has queued so far
        self._lib.check(self._lib.plk_stager_acquire(self._h, s, release, stream), "plk_stager_acquire")
        self._taken += 1
        while self._issued < self._taken - 1 + self._depth and self._issue():
            pass
        x, y = self._slots[s]
        return x.detach(), y.detach()

    # -- device -> host -------------------------------------------------------------------------
    def read_async(self, value: torch.Tensor) -> Callable[[], float]:
        """Queue a device->host copy of an fp32 scalar on the current stream; the returned callable
        waits for that copy alone and returns the number.  At most 2*depth reads may be pending."""
        v = value.detach()
        if v.dtype is not torch.float32:
            v = v.float()
        k = self._reads % self._n_read
        self._reads += 1
        stream = torch._C._cuda_getCurrentRawStream(self.device.index)
        self._lib.check(self._lib.plk_stager_read_async(self._h, k, self._host_ptr + 4 * k, v.data_ptr(), 4,
                                                        stream), "plk_stager_read_async")
        lib, h, host = self._lib, self._h, self._host

        def wait() -> float:
            lib.check(lib.plk_stager_read_wait(h, k), "plk_stager_read_wait")
            return host[k].item()

        return wait
