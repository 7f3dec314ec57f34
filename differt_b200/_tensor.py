"""Host-side plumbing shared by the public functions: device placement, broadcast → strided batch.

PyTorch is used for device memory, streams and autograd bookkeeping only; every arithmetic result
comes from the CUDA library.
"""

from __future__ import annotations

import ctypes as C
from typing import Sequence

import numpy as np
import torch

from ._lib import DRT_MAX_BATCH_DIMS, begin_call, i64_array, note_device

F32_EPS = float(np.finfo(np.float32).eps)


def require_cuda() -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError(
            "differt_b200 needs a CUDA device (sm_100a); there is no CPU fallback by design"
        )
    return torch.device("cuda", torch.cuda.current_device())


def stream_ptr():
    """First argument of every stream-ordered library call: the current stream of the device the
    call's operands live on (resolved by ``_lib._Library`` once ``ptr`` has seen them)."""
    return begin_call()


def ptr(t: torch.Tensor | None) -> C.c_void_p:
    if t is None or t.numel() == 0:
        return C.c_void_p(0)
    if t.is_cuda:
        note_device(t.device.index)
    return C.c_void_p(t.data_ptr())


class Placement:
    """Remembers whether the caller passed host arrays, so results can be handed back there."""

    def __init__(self) -> None:
        self.device: torch.device | None = None
        self.saw_cuda = False

    def put(self, x, dtype: torch.dtype) -> torch.Tensor:
        if isinstance(x, torch.Tensor):
            if x.is_cuda:
                self.saw_cuda = True
                if self.device is None:
                    self.device = x.device
                return x if x.dtype == dtype else x.to(dtype)
            t = x
        else:
            t = torch.as_tensor(np.asarray(x))
        if self.device is None:
            self.device = require_cuda()
        if t.dtype != dtype:
            t = t.to(dtype)
        if t.numel() > (1 << 16) and not t.is_pinned():
            # large host inputs: go through pinned memory so the copy is a real async DMA
            t = t.contiguous().pin_memory()
        return t.to(self.device, non_blocking=True)

    def out(self, t: torch.Tensor) -> torch.Tensor:
        return t if self.saw_cuda else t.cpu()


def batch_strides(
    batch: Sequence[int], operands: Sequence[tuple[torch.Tensor, int]]
) -> tuple[int, C.Array, list[C.Array], list[torch.Tensor]]:
    """Describe ``operands`` broadcast over ``batch`` as ≤4 strided batch dims.

    ``operands`` = [(tensor, core_ndim)]: the trailing ``core_ndim`` dims are the per-element payload
    and are made contiguous; leading dims broadcast against ``batch``.  Returns
    ``(ndim, shape, [strides per operand], [tensors to keep alive])``; strides are in elements.
    """
    batch = [int(b) for b in batch]
    views = []
    for t, core in operands:
        core_shape = tuple(t.shape[t.ndim - core:])
        t = t.contiguous()
        views.append(t.expand(*batch, *core_shape) if batch else t)
    nb = len(batch)
    strides = [[int(v.stride(i)) for i in range(nb)] for v in views]
    dims = [(batch[i], [s[i] for s in strides]) for i in range(nb) if batch[i] != 1]
    merged: list[tuple[int, list[int]]] = []
    for size, st in dims:
        if merged:
            psize, pst = merged[-1]
            if all(pst[j] == st[j] * size for j in range(len(views))):
                merged[-1] = (psize * size, st)
                continue
        merged.append((size, st))
    if len(merged) > DRT_MAX_BATCH_DIMS:
        views = [v.contiguous() for v in views]
        n = 1
        for b in batch:
            n *= b
        merged = [(n, [int(np.prod(v.shape[nb:], dtype=np.int64)) for v in views])]
    shape = i64_array([m[0] for m in merged])
    per_op = [i64_array([m[1][j] for m in merged]) for j in range(len(views))]
    return len(merged), shape, per_op, views


def numel(shape: Sequence[int]) -> int:
    n = 1
    for s in shape:
        n *= int(s)
    return n


def sum_to_shape(g: torch.Tensor, shape: Sequence[int]) -> torch.Tensor:
    """Reduce a broadcast gradient back to the shape of the operand it belongs to."""
    shape = tuple(shape)
    if tuple(g.shape) == shape:
        return g
    return g.sum_to_size(shape) if len(shape) > 0 else g.sum()
