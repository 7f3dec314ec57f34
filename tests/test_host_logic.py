"""Host-side plumbing that needs no GPU: the broadcast → strided-batch descriptors handed to the
element-wise kernels, gradient reduction over broadcast axes, API-level argument checks."""

from __future__ import annotations

import itertools

import numpy as np
import pytest
import torch

from differt_b200 import _tensor


def _emulate(shape, strides, flat_len):
    """Offsets the kernels compute (common.cuh Batch4::offsets): row-major over `shape`."""
    shape, strides = list(shape), list(strides)
    offs = []
    for idx in itertools.product(*[range(n) for n in shape]):
        offs.append(sum(i * s for i, s in zip(idx, strides)))
    assert all(0 <= o < flat_len for o in offs)
    return offs


@pytest.mark.parametrize(
    "batch,shapes",
    [
        ((), [(3,), (3,), (3, 3)]),
        ((7,), [(7, 3), (1, 3), (3, 3)]),
        ((4, 5), [(4, 1, 3), (5, 3), (4, 5, 3, 3)]),
        ((2, 3, 4), [(3, 1, 3), (2, 1, 4, 3), (4, 3, 3)]),
        ((2, 3, 4, 5, 6), [(2, 1, 4, 1, 6, 3), (3, 1, 5, 1, 3), (6, 3, 3)]),  # > 4 dims: contiguous fallback
        ((0, 4), [(0, 4, 3), (4, 3), (3, 3)]),
    ],
)
def test_batch_strides_describe_the_broadcast(batch, shapes):
    rng = np.random.default_rng(0)
    ops = [torch.from_numpy(rng.normal(size=s).astype(np.float32)) for s in shapes]
    cores = [1, 1, 2]
    ndim, shape, strides, keep = _tensor.batch_strides(batch, list(zip(ops, cores)))
    assert ndim <= 4
    dims = [int(shape[i]) for i in range(ndim)]
    assert int(np.prod(dims, dtype=np.int64)) == int(np.prod(batch, dtype=np.int64))
    if 0 in batch:
        return
    for op, core, st, kept in zip(ops, cores, strides, keep):
        core_shape = tuple(op.shape[op.ndim - core:])
        want = op.expand(*batch, *core_shape).reshape(-1, *core_shape) if batch else op.reshape(1, *core_shape)
        base = kept.contiguous().reshape(-1) if kept.numel() else kept.reshape(-1)
        # the kept tensor is what the kernel indexes: data_ptr + offset (in elements)
        flat = torch.as_strided(kept, (kept.untyped_storage().nbytes() // 4,), (1,), 0) if kept.numel() else base
        offs = _emulate(dims, [int(st[i]) for i in range(ndim)], flat.numel()) if ndim else [0]
        n_core = int(np.prod(core_shape))
        got = torch.stack([flat[o:o + n_core] for o in offs]).reshape(want.shape)
        torch.testing.assert_close(got, want, rtol=0, atol=0)


def test_sum_to_shape_reduces_broadcast_gradients():
    g = torch.arange(24, dtype=torch.float32).reshape(2, 3, 4)
    assert _tensor.sum_to_shape(g, (2, 3, 4)) is g
    torch.testing.assert_close(_tensor.sum_to_shape(g, (3, 1)), g.sum(dim=(0, 2)).reshape(3, 1))
    torch.testing.assert_close(_tensor.sum_to_shape(g, (4,)), g.sum(dim=(0, 1)))
    torch.testing.assert_close(_tensor.sum_to_shape(g, ()), g.sum())


def test_numel_and_defaults():
    assert _tensor.numel(()) == 1 and _tensor.numel((3, 0, 2)) == 0 and _tensor.numel((2, 5)) == 10
    assert _tensor.F32_EPS == float(np.finfo(np.float32).eps)


def test_public_api_refuses_to_run_without_a_gpu():
    """No CPU fallback: without a CUDA device the host API raises instead of computing elsewhere."""
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import differt_b200 as drt

    with pytest.raises(RuntimeError, match="no CPU fallback"):
        drt.ray_intersect_triangle(np.zeros(3, np.float32), np.ones(3, np.float32), np.zeros((3, 3), np.float32))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        drt.Mesh.box()


def test_traced_paths_float_mask_semantics_and_reduce():
    # _paths.py:101-103, 270-283, 461-479: a float mask is a confidence, valid when >= the threshold
    # (NaN never is); reduce() weights by it, or skips invalid paths entirely for a boolean mask
    from differt_b200.mesh import TracedPaths

    v = torch.arange(2 * 3 * 3 * 3, dtype=torch.float32).reshape(2, 3, 3, 3)
    o = torch.zeros((2, 3, 3), dtype=torch.int32)
    it = torch.zeros((2, 3, 1), dtype=torch.int32)
    conf = torch.tensor([[0.5, 0.49999, float("nan")], [1.0, 0.0, 0.75]])
    soft = TracedPaths(vertices=v, objects=o, mask=conf, interaction_types=it)
    assert soft.num_valid_paths == 3 and soft._valid().tolist() == [[True, False, False], [True, False, True]]
    assert TracedPaths(vertices=v, objects=o, mask=conf, interaction_types=it, confidence_threshold=0.8).num_valid_paths == 1
    length = lambda p: torch.linalg.norm(p[..., 1:, :] - p[..., :-1, :], dim=-1).sum(-1)  # noqa: E731
    w = torch.nan_to_num(conf)
    soft0 = TracedPaths(vertices=v, objects=o, mask=w, interaction_types=it)
    torch.testing.assert_close(soft0.reduce(length), (length(v) * w).sum())
    torch.testing.assert_close(soft0.reduce(length, axis=-1), (length(v) * w).sum(-1))
    hard = TracedPaths(vertices=v, objects=o, mask=conf >= 0.5, interaction_types=it)
    poisoned = lambda p: torch.where(hard.mask, length(p), torch.full((2, 3), float("inf")))  # noqa: E731
    torch.testing.assert_close(hard.reduce(poisoned), length(v)[hard.mask].sum())
    assert hard.num_valid_paths == 3 and hard.reshape(6).mask.shape == (6,)
    # differentiable through both the vertices and the confidences
    vg, cg = v.clone().requires_grad_(True), w.clone().requires_grad_(True)
    TracedPaths(vertices=vg, objects=o, mask=cg, interaction_types=it).reduce(length).backward()
    torch.testing.assert_close(cg.grad, length(v))
    assert vg.grad.abs().sum() > 0
