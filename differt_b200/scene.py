"""``Scene``: the caller of the hot path, with the reference's call signatures.

Only what drives the hot path is mirrored (reference
``differt/src/differt/geometry/_scene.py:305-407, 650-835, 1250-1371``): transmitters / receivers / mesh,
``trace_paths``, ``launch_paths``, ``compute_tx_mlm`` and the receiver / transmitter grids.  Loading,
plotting and the EM layer are out of scope.
"""

from __future__ import annotations

import dataclasses
import math
import warnings
from typing import Iterator

import torch

from . import launch, solvers
from .mesh import Mesh, TracedPaths

__all__ = ["Scene"]


@dataclasses.dataclass
class Scene:
    """Reference ``Scene`` (``_scene.py:305-320``): ``transmitters [*tx_batch, 3]``,
    ``receivers [*rx_batch, 3]`` and one ``Mesh``, all on the mesh's device."""

    transmitters: torch.Tensor
    receivers: torch.Tensor
    mesh: Mesh

    def __post_init__(self) -> None:
        dev = self.mesh.vertices.device
        self.transmitters = torch.as_tensor(self.transmitters).to(dev, torch.float32)
        self.receivers = torch.as_tensor(self.receivers).to(dev, torch.float32)
        if self.transmitters.shape[-1:] != (3,) or self.receivers.shape[-1:] != (3,):
            raise TypeError("transmitters and receivers must have shape [*batch, 3]")

    @property
    def num_transmitters(self) -> int:
        return math.prod(self.transmitters.shape[:-1])

    @property
    def num_receivers(self) -> int:
        return math.prod(self.receivers.shape[:-1])

    def set_assume_quads(self, flag: bool = True) -> "Scene":
        return dataclasses.replace(self, mesh=self.mesh.set_assume_quads(flag))

    def _grid(self, m: int, n: int | None, height: float) -> torch.Tensor:
        n = m if n is None else n
        lo, hi = self.mesh.vertices.amin(dim=0), self.mesh.vertices.amax(dim=0)
        x = torch.linspace(float(lo[0]), float(hi[0]), m, device=lo.device)
        y = torch.linspace(float(lo[1]), float(hi[1]), n, device=lo.device)
        xx, yy = torch.meshgrid(x, y, indexing="xy")  # jnp.meshgrid default (_scene.py:398-401)
        return torch.stack((xx, yy, torch.full_like(xx, height)), dim=-1)

    def with_transmitters_grid(self, m: int = 50, n: int | None = 50, *, height: float = 1.5) -> "Scene":
        """``_scene.py:343-375``."""
        return dataclasses.replace(self, transmitters=self._grid(m, n, height))

    def with_receivers_grid(self, m: int = 50, n: int | None = 50, *, height: float = 1.5) -> "Scene":
        """``_scene.py:377-407``."""
        return dataclasses.replace(self, receivers=self._grid(m, n, height))

    # -- the three hot-path entry points ---------------------------------------------------------

    def trace_paths(self, order: int | None = None, *, solver: str = "exhaustive", path_candidates=None,
                    chunk_size: int | None = None, **solver_kwargs) -> TracedPaths | Iterator[TracedPaths]:
        """``Scene.trace_paths`` (``_scene.py:650-764``): batch shape ``(*tx_batch, *rx_batch, C)``.
        ``solver`` is ``"exhaustive"`` or ``"hybrid"``; ``chunk_size`` returns an iterator of chunks."""
        if (order is None) == (path_candidates is None):
            raise ValueError("You must specify one of 'order' or `path_candidates`, not both.")
        if solver not in ("exhaustive", "hybrid"):
            raise ValueError(f"Unknown solver: {solver}")
        if solver == "hybrid" and order is None:
            raise ValueError("Argument 'order' is required when using HybridPathTracer.")
        if solver == "hybrid" and solver_kwargs.get("smoothing_factor") is not None:
            warnings.warn("Argument 'smoothing' is currently ignored when using HybridPathTracer.", UserWarning,
                          stacklevel=2)
            solver_kwargs = {**solver_kwargs, "smoothing_factor": None}
        tx_batch, rx_batch = tuple(self.transmitters.shape[:-1]), tuple(self.receivers.shape[:-1])
        tx, rx = self.transmitters.reshape(-1, 3), self.receivers.reshape(-1, 3)
        gen_kwargs = {k: solver_kwargs.pop(k) for k in ("num_rays", "accel") if k in solver_kwargs}

        def shaped(p: TracedPaths) -> TracedPaths:
            return p.reshape(*tx_batch, *rx_batch, int(p.mask.shape[-1]))

        if path_candidates is not None:
            if chunk_size is not None:
                warnings.warn("Argument 'chunk_size' is ignored when 'path_candidates' is provided.", UserWarning,
                              stacklevel=2)
            cand = torch.as_tensor(path_candidates).to(self.mesh.vertices.device, torch.int32)
            if self.mesh.assume_quads:
                cand = cand - cand % 2  # _scene.py:753-756
            return shaped(solvers.trace_path_candidates(self.mesh, tx, rx, cand, **solver_kwargs))
        if chunk_size is not None:
            return (shaped(p) for p in solvers.trace_paths_chunks_iter(
                self.mesh, tx, rx, order, chunk_size=chunk_size, solver=solver, **gen_kwargs, **solver_kwargs))
        return shaped(solvers.trace_paths(self.mesh, tx, rx, order, solver=solver, **gen_kwargs, **solver_kwargs))

    def launch_paths(self, order: int, *, num_rays: int = 1_000_000, epsilon=None, max_dist: float = 1e-3,
                     accel: str = "brute") -> launch.LaunchedPaths:
        """``Scene.launch_paths`` with the SBR launcher (``_scene.py:783-835``, ``_solvers.py:358-491``);
        transmitters and receivers are flattened like the reference does."""
        return launch.launch_paths(self.mesh, self.transmitters.reshape(-1, 3), self.receivers.reshape(-1, 3),
                                   order, num_rays=num_rays, epsilon=epsilon, max_dist=max_dist, accel=accel)

    def compute_tx_mlm(self, max_order: int, dim_x: int, dim_y: int, num_rays: int = 1_000_000,
                       min_order: int = 0, height: float | None = None, *, accel: str = "brute") -> torch.Tensor:
        """``Scene.compute_tx_mlm`` (``_scene.py:1250-1371``): grid over the mesh's bounding box at
        ``height`` (default: the first receiver's, else 1.5) → ``[*tx_batch, dim_x, dim_y]``."""
        if height is not None:
            receiver_height = float(height)
        elif self.receivers.numel() > 0:
            receiver_height = float(self.receivers.reshape(-1, 3)[0, 2])
        else:
            receiver_height = 1.5
        lo, hi = self.mesh.vertices.amin(dim=0), self.mesh.vertices.amax(dim=0)
        out = launch.compute_tx_mlm(
            self.mesh, self.transmitters.reshape(-1, 3), max_order=max_order, min_order=min_order, dim_x=dim_x,
            dim_y=dim_y, num_rays=num_rays, receiver_height=receiver_height, min_x=float(lo[0]), max_x=float(hi[0]),
            min_y=float(lo[1]), max_y=float(hi[1]), accel=accel,
        )
        return out.reshape(*self.transmitters.shape[:-1], dim_x, dim_y)
