"""Shared pytest configuration: marker registration, import paths and common fixtures."""

from __future__ import annotations

import json
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config: pytest.Config) -> None:
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def kats() -> dict:
    return json.loads((GOLDEN / "reference_kats.json").read_text())


@pytest.fixture(scope="session")
def two_buildings() -> tuple[np.ndarray, np.ndarray]:
    data = np.load(GOLDEN / "two_buildings.npz")
    return data["vertices"], data["triangles"]


@pytest.fixture(scope="session")
def bruxelles() -> tuple[np.ndarray, np.ndarray]:
    data = np.load(GOLDEN / "bruxelles.npz")
    return data["vertices"], data["triangles"]


@pytest.fixture(scope="session")
def golden_dir() -> Path:
    return GOLDEN


@pytest.fixture()
def rng() -> np.random.Generator:
    return np.random.default_rng(1234)
