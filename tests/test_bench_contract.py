"""bench.py's reference arm runs on CPU: check the JSON contract of its line (the GPU arm prints the
same keys plus roofline / clocks and is exercised on the GPU box)."""

from __future__ import annotations

import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    import os

    # the reference arm must not load the product: with the library path pointing nowhere, importing
    # the package would raise
    env = dict(os.environ, DIFFERT_B200_LIB="/nonexistent/libdiffert_b200.so")
    res = subprocess.run(
        [sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
         "--cpu-seconds", "0.3", "--workload", "urban10k_small"],
        capture_output=True, text=True, timeout=600, cwd=ROOT, env=env,
    )
    assert res.returncode == 0, res.stderr
    lines = [ln for ln in res.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "impl", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["metric"] == "ray_triangle_tests_per_s" and d["unit"] == "tests/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0
    assert d["value"] > 1e6 and "workload" in d["config"]
    # same unit as the GPU arm: decided pairs per second with the same early exit, executed rate beside it
    assert d["executed_tests_per_s"] <= d["value"] and 0 < d["executed_fraction_of_algorithmic"] <= 1
    assert d["dense_no_early_exit_tests_per_s"] > 0


def test_non_zero_ranks_of_the_reference_arm_do_nothing():
    import os

    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    res = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--gpus", "2"],
                         capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert res.returncode == 0 and res.stdout.strip() == ""


def test_workload_is_deterministic_and_sharded_per_rank():
    sys.path.insert(0, str(ROOT))
    import numpy as np

    import bench

    a = bench.build_workload("urban10k_small", 0, 2)
    b = bench.build_workload("urban10k_small", 1, 2)
    a2 = bench.build_workload("urban10k_small", 0, 2)
    one = bench.build_workload("urban10k_small", 0, 1)
    np.testing.assert_array_equal(a["cand"], a2["cand"])
    # strong scaling: ONE fixed problem, its receivers dealt round-robin, every candidate on every rank
    np.testing.assert_array_equal(a["cand"], b["cand"])
    np.testing.assert_array_equal(a["cand"], one["cand"])
    from differt_b200.distributed import RX_BLOCK, receiver_shard

    assert bench.RX_BLOCK == RX_BLOCK
    for w_, r_ in ((a, 0), (b, 1)):
        mine = receiver_shard(one["rx"].shape[0], 2, r_).numpy()
        np.testing.assert_array_equal(w_["rx"], one["rx"][mine])
        np.testing.assert_array_equal(w_["rx_index"], mine)
    assert a["rx_global"] == one["rx"].shape[0] and a["rx"].shape[0] + b["rx"].shape[0] == one["rx"].shape[0]
    assert (a["cand"][:, 1:] != a["cand"][:, :-1]).all()
    # the weak-scaling leg: every rank its own candidates, all receivers
    wa, wb = bench.build_workload("urban10k_small", 0, 2, weak=True), bench.build_workload("urban10k_small", 1, 2, weak=True)
    assert wa["cand_start"] == 0 and wb["cand_start"] == wa["cand"].shape[0] and wa["cand_global"] == 2 * wa["cand"].shape[0]
    assert not np.array_equal(wa["cand"], wb["cand"]) and wa["rx"].shape == one["rx"].shape
