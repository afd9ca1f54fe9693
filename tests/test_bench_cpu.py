"""bench.py's reference arm (--impl reference: the C port of the reference algorithm on the host
cores) keeps the JSON contract the driver parses; rank != 0 of a multi-rank launch prints nothing."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = {"impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
        "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"}


def _run(extra, env=None):
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"] + extra
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, **(env or {})))
    assert res.returncode == 0, res.stderr[-2000:]
    return res.stdout.strip().splitlines()


@pytest.mark.parametrize("extra,metric", [
    (["--workload", "indexer", "--kmer", "11", "--scale", "0.004", "--cpu-sample-mbp", "2"], "indexer_bp_per_s_K11"),
    (["--workload", "indexer", "--kmer", "11", "--scale", "0.004"], "indexer_bp_per_s_K11"),
    (["--workload", "merger", "--kmer", "9", "--samples", "6"], "merger_bitmask_GB_per_s"),
])
def test_reference_arm_prints_one_contract_line(extra, metric):
    lines = _run(extra)
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert KEYS <= set(line) and line["impl"] == "reference" and line["metric"] == metric
    assert line["value"] > 0 and line["gpu_launches"] == 0 and line["vs_baseline"] is None
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0,
                           "d2h_bytes_per_step": 0}
    assert "workload" in line["config"] and "model" not in line["config"]


def test_reference_arm_default_line_carries_every_named_metric():
    """No --workload: the headline indexer line + the merger sub-objects (BASELINE.json names indexer
    bp/s AND merger GB/s); the workload strings are the ones the CUDA arm prints (same_config)."""
    import bench
    lines = _run(["--kmer", "11", "--scale", "0.004", "--no-reference-python", "--sub-steps", "1"])
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert KEYS <= set(line) and line["metric"] == "indexer_bp_per_s_K11"
    assert line["config"]["workload"] == bench.workload_name("indexer", 11, bp=3130073)
    for tag, n, mc in (("n50", 50, 50), ("n255", 255, 255)):
        sub = line["merger"][tag]
        assert KEYS <= set(sub) and sub["metric"] == "merger_bitmask_GB_per_s" and sub["value"] > 0
        assert sub["config"]["workload"] == bench.workload_name("merger", 15, N=n, max_count=mc)


def test_reference_python_leg_reports_unavailable_without_a_copy(tmp_path, monkeypatch):
    """cpu_baseline.reference_python never takes the line down: without baseline/_ref it says so."""
    import bench
    monkeypatch.setattr(bench, "ROOT", str(tmp_path))
    assert "unavailable" in bench.reference_python()


def test_reference_arm_other_ranks_exit_quietly():
    assert _run(["--kmer", "11", "--scale", "0.004"], env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}) == []
