"""CPU checks of bench.py's reference arm (`--impl reference`: the reference algorithm on the host cores; no GPU, no
product code on that path) and of the pure helpers the GPU arm shares with it."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run_reference(*extra, env=None):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", *extra], capture_output=True, text=True, timeout=600,
                         env=dict(os.environ, **(env or {})))
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    return lines


def test_reference_arm_prints_one_contract_line():
    lines = _run_reference("--config", "c2", "--slices", "24")
    assert len(lines) == 1  # ONE JSON line on stdout, everything else on stderr
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "iter/s" and d["higher_is_better"] is True
    assert d["metric"] == "AO-ADMM outer iterations per second" and d["value"] > 0
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 0 and d["vs_baseline"] is None
    assert abs(d["ms_per_step"] * d["value"] - 1000.0) < 1e-6 * 1000.0
    assert "workload" in d["config"] and "REDUCED to 24 slices" in d["config"]["workload"]
    assert d["e2e"] == {"value": d["value"], "unit": "iter/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    # the unmodified reference package when __graft_entry__.build() staged it under oracle/_ref, else the oracle port
    staged = os.path.isfile(os.path.join(ROOT, "oracle", "_ref", "matcouply", "decomposition.py"))
    assert cb["kind"] == ("reference" if staged else "port")
    assert cb["value"] == d["value"] and cb["cores"] >= 1 and "slices" in cb["sample"]
    # forcing the port (what a box without the staged package runs) keeps the same contract line
    dp = json.loads(_run_reference("--config", "c2", "--slices", "24", env={"B2_REFERENCE_ARM": "port"})[0])
    assert dp["cpu_baseline"]["kind"] == "port" and dp["impl"] == "reference"
    if staged:  # same algorithm, same sample: the two arms agree within timing noise and overhead (< 2x)
        assert 0.5 < dp["value"] / d["value"] < 2.0, (dp["value"], d["value"])
    tc = d["cpu_baseline_torch"]  # the torch-CPU column: config 2's penalties run under the reference's torch backend
    assert tc["kind"] == "port" and tc["value"] == tc["value_f64"] > 0 and tc["value_f32"] > 0 and tc["cores"] >= 1


def test_reference_arm_other_ranks_exit_silently_and_threads_are_restored():
    # under torchrun every rank but 0 returns without work; OMP_NUM_THREADS=1 (torchrun's default) must not make
    # rank 0's baseline single-threaded
    assert _run_reference("--config", "c1", "--slices", "16", env={"RANK": "1", "WORLD_SIZE": "2"}) == []
    lines = _run_reference("--config", "c1", "--slices", "16", "--gpus", "2",
                           env={"RANK": "0", "WORLD_SIZE": "2", "OMP_NUM_THREADS": "1"})
    d = json.loads(lines[0])
    try:
        usable = len(os.sched_getaffinity(0))
    except AttributeError:
        usable = os.cpu_count() or 1
    assert d["n_gpus"] == 2 and d["cpu_baseline"]["cores"] == usable
    # config 3 has penalties the reference's torch backend refuses: the column says so instead of a number
    d3 = json.loads(_run_reference("--config", "c3", "--slices", "8")[0])
    assert "unavailable" in d3["cpu_baseline_torch"] and "unimodal" in d3["cpu_baseline_torch"]["unavailable"]


def test_shard_bounds_balance_rows_and_cover_all_slices():
    sys.path.insert(0, ROOT)
    import bench

    sizes = bench.slice_sizes(bench.CONFIGS["c2"])
    assert sizes.min() >= 256 and sizes.max() <= 2048 and len(sizes) == 16384
    for world in (1, 2, 4, 8):
        b = bench.shard_bounds(sizes, world)
        assert b[0][0] == 0 and b[-1][1] == len(sizes) and all(x[1] == y[0] for x, y in zip(b, b[1:]))
        rows = np.array([sizes[lo:hi].sum() for lo, hi in b], dtype=np.float64)
        assert rows.max() / rows.mean() < 1.001  # balanced by row count, not by slice count


def test_weak_scaling_config_grows_with_the_rank_count():
    """--config c4w (BASELINE config[4] as the weak-scaling sweep): slices per GPU fixed, problem x WORLD_SIZE."""
    d1 = json.loads(_run_reference("--config", "c4w", "--slices", "8")[0])
    d2 = json.loads(_run_reference("--config", "c4w", "--slices", "8", "--gpus", "2",
                                   env={"RANK": "0", "WORLD_SIZE": "2"})[0])
    assert d1["scaling"] == d2["scaling"] == "weak"
    assert "[8 slices on 1 GPU(s)]" in d1["config"]["workload"] and "[16 slices on 2 GPU(s)]" in d2["config"]["workload"]
    assert abs(d2["slice_iterations_per_s"] - 16 * d2["value"]) < 1e-9 * d2["slice_iterations_per_s"]
