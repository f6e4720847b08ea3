"""CPU suite, part 4: bench.py's contract on the legs that run without a GPU (the reference arm)."""
from __future__ import annotations

import json
import os
import subprocess
import sys

from conftest import ROOT


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["unit"] == "Mpix/s" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["config"]["workload"].startswith("3840x2160") and "model" not in d["config"]      # default = c3, BASELINE's 1/2/4/8-GPU config
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_other_ranks_of_the_reference_arm_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1"],
                         capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_extras_tables_name_known_workloads():
    """The extra records of a default run (other sharded workloads, CONV mode) are table-driven: every entry names a
    BASELINE workload and a mode, and the N > 1 set contains the row-band CONV cases the north_star asks for."""
    sys.path.insert(0, ROOT)
    import bench
    for key, wl, mode, steps, warm, e2e in bench.EXTRAS_N1:
        assert wl in bench.WORKLOADS and mode in ("ref", "conv") and steps >= 1 and warm >= 3
    keys = [k for k, *_ in bench.EXTRAS_NX]
    assert {"c4_ref_rowband", "c4_conv_rowband", "c5_conv_rowband"} <= set(keys)
    for key, wl, mode, steps, warm in bench.EXTRAS_NX:
        assert bench.WORKLOADS[wl][4] in ("rowband", "batch") and warm >= 3
    assert bench.WORKLOADS["c3"][3] == 256 and bench.parse_args.__module__ == "bench"


def test_host_side_helpers_never_break_a_run():
    """bind_to_gpu_numa_node reports what it finds and never raises (here: no GPU at all); Solo offers every method of
    Dist that measure() / run_e2e() call, so a one-rank measurement inside an N-rank job takes the same code path."""
    sys.path.insert(0, ROOT)
    import bench
    info = bench.bind_to_gpu_numa_node(0)
    assert isinstance(info, dict) and "numa_node" in info
    solo = bench.Solo()
    for name in ("barrier", "max", "sum", "gather"):
        assert callable(getattr(solo, name)) and callable(getattr(bench.Dist, name))
    assert solo.max(3.5) == 3.5 and solo.sum(2.0) == 2.0 and solo.gather(1.25) == [1.25] and solo.world == 1
    assert bench.square_equivalent(2160, 3840) == 2880 == bench.REF_SIDE_CAP          # c3's reference arm runs at full size
