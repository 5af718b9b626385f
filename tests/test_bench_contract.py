"""bench.py's contract: the JSON line of both arms, the reference arm's independence from the
launcher's environment (torchrun exports OMP_NUM_THREADS=1), and the device-side synthetic generators."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT

BENCH = os.path.join(ROOT, "bench.py")


def run_bench(args, env=None, timeout=900):
    r = subprocess.run([sys.executable, BENCH] + args, capture_output=True, text=True, timeout=timeout,
                       env=dict(os.environ, **(env or {})))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    return r.stdout


def test_reference_arm_line(oracle_libs):
    out = run_bench(["--impl", "reference", "--grid", "256", "--steps", "2", "--warmup", "1"])
    d = json.loads(out.strip().splitlines()[-1])
    assert d["impl"] == "reference" and d["metric"] == "gfmd_force_eval_steps_per_sec" and d["unit"] == "steps/s"
    assert d["value"] > 0 and d["higher_is_better"] is True and d["gpu_launches"] == 0
    assert d["extrapolated"] is False and "256x256" in d["config"]["workload"]
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]


def test_reference_arm_ignores_torchrun_thread_default(oracle_libs):
    """Under torchrun rank 0 alone runs, on ALL host cores although the launcher exports
    OMP_NUM_THREADS=1; the other ranks exit 0 without output."""
    env = {"OMP_NUM_THREADS": "1", "WORLD_SIZE": "2", "RANK": "0", "LOCAL_RANK": "0"}
    d = json.loads(run_bench(["--impl", "reference", "--gpus", "2", "--grid", "128", "--steps", "1", "--warmup", "1"],
                             env).strip().splitlines()[-1])
    assert d["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))
    assert d["n_gpus"] == 2
    out = run_bench(["--impl", "reference", "--gpus", "2", "--grid", "128", "--steps", "1"], dict(env, RANK="1"))
    assert out.strip() == ""


def test_reference_arm_flags_extrapolation(oracle_libs, monkeypatch):
    """A surface beyond what the reference arm runs in full is sampled and FLAGGED, not silently scaled."""
    sys.path.insert(0, ROOT)
    import bench
    monkeypatch.setattr(bench, "REF_SAMPLE_GRID", 64)
    r = bench.reference_run(256, 256, 3, 1, 1)
    assert r["extrapolated"] is True and "extrapolated" in r["sample"] and r["measured"]["grid"] == "64x64"
    assert abs(r["value"] - r["measured"]["value"] / 16.0) <= 1e-9 * r["value"]


def test_device_generators_match_host_ones():
    import torch
    from gfmd_b200 import synthetic as S
    a = S.sc100_dynamical_matrices(24, 18, 3, 5)
    b = S.sc100_dynamical_matrices_torch(24, 18, 3, 5, "cpu").numpy()
    assert np.abs(a - (b[..., 0] + 1j * b[..., 1])).max() < 1e-15
    # slabs of a field are slices of the whole field (every value a function of the global cell index)
    u = S.displacement_field_torch(64, 48, 0, 64, "cpu")
    for x0, n in ((0, 16), (16, 16), (48, 16)):
        assert torch.equal(S.displacement_field_torch(64, 48, x0, n, "cpu"), u[:, x0:x0 + n])
    assert 1e-4 < float(u.abs().max()) < 2e-2
    x, xeq, gid, mask = S.atoms_for_slab_torch(64, 48, 16, 16, u[:, 16:32].contiguous())
    assert x.shape == (16 * 48, 3) and int(gid[:, 0].min()) == 16 and int(gid[:, 0].max()) == 31
    assert torch.allclose((x - xeq).t().reshape(3, 16, 48), u[:, 16:32], atol=1e-15)


@pytest.mark.gpu
def test_bench_line_small_surface():
    """The b200 arm end to end on a small surface: every key the contract names is present."""
    out = run_bench(["--grid", "2048", "--steps", "3", "--warmup", "3", "--no-cpu-baseline", "--no-4096"])
    d = json.loads(out.strip().splitlines()[-1])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["gpu_launches"] >= 3 * 6 and d["value"] > 0 and d["scaling"] == "strong"
    assert d["e2e"]["h2d_bytes_per_step"] == 3 * 2048 * 2048 * 8
    assert d["roofline"]["bound"] == "hbm" and 0 < d["roofline"]["frac"] < 1.5
    assert d["parity_max_rel_err"] < 1e-11            # E = -1/2 sum f.u on the timed surface
    lat = d["config"]["latency"]
    assert set(lat) >= {"C1_sc100_128x128", "C2_fcc111_64x37", "C3_fcc100_two_layers_10x10"}
