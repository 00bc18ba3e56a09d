"""The reference arm of bench.py (CPU only: the oracle on a bounded, density-matched sample of the bench workload) prints
one JSON line that follows the driver's contract.  Uses the smallest workload so that the test stays within seconds."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def test_reference_arm_prints_the_contract_line():
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "c1_cpu",
                          "--steps", "2", "--warmup", "1"], capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "it/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["n_gpus"] == 1 and d["steps"] == 2 and d["gpu_launches"] == 0
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "oracle" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "it/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]
    # honest bookkeeping: the frames actually run, an explicit extrapolation flag, the workload's own counts
    assert d["extrapolated"] is True and d["warmup"] == 1 and d["requested_steps"] == 2
    wc = d["config"]["workload_counts"]
    assert wc["n"] == 10000 and 0 < wc["visible"] <= wc["n"] and wc["duplicates"] >= wc["visible"]
    assert d["metric"].startswith("train iters/sec (raster fwd+bwd+loss) at 256x256/10K Gaussians")


def test_reference_arm_caps_the_frames_it_runs():
    sys.path.insert(0, ROOT)
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    assert b.metric_name("c4_iphone") == "train iters/sec (raster fwd+bwd+loss) at 1080p/2M Gaussians; % HBM roofline"   # BASELINE.json
    assert "960x540/1M" in b.metric_name("c3_nvidia") and "512x512/300K" in b.metric_name("c2_kubric")
    assert b.sample_shape("c4_iphone")[:3] == (62745, 256, 256)


def test_cpu_sample_matches_the_workload_density():
    """The bounded sample keeps the workload's Gaussians per tile: N' = N * 256 / tiles on a 256x256 window."""
    sys.path.insert(0, ROOT)
    from rodygs_b200 import synthetic
    for cfg, (N, H, W, T, _) in synthetic.CONFIGS.items():
        if cfg == "c1_cpu":
            continue
        tiles = ((H + 15) // 16) * ((W + 15) // 16)
        n_sample = int(round(N * 256 / tiles))
        assert abs(n_sample / 256 - N / tiles) < 1.0
        assert n_sample < 200_000            # seconds per frame on the host cores, not minutes
