"""CPU tests of bench.py's contract: the reference arm's JSON line (runs the CPU oracle, no CUDA needed) and the
algorithmic-bytes table used for the roofline block."""
import json
import os
import subprocess
import sys

import helpers as h

ROOT = h.ROOT


def test_reference_arm_prints_one_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "1",
                          "--steps", "1", "--warmup", "0", "--gaussians", "6000", "--width", "160", "--height", "120"],
                         capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "views/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("views/sec fwd+bwd") and d["n_gpus"] == 1 and d["steps"] == 1
    assert d["value"] > 0 and abs(d["value"] - d["cpu_baseline"]["value"]) < 1e-9
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and "sample" in d["cpu_baseline"]
    assert d["e2e"] == {"value": d["value"], "unit": "views/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["vs_baseline"] is None and "workload" in d["config"]


def test_reference_arm_other_ranks_exit_without_work():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                         capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_ours_arm_fails_loudly_without_cuda():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=120, cwd=ROOT)
    assert out.returncode != 0 and "no CUDA device" in (out.stderr + out.stdout)


def test_algorithmic_bytes_match_design_table():
    sys.path.insert(0, ROOT)
    import bench
    N, K, P, T = 300_000, 1_295_000, 1920 * 1080, 8160
    ab = bench.algorithmic_bytes(N, K, P, T)
    assert ab["blend_fwd"] == 48 * K + 28 * P and ab["blend_bwd"] == 48 * K + 28 * P + 40 * N
    assert ab["sh_color"] == (12 + 12 * 16) * N + 12 * N
    assert ab["preprocess_bwd"] == (84 + 192) * N + (56 + 192) * N
    assert set(ab) >= {"project", "tile_scan", "sh_color", "emit", "sort_pack", "blend_fwd", "blend_bwd",
                       "preprocess_bwd", "photometric_fwd", "photometric_bwd"}


def test_both_arms_report_the_same_config_dict():
    """the driver compares the two arms' `config` dicts key by key -- one function builds both."""
    sys.path.insert(0, ROOT)
    import argparse
    import bench
    a = argparse.Namespace(gaussians=300_000, width=1920, height=1080, gpus=4)
    ours, ref = bench.config_dict(a, 4), bench.config_dict(a, max(1, a.gpus))
    assert ours == ref and ours["workload"] == bench.WORKLOAD and ours["views_per_step"] == 4
    src = open(os.path.join(ROOT, "bench.py")).read()
    assert src.count('"config": config_dict(') == 2
