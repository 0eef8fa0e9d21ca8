"""bench.py contract checks that need no GPU: the reference arm (`--impl reference`) runs the compiled reference (or the
oracle port) on the host cores and prints exactly ONE JSON line on stdout with the keys the driver reads; the CUDA arm's
byte model and ncu-traffic lookup are plain functions."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1",
                        "--config", "c1_sphere100k"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "Mtris/s" and d["unit"] == "Mtris/s"
    assert d["higher_is_better"] is True and d["steps"] == 2 and d["warmup"] == 1 and d["n_gpus"] == 1
    assert d["value"] > 0 and d["ms_per_step"] > 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "Mtris/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "c1_sphere100k" in d["config"]["workload"]


def test_reference_arm_other_ranks_stay_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0",
                        "--config", "c1_sphere100k"], capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_byte_model_and_traffic_lookup():
    sys.path.insert(0, ROOT)
    import bench
    from softlight_b200 import scenes as S

    sc = S.SCENES["c1_sphere100k"]()
    alg = bench.algorithmic_bytes("c1_sphere100k", sc, 362443)
    # SURVEY.md section 8d worked example: indices 1.204 MB + vertices 1.620 MB + clear 6.29 MB + fragments x 12 B = 13.46 MB
    assert alg["clear"] == 1024 * 768 * 8 and alg["vertex"] == 100352 * 12 + 50625 * 32
    assert alg["raster"] + alg["shade"] == 362443 * 12 and abs(alg["frame"] - 13464996) < 16
    # ncu figures are only quoted for the kernel sources they were captured on
    n = bench.ncu_counters("raster", "c3_indoor2m")
    assert n is None or (n["inst_executed"] > 0 and n["source"].endswith(bench.kernel_source_sha()))


def test_both_arms_describe_the_workload_identically():
    sys.path.insert(0, ROOT)
    import bench
    from softlight_b200 import scenes as S

    assert bench.DEFAULT_CONFIG == "c3_indoor2m"  # the configuration BASELINE.json's metric is quoted on (4K)
    sc = S.SCENES[bench.DEFAULT_CONFIG]()
    assert (sc.width, sc.height) == (3840, 2160)
    cfg = bench.workload_config(bench.DEFAULT_CONFIG, sc)
    assert list(cfg) == ["workload"] and cfg["workload"].startswith("c3_")
