"""bench.py's driver contract that can be checked without a GPU: the reference arm prints exactly one JSON line with
the agreed keys (and runs the unmodified reference binary when it is present), and the b200 arm refuses to run
without a device instead of falling back."""
import json
import os
import subprocess
import sys

import helpers as H

BENCH = os.path.join(H.ROOT, "bench.py")


def test_reference_arm_prints_one_json_line():
    r = subprocess.run([sys.executable, BENCH, "--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-seconds", "2",
                        "--n-ind", "60"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "snp_pairs_per_sec" and d["unit"] == "pairs/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["gpu_launches"] == 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"] == d["e2e"]["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"]


def test_reference_arm_other_ranks_stay_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, BENCH, "--impl", "reference", "--gpus", "2"], capture_output=True, text=True, env=env,
                       timeout=60)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_b200_arm_needs_a_gpu():
    if os.path.exists("/dev/nvidia0"):
        return
    r = subprocess.run([sys.executable, BENCH, "--steps", "1", "--warmup", "3"], capture_output=True, text=True, timeout=300)
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)


def test_measured_peaks_lookup(tmp_path):
    sys.path.insert(0, H.ROOT)
    import bench
    assert bench.measured_hbm_peak(str(tmp_path / "absent.json")) == (6650.0, "fallback (B200_PROFILING.md)")
    p = tmp_path / "MEASURED_PEAKS.json"
    p.write_text(json.dumps({"hbm_gbs": 6553.9, "bf16_tflops": 1637.0}))
    assert bench.measured_hbm_peak(str(p))[0] == 6553.9
    p.write_text(json.dumps({"peaks": [{"name": "x"}, {"hbm_gbs_sustained": 6400}]}))
    assert bench.measured_hbm_peak(str(p))[0] == 6400.0
    p.write_text("not json")
    assert bench.measured_hbm_peak(str(p))[0] == 6650.0
