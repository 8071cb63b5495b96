"""bench.py contract on the CPU side: the reference arm (the reference's own CPU code on the host cores) runs without a GPU and prints
one JSON line with the keys the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--size", "512x384"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    d = json.loads(out.stdout.strip().splitlines()[-1])
    assert d["impl"] == "reference" and d["value"] > 0 and d["unit"] == "pairs/s" and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "MSER" in d["config"]["workload"]
    # both arms print the SAME metric string (the driver forms vs_reference only then), and the arm reports what it ran
    import bench
    assert d["metric"] == bench.METRIC and d["steps"] == 1 and d["warmup"] >= 1
    assert "nothing cropped or extrapolated" in d["cpu_baseline"]["sample"] and d["config"]["tentatives"] > 0


def test_reference_arm_honours_steps_within_its_budget():
    env = dict(os.environ, MB2_REF_BUDGET_S="1000")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "3", "--warmup", "2", "--size", "384x288"],
                         capture_output=True, text=True, timeout=600, env=env)
    d = json.loads(out.stdout.strip().splitlines()[-1])
    assert (d["steps"], d["warmup"]) == (3, 2)
    env["MB2_REF_BUDGET_S"] = "0"      # nothing fits: one warm-up, one step, and the line says so
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "3", "--warmup", "2", "--size", "384x288"],
                         capture_output=True, text=True, timeout=600, env=env)
    d = json.loads(out.stdout.strip().splitlines()[-1])
    assert (d["steps"], d["warmup"]) == (1, 1) and d["config"]["requested_steps"] == 3


def test_ours_refuses_to_run_without_a_gpu():
    try:
        import torch
        if torch.cuda.is_available():
            return
    except Exception:
        pass
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1"], capture_output=True, text=True, timeout=600)
    assert out.returncode != 0 and "no CPU fallback" in (out.stderr + out.stdout)


def test_clock_sampler_degrades_without_a_gpu():
    """The clocks sampler (NVML in-process, nvidia-smi as fallback) must never take the bench down: without a driver it reports no source."""
    import bench
    s = bench.ClockSampler(0)
    s.start()
    out = s.stop()
    assert set(("sm_mhz", "sm_max_mhz", "reasons")) <= set(out)
    if out["sm_mhz"] is None:
        assert out["reasons"] and "unavailable" in out["reasons"][0] or out.get("samples", 0) == 0


def test_workload_strings_name_the_baseline_configs():
    import bench
    assert bench.workload_string(4096, 3072, False).startswith("C3:") and "MSER" in bench.workload_string(4096, 3072, False)
    assert bench.workload_string(1920, 1080, False, True).startswith("C5:") and "HalfRootSIFT" in bench.workload_string(1920, 1080, False, True)
    assert (bench.C4_WORKLOAD % (4096, 3072)).startswith("C4:")
