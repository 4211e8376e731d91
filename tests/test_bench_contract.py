"""bench.py's reference arm (the CPU leg the driver runs beside the GPU arm) on a reduced sample: one JSON line on
stdout with the contract's keys, and the algorithmic-byte model against the SURVEY 8d worked figures."""
import json
import os
import subprocess
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    # 8 x 2048 cells per step instead of 192 x 2048; OMP_NUM_THREADS=1 is what torch.distributed.run exports at N > 1:
    # the arm must still use every core (round 1's N > 1 reference lines ran on one thread)
    env = dict(os.environ, CQ_BENCH_SAMPLE_COLS="8", OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, env=env, timeout=600)  # fmt: skip
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "element_tensors_per_s" and d["unit"] == "element tensors/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["steps"] == 1
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0)) and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_algorithmic_bytes_match_the_survey_figures():
    sys.path.insert(0, ROOT)
    import bench

    C, P = 4194304, 4194304 * 40  # config 2: 2048^2 cells, 40 points per cell on average
    q1 = bench.algorithmic_bytes(C, P, 2, 4, 4, 2049 * 2049, 37761025)
    q2 = bench.algorithmic_bytes(C, P, 2, 4, 9, 2049 * 2049, 268468225)
    assert abs(q1 - (4.30e9 + 0.30e9)) < 0.05e9  # in 4.30 GB + out 0.30 GB
    assert abs(q2 - (4.38e9 + 2.15e9)) < 0.05e9  # in 4.38 GB + out 2.15 GB
