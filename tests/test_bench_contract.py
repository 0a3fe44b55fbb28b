"""bench.py's host-side pieces that run without a GPU: the reference arm (the restated CPU path timed on the host
cores, the one leg of bench.py that may execute oracle/), the bounded cpu_baseline sample, and the algorithmic-byte
formula of SURVEY.md 8(d)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line(orc):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "12", "--warmup", "3",
                          "--roots", "48"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "mcts_simulations_per_sec" and d["unit"] == "simulations/s"
    assert d["higher_is_better"] is True and d["steps"] == 12 and d["warmup"] == 3 and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0 and d["config"]["steps_run"] == 12


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"], capture_output=True,
                         text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_cpu_sample_is_bounded_in_time(orc):
    import bench

    val, dt, k = bench.cpu_reference_run(19, 64, 400, 2, 0, 2, budget_s=0.05)
    assert k["steps_run"] < 400 and k["steps_run"] % 50 == 0 and val > 0  # stopped after the chunk that crossed the budget
    val2, dt2, k2 = bench.cpu_reference_run(19, 64, 30, 2, 0, 2)
    assert k2["steps_run"] == 30 and k2["n_live"] > 0


def test_algorithmic_bytes_formula():
    import bench

    # SURVEY 8(d) at N = 19: NODE 32, EDGE 20, PRED 12, KEY 24, ST 40, VEC 1216, H 608
    zero = dict(n_sel=0, d_sel=0, n_cand=0, n_probe=0, n_ins=0, n_arc=0, n_pred=0, n_cn=0, d_cn=0, n_reset=0, n_live=0)
    assert bench.algorithmic_bytes(dict(zero, n_live=1), 19) == 2 * 40 + 1216 + 608
    assert bench.algorithmic_bytes(dict(zero, n_probe=1, n_ins=1), 19) == 24 + 32 + 24
    assert bench.algorithmic_bytes(dict(zero, n_sel=1, d_sel=3, n_cand=5), 19) == 32 + 3 * 52 + 5 * 12
    assert bench.algorithmic_bytes(dict(zero, n_cn=2, d_cn=1, n_arc=1, n_pred=7, n_reset=1), 19) == 128 + 20 + 24 + 84 + 40
    assert bench.algorithmic_bytes(dict(zero, n_live=1), 64) == 2 * 308 + 15616 + 7808
