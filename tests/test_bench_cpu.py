"""bench.py's CPU arm (`--impl reference`) runs without a GPU: the line carries the contract's keys, and the arm uses every
host core even when the launcher exported OMP_NUM_THREADS=1 (torchrun does that to every rank)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_line_and_threads():
    env = dict(os.environ, OMP_NUM_THREADS="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--n", "96", "--steps", "1",
                          "--warmup", "1"], capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads([ln for ln in out.stdout.splitlines() if ln.startswith("{")][-1])
    assert line["impl"] == "reference" and line["metric"] == "newton_steps_per_s" and line["higher_is_better"] is True
    assert line["steps"] == 1 and line["warmup"] == 1 and line["value"] > 0 and line["dtype"] == "f64"
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = line["cpu_baseline"]
    # the reference's own modules when oracle/build_ref.py has been run (this container, and the GPU box via the
    # snapshot), else the oracle port
    have_ref = os.path.exists(os.path.join(ROOT, "oracle", "_ref", "LPSolver.py"))
    assert cb["kind"] == ("reference" if have_ref else "port") and cb["value"] == line["value"]
    assert line["config"]["workload"].startswith("dense LP n=96")
    assert cb["cores"] == len(os.sched_getaffinity(0))


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                         capture_output=True, text=True, timeout=120, env=env, cwd=ROOT)
    assert out.returncode == 0 and out.stdout.strip() == ""
