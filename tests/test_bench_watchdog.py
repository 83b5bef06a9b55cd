"""bench.py's watchdog (host logic, no GPU): a leg that overruns its deadline must not cost the JSON line."""
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]

SCRIPT = """
import sys, time
sys.path.insert(0, {root!r})
import bench
wd = bench.Watchdog()
wd.start()
{setup}
wd.leg("a leg that never returns", 0.5)
time.sleep(60)
print("not reached")
"""


def _run(setup):
    return subprocess.run([sys.executable, "-c", SCRIPT.format(root=str(ROOT), setup=setup)], capture_output=True,
                          text=True, timeout=50)


def test_overrun_prints_what_was_measured_before():
    r = _run('wd.line = {"metric": "pair-distances/s (fused top-k)", "value": 1.5e14, "e2e": {"value": None}}')
    assert r.returncode == 0
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["value"] == 1.5e14 and "a leg that never returns" in line["watchdog"]
    assert "not reached" not in r.stdout and "Thread" in r.stderr      # stacks of all threads on stderr


def test_overrun_before_any_measurement_fails_loudly():
    r = _run("")
    assert r.returncode == 3 and r.stdout.strip() == ""
    assert "did not finish in time" in r.stderr


def test_finished_run_never_sees_it():
    code = SCRIPT.format(root=str(ROOT), setup="").replace('time.sleep(60)', 'wd.done = True; time.sleep(1.5)')
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=50)
    assert r.returncode == 0 and "not reached" in r.stdout and "watchdog" not in r.stderr
