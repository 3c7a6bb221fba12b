"""bench.py --impl reference needs no GPU: the contract of its JSON line is checked here (one line on stdout; the base keys;
`impl`, `cpu_baseline` and a zero-copy `e2e`; rank != 0 prints nothing and exits 0).  The arm times the oracle's two CPU
restatements of the reference's fg! -- the BLAS route Julia's mul! takes and the OpenMP loop nest -- and reports the faster."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra):
    env = dict(os.environ, **env_extra)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "3"],
                          capture_output=True, text=True, cwd=ROOT, env=env, timeout=600)


@pytest.mark.timeout(900)
def test_reference_arm_json_line():
    r = _run({})
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["unit"] == "evals/s" and d["n_gpus"] == 1
    assert "loglikelihood+gradient" in d["metric"] and "loglikelihood+gradient" in base["metric"]
    assert d["steps"] == 2 and d["warmup"] == 3 and d["value"] > 0 and d["ms_per_step"] == pytest.approx(1e3 / d["value"], rel=1e-6)
    assert d["vs_baseline"] is None and d["data"] == "synthetic" and d["dtype"] == "f64" and "workload" in d["config"]
    c = d["cpu_baseline"]
    assert c["kind"] == "port" and c["route"] in ("blas", "openmp") and c["value"] == d["value"] and c["unit"] == d["unit"]
    assert 1 <= c["cores"] <= (os.cpu_count() or 1) and "60000x2400" in c["sample"] and "slower restatement" in c["sample"]
    one = c["blas_one_thread"]                 # the reference benchmark suite's own setting (benchmark/benchmarks.jl:5)
    assert one is None or 0 < one <= 1.5 * d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_quietly():
    r = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""


@pytest.mark.timeout(900)
def test_reference_arm_ignores_the_launchers_thread_cap():
    """torchrun exports OMP_NUM_THREADS=1 to its workers; round 1's N > 1 reference arm therefore ran on ONE core and the
    driver's vs_reference ratios at N = 2, 4, 8 were void.  The arm now sets its own thread counts from the affinity mask."""
    ncores = len(os.sched_getaffinity(0))
    if ncores < 2:
        pytest.skip("one core: nothing to cap")
    r = _run({"OMP_NUM_THREADS": "1", "OPENBLAS_NUM_THREADS": "1", "RANK": "0", "WORLD_SIZE": "2", "LOCAL_RANK": "0"})
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads([ln for ln in r.stdout.splitlines() if ln.strip()][0])
    assert d["cpu_baseline"]["cores"] == ncores, d["cpu_baseline"]
