"""The library's host-only C++ -- native BFGS / L-BFGS-B loops (csrc/sfh_drivers.h), the multi-threaded NUTS around one batched
evaluation (csrc/sfh_nuts.h), the on-disk container (csrc/sfh_file.h) and the host side of the packet completion protocol
(csrc/sfh_packets.h: a thread stands in for the finalize kernel and delivers torn, reordered packets over stale ones) -- built by itself and run under AddressSanitizer +
UndefinedBehaviorSanitizer and under ThreadSanitizer (tests/native_host_sanitize.cpp; a CPU Poisson objective stands in for the
device evaluations).  The reference has no race detection of its own; its chains run on Julia threads
(hmc_sample.jl:123-141, generic_fitting.jl:617-626), here they are host threads parked on a condition variable, which is what
ThreadSanitizer watches.  No device needed."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "native_host_sanitize.cpp")
INC = os.path.join(ROOT, "starformationhistories.jl_b200", "csrc")

BUILDS = {
    "asan_ubsan": ["-fsanitize=address,undefined", "-fno-sanitize-recover=undefined", "-fno-omit-frame-pointer"],
    "tsan": ["-fsanitize=thread"],
}


@pytest.mark.timeout(600)
@pytest.mark.parametrize("name", list(BUILDS))
def test_host_native_code_under_sanitizers(name, tmp_path):
    exe = tmp_path / name
    cmd = ["g++", "-std=c++17", "-O1", "-g", "-Wall", "-Werror", *BUILDS[name], "-I", INC, SRC, "-o", str(exe), "-pthread"]
    b = subprocess.run(cmd, capture_output=True, text=True)
    if b.returncode != 0 and ("cannot find -l" in b.stderr or "unrecognized" in b.stderr):
        pytest.skip(f"this g++ has no {name} runtime: {b.stderr.strip().splitlines()[-1]}")
    assert b.returncode == 0, b.stderr
    env = dict(os.environ, ASAN_OPTIONS="detect_leaks=1:abort_on_error=0", TSAN_OPTIONS="halt_on_error=1",
               UBSAN_OPTIONS="print_stacktrace=1")
    r = subprocess.run([str(exe), str(tmp_path)], capture_output=True, text=True, env=env, timeout=540)
    report = r.stdout[-3000:] + r.stderr[-6000:]
    assert r.returncode == 0, report
    assert "all checks passed" in r.stdout, report
    for marker in ("ThreadSanitizer", "AddressSanitizer", "LeakSanitizer", "runtime error"):
        assert marker not in r.stderr, report
