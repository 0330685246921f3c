"""rm_pool.h -- the library's host threads (delivery of frames: clearing black tiles, scattering busy ones) -- under
ThreadSanitizer: every item of every job runs exactly once whatever the thread limit, back to back and after the workers
have gone to sleep, and the tool sees no data race in the pool's hand-over of jobs."""
import os
import subprocess

from tests.conftest import ROOT


def test_host_pool_under_thread_sanitizer(tmp_path):
    src = os.path.join(ROOT, "tests", "native", "pool_check.cpp")
    exe = str(tmp_path / "pool_check")
    subprocess.run(["g++", "-std=c++17", "-O1", "-g", "-fsanitize=thread", "-pthread", src, "-o", exe], check=True)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300, env=dict(os.environ, TSAN_OPTIONS="halt_on_error=1"))
    assert r.returncode == 0, r.stdout + r.stderr
    assert "pool ok" in r.stdout and "WARNING: ThreadSanitizer" not in r.stderr
