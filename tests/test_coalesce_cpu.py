"""Host-side batching of concurrent searches (gamma_b200/csrc/coalesce.h, used by gb200_ivfpq_search): the policy class is
CUDA-free, so it is stress-tested here with a fake device under ThreadSanitizer — every request executed exactly once by a
batch of compatible requests, results and errors delivered to the right caller, never more batches in flight than the
policy allows, and no data race (the reference runs one Search per request thread: tests/test.h:1033-1062)."""
import json
import os
import shutil
import subprocess
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("sanitizer", ["thread", "address"])
def test_coalescer_stress_under_sanitizer(sanitizer):
    gxx = shutil.which("g++")
    assert gxx, "g++ is part of the image"
    tmp = tempfile.mkdtemp(prefix="gb200_coalesce_")
    exe = os.path.join(tmp, "coalesce_stress")
    subprocess.run([gxx, "-std=c++17", "-O1", "-g", "-fsanitize=" + sanitizer, "-pthread",
                    os.path.join(ROOT, "tests", "coalesce_stress.cc"), "-o", exe], check=True, cwd=ROOT)
    env = dict(os.environ, TSAN_OPTIONS="halt_on_error=1 exitcode=66", ASAN_OPTIONS="detect_stack_use_after_return=1")
    r = subprocess.run([exe, "12", "300"], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, (r.returncode, r.stdout[-2000:], r.stderr[-4000:])
    cases = json.loads(r.stdout.strip().splitlines()[-1])
    assert len(cases) == 16
    for c in cases:
        assert c["bad"] == 0 and c["requests"] == 12 * 300 and c["max_running"] <= c["slots"], c
    # with one slot and many callers the requests do travel together
    assert any(c["slots"] == 1 and c["batches"] < c["requests"] for c in cases)
