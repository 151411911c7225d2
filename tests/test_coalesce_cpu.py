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


def test_exchange_protocol_model():
    """The buffer / flag protocol of the multi-GPU exchange (gamma_b200/csrc/comm.cu header: four window buffers indexed by
    epoch, per-peer epoch flags, immediate or deferred wait) played by host threads with plain-memory payloads under
    ThreadSanitizer: with four buffers no window is overwritten before its consumer has read it, whatever mix of wait
    forms the ranks use; with three buffers and the deferred wait the overwrite the header describes is reported."""
    gxx = shutil.which("g++")
    tmp = tempfile.mkdtemp(prefix="gb200_xsim_")
    exe = os.path.join(tmp, "xsim")
    subprocess.run([gxx, "-std=c++17", "-O1", "-g", "-fsanitize=thread", "-pthread",
                    os.path.join(ROOT, "tests", "exchange_protocol_sim.cc"), "-o", exe], check=True, cwd=ROOT)
    env = dict(os.environ, TSAN_OPTIONS="halt_on_error=1 exitcode=66")

    def run(ranks, epochs, nbuf, mode):
        return subprocess.run([exe, str(ranks), str(epochs), str(nbuf), str(mode)], capture_output=True, text=True,
                              timeout=600, env=env)

    for mode in (0, 1, 2):
        r = run(4, 3000, 4, mode)
        assert r.returncode == 0, (mode, r.stdout[-500:], r.stderr[-3000:])
        assert json.loads(r.stdout.strip().splitlines()[-1])["bad"] == 0
    r = run(2, 3000, 2, 0)  # the immediate form alone needs only two buffers
    assert r.returncode == 0, (r.stdout[-500:], r.stderr[-3000:])
    # the model has teeth: three buffers are not enough for the deferred wait
    assert any(run(4, 3000, 3, 1).returncode != 0 for _ in range(3))
