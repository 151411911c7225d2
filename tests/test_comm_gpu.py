"""Multi-GPU result exchange behind the C-ABI (gb200_comm_*, gb200_ivfpq_search_sharded): one process per GPU, every
rank searches its shard of the batch and pushes its top-k into every peer's window over NVLink; the gathered result on
every rank must equal the single-GPU search of the whole batch.  With >= 2 GPUs every rank gets its own device (run with
`gpurun --gpus 2`); on a one-GPU box the two ranks share device 0 — the windows are still mapped through CUDA IPC and the
flag protocol is the same, the two exchange kernels just take turns on the GPU."""
import json
import os
import subprocess
import sys
import tempfile

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sharded_search_gathers_on_every_rank():
    from gamma_b200 import api
    ndev = api.lib().gb200_device_count()
    world = max(2, min(ndev, 4))
    tmp = tempfile.mkdtemp(prefix="gb200_comm_")
    procs = [subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "comm_worker.py"), str(r), str(world), tmp, str(r % ndev)],
                              stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True) for r in range(world)]
    outs = [p.communicate(timeout=600) for p in procs]
    for r, (p, (so, se)) in enumerate(zip(procs, outs)):
        lines = [json.loads(l) for l in so.splitlines() if l.startswith("{")]
        assert p.returncode == 0 and lines and lines[-1]["ok"] and lines[-1]["status"] == 0, (r, so[-2000:], se[-2000:])
