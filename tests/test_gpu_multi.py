"""Multi-GPU checks (need >= 2 GPUs on the box; skipped otherwise)."""
import json
import os
import socket
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _ngpu():
    from igdm_b200 import capi
    return capi.device_count()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("prec", ["fp32", "fp16"])
def test_data_parallel_matches_single_gpu(prec):
    n = _ngpu()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2
    env = dict(os.environ, DP_PREC=prec)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "dp_check.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    lines = [l for l in res.stdout.splitlines() if l.startswith("{")]
    assert lines, res.stdout[-2000:] + res.stderr[-2000:]
    v = json.loads(lines[-1])
    print(v)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", f"dp_check_{prec}.json"), "w") as fh:
        json.dump(v, fh, indent=1)
    assert v["ok"] and res.returncode == 0, v
