"""Multi-GPU parity (needs >= 2 B200s: `gpurun --gpus 2 -- python -m pytest tests/test_multigpu.py -m gpu`).
Fixed-mesh runs must be bit-identical to the undivided oracle.  Moving-mesh runs are bit-identical too when the ownership
is chunk-aligned (cfd_b200/partition.py: every rank owns >= 2 reduction chunks of 4096 nodes; cases *_aligned, 6 steps,
mesh solve included).  The tiny `ale` case (672 nodes) cannot be aligned: its inner products are summed per rank and then
across ranks, and this impulsively started case amplifies a one-ulp difference to percent level within five steps even
inside the oracle (measured: dU 2e-9 at step 4, 4e-2 at step 5), so it is compared after two steps to round-off."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    from cfd_b200 import capi

    return capi.lib().cfdb_device_count()


@pytest.mark.parametrize("case", ["square_visc", "channel_itlocal", "ale", "ale_aligned", "ale_visc_aligned", "square_aligned"])
def test_two_ranks_match_undivided_oracle(case):
    n = _ngpu()
    if n < 2:
        pytest.skip("needs 2 GPUs")
    world = 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29611", os.path.join(ROOT, "tests", "mg_worker.py"), case, "2" if case == "ale" else "6"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=180)
    assert r.returncode == 0 and "MULTIGPU_OK" in r.stdout, r.stdout[-4000:]
