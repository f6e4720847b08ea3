"""GPU suite, more than one GPU (skipped on a single-GPU box; run with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`):
the row-band paths with a REAL neighbour on another device -- per-level NCCL halo exchange, halos read over NVLink
peer memory inside the blur kernel (CUDA IPC, ld.acquire.sys / st.release.sys counters, device-side build epochs, CUDA
graph replay) -- one process per GPU under torch.distributed.run, as bench.py launches them."""
from __future__ import annotations

import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _gpus() -> int:
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize("world", [2, 4, 8])
def test_row_bands_across_gpus(world):
    """scripts/check_conv_bands_nccl.py --c4 on `world` GPUs: CONV bands vs the specification (NCCL and peer-memory halos),
    REF bands bit-exact, then C4 band geometry with three slots in flight over four rounds, bit-identical to the unbanded
    build and free of time-outs.  Exit code 0 only if every rank passed."""
    if _gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29640 + world), os.path.join(ROOT, "scripts", "check_conv_bands_nccl.py"), "--c4"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert out.stdout.count("PASS (all ranks: PASS)") == world, out.stdout[-3000:]
