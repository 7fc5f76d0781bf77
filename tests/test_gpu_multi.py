"""Multi-GPU parity (needs >= 2 GPUs on the box; skipped otherwise): targets sharded over the ranks, packed leaf source
records all-gathered over NCCL after every stage, results equal to the CPU oracle on every rank."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_gpus():
    try:
        import torch
        return torch.cuda.device_count() if torch.cuda.is_available() else 0
    except Exception:
        return 0


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 4])
def test_sharded_steppers_equal_oracle(world):
    if _n_gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr",
           "127.0.0.1", "--master-port", str(29500 + 7 * world), os.path.join(ROOT, "tests", "multi_gpu_check.py")]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    # keep the per-rank lines where a GPU visit picks them up (gpurun_out/ is merged back; copied to profiles/ from there)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", f"multi_gpu_check_n{world}.log"), "w") as f:
        f.write(p.stdout + ("\n--- stderr tail ---\n" + p.stderr[-2000:] if p.returncode else ""))
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    assert p.stdout.count("swe_rk2 cubed-3: ok") == world
    assert p.stdout.count("sharded host I/O") == world and "FAILED" not in p.stdout


@pytest.mark.gpu
def test_peer_exchange_equals_nccl_exchange():
    """lpmx_peer.cu (default for world <= 8; LPMX_PEER_EXCHANGE=0 turns it off): the one-kernel exchange over NVLink peer memory must leave the
    steppers' results bit-identical to the NCCL exchange.  Needs >= 2 GPUs."""
    if _n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    env = dict(os.environ, LPMX_PEER_TIMEOUT_S="10")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
           "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tools", "peer_exchange_check.py"), "--time", ""]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "peer_exchange_check_n2.log"), "w") as f:
        f.write(p.stdout + ("\n--- stderr tail ---\n" + p.stderr[-2000:] if p.returncode else ""))
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    assert p.stdout.count("peer == nccl bitwise: True") == 4
