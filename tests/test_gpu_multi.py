"""N > 1 paths on real GPUs (skipped on a single-GPU box): torchrun-spawned verification of the partitioned multi-GPU Q3 plan and of
the fused partition + peer exchange + join (PartitionedHashJoin) against the oracle."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    from datafusion_b200 import capi
    return capi.load_library().dfgpu_device_count()


def _torchrun(script, n, *args, timeout=900):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1", "--master-port", "29541",
           os.path.join(ROOT, "scripts", script), *args]
    return subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=timeout)


@pytest.mark.parametrize("script,args,marker", [("verify_q3_multi_gpu.py", ["0.5"], "VERIFY_Q3_MULTI_GPU OK"), ("verify_peer_exchange.py", [], "exchange_identical=True")])
def test_two_rank_paths_match_the_oracle(script, args, marker):
    if _gpus() < 2:
        pytest.skip("needs >= 2 GPUs on the box")
    r = _torchrun(script, 2, *args)
    assert r.returncode == 0 and marker in r.stdout and "=False" not in r.stdout, (r.stdout[-2000:], r.stderr[-2000:])


def test_library_communicator_and_exchange_without_nccl():
    """dfgpu_comm / dfgpu_exchange: rendezvous over POSIX shared memory, rows over CUDA IPC peer stores — no torch.distributed at all"""
    if _gpus() < 2:
        pytest.skip("needs >= 2 GPUs on the box")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "verify_comm_exchange.py"), "2"], cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "VERIFY_COMM_EXCHANGE OK" in r.stdout, (r.stdout[-2000:], r.stderr[-2000:])
