"""GPU checks of the nested-dissection plan of the reduced system (the default for large sparse blocks).

Each case runs tests/sparse_gpu_check.py in a process of its own (plan parameters come from the environment at handle
creation): plan form vs the dense form of the library vs an oracle, free and control-point networks.
  * a small block with the plan forced and cut deep (leaves of 12 images, supertiles of 2 blocks);
  * u_c = 6,610 (1,100 images) -- above the 6,144 unknowns where the dense form switches to its task graph -- with
    the library's own automatic choice and default parameters, against the C restatement of the reference
    algorithm: oracle parity for BOTH factorisations at the size class of the headline benchmark.
The host parts (ordering, pattern, owners, the group algorithm) are verified on the CPU in
tests/test_reduced_plan_host.py."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def check(*args, timeout=900):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "sparse_gpu_check.py"), *map(str, args)], cwd=ROOT,
                       capture_output=True, text=True, timeout=timeout)
    sys.stdout.write(r.stdout[-4000:])
    sys.stderr.write(r.stderr[-4000:])
    assert r.returncode == 0


@pytest.mark.gpu
def test_forced_plan_matches_dense_form_and_oracle():
    check(150, 12000, 12, 2, 1)


@pytest.mark.gpu
def test_automatic_plan_at_6610_unknowns_matches_dense_task_graph_and_c_oracle():
    check(1100, 110000, 96, 8, 0)


@pytest.mark.gpu
def test_group_of_two_gpus_equals_one_gpu():
    """feba_create_shard over two ranks (NCCL) against one GPU and the oracle: scripts/group_check.py."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29543", os.path.join(ROOT, "scripts", "group_check.py")]
    out = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)
    sys.stdout.write(out.stdout[-4000:])
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
