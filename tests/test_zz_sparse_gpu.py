"""GPU check of the opt-in block-sparse form of the reduced system (FEBA_SPARSE=1).

Runs tests/sparse_gpu_check.py in a process of its own (named test_zz_*: after every other GPU test).  The device
side of this path was written after this round's GPU budget was spent -- its host parts (supertile pattern, symbolic
fill, datum images, the 14x14 border) are verified on the CPU in tests/test_sparse_reduced_host.py, the kernels
and their wiring have not met a GPU yet -- hence xfail(strict=False): a pass shows as XPASS, a failure does not
turn the suite red for a path nothing uses by default."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
@pytest.mark.xfail(reason="opt-in FEBA_SPARSE=1 path: first GPU contact pending (round-1 GPU budget spent)", strict=False)
def test_sparse_form_matches_dense_form_and_oracle():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "sparse_gpu_check.py")], cwd=ROOT,
                       capture_output=True, text=True, timeout=420)
    sys.stdout.write(r.stdout[-4000:])
    sys.stderr.write(r.stderr[-4000:])
    assert r.returncode == 0
