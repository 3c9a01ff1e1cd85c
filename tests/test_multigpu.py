"""Multi-GPU parity on real devices (needs >= 2 GPUs: `gpurun --gpus 2 -- python -m pytest tests -m gpu`):
the same worker as tests/test_multirank.py with the PRODUCT library on one B200 per rank and the
halo exchange / reductions over NCCL. Compared with the single-partition oracle."""
import os

import pytest

from test_multirank import check, run_world


def _n_gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.gpu
@pytest.mark.parametrize("case,solver,merge", [("hex_slabs", "amg", 64), ("tet_rcb", "bcgstab", 200),
                                                ("tet_rcb", "model", 200), ("hex_slabs", "flow", 200),
                                                ("tet_rcb", "electric", 200)])
def test_two_gpus_match_single_partition_oracle(case, solver, merge):
    if _n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    os.environ["FVM_WORKER_GPU"] = "1"
    try:
        res = run_world(2, case, solver, merge)
    finally:
        del os.environ["FVM_WORKER_GPU"]
    check(res)
