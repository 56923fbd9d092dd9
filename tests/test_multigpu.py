"""Multi-rank tests.  CPU: world_size-2 gloo run of the host-side slab logic.  GPU (needs >= 2 devices):
N-GPU == 1-GPU bitwise for the unfused and fused solvers."""
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent


def _torchrun(script, nproc, *args, timeout=600):
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
           "--master-port", "29517", str(script), *map(str, args)]
    return subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env, cwd=str(ROOT))


def test_slab_partition_logic_gloo_world2():
    """Two CPU ranks (gloo): slabs built by slab_of() tile the global case exactly, halos hold the periodic
    neighbours' rows, and the halo width follows 2K+3 (se.jl:55-56)."""
    r = _torchrun(ROOT / "tests" / "gloo_slab_check.py", 2)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "GLOO_SLABS_OK" in r.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("solver,K,topo", [("unfused", 3, "periodic"), ("auto", 3, "periodic"), ("auto", 10, "periodic"),
                                            ("unfused", 3, "bounded_y"), ("auto", 4, "bounded_y"), ("auto", 3, "arctic"), ("unfused", 3, "arctic"),
                                            ("fused", 3, "coastline"), ("unfused", 3, "coastline"), ("auto", 3, "curvilinear"), ("unfused", 3, "curvilinear"),
                                            ("auto", 3, "folded"), ("unfused", 1, "folded")])
def test_two_gpus_equal_one_gpu(solver, K, topo):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    r = _torchrun(ROOT / "tests" / "multigpu_check.py", 2, solver, K, 2, topo)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "MULTIGPU_OK" in r.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("solver,K,topo", [("unfused", 3, "periodic"), ("auto", 3, "periodic"), ("auto", 4, "bounded_x"), ("unfused", 3, "bounded_x"),
                                            ("fused", 3, "coastline"), ("unfused", 3, "coastline")])
def test_two_gpus_split_along_x_equal_one_gpu(solver, K, topo):
    """Partition(2, 1): packed west/east strips instead of zero-copy rows (test/distributed_tests_utils.jl:60-62 runs (4,1))."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    r = _torchrun(ROOT / "tests" / "multigpu_check.py", 2, solver, K, 2, topo, 2)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "MULTIGPU_OK" in r.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("solver,K,topo", [("auto", 3, "periodic"), ("unfused", 3, "periodic"), ("auto", 3, "bounded_x"), ("unfused", 3, "bounded_x"),
                                            ("fused", 3, "coastline"), ("unfused", 3, "coastline")])
def test_four_gpus_2x2_equal_one_gpu(solver, K, topo):
    """Partition(2, 2): strips, rows and the corners the rows carry (test/distributed_tests_utils.jl:60-62 runs (2,2))."""
    if torch.cuda.device_count() < 4:
        pytest.skip("needs 4 GPUs")
    r = _torchrun(ROOT / "tests" / "multigpu_check.py", 4, solver, K, 2, topo, 2)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "MULTIGPU_OK" in r.stdout
