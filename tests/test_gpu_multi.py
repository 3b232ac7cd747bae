"""Multi-GPU path behind the C ABI (`pmc_comm_init` / `pmc_allreduce_sums`, NCCL over NVLink) and the C++ managers
sharded over ranks.  NCCL refuses two ranks on one device, so these tests need two GPUs (`gpurun --gpus 2`); with one
they skip.  The CPU twin of the sharding logic is tests/test_host_logic.py (gloo, world size 2)."""
import os
import re
import subprocess
import threading

import numpy as np
import pytest

from common import hex_problem, make_context
from parelagmc_b200 import hierarchy as H

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "parelagmc_b200", "lib")


def _ngpu():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.gpu
def test_allreduce_sums_world2_through_c_abi():
    """One process, one handle and host thread per GPU (the model include/pmc_b200.h describes): each rank runs its slice
    of a level batch, the per-level sums are combined by pmc_allreduce_sums, and the result equals the 1-rank sums."""
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    from parelagmc_b200 import capi
    from parelagmc_b200.managers import split_samples
    p = hex_problem(8, 2)
    n, lev = 24, 0
    Ne = p["sampler"][lev].Ne
    ref_ctx = make_context(p, rel=1e-10, device=0)
    ref, _, _ = ref_ctx.mlmc_level_batch(lev, n, 100)
    ref_ctx.close()
    uid = capi.comm_unique_id()
    out, errs = [None, None], []

    def rank_main(r):
        try:
            c = make_context(p, rel=1e-10, device=r)
            c.comm_init(2, r, uid)
            first, count = split_samples(n, r, 2)
            local = np.zeros(9)
            c.mlmc_level_batch(lev, count, 100 + first * Ne, sums=local)
            out[r] = (local.copy(), c.allreduce_sums(local).copy())
            c.comm_destroy()
            c.close()
        except Exception as e:      # surfaced in the main thread
            errs.append(e)

    th = [threading.Thread(target=rank_main, args=(r,)) for r in range(2)]
    for t in th:
        t.start()
    for t in th:
        t.join(timeout=600)
    assert not errs, errs
    assert np.array_equal(out[0][1], out[1][1])                       # every rank holds the same reduced sums
    assert np.allclose(out[0][1], out[0][0] + out[1][0], rtol=1e-15)  # = the sum of the two contributions
    assert np.allclose(out[0][1], ref, rtol=1e-9)                     # = the 1-rank sums (another summation order)
    assert out[0][0][6] + out[1][0][6] == ref[6]                      # the cost column is an exact integer sum


@pytest.mark.gpu
def test_cpp_mlmc_driver_two_ranks(tmp_path):
    """MLMC.exe as two processes (one per GPU, PMC_RANK / PMC_WORLD_SIZE, NCCL id through a file): rank 0's table equals
    the single-process run's."""
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    p = hex_problem(8, 3)
    path = str(tmp_path / "hex8_3.pmch")
    H.dump_problem(path, p["sampler"], p["darcy"], 3, 0.1)
    args = [os.path.join(LIB, "MLMC.exe"), "--hierarchy", path, "--samples", "7,13,21", "--mse", "1e6", "--rel-tol",
            "1e-10", "--dof-cost", "--log", ""]
    single = subprocess.run(args, capture_output=True, text=True, timeout=600,
                            env={**os.environ, "PMC_WORLD_SIZE": "1", "PMC_RANK": "0"}).stdout
    procs = []
    for r in range(2):
        env = {**os.environ, "PMC_WORLD_SIZE": "2", "PMC_RANK": str(r), "PMC_ID_FILE": str(tmp_path / "nccl_id")}
        procs.append(subprocess.Popen(args, stdout=subprocess.PIPE, text=True, env=env))
    outs = [pr.communicate(timeout=600)[0] for pr in procs]
    assert "FINAL MLMC ERRORS" in outs[0] and "FINAL MLMC ERRORS" not in outs[1]      # only rank 0 prints
    est1 = float(re.findall(r"Estimate\s+([-0-9.e+]+)", single)[-1])
    est2 = float(re.findall(r"Estimate\s+([-0-9.e+]+)", outs[0])[-1])
    assert est2 == pytest.approx(est1, rel=1e-7)
    n1 = re.findall(r"NumSamples\s+(.*)", single)[-1].split()
    n2 = re.findall(r"NumSamples\s+(.*)", outs[0])[-1].split()
    assert n1 == n2 == ["7", "13", "21"]
