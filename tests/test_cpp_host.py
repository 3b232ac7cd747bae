"""The C++ host layer (parelagmc_b200/host): the reference's class interfaces over the C ABI and reference-style
drivers.  CPU: it builds and links.  GPU: the drivers reproduce the reference's ctest output (DarcyDeterministicTest)
and the C++ MLMC_Manager agrees with the Python twin."""
import os
import re
import subprocess

import numpy as np
import pytest

from common import hex_problem, make_context
from parelagmc_b200 import hierarchy as H

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "parelagmc_b200", "lib")


def _dump(tmp_path, n, nl):
    p = hex_problem(n, nl)
    path = str(tmp_path / f"hex{n}_{nl}.pmch")
    H.dump_problem(path, p["sampler"], p["darcy"], 3, 0.1)
    return p, path


def test_host_layer_builds_and_roundtrips_dump(tmp_path):
    from parelagmc_b200 import capi
    capi.build()
    for f in ("libparelagmc_b200_host.so", "MLMC.exe", "SLMC.exe", "DarcyTest.exe", "PDESamplerTest.exe", "LikelihoodExample.exe",
              "RatioEstimator_MLMC.exe"):
        assert os.path.exists(os.path.join(LIB, f))
    p, path = _dump(tmp_path, 4, 2)
    assert os.path.getsize(path) > 1000
    # without a GPU the driver must fail loudly through the reference's error convention (message, exit 0)
    import torch
    if not torch.cuda.is_available():
        out = subprocess.run([os.path.join(LIB, "DarcyTest.exe"), "--hierarchy", path], capture_output=True, text=True)
        assert out.returncode == 0 and "no CPU fallback" in out.stdout


def test_reference_signatures_compile_against_stub_headers():
    """The -DPARELAGMC_B200_WITH_PARELAG branch of the host layer -- the reference's verbatim constructors
    (`PDESampler(const std::shared_ptr<mfem::ParMesh>&, NormalDistributionSampler&, parelag::ParameterList&)`,
    `DarcySolver(const std::shared_ptr<mfem::ParMesh>&, parelag::ParameterList&)`, `NormalDistributionSampler(double,
    double)`), the set-up calls, all MLSampler / PhysicalMLSolver virtuals, the ParELAG -> plain-array extraction and a
    driver with the call sequence of the reference's MLMC.cpp -- compiles (g++ -fsyntax-only) against declaration-only
    MFEM / ParELAG / MPI headers (parelagmc_b200/host/stubs); the real libraries are not in this image."""
    out = subprocess.run(["make", "-C", os.path.join(ROOT, "parelagmc_b200", "host"), "parelag-syntax"], capture_output=True,
                         text=True)
    assert out.returncode == 0 and "parelag-syntax: ok" in out.stdout, out.stdout + out.stderr


@pytest.mark.gpu
def test_darcy_deterministic_ctest_regex(tmp_path):
    """PASS_REGULAR_EXPRESSION of DarcyDeterministicTest (/root/reference/examples/CMakeLists.txt:62-66)."""
    p, path = _dump(tmp_path, 16, 3)
    out = subprocess.run([os.path.join(LIB, "DarcyTest.exe"), "--hierarchy", path, "--rel-tol", "1e-10"],
                         capture_output=True, text=True, timeout=600).stdout
    assert re.search(r"0  2         17152", out), out
    assert re.search(r"1  2         2240", out), out
    assert re.search(r"2  2         304", out), out


@pytest.mark.gpu
def test_cpp_mlmc_manager_matches_python_twin(tmp_path):
    from parelagmc_b200 import managers as MG
    p, path = _dump(tmp_path, 8, 3)
    log = str(tmp_path / "MLMC.dat")
    out = subprocess.run([os.path.join(LIB, "MLMC.exe"), "--hierarchy", path, "--samples", "6,12,20", "--mse", "1e6",
                          "--rel-tol", "1e-10", "--dof-cost", "--log", log], capture_output=True, text=True,
                         timeout=600).stdout
    assert "FINAL MLMC ERRORS" in out                    # first alternative of the MLMC_PDESampler ctest regex
    est = float(re.findall(r"Estimate\s+([-0-9.e+]+)", out)[-1])
    c = make_context(p, rel=1e-10)
    try:
        m = MG.MLMC_Manager(None, 3, c, {"Use array samples": True, "Array number of samples": [6, 12, 20],
                                         "Mean square error": 1e6, "Output filename for MC managers": ""}, out=None)
        m.wallTime = False
        m.Run()
    finally:
        c.close()
    assert est == pytest.approx(float(np.sum(m.eY)), rel=1e-6)
    rows = [l.split() for l in open(log).read().splitlines()[1:]]
    assert len(rows) == 38 and [int(r[0]) for r in rows[:20]] == [2] * 20


@pytest.mark.gpu
def test_cpp_slmc_driver(tmp_path):
    p, path = _dump(tmp_path, 4, 2)
    out = subprocess.run([os.path.join(LIB, "SLMC.exe"), "--hierarchy", path, "--nsamples", "16", "--mse", "1e6",
                          "--log", str(tmp_path / "SLMC.dat")], capture_output=True, text=True, timeout=600).stdout
    assert "FINAL SLMC ERRORS" in out and "SLMC Manager Errors:" in out


@pytest.mark.gpu
def test_cpp_drivers_reproduce_reference_ctest_regexes(tmp_path):
    """The reference's ctest PASS_REGULAR_EXPRESSIONs (/root/reference/examples/CMakeLists.txt:83-115) applied to the
    output of the C++ drivers of the host layer (per-sample Sample / Eval / SolveFwd_RtnPressure through the reference's
    class interfaces and BayesianInverseProblem), on the reference's default problems in MFEM's element numbering."""
    from common import reference_enlarged_problem, reference_sampler_problem
    p = reference_sampler_problem()
    path = str(tmp_path / "sampler.pmch")
    H.dump_problem(path, p["sampler"], p["darcy"], 3, 0.1)
    out = subprocess.run([os.path.join(LIB, "PDESamplerTest.exe"), "--hierarchy", path], capture_output=True, text=True,
                         timeout=600).stdout
    for rx in (r"1.2593[0-9]*", r"9.3103[0-9]*", r"6.3853[0-9]*"):        # PDESamplerTest (:83-87)
        assert re.search(rx, out), out
    q = reference_enlarged_problem()
    path = str(tmp_path / "bayes.pmch")
    H.dump_problem(path, q["sampler"], q["darcy"], 3, 0.1, gobs=q["gobs"])
    out = subprocess.run([os.path.join(LIB, "LikelihoodExample.exe"), "--hierarchy", path], capture_output=True, text=True,
                         timeout=600).stdout
    for rx in (r"L = 0 : 0.9279[0-9]*", r"L = 1 : 0.9578[0-9]*", r"L = 2 : 0.9269[0-9]*"):     # LikelihoodEvaluation (:97-102)
        assert re.search(rx, out), out
    out = subprocess.run([os.path.join(LIB, "LikelihoodExample.exe"), "--hierarchy", path, "--ratio-mc"], capture_output=True,
                         text=True, timeout=600).stdout
    # BayesianInverseProblem_MC_RatioEstimator (:110-115)
    assert re.search(r"0 [ ]*  1.987[0-9 ]*  0.07749[0-9 ]* 0.8569[0-9 ]*  0.009691[0-9 ]* 2.319[0-9 ]*  2.332[0-9 ]*", out), out


@pytest.mark.gpu
def test_cpp_bayes_ratio_manager(tmp_path):
    """C++ ML_BayesRatio_Manager: the batched level loops (pmc_bayes_level_batch) against the reference's per-sample loop
    through BayesianInverseProblem (same stream positions), and against the Python twin."""
    from common import bayes_problem
    from parelagmc_b200 import managers as MG
    p = bayes_problem()
    path = str(tmp_path / "bayes8.pmch")
    H.dump_problem(path, p["sampler"], p["darcy"], 3, 0.1, gobs=p["gobs"])
    args = [os.path.join(LIB, "RatioEstimator_MLMC.exe"), "--hierarchy", path, "--samples", "5,9", "--mse", "1e6", "--rel-tol",
            "1e-10", "--dof-cost"]
    outs = [subprocess.run(args + extra, capture_output=True, text=True, timeout=600).stdout for extra in ([], ["--per-sample"])]
    est = [float(re.findall(r"Ratio Estimate\s+([-0-9.e+]+)", o)[-1]) for o in outs]
    assert "FINAL ML_BayesRatio_Manager ERRORS" in outs[0]
    assert est[0] == pytest.approx(est[1], rel=1e-6)
    ey = [[float(x) for x in re.findall(r"E\[Y_R\]\s+(.*)", o)[-1].split()] for o in outs]
    assert np.allclose(ey[0], ey[1], rtol=1e-6)
    assert 1.0 < est[0] < 4.0
