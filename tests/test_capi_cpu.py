"""CPU checks of the boundary: the shared library loads, exports every symbol the public header declares, and
refuses to compute without a B200 (no CPU fallback, no route through the oracle)."""
import ctypes
import os
import re

import pytest

import depecher_b200 as dp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    lib = dp.load_library()
    names = dp.declared_symbols()
    assert len(names) >= 35
    for required in ("dpr_allocate_points", "dpr_sparse_k_means", "dpr_grid_search", "dpr_rand_index", "dpr_last_error"):
        assert required in names
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_header_has_no_framework_types():
    txt = open(os.path.join(ROOT, "include", "depecher_b200.h")).read()
    assert 'extern "C"' in txt
    assert not re.search(r"torch|at::|std::|Rcpp|Eigen::", re.sub(r"/\*.*?\*/", "", txt, flags=re.S))


def test_product_does_not_reference_the_oracle():
    for base, _dirs, files in os.walk(os.path.join(ROOT, "depecher_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")) or f == "Makefile":
                src = open(os.path.join(base, f), errors="ignore").read()
                assert "liboracle" not in src and "oracle." not in src.replace("the oracle.", ""), os.path.join(base, f)


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    eng = dp.Engine()
    import numpy as np
    with pytest.raises(dp.DeviceError):
        eng.allocate_points(np.zeros((10, 3)), np.ones((2, 3)), False)
    with pytest.raises(dp.DeviceError):
        eng.rand_index(np.zeros(5, int), np.zeros(5, int), 2)


def test_philox_header_is_shared_between_oracle_and_library():
    """one randomness contract: both sides include include/dpr_philox.h"""
    assert "dpr_philox.h" in open(os.path.join(ROOT, "oracle", "clusterer_oracle.cpp")).read()
    assert "dpr_philox.h" in open(os.path.join(ROOT, "depecher_b200", "csrc", "aux_kernels.cu")).read()


def test_version_and_error_text():
    lib = dp.load_library()
    assert lib.dpr_version() == 100
    lib.dpr_set_ari_mode.argtypes = [ctypes.c_int32]
    assert lib.dpr_set_ari_mode(7) == -1
    assert b"ARI mode" in lib.dpr_last_error()


def test_rcpp_layer_compiles_against_the_abi():
    """rpkg/src/InterfaceUtils.cpp (the Rcpp host layer a DepecheR maintainer drops in) is checked for syntax and against the
    declarations of include/depecher_b200.h; R and Rcpp are not in the image, so a stub of the few names it uses stands in
    (rpkg/compile_check/Rcpp.h).  Every [[Rcpp::export]] must have its wrapper in rpkg/R/RcppExports.R."""
    import re
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = os.path.join(root, "rpkg", "src", "InterfaceUtils.cpp")
    subprocess.run(["g++", "-std=c++11", "-fsyntax-only", "-Wall", "-Werror", "-I" + os.path.join(root, "rpkg", "compile_check"),
                    "-I" + os.path.join(root, "include"), src], check=True)
    text = open(src).read()
    exports = re.findall(r"// \[\[Rcpp::export\]\]\n[^\n(]*?(\w+)\(", text)
    assert {"sparse_k_means", "grid_search", "allocate_points", "rand_index"} <= set(exports)          # the reference's four
    assert {"dataset_create", "grid_search_each", "penalty_opt", "select_solution", "set_engine_seed"} <= set(exports)
    rside = open(os.path.join(root, "rpkg", "R", "RcppExports.R")).read()
    for name in exports:
        assert f"'_DepecheR_{name}'" in rside, name
    # the seed handed to the library varies from call to call (the unchanged R caller repeats seed_off_set = 1..nCores in every set)
    assert "next_seed(seed_off_set)" in text and "(uint64_t)seed_off_set," not in text
