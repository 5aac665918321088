"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol
include/jitr_b200.h declares; argument errors come back as codes, not crashes.  No GPU needed."""
import ctypes
import os
import re

import pytest

from jitr_b200 import _cabi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_loads_and_exports_header_symbols():
    hdr = open(os.path.join(ROOT, "include", "jitr_b200.h")).read()
    declared = set(re.findall(r"\b(jitr_b200_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    L = _cabi.lib()
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/jitr_b200.h but not exported"
    assert declared == set(_cabi.SIGNATURES), "ctypes SIGNATURES out of sync with the header"
    assert L.jitr_b200_version() >= 100


def test_bad_arguments_return_codes_without_device():
    L = _cabi.lib()
    rc = L.jitr_b200_elastic_sweep_params(40, 30, 1, 1, 1, None, None, None, None, None, None, 40, 1e-6, None, None, None, None)
    assert rc == _cabi.ERR_BAD_ARG
    assert b"null" in L.jitr_b200_last_error()
    with pytest.raises(ValueError):
        _cabi.check(rc, "sweep")
    rc = L.jitr_b200_solve_batched(40, 99, 1, None, 0, None, None, None, None, None, 0, None, None, 0, None, None, None, None, None, None, None, None)
    assert rc == _cabi.ERR_BAD_ARG


def test_no_cpu_fallback():
    """Without a CUDA device the product path must fail loudly."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import numpy as np

    import jitr_b200
    from jitr_b200.reactions import ElasticReaction

    rx = ElasticReaction((40, 20), (1, 0), mass_target=37214.7)
    ws = jitr_b200.IntegralWorkspace(rx, rx.kinematics(10.0), 10.0, jitr_b200.Solver(20), 5)
    with pytest.raises(_cabi.JitrB200Error):
        ws.smatrix(np.zeros(20, dtype=complex))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "jitr_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, fn)).read()
                assert "oracle" not in txt.replace("0+oracle", ""), f"{fn} mentions the oracle"
