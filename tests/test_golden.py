"""Committed golden fixture (tests/golden/allsky_golden.npz, made by tools/gen_golden_allsky.py from the CPU
oracle).  CPU: the oracle still reproduces it bit for bit.  GPU: the CUDA path matches it within the flux
tolerance without needing the oracle at run time."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import gen_golden_allsky as gg  # noqa: E402

GOLD = np.load(os.path.join(ROOT, "tests", "golden", "allsky_golden.npz"))


def test_oracle_reproduces_golden_bitwise(oracle_lib):
    got = gg.compute(oracle_lib)
    assert set(got) == set(GOLD.files)
    for k in GOLD.files:
        np.testing.assert_array_equal(got[k], GOLD[k], err_msg=k)


@pytest.mark.gpu
def test_cuda_matches_golden(cuda_lib):
    got = gg.compute(cuda_lib, "cuda:0")
    for k in GOLD.files:
        assert np.max(np.abs(got[k] - GOLD[k])) <= 1.0e-5, k  # W/m2, examples/compare-to-reference.py:58
