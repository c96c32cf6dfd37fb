"""CPU: pin the re-ranking oracle to the real reference's outputs (tests/golden/rerank_*.npz, oracle/make_golden.py)."""
import os

import numpy as np
import pytest

from oracle import rerank_oracle as ro


@pytest.mark.parametrize("name", ["rerank_k20", "rerank_k6", "rerank_k5_noqe"])
def test_rerank_oracle_matches_reference_golden(golden_dir, name):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    final = ro.re_ranking(g["q_g"], g["q_q"], g["g_g"], int(g["k1"]), int(g["k2"]), float(g["lam"]))
    assert final.dtype == np.float32 and final.shape == g["final"].shape
    assert np.array_equal(final, g["final"])            # bit for bit (same float32 operation order)


def test_rerank_oracle_stage_properties(golden_dir):
    g = np.load(os.path.join(golden_dir, "rerank_k20.npz"))
    final, st = ro.re_ranking(g["q_g"], g["q_q"], g["g_g"], 20, 6, 0.3, return_stages=True)
    V = st["V"]
    assert np.allclose(V.sum(1), 1.0, atol=1e-5)        # every expanded row is a mean of unit-sum rows
    assert st["rank"].shape[1] == 21
    lo, hi = 0.3 * st["O"][:48, 48:], 0.7 + 0.3 * st["O"][:48, 48:]
    assert (final >= lo - 1e-6).all() and (final <= hi + 1e-6).all()   # Jaccard distance lies in [0, 1]
