"""Distributional network restatement (oracle/distnet_oracle.c) vs the reference's own torch module (golden vector)."""
import os

import numpy as np


def test_distnet_oracle_matches_reference_golden(oracle):
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "distnet_golden.npz"))
    got = oracle.distnet_forward(oracle.seeded_dist_weights(int(z["seed"])), z["states"])
    assert np.allclose(got, z["dist"], rtol=1e-5, atol=1e-7)


def test_product_and_oracle_generate_the_same_seeded_weights(oracle):
    from tetris_mcts_b200.agents.DistValueSimOnline import init_dist_weights
    from tetris_mcts_b200.model.model_vv import init_weights
    assert np.array_equal(init_dist_weights(2), oracle.seeded_dist_weights(2))
    assert np.array_equal(init_weights(2), oracle.seeded_weights(2))
