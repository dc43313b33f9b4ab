import importlib

import numpy as np

from oracle import memory_port as mp


def test_product_and_oracle_generators_agree():
    synth = importlib.import_module("bsc-nav_b200.synth")
    for f in (0, 7):
        a, b = synth.synthetic_frame(f, 48, 64, n_frames=12, seed_base=3), mp.synthetic_frame(f, 48, 64, n_frames=12, seed_base=3)
        for x, y in zip(a, b):
            assert np.array_equal(x, y)
