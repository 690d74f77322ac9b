"""Fit the prob/dist head weights used by bench.py (and smoke tests) so that a seeded
Glorot-uniform U-Net body produces StarDist-like maps on synthetic cell images:
  python tests/golden/make_bench_heads.py   ->  tests/golden/bench_heads_2d.npz
See oracle/synth.py::calibrated_weights.  The body weights are NOT stored: they are regenerated
from the seed (stardist_b200.models.weights.glorot_uniform_weights(config, seed=0))."""
import os, sys
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import synth
from stardist_b200.models.config import Config2D, Config3D
from stardist_b200.rays3d import Rays_GoldenSpiral

if __name__ == "__main__":
    cfg = Config2D(n_rays=32)
    w = synth.calibrated_weights(cfg, seed=0)
    if "--only-3d" not in sys.argv: np.savez(os.path.join(HERE, "bench_heads_2d.npz"), prob_kernel=w['prob'][0], prob_bias=w['prob'][1],
             dist_kernel=w['dist'][0], dist_bias=w['dist'][1])
    print("saved", w['prob'][0].shape, w['dist'][0].shape)
    # configs[2]: default Config3D, Rays_GoldenSpiral(96)  ->  tests/golden/bench_heads_3d.npz
    cfg3 = Config3D(rays=Rays_GoldenSpiral(96))
    w = synth.calibrated_weights(cfg3, seed=0)
    np.savez(os.path.join(HERE, "bench_heads_3d.npz"), prob_kernel=w['prob'][0], prob_bias=w['prob'][1],
             dist_kernel=w['dist'][0], dist_bias=w['dist'][1])
    print("saved 3d", w['prob'][0].shape, w['dist'][0].shape)
