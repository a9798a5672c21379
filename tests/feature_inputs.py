"""Synthetic inputs for the feature pre-processing tests (src/featurespace.cc:18-108): a data mesh that is NOT the ico template
(a smoothly warped icosphere of another resolution), multi-dimensional noisy fields with different offsets / scales per mesh."""
import numpy as np

import problems as pb
from oracle import pyorc


def make(level_in=4, level_ref=4, D=2, seed=0, warp=1.5, noise=0.3):
    xi, ti = pyorc.icosphere(level_in)
    xr, tr = pyorc.icosphere(level_ref)
    xi = pb.smooth_warp(xi, seed + 5, warp) if warp > 0 else xi
    rng = np.random.default_rng(seed)
    di = np.stack([2.0 + 1.5 * pb.random_field(xi, seed * 7 + d) + noise * rng.normal(size=len(xi)) for d in range(D)])
    dr = np.stack([-1.0 + 0.7 * pb.random_field(xr, seed * 7 + 50 + d) + noise * rng.normal(size=len(xr)) for d in range(D)])
    return (xi, ti), di, (xr, tr), dr
