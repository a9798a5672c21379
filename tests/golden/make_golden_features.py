"""Generates tests/golden/features.npz from THE REFERENCE ITSELF: oracle/_ref/libref_feat.so = the reference's own
src/featurespace.cc + src/reg_tools.cpp compiled unmodified where they lie (oracle/Makefile `reffeat`; the FSL pieces it links —
metric_resample, smooth_data, create_exclusion, MISCMATHS::Histogram — stood in for by oracle/ext/fsl_ext.h over oracle/feat.h).

The committed fixture is what travels: the CPU test checks the oracle port against it, the GPU test the product's
featurespace::initialize.  Re-run (dev container only, needs /root/reference at build time):
    python tests/golden/make_golden_features.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import feature_inputs as fi  # noqa: E402

CASES = {
    # name: (feature_inputs.make kwargs, ico, featurespace options)
    "defaults_ico3": (dict(level_in=4, level_ref=3, D=2, seed=31), 3, dict()),
    "normalised_ico3": (dict(level_in=4, level_ref=3, D=2, seed=32), 3, dict(sigma=(4.0, 2.0), intensitynorm=True, varnorm=True)),
    "masked_own_vertices": (dict(level_in=3, level_ref=3, D=2, seed=33), 0,
                            dict(sigma=(6.0, 8.0), exclude=True, thresholds=(1.5, 2.5), intensitynorm=True, varnorm=True)),
}


def inputs(name):
    kw, ico, opt = CASES[name]
    return fi.make(**kw) + (ico, opt)


def main():
    from oracle import pyreffeat
    if not pyreffeat.available():
        raise SystemExit("oracle/_ref/libref_feat.so is not built (needs /root/reference): make -C oracle reffeat")
    out = {}
    for name in CASES:
        mi, di, mr, dr, ico, opt = inputs(name)
        a, b, ea, eb = pyreffeat.featurespace_initialize(ico, mi, di, mr, dr, **opt)
        out[f"{name}/input"] = a; out[f"{name}/reference"] = b
        if ea is not None:
            out[f"{name}/excl_input"] = ea; out[f"{name}/excl_reference"] = eb
    np.savez_compressed(os.path.join(HERE, "features.npz"), **out)
    print("wrote", os.path.join(HERE, "features.npz"), {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
