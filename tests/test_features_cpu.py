"""Feature pre-processing, CPU side: known answers for the oracle port (oracle/feat.h) and the pin of
that port against the REFERENCE ITSELF — oracle/_ref/libref_feat.so = /root/reference/src/featurespace.cc + src/reg_tools.cpp
compiled where they lie (oracle/Makefile `reffeat`).  What the reference calls into FSL for (metric_resample, smooth_data,
create_exclusion, MISCMATHS::Histogram: [EXT] A8-A11) is restated once in oracle/feat.h and shared by both sides, so the pin
covers the reference's own code: the driver loop of initialize(), variance_normalise, multivariate_histogram_normalization,
set_data."""
import numpy as np
import pytest

import feature_inputs as fi
from oracle import pyorc, pyreffeat


# ------------------------------------------------------------------------------------------------ known answers
def test_smoothing_known_answers():
    xyz, tris = pyorc.icosphere(3)
    const = np.full((2, len(xyz)), 3.25)
    assert np.abs(pyorc.smooth_data(xyz, const, xyz, 6.0) - 3.25).max() < 1e-13           # weights sum to one
    f = np.random.default_rng(0).normal(size=(1, len(xyz)))
    # a kernel narrower than the vertex spacing (ico3: ~13.8) sees only the vertex itself
    assert np.abs(pyorc.smooth_data(xyz, f, xyz, 2.0) - f).max() == 0.0
    # direct evaluation of the stated kernel at one vertex
    k, sigma = 17, 12.0
    chord = np.linalg.norm(xyz - xyz[k], axis=1)
    d = 200.0 * np.arcsin(np.minimum(1.0, chord / 200.0))
    w = np.where(chord < 2.0 * sigma, np.exp(-d * d / (2 * sigma * sigma)), 0.0)
    assert abs(pyorc.smooth_data(xyz, f, xyz, sigma)[0, k] - (w * f[0]).sum() / w.sum()) < 1e-13
    # excluded sources carry no weight; a vertex with nothing in range gets 0
    mask = np.ones(len(xyz)); mask[k] = 0.0
    w2 = w.copy(); w2[k] = 0.0
    assert abs(pyorc.smooth_data(xyz, f, xyz, sigma, mask)[0, k] - (w2 * f[0]).sum() / w2.sum()) < 1e-13
    assert pyorc.smooth_data(xyz, f, xyz, 2.0, mask)[0, k] == 0.0


def test_variance_normalise_known_answers():
    rng = np.random.default_rng(1)
    x = 5.0 + 3.0 * rng.normal(size=(3, 1000))
    y = pyorc.variance_normalise(x)
    assert np.abs(y.mean(1)).max() < 1e-13 and np.abs(y.var(1, ddof=1) - 1.0).max() < 1e-12
    mask = (rng.uniform(size=1000) > 0.3).astype(float)
    z = pyorc.variance_normalise(x, mask)
    keep = mask > 0
    assert np.array_equal(z[:, ~keep], x[:, ~keep])                                      # excluded vertices untouched
    assert np.abs(z[:, keep].mean(1)).max() < 1e-13 and np.abs(z[:, keep].var(1, ddof=1) - 1.0).max() < 1e-12
    c = pyorc.variance_normalise(np.full((1, 50), 2.0))                                  # zero variance: mean removed, no division
    assert np.array_equal(c, np.zeros((1, 50)))


def test_histogram_match_known_answers():
    rng = np.random.default_rng(2)
    a = rng.normal(size=(1, 5000)); b = 10.0 + 2.0 * rng.normal(size=(1, 4000))
    m = pyorc.histogram_normalization(a, b)
    lo, hi = np.float32(b.min()), np.float32(b.max())
    width = np.float32((hi - lo) / np.float32(256))
    edges = (lo + np.arange(256, dtype=np.float32) * width).astype(np.float64)
    assert np.isin(m, edges).all()                                                       # every value lands on a template bin edge
    order = np.argsort(a[0])
    assert (np.diff(m[0, order]) >= 0).all()                                             # monotone map
    assert abs(np.median(m) - np.median(b)) < 0.1 and abs(m.std() - b.std()) < 0.1       # distribution follows the template
    assert m[0, np.argmax(a[0])] == edges[-1]                                            # the last bin stays the last bin
    # masks: excluded values keep their value and do not count
    ex = np.ones(5000); ex[:100] = 0.0
    m2 = pyorc.histogram_normalization(a, b, excl_in=ex)
    assert np.array_equal(m2[0, :100], a[0, :100])


def test_create_exclusion_known_answers():
    d = np.array([[0.0, 0.5, 2.0, 0.0], [0.0, 0.2, 0.1, 3.0]])
    assert np.array_equal(pyorc.create_exclusion(d, 0.0, 0.5), [0.0, 0.0, 1.0, 1.0])
    assert np.array_equal(pyorc.create_exclusion(d[:1], 0.0, 0.0), [0.0, 1.0, 1.0, 0.0])


def test_masked_resampling_known_answers():
    (xi, ti), di, (xr, tr), _ = fi.make(3, 3, D=1, seed=4)
    ones = np.ones(len(xi))
    assert np.abs(pyorc.metric_resample_excl(xi, ti, di, ones, xr, tr) - pyorc.metric_resample(xi, ti, di, xr, tr)).max() < 1e-14
    mask = ones.copy(); mask[::3] = 0.0
    const = np.full((1, len(xi)), 2.5)
    out = pyorc.metric_resample_excl(xi, ti, const, mask, xr, tr)
    assert np.isin(np.round(out, 12), [0.0, 2.5]).all() and (out == 0).sum() < len(xr) // 4   # renormalised over what is left


# ------------------------------------------------------------------------------------------------ pin against the reference build
CASES = [
    dict(ico=3),                                                            # defaults: sigma 5 on both, no normalisation
    dict(ico=0),                                                            # every mesh keeps its own vertices
    dict(ico=4, sigma=(3.0, 0.0)),                                          # no smoothing of the reference (sigma 0)
    dict(ico=3, intensitynorm=True),
    dict(ico=3, varnorm=True),
    dict(ico=4, sigma=(4.0, 2.0), intensitynorm=True, varnorm=True),
    dict(ico=0, sigma=(2.0, 3.0), exclude=True, thresholds=(1.5, 2.5), intensitynorm=True, varnorm=True),   # masks (same vertex sets)
    dict(ico=4, cut=True, thresholds=(-10.0, -9.0), varnorm=True, same=True),   # --cut with an empty band: masks of ones
]


@pytest.mark.skipif(not pyreffeat.available(), reason="oracle/_ref/libref_feat.so not built (needs /root/reference at build time)")
@pytest.mark.parametrize("kw", CASES)
def test_port_equals_the_reference(kw):
    kw = dict(kw)
    ico = kw.pop("ico"); same = kw.pop("same", False)
    mi, di, mr, dr = fi.make(4, 4 if (same or ico == 0) else 3, D=2, seed=11, warp=0.0 if same else 1.5)
    a = pyorc.featurespace_initialize(ico, mi, di, mr, dr, **kw)
    b = pyreffeat.featurespace_initialize(ico, mi, di, mr, dr, **kw)
    for u, v in zip(a[:2], b[:2]):
        assert u.shape == v.shape and np.isfinite(u).all()
        assert np.abs(u - v).max() < 1e-12
    for u, v in zip(a[2:], b[2:]):
        assert (u is None) == (v is None)
        if u is not None:
            assert np.array_equal(u, v)
    if kw.get("varnorm") and not (kw.get("exclude") or kw.get("cut")):
        assert np.abs(a[0].mean(1)).max() < 1e-12 and np.abs(a[0].var(1, ddof=1) - 1).max() < 1e-12


@pytest.mark.skipif(not pyreffeat.available(), reason="oracle/_ref/libref_feat.so not built")
def test_reference_rejects_masks_on_foreign_vertices_like_the_port():
    """exclude with a data mesh that is not on the template's vertices: the reference reads its mask out of range there; the
    stand-in's bounds check and the port both refuse."""
    mi, di, mr, dr = fi.make(4, 3, D=1, seed=12)
    with pytest.raises(RuntimeError):
        pyorc.featurespace_initialize(3, mi, di, mr, dr, exclude=True)
    with pytest.raises(RuntimeError):
        pyreffeat.featurespace_initialize(3, mi, di, mr, dr, exclude=True)


# ------------------------------------------------------------------------------------------------ committed golden vectors
def _golden():
    import importlib.util
    import os
    here = os.path.dirname(os.path.abspath(__file__))
    spec = importlib.util.spec_from_file_location("make_golden_features", os.path.join(here, "golden", "make_golden_features.py"))
    mg = importlib.util.module_from_spec(spec); spec.loader.exec_module(mg)
    return mg, np.load(os.path.join(here, "golden", "features.npz"))


@pytest.mark.parametrize("name", ["defaults_ico3", "normalised_ico3", "masked_own_vertices"])
def test_port_reproduces_the_committed_reference_output(name):
    """tests/golden/features.npz was written by the reference's own featurespace (tests/golden/make_golden_features.py); it
    travels to boxes where oracle/_ref cannot be rebuilt."""
    mg, gold = _golden()
    mi, di, mr, dr, ico, opt = mg.inputs(name)
    a, b, ea, eb = pyorc.featurespace_initialize(ico, mi, di, mr, dr, **opt)
    assert np.abs(a - gold[f"{name}/input"]).max() < 1e-12 and np.abs(b - gold[f"{name}/reference"]).max() < 1e-12
    if ea is not None:
        assert np.array_equal(ea, gold[f"{name}/excl_input"]) and np.array_equal(eb, gold[f"{name}/excl_reference"])
