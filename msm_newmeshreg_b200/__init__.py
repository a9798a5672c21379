"""msm_newmeshreg_b200 — B200-native (sm_100a) cost evaluation for newMSM discrete surface registration.

Only the hot path of rbesenczi/msm-newmeshreg lives here: the unary data term, the triplet / pairwise
regularisers and spherical point location, as hand-written CUDA behind the C ABI of include/msm_b200.h.
There is no CPU fallback: importing works anywhere, but creating a Context requires libmsm_b200.so and a CUDA device.
"""
from ._capi import Context, MsmError, MsmParams, EXPORTS, LIB_PATH, lib  # noqa: F401
