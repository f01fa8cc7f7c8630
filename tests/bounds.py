"""Hard-bound comparison used by the GPU parity tests (quantile asserts let 1 % of the pixels be
arbitrarily wrong).

``assert_bounded`` compares a set of per-pixel outputs of two renders.  A pixel is a THRESHOLD-FLIP pixel if any of its
outputs differs by more than the bound: with fp16 network outputs that differ in the last bit, a sample can fall on the
other side of ``alpha >= alpha_thre`` or a ray on the other side of ``opacity <= 1 - early_stop_eps``, which changes
the pixel by up to one sample's weight.  Their NUMBER is asserted (<= 0.1 % of the pixels, at least ``floor`` so that
a 24x32 golden image may contain a couple); every other pixel of every output is within its bound, by construction
of the count.  tests/test_parity_baseline_gpu.py identifies flips exactly, from per-ray sample counts, at the BASELINE
sizes; this helper is for the comparisons against fixtures that carry only the rendered images."""
import numpy as np


def flip_pixels(pairs, n_pixels):
    """pairs: iterable of (name, got, ref, bound); returns the boolean mask of pixels with any output out of bound."""
    bad = np.zeros(n_pixels, bool)
    worst = {}
    for name, a, ref, bound in pairs:
        err = np.abs(np.asarray(a, np.float64).reshape(n_pixels, -1) - np.asarray(ref, np.float64).reshape(n_pixels, -1))
        bad |= (err > bound).any(1)
        worst[name] = (float(err.max()), bound)
    return bad, worst


def assert_bounded(pairs, n_pixels, what="", frac=1e-3, floor=2):
    pairs = list(pairs)
    bad, worst = flip_pixels(pairs, n_pixels)
    allowed = max(floor, int(frac * n_pixels))
    msg = (f"{what}: {int(bad.sum())} of {n_pixels} pixels out of bound (allowed {allowed}); worst |err| per output "
           + ", ".join(f"{k} {v[0]:.2e}/{v[1]:.1e}" for k, v in worst.items()))
    print(msg)
    assert bad.sum() <= allowed, msg
    return int(bad.sum())


def scale_of(ref):
    return max(1.0, float(np.abs(ref).max()))
