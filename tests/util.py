import numpy as np

DT600 = (1.0 * (1.0 / 60.0)) / 10     # src/main.js:79 with numSubsteps = 10
DT1200 = (1.0 * (1.0 / 60.0)) / 20    # numSubsteps = 20 (the WebGL default, src/main.js:26)


def vec_rel_err(x, ref):
    """max_i ||x_i - ref_i|| / ||ref_i||  (BASELINE.md section 2)."""
    x = np.asarray(x, np.float64).reshape(-1, 3)
    r = np.asarray(ref, np.float64).reshape(-1, 3)
    return float(np.max(np.linalg.norm(x - r, axis=1) / np.maximum(np.linalg.norm(r, axis=1), 1e-30)))


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def assert_bit_equal(a, b, what=""):
    a, b = np.ascontiguousarray(a, np.float32).reshape(-1), np.ascontiguousarray(b, np.float32).reshape(-1)
    bad = np.flatnonzero(bits(a) != bits(b))
    # +0 / -0 are the same number; the reference never branches on the sign of zero
    bad = bad[~((a[bad] == 0) & (b[bad] == 0))]
    assert bad.size == 0, "%s: %d of %d floats differ, first at %d: %r vs %r (max abs %g)" % (
        what, bad.size, a.size, bad[0], a[bad[0]], b[bad[0]], np.nanmax(np.abs(a[bad] - b[bad])))
