"""TEST INFRASTRUCTURE ONLY.  numpy restatement of libm's hypot as numpy.hypot evaluates it on the reference's platform.

The reference computes radii with ``np.hypot`` (torchbox3d/math/numpy/conversions.py:61-62), i.e. glibc's ``hypot``
(third-party, not vendored: glibc 2.39 in this image).  Since glibc 2.35 the generic (non-FMA) double routine is
``sysdeps/ieee754/dbl-64/e_hypot.c``: ``h = sqrt(ax*ax + ay*ay)`` plus one correction step (C. Borges, "An Improved
Algorithm for hypot(a,b)").  It is not correctly rounded, so the CUDA rasterizer restates it
(csrc/fastmath.cuh ``libm_hypot``) where a last-bit difference could change a pixel assignment; this file is the same
restatement on the CPU, pinned against ``np.hypot`` itself by tests/test_oracle_golden.py::test_libm_hypot_restatement."""
import numpy as np

_EPS = 2.0 ** -54


def libm_hypot(x: np.ndarray, y: np.ndarray) -> np.ndarray:
    x = np.abs(np.asarray(x, dtype=np.float64)); y = np.abs(np.asarray(y, dtype=np.float64))
    ax, ay = np.maximum(x, y), np.minimum(x, y)
    with np.errstate(all="ignore"):
        h = np.sqrt(ax * ax + ay * ay)
        d1 = h - ay
        a1 = ax * (2.0 * d1 - ax); b1 = (d1 - 2.0 * (ax - ay)) * d1
        d2 = h - ax
        a2 = 2.0 * d2 * (ax - 2.0 * ay); b2 = (4.0 * d2 - ay) * ay + d2 * d2
        first = h <= 2.0 * ay
        t = np.where(first, a1 + b1, a2 + b2)
        out = h - t / (2.0 * h)
        return np.where(ax >= ay / _EPS, ax + ay, out)
