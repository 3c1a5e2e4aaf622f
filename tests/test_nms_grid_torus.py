"""The exactness argument of the NMS grids' bucket mapping (csrc/nms.cu torus_bucket), restated in numpy.

The walks test every entry of every visited bucket with the padded-circle test and nothing else.  That is exact iff
(1) the cells of one query window map to distinct buckets (an entry is met at most once per query) and (2) an entry that
sits in a visited bucket but belongs to another cell of the plane can never pass the circle test.  Both follow from the
window limit (kMaxCellsPerQuery = 25 cells) and the torus dimensions (>= 32 cells per axis); this test replays the
argument on random geometry, including the clamped cell range and the three grid sizes the kernels use."""
import numpy as np
import pytest

K_MAX_CELLS = 25


def cell_of(v, inv_cell):
    return np.clip(np.floor(np.float32(v) * np.float32(inv_cell)), -32768.0, 32767.0).astype(np.int64)


def torus_bucket(ix, iy, n_buckets):
    bits = int(n_buckets).bit_length() - 1
    bx = bits >> 1
    return (((iy << bx) | (ix & ((1 << bx) - 1))) & (n_buckets - 1)).astype(np.int64)


@pytest.mark.parametrize("n_buckets", [1024, 4096, 8192, 1 << 18])
def test_window_cells_never_share_a_bucket_and_aliases_cannot_touch(n_buckets):
    rng = np.random.default_rng(n_buckets)
    bits = n_buckets.bit_length() - 1
    dims = (1 << (bits >> 1), 1 << (bits - (bits >> 1)))
    assert min(dims) >= 32 > K_MAX_CELLS
    checked = 0
    for _ in range(4000):
        inv_cell = np.float32(10.0 ** rng.uniform(-2, 2))
        cell = 1.0 / float(inv_cell)
        span = rng.uniform(-1.0, 1.0) * 10.0 ** rng.uniform(0, 5.5)       # up to beyond the clamp for small cells
        x, y = np.float32(span), np.float32(rng.uniform(-1, 1) * abs(span))
        reach = np.float32(rng.uniform(0.05, 12.4) * cell)
        ix0, ix1 = cell_of(x - reach, inv_cell), cell_of(x + reach, inv_cell)
        iy0, iy1 = cell_of(y - reach, inv_cell), cell_of(y + reach, inv_cell)
        if (ix1 - ix0 + 1) * (iy1 - iy0 + 1) > K_MAX_CELLS:
            continue                                                      # the kernels scan linearly then
        gx, gy = np.meshgrid(np.arange(ix0, ix1 + 1), np.arange(iy0, iy1 + 1))
        b = torus_bucket(gx.ravel(), gy.ravel(), n_buckets)
        assert len(np.unique(b)) == b.size                                # (1) one bucket per window cell
        # (2) entries whose cell aliases into a visited bucket: shifted by a non-zero multiple of the torus period
        k = rng.integers(0, gx.size)
        for dxp, dyp in ((1, 0), (-1, 0), (0, 1), (0, -1), (1, -1), (2, 0)):
            ex_c, ey_c = gx.ravel()[k] + dxp * dims[0], gy.ravel()[k] + dyp * dims[1]
            if not (-32768 < ex_c < 32767 and -32768 < ey_c < 32767):
                continue                                                  # clamp cells collect everything beyond: same cell, no alias
            assert torus_bucket(np.int64(ex_c), np.int64(ey_c), n_buckets) == b[k]
            # the closest point of that cell to the query centre is farther away than the reach (= r + r_partner_max)
            lo_x, hi_x = ex_c * cell, (ex_c + 1) * cell
            lo_y, hi_y = ey_c * cell, (ey_c + 1) * cell
            dx = max(lo_x - float(x), 0.0, float(x) - hi_x)
            dy = max(lo_y - float(y), 0.0, float(y) - hi_y)
            assert np.hypot(dx, dy) > float(reach) * 1.2                   # not even close: >= 19 cells vs < 13
            checked += 1
    assert checked > 1000
