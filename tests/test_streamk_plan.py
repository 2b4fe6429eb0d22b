"""Host logic of the stream-K work split (gstools_b200/csrc/gsb_sepk.cuh, sk_plan) through the C ABI: the
shares must tile the (output tile, stage) iteration space exactly once, in order, with balanced cost."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from gstools_b200 import _lib


def _cost(ly, lc, n_yt, n_ct, tile):
    p = tile % (n_yt * n_ct)
    yt, ct = p // n_ct, p % n_ct
    rows = min(128, ly - yt * 128)
    cols = min(128, lc - ct * 128)
    rowq = (-(-rows // 8) + 3) // 4
    return max(40, rowq * 2 * -(-cols // 16))


def _check(tile_begin, tile_end, ly, lc, n_stages, max_grid):
    bnd = _lib.streamk_plan(tile_begin, tile_end, ly, lc, n_stages, max_grid)
    grid = len(bnd) - 1
    n_yt, n_ct = -(-ly // 128), -(-lc // 128)
    assert 1 <= grid <= min(max_grid, 160)
    assert bnd[0] == (tile_begin, 0) and bnd[-1] == (tile_end, 0)
    for a, b in zip(bnd, bnd[1:]):
        assert a <= b and 0 <= a[1] < n_stages and tile_begin <= a[0] <= tile_end
    # cost of every share; equal up to one stage of the most expensive tile (64 units) on either side
    if tile_end == tile_begin:
        return
    shares = []
    for (t0, s0), (t1, s1) in zip(bnd, bnd[1:]):
        c = 0
        for t in range(t0, t1 + (1 if s1 > 0 else 0)):
            lo = s0 if t == t0 else 0
            hi = s1 if t == t1 else n_stages
            c += (hi - lo) * _cost(ly, lc, n_yt, n_ct, t)
        shares.append(c)
    total = sum(_cost(ly, lc, n_yt, n_ct, t) for t in range(tile_begin, tile_end)) * n_stages
    assert sum(shares) == total
    assert max(shares) - min(shares) <= 2 * 64 + 1, (max(shares), min(shares))
    # no share smaller than an eighth of a full tile (bounded number of contributors per tile), unless there is
    # only one share
    if grid > 1:
        assert min(shares) >= max(64, n_stages * 64 // 8) - 64, (min(shares), n_stages)


@pytest.mark.parametrize("args", [
    (0, 128, 128, 128, 125, 148),          # one 128^3 field: 128 tiles on 148 SMs
    (0, 626, 200, 200, 125, 148),          # 200^3: partial row and column tiles
    (0, 4 * 4 * 512, 512, 512, 125, 148),  # 512^3
    (0, 1, 5, 9, 2, 148),                  # fewer (tile, stage) units than CTAs
    (37, 91, 300, 130, 13, 7),             # a piece in the middle of the numbering
    (5, 5, 128, 128, 4, 16),               # empty
])
def test_streamk_plan_cases(args):
    _check(*args)


@settings(max_examples=150, deadline=None)
@given(st.integers(0, 50), st.integers(0, 400), st.integers(1, 700), st.integers(1, 700), st.integers(1, 40),
       st.integers(1, 200))
def test_streamk_plan_property(tile_begin, n_tiles, ly, lc, n_stages, max_grid):
    _check(tile_begin, tile_begin + n_tiles, ly, lc, n_stages, max_grid)


def test_streamk_plan_rejects_bad_geometry():
    with pytest.raises(ValueError):
        _lib.streamk_plan(0, 4, 0, 128, 4, 8)
    with pytest.raises(ValueError):
        _lib.streamk_plan(4, 2, 128, 128, 4, 8)
