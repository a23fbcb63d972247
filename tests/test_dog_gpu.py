"""GPU parity, next row 8f-4: bs_dog_detect (through the C ABI) against oracle/dog_oracle.py.
Bars: the SAME set of extremum voxels (integer, bit-identical); sub-pixel location within 1e-3 px; value 1e-4 relative + 2e-6 (float32 DoG of a unit-range image)."""
import numpy as np
import pytest

from oracle import dog_oracle as do
from tests import synth
from tests.test_dog_oracle import _beads

pytestmark = pytest.mark.gpu


def _compare(got, want, loc_tol=1e-3):
    assert [g[2] for g in got] == [w[2] for w in want]
    for g, w in zip(got, want):
        assert g[3] == w[3]
        assert np.allclose(g[0], w[0], atol=loc_tol), (g, w)
        # the DoG is a float32 difference of two blurs of a [0, 1] image scaled by 1 / (k - 1) = 5.3: absolute floor 2e-6
        assert abs(g[1] - w[1]) <= 1e-4 * abs(w[1]) + 2e-6


@pytest.mark.parametrize("interval", [((0, 0, 0), (56, 48, 40)), ((10, 8, 4), (30, 32, 28)), ((28, 24, 0), (28, 24, 40))])
def test_beads_match_oracle(ctx, interval):
    img, centers = _beads()
    h = ctx.volume_upload(img)
    got = ctx.dog_detect(h, interval[0], interval[1], sigma=1.8, threshold=0.004, max_intensity=4000.0)
    want = do.detect(img, interval[0], interval[1], sigma=1.8, threshold=0.004, max_intensity=4000.0)
    _compare(got, want)
    ctx.volume_free(h)


def test_noise_field_many_detections_min_max_and_no_localization(ctx):
    vol = synth.tile_from(synth.field((48, 64, 72), seed=41, sigma=2.0), (0, 0, 0), (48, 64, 72), 41, noise=10)
    h = ctx.volume_upload(vol)
    kw = dict(sigma=1.8, threshold=0.002, min_intensity=0.0, max_intensity=3000.0)
    got = ctx.dog_detect(h, (0, 0, 0), (72, 64, 48), find_max=True, find_min=True, **kw)
    want = do.detect(vol, (0, 0, 0), (72, 64, 48), find_max=True, find_min=True, **kw)
    assert len(want) > 50
    # a voxel whose DoG sits within float rounding of the threshold may fall on either side: compare the common set
    gs, ws = {g[2]: g for g in got}, {w[2]: w for w in want}
    common = sorted(set(gs) & set(ws))
    assert len(common) >= 0.98 * max(len(gs), len(ws))
    _compare([gs[k] for k in common], [ws[k] for k in common], loc_tol=2e-3)
    got = ctx.dog_detect(h, (8, 8, 8), (40, 40, 30), localization=False, **kw)
    want = do.detect(vol, (8, 8, 8), (40, 40, 30), localization=False, **kw)
    assert len(want) > 5 and [g[2] for g in got] == [w[2] for w in want]
    assert all(g[0] == tuple(float(v) for v in g[2]) for g in got)
    # buffer growth: a tiny first buffer still returns everything
    assert len(ctx.dog_detect(h, (0, 0, 0), (72, 64, 48), find_max=True, find_min=True, max_points=3, **kw)) == len(gs)
    ctx.volume_free(h)


def test_bad_arguments(ctx):
    import bsgpu
    h = ctx.volume_upload(np.zeros((8, 8, 8), np.uint16))
    with pytest.raises(bsgpu.BsError):
        ctx.dog_detect(h, (0, 0, 0), (9, 8, 8))
    with pytest.raises(bsgpu.BsError):
        ctx.dog_detect(h, (0, 0, 0), (8, 8, 8), sigma=0.3)
    ctx.volume_free(h)
