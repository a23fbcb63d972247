"""Host-side mirror of the reference interface (geometry / bookkeeping), CPU only."""
import numpy as np
import pytest

import bsgpu
from bsgpu import fusion, stitching
from oracle import pcm_oracle as po
from tests import synth


def test_grid_create_matches_reference_usage():
    # J/SparkAffineFusion.java:457-461: 2048^3, blockSize 128^3, blockScale 2,2,1 -> 1024 jobs
    g = fusion.grid_create((2048, 2048, 2048), (256, 256, 128), (128, 128, 128))
    assert len(g) == 8 * 8 * 16
    assert g[0] == ((0, 0, 0), (256, 256, 128), (0, 0, 0))
    assert g[1] == ((256, 0, 0), (256, 256, 128), (2, 0, 0))      # x fastest, grid pos in storage blocks
    assert g[8] == ((0, 256, 0), (256, 256, 128), (0, 2, 0))
    g = fusion.grid_create((300, 100, 50), (256, 256, 128), (128, 128, 128))
    assert g == [((0, 0, 0), (256, 100, 50), (0, 0, 0)), ((256, 0, 0), (44, 100, 50), (2, 0, 0))]


def test_adjust_all_transforms_anisotropy():
    regs = {0: synth.translation((10, 20, 30))}
    out = fusion.adjust_all_transforms(regs, anisotropy_factor=2.0)
    assert np.allclose(out[0], [[1, 0, 0, 10], [0, 1, 0, 20], [0, 0, 0.5, 15]])
    assert np.allclose(fusion.adjust_all_transforms(regs)[0], regs[0])


def test_find_overlapping_views_expand_by_two():
    dims = {0: (100, 100, 100), 1: (100, 100, 100)}
    regs = {0: synth.translation((0, 0, 0)), 1: synth.translation((200, 0, 0))}
    # view 0 spans x [0, 99]; block [101, 150] is 2 px away -> still "overlapping" (expand 2)
    assert fusion.find_overlapping_views(dims, regs, (101, 0, 0), (150, 50, 50)) == [0]
    assert fusion.find_overlapping_views(dims, regs, (102, 0, 0), (150, 50, 50)) == []
    assert fusion.find_overlapping_views(dims, regs, (90, 0, 0), (210, 50, 50)) == [0, 1]


def test_best_mipmap_level_rule():
    res = [(1, 1, 1), (2, 2, 1), (4, 4, 2)]
    mts = [np.array([[f[0], 0, 0, (f[0] - 1) / 2], [0, f[1], 0, (f[1] - 1) / 2], [0, 0, f[2], (f[2] - 1) / 2]], float) for f in res]
    # full-res registration: any downsampled level would step > 1.02 px in x/y -> level 0
    assert fusion.best_mipmap_level(synth.translation((0, 0, 0)), res, mts) == 0
    # registration that shrinks the view by 4: level 2 steps are (1, 1, 0.5) -> accepted, best scaling
    M = np.diag([0.25, 0.25, 0.25]) @ synth.translation((0, 0, 0))
    assert fusion.best_mipmap_level(M, res, mts) == 2


def test_local_raster_overlaps_integer_and_real():
    a1, a2, size, s1, s2 = stitching.local_raster_overlaps((256, 256, 256), (256, 256, 256), (0, 0, 0), (205, 0, 0))
    assert list(a1) == [205, 0, 0] and list(a2) == [0, 0, 0] and list(size) == [51, 256, 256]
    a1, a2, size, s1, s2 = stitching.local_raster_overlaps((256, 256, 256), (256, 256, 256), (0, 0, 0), (204.6, 0, 0))
    assert list(a1) == [205, 0, 0] and list(a2) == [0, 0, 0] and list(size) == [51, 256, 256]
    assert np.allclose(s1, [0.4, 0, 0]) and np.allclose(s2, [0, 0, 0])
    assert stitching.local_raster_overlaps((10, 10, 10), (10, 10, 10), (0, 0, 0), (10, 0, 0)) is None


def test_filters_and_transform_equality():
    R = stitching.PairwiseStitchingResult
    rs = [R((0, 1), synth.translation((3, -2, 1)), 0.9, (0, 0, 0), (1, 1, 1)),
          R((0, 2), synth.translation((1, 0, 0)), 0.2, (0, 0, 0), (1, 1, 1)),
          R((1, 2), synth.translation((50, 0, 0)), 0.8, (0, 0, 0), (1, 1, 1)), None]
    assert [r.pair for r in stitching.filter_results(rs)] == [(0, 1), (1, 2)]
    assert [r.pair for r in stitching.filter_results(rs, max_shift_xyz=(10, 10, 10))] == [(0, 1)]
    assert [r.pair for r in stitching.filter_results(rs, max_shift_total=3.0)] == []
    assert stitching.non_translations_equal(synth.translation((1, 2, 3)), synth.translation((9, 9, 9)))
    assert not stitching.non_translations_equal(synth.translation((1, 2, 3)), synth.rot_z(1.0))


class _OracleCtx:
    """Stand-in for native.Context in CPU tests of the host logic: routes pcm_pair to the
    oracle (tests only -- the product path has no such fallback)."""

    @staticmethod
    def pcm_params(peaks, sub, mo, ext):
        return dict(peaks_to_check=peaks, do_subpixel=sub, min_overlap_frac=mo, extension=ext)

    def pcm_pair(self, a, b, p):
        return po.pcm_shift(a, b, **p)


def test_compute_stitching_sign_convention_config1_style():
    """B registered at the nominal (205,0,0)-style offset but truly at
    nominal + (3,-2,1) -> the recovered correction of B is (+3,-2,+1)."""
    n, ov = 96, 40
    G = synth.field((n + 16, n + 16, 2 * n + 16), seed=3, sigma=1.0)
    A = synth.tile_from(G, (8, 8, 8), (n, n, n), 1)
    nominal = n - ov
    B = synth.tile_from(G, (8 + 1, 8 - 2, 8 + nominal + 3), (n, n, n), 2)
    res = stitching.compute_stitching(A, B, synth.translation((0, 0, 0)), synth.translation((nominal, 0, 0)),
                                      stitching.PairwiseStitchingParameters(), (1, 1, 1), _OracleCtx())
    assert res is not None
    (T, r), (bmin, bmax) = res
    assert np.allclose(T[:, :3], np.eye(3))
    assert np.allclose(T[:, 3], (3, -2, 1), atol=0.3) and np.all(np.rint(T[:, 3]) == (3, -2, 1))
    assert r > 0.9 and bmin[0] == nominal and bmax[0] == n - 1


def test_compute_stitching_downsampled_scales_shift_back():
    n, ov = 96, 48
    G = synth.field((n + 16, n + 16, 2 * n + 16), seed=4, sigma=2.0)
    A = synth.tile_from(G, (8, 8, 8), (n, n, n), 1, noise=5)
    nominal = n - ov
    B = synth.tile_from(G, (8, 8 + 2, 8 + nominal - 4), (n, n, n), 2, noise=5)
    res = stitching.compute_stitching(A, B, synth.translation((0, 0, 0)), synth.translation((nominal, 0, 0)),
                                      stitching.PairwiseStitchingParameters(), (2, 2, 1), _OracleCtx())
    (T, r), _ = res
    assert np.all(np.abs(T[:, 3] - (-4, 2, 0)) <= 1.0)


def test_aggregate_group_actions():
    a = np.full((2, 2, 2), 10, np.uint16)
    b = np.full((2, 2, 2), 30, np.uint16)
    assert stitching.aggregate_group([a]) is a
    avg = stitching.aggregate_group([a, b], "AVERAGE")
    assert avg.dtype == np.float32 and np.all(avg == 20)
    assert stitching.aggregate_group([a, b], "PICK_BRIGHTEST") is b
    with pytest.raises(ValueError):
        stitching.aggregate_group([a, b], "NOPE")


def test_bench_reference_arm_json_contract():
    """`bench.py --impl reference` prints one JSON line with the keys the driver reads (tiny size here)."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--size", "32",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-500:]
    d = json.loads(out.stdout.strip().splitlines()[-1])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
              "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "pairs/s" and d["value"] > 0 and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and "sample" in d["cpu_baseline"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"] and "workload" in d["config"]


def test_stitch_pairs_resident_batches_respect_the_budget():
    """The resident-tile work queue: with a budget of two tiles every pair still gets its result (tiles are re-uploaded
    across batches), identical to the unconstrained run."""
    from bsgpu import stitching as bst
    from tests import synth
    from tests.fake_ctx import FakeContext
    G = synth.field((40, 40, 150), seed=4, sigma=1.0)
    tiles, models = {}, {}
    for i in range(4):
        tiles[(0, i)] = synth.tile_from(G, (4, 4, 4 + 30 * i + (i % 2)), (32, 32, 48), 80 + i, noise=3)
        models[(0, i)] = synth.translation((30 * i, 0, 0))
    pairs = [((0, 0), (0, 1)), ((0, 1), (0, 2)), ((0, 2), (0, 3))]

    class Counting(FakeContext):
        uploads = 0

        def volume_upload(self, vol):
            Counting.uploads += 1
            return super().volume_upload(vol)

    ctx = Counting()
    full = bst.stitch_pairs(pairs, tiles, models, None, (1, 1, 1), ctx)
    assert Counting.uploads == 4 and not ctx.vols
    Counting.uploads = 0
    small = bst.stitch_pairs(pairs, tiles, models, None, (1, 1, 1), ctx, max_resident_bytes=2 * tiles[(0, 0)].nbytes)
    assert Counting.uploads == 6 and not ctx.vols          # three batches of two tiles
    for a, b in zip(full, small):
        assert a is not None and np.array_equal(a.transform, b.transform) and a.r == b.r
    assert [tuple(np.rint(r.transform[:, 3]).astype(int)) for r in full] == [(1, 0, 0), (-1, 0, 0), (1, 0, 0)]
