"""Host logic end to end on the CPU: the command bodies (XML + BDV-N5 in, <StitchingResults> and
N5 / OME-Zarr / multi-resolution containers out) driven through an oracle-backed fake context."""
import numpy as np

import bsgpu
from bsgpu import commands, fusion, n5 as bn5, spimdata
from bsgpu import zarr as bz
from oracle import fusion_oracle as fo
from tests import synth
from tests.fake_ctx import FakeContext


def _dataset(tmp_path, n=48, ov=20):
    nominal = n - ov
    G = synth.field((n + 16, 2 * n + 16, 2 * n + 16), seed=9, sigma=1.0)
    tiles, vols = [], {}
    planted = {1: (2, -1, 1), 2: (-1, 2, 0)}
    for s, (gx, gy) in enumerate([(0, 0), (1, 0), (0, 1)]):
        e = planted.get(s, (0, 0, 0))
        vols[s] = synth.tile_from(G, (8 + e[2], 8 + gy * nominal + e[1], 8 + gx * nominal + e[0]), (n, n, n), 10 + s, noise=5)
        tiles.append(dict(setup=s, size_xyz=(n, n, n), tile=s, translation_xyz=(gx * nominal, gy * nominal, 0)))
    store = bn5.N5Store(str(tmp_path / "dataset.n5"), create=True)
    for s, v in vols.items():
        bn5.write_bdv_setup(store, s, 0, v, (32, 32, 32), compression="gzip" if s else "raw")
    xml = spimdata.write_dataset_xml(str(tmp_path / "dataset.xml"), "dataset.n5", tiles)
    return xml, vols, tiles, planted, nominal


def test_stitching_command_writes_filtered_results(tmp_path):
    xml, vols, tiles, planted, nominal = _dataset(tmp_path)
    ctx = FakeContext()
    raw = commands.stitching(xml, ctx, downsampling=(1, 1, 1), min_r=0.5)
    d = spimdata.SpimData2.load(xml)
    assert [p for p in d.stitching_pairs()] == [((0, 0), (0, 1)), ((0, 0), (0, 2)), ((0, 1), (0, 2))]
    assert ctx.calls["pcm"] == 3 and len(raw) == 3
    res = {r["pair"]: r for r in d.stitching_results()}
    # pair (0,1): correction of tile 1 relative to tile 0 == its planted registration error
    assert np.all(np.rint(res[((0, 0), (0, 1))]["shift"][:, 3]) == planted[1])
    assert np.all(np.rint(res[((0, 0), (0, 2))]["shift"][:, 3]) == planted[2])
    # pair (1,2): relative error = planted[2] - planted[1]
    assert np.all(np.rint(res[((0, 1), (0, 2))]["shift"][:, 3]) == np.subtract(planted[2], planted[1]))
    assert all(r["r"] >= 0.5 for r in res.values())
    # dry run leaves the XML untouched
    before = open(xml).read()
    commands.stitching(xml, ctx, downsampling=(1, 1, 1), dry_run=True)
    assert open(xml).read() == before


def test_fusion_commands_n5_zarr_and_multires(tmp_path):
    xml, vols, tiles, planted, nominal = _dataset(tmp_path)
    ctx = FakeContext()
    views = []
    for t in tiles:
        M = synth.translation(t["translation_xyz"])
        border, rng = fo.adjust_blending(M)
        views.append(fo.View(vols[t["setup"]], M, border, rng))
    ext = (nominal + 48, nominal + 48, 48)
    want = fo.fuse_block(views, (0, 0, 0), ext, fo.AVG_BLEND)

    out = str(tmp_path / "fused.n5")
    commands.create_fusion_container(xml, out, block_size=(16, 16, 16), downsamplings=[(2, 2, 1)])
    ds = commands.affine_fusion(out, ctx, "AVG_BLEND", block_scale=(2, 2, 2))
    st, meta = bn5.read_fusion_container(out)
    assert ds == "ch0tp0/s0" and meta["bb_min"] == [0, 0, 0] and meta["bb_max"] == [e - 1 for e in ext]
    s0 = st.read_volume("ch0tp0/s0")
    assert np.array_equal(s0, want)                                           # block seams invisible
    assert np.array_equal(st.read_volume("ch0tp0/s1"), fo.downsample2x(want, (2, 2, 1)))
    assert ctx.calls["downsample"] == len(fusion.grid_create(ext, (32, 32, 32), (16, 16, 16)))
    assert not ctx.vols                                                      # every resident volume was released

    outz = str(tmp_path / "fused.zarr")
    commands.create_fusion_container(xml, outz, block_size=(16, 16, 16), dtype="uint16", min_intensity=0.0,
                                     max_intensity=4000.0)
    dz = commands.affine_fusion(outz, ctx, "AVG_BLEND")
    stz, mz = bz.read_fusion_container_zarr(outz)
    assert dz == "0" and mz["dtype"] == "uint16"
    assert np.array_equal(stz.read_volume("0"), fo.convert_output(want, "uint16", 0.0, 4000.0))


def test_fuse_volume_retries_failed_blocks(tmp_path):
    """RetryTrackerSpark semantics: blocks whose native call fails are re-queued, at most 5 attempts."""
    import pytest
    vol = synth.tile_from(synth.field((16, 16, 16), seed=1), (0, 0, 0), (16, 16, 16), 1)

    class Flaky(FakeContext):
        fails = 2

        def fuse_block(self, *a, **k):
            if Flaky.fails > 0:
                Flaky.fails -= 1
                raise bsgpu.BsError(-2, "injected")
            return super().fuse_block(*a, **k)

    ctx = Flaky()
    sup = fusion.BlkAffineFusion.init(ctx, {0: vol}, {0: synth.translation((0, 0, 0))}, "AVG", bounding_box=((0, 0, 0), (15, 15, 15)))
    out = fusion.fuse_volume(sup, (16, 16, 16), (8, 8, 8), (1, 1, 1))
    assert np.array_equal(out, vol.astype(np.float32))
    Flaky.fails = 10 ** 6
    with pytest.raises(RuntimeError):
        fusion.fuse_volume(sup, (16, 16, 16), (8, 8, 8), (1, 1, 1), retries=2)
