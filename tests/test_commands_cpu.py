"""Host logic end to end on the CPU: the command bodies (XML + BDV-N5 in, <StitchingResults> and
N5 / OME-Zarr / multi-resolution containers out) driven through an oracle-backed fake context."""
import numpy as np
import pytest

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
    assert ctx.calls["pcm"] == 3 and len(raw) == 3         # three pcm jobs on tiles that were uploaded once each
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
    assert ds == ["ch0tp0/s0"] and meta["bb_min"] == [0, 0, 0] and meta["bb_max"] == [e - 1 for e in ext]
    assert st.dataset_attributes("ch0tp0/s0")["compression"]["type"] == "zstd"      # the reference's default codec
    s0 = st.read_volume("ch0tp0/s0")
    assert np.array_equal(s0, want)                                           # block seams invisible
    assert np.array_equal(st.read_volume("ch0tp0/s1"), fo.downsample2x(want, (2, 2, 1)))
    # super-blocks of 32^3 are whole multiples of the pyramid step: levels come from the resident fused block
    assert ctx.calls["downsample"] == len(fusion.grid_create(ext, (32, 32, 32), (16, 16, 16)))
    assert not ctx.vols                                                      # every resident volume was released
    # the reference's metadata form: one nested "Bigstitcher-Spark" object, lowercase DataType
    import json
    root = json.load(open(out + "/attributes.json"))
    assert root["Bigstitcher-Spark"]["DataType"] == "float32" and root["Bigstitcher-Spark"]["NumChannels"] == 1
    assert "Bigstitcher-Spark/InputXML" not in root

    outz = str(tmp_path / "fused.zarr")
    commands.create_fusion_container(xml, outz, block_size=(16, 16, 16), dtype="uint16", min_intensity=0.0,
                                     max_intensity=4000.0)
    dz = commands.affine_fusion(outz, ctx, "AVG_BLEND")
    stz, mz = bz.read_fusion_container_zarr(outz)
    assert dz == ["0"] and mz["dtype"] == "uint16"
    assert np.array_equal(stz.read_volume("0"), fo.convert_output(want, "uint16", 0.0, 4000.0))

    # storage blocks that are whole kernel tiles (64 x 16 x 8) leave the device as finished chunks: big-endian N5
    # payloads, little-endian zarr chunks, no host re-striding
    for name, dtype in (("packed.n5", "float32"), ("packed.zarr", "uint16")):
        outp = str(tmp_path / name)
        commands.create_fusion_container(xml, outp, block_size=(64, 16, 8), dtype=dtype, min_intensity=0.0, max_intensity=4000.0)
        dp = commands.affine_fusion(outp, ctx, "AVG_BLEND", block_scale=(1, 2, 3))
        if name.endswith(".n5"):
            got = bn5.read_fusion_container(outp)[0].read_volume(dp[0])
        else:
            got = bz.read_fusion_container_zarr(outp)[0].read_volume(dp[0])
        assert np.array_equal(got, want if dtype == "float32" else fo.convert_output(want, "uint16", 0.0, 4000.0))

    # multi-resolution OME-ZARR (the reference's default container + --multiRes / -ds), with a pyramid step the
    # super-blocks are NOT multiples of: levels are rebuilt from the container's level l-1 (the reference's way)
    outm = str(tmp_path / "fused_mr.zarr")
    commands.create_fusion_container(xml, outm, block_size=(16, 16, 16), downsamplings=[(2, 2, 1), (2, 2, 2)])
    commands.affine_fusion(outm, ctx, "AVG_BLEND", block_scale=(1, 1, 1))
    stm, mm = bz.read_fusion_container_zarr(outm)
    lv = mm["mr_infos"][0]
    assert [l["dataset"] for l in lv] == ["0", "1", "2"] and lv[2]["absoluteDownsampling"] == [4, 4, 2, 1, 1]
    assert np.array_equal(stm.read_volume("0"), want)
    s1 = fo.downsample2x(want, (2, 2, 1))
    assert np.array_equal(stm.read_volume("1"), s1)
    assert np.array_equal(stm.read_volume("2"), fo.downsample2x(s1, (2, 2, 2)))
    ms = stm.get_attributes("")["multiscales"][0]
    assert [d["path"] for d in ms["datasets"]] == ["0", "1", "2"]
    assert ms["datasets"][2]["coordinateTransformations"][0]["scale"] == [1.0, 1.0, 2.0, 4.0, 4.0]
    assert ms["datasets"][2]["coordinateTransformations"][1]["translation"] == [0.0, 0.0, 0.5, 1.5, 1.5]
    assert not ctx.vols


def test_flat_round1_metadata_still_readable(tmp_path):
    """Containers written by round 1 of this build used flat keys with a literal slash; the reader accepts both."""
    import json
    import os
    root = tmp_path / "old.n5"
    os.makedirs(root)
    json.dump({"n5": "2.5.1", "Bigstitcher-Spark/FusionFormat": "N5", "Bigstitcher-Spark/InputXML": "x.xml",
               "Bigstitcher-Spark/Boundingbox_min": [0, 0, 0], "Bigstitcher-Spark/Boundingbox_max": [7, 7, 7],
               "Bigstitcher-Spark/DataType": "FLOAT32", "Bigstitcher-Spark/BlockSize": [8, 8, 8],
               "Bigstitcher-Spark/MultiResolutionInfos": [[{"dataset": "ch0tp0/s0"}]]}, open(root / "attributes.json", "w"))
    _, meta = bn5.read_fusion_container(str(root))
    assert meta["dtype"] == "float32" and meta["input_xml"] == "x.xml" and meta["num_channels"] == 1


def test_multichannel_project_fuses_each_channel_from_its_own_views(tmp_path):
    """views are selected per (channel, timepoint) (J/SparkAffineFusion.java:425-440) and the container
    carries NumChannels from the XML; `stitching` groups a tile's channels (AVERAGE) before correlating."""
    n, ov = 48, 20
    nominal = n - ov
    G = synth.field((n + 16, n + 16, 2 * n + 16), seed=19, sigma=1.0)
    tiles, vols = [], {}
    for tile in range(2):
        for ch in range(2):
            s = tile * 2 + ch
            e = (2, -1, 1) if tile == 1 else (0, 0, 0)
            v = synth.tile_from(G, (8 + e[2], 8 + e[1], 8 + tile * nominal + e[0]), (n, n, n), 30 + s, noise=5)
            vols[s] = (v // (1 + ch)).astype(np.uint16)          # channel 1 is dimmer
            tiles.append(dict(setup=s, size_xyz=(n, n, n), tile=tile, channel=ch, translation_xyz=(tile * nominal, 0, 0)))
    store = bn5.N5Store(str(tmp_path / "dataset.n5"), create=True)
    for s, v in vols.items():
        bn5.write_bdv_setup(store, s, 0, v, (32, 32, 32), compression="zstd")
    xml = spimdata.write_dataset_xml(str(tmp_path / "dataset.xml"), "dataset.n5", tiles)
    ctx = FakeContext()
    d = spimdata.SpimData2.load(xml)
    assert d.channels_ordered() == [0, 1]
    assert d.stitching_groups() == [([(0, 0), (0, 1)], [(0, 2), (0, 3)])]
    raw = commands.stitching(xml, ctx, downsampling=(1, 1, 1))
    assert len(raw) == 1 and np.all(np.rint(raw[0].transform[:, 3]) == (2, -1, 1))
    res = spimdata.SpimData2.load(xml).stitching_results()
    assert res[0]["pair"] == (((0, 0), (0, 1)), ((0, 2), (0, 3)))           # grouped ids survive the XML round trip
    # a second run replaces (never duplicates) the stored result of every compared pair
    commands.stitching(xml, ctx, downsampling=(1, 1, 1), min_r=2.0)         # nothing passes the filter now
    assert spimdata.SpimData2.load(xml).stitching_results() == []

    out = str(tmp_path / "fused.n5")
    commands.create_fusion_container(xml, out, block_size=(16, 16, 16), compression="gzip")
    st, meta = bn5.read_fusion_container(out)
    assert meta["num_channels"] == 2 and len(meta["mr_infos"]) == 2
    assert commands.affine_fusion(out, ctx, "AVG_BLEND") == ["ch0tp0/s0", "ch1tp0/s0"]
    for ch in range(2):
        views = []
        for t in tiles:
            if t["channel"] == ch:
                M = synth.translation(t["translation_xyz"])
                border, rng = fo.adjust_blending(M)
                views.append(fo.View(vols[t["setup"]], M, border, rng))
        want = fo.fuse_block(views, (0, 0, 0), (nominal + n, n, n), fo.AVG_BLEND)
        assert np.array_equal(st.read_volume(f"ch{ch}tp0/s0"), want)
    assert commands.affine_fusion(out, ctx, "AVG_BLEND", channel=1, timepoint=0) == ["ch1tp0/s0"]

    # view selection (AbstractSelectableViews): id lists OR explicit ViewIds, never both; empty selections are errors
    assert d.select_views(channel_ids="1") == [(0, 1), (0, 3)]
    assert d.select_views(tile_ids=[1], channel_ids=[0]) == [(0, 2)]
    assert d.select_views(vi=["0,3", "0,0", "5,9"]) == [(0, 0), (0, 3)]          # only the ones that exist
    with pytest.raises(ValueError):
        d.select_views(vi=["0,0"], tile_ids="0")
    with pytest.raises(ValueError):
        d.select_views(angle_ids="7")
    # stitching only the selected channel: the groups shrink to that channel's views
    raw1 = commands.stitching(xml, ctx, downsampling=(1, 1, 1), view_selection=dict(channel_ids="1"), dry_run=True)
    assert len(raw1) == 1 and raw1[0].pair == ((0, 1), (0, 3)) and np.all(np.rint(raw1[0].transform[:, 3]) == (2, -1, 1))
    # fusing only tile 1: channel volumes contain that tile alone
    out1 = str(tmp_path / "fused_tile1.n5")
    commands.create_fusion_container(xml, out1, block_size=(16, 16, 16), compression="raw")
    commands.affine_fusion(out1, ctx, "AVG_BLEND", view_selection=dict(tile_ids="1"))
    st1, _ = bn5.read_fusion_container(out1)
    M = synth.translation((nominal, 0, 0))
    border, rng = fo.adjust_blending(M)
    want1 = fo.fuse_block([fo.View(vols[2], M, border, rng)], (0, 0, 0), (nominal + n, n, n), fo.AVG_BLEND)
    assert np.array_equal(st1.read_volume("ch0tp0/s0"), want1) and not want1[:, :, :nominal].any()


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


def test_masks_mode_writes_coverage_and_its_pyramid(tmp_path):
    """`affine-fusion --masks [--maskOffset]`: geometry only, 255 where a view covers the voxel, pyramid from that s0."""
    xml, vols, tiles, planted, nominal = _dataset(tmp_path)
    ctx = FakeContext()
    out = str(tmp_path / "masks.n5")
    commands.create_fusion_container(xml, out, block_size=(16, 16, 16), dtype="uint8", min_intensity=0.0, max_intensity=255.0,
                                     downsamplings=[(2, 2, 1)])
    ds = commands.affine_fusion(out, ctx, "AVG_BLEND", block_scale=(2, 2, 1), masks=True, mask_offset=(-1.0, 0.0, 0.0))
    st, meta = bn5.read_fusion_container(out)
    ext = (nominal + 48, nominal + 48, 48)
    geom = [(synth.translation(t["translation_xyz"]), vols[t["setup"]].shape[::-1]) for t in tiles]
    want = fo.mask_block(geom, (0, 0, 0), ext, (-1.0, 0.0, 0.0), "uint8")
    got = st.read_volume(ds[0])
    assert np.array_equal(got, want) and 0 < np.count_nonzero(want) < want.size
    assert np.array_equal(st.read_volume("ch0tp0/s1"), fo.downsample2x(want, (2, 2, 1)))
    assert ctx.calls["fuse"] == 0 and not ctx.vols          # nothing was fused, no image data touched
