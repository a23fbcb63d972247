"""The `call()` bodies of the reference's `stitching`, `create-fusion-container` and
`affine-fusion` commands with the Spark RDD collapsed to a plain host work queue over one
`native.Context` (BASELINE north_star), SpimData2 XML in, N5 blocks out.

  stitching               J/SparkPairwiseStitching.java:110-392
  create_fusion_container J/CreateFusionContainer.java:122-519   (N5 storage, one channel/timepoint)
  affine_fusion           J/SparkAffineFusion.java:179-800       (s0 level; no pyramid)

No argument parsing here (the picocli layer is out of scope); keyword names follow the CLI flags.
"""
from __future__ import annotations

import math
import os

import numpy as np

from . import fusion as bf
from . import n5 as bn5
from . import zarr as bzarr
from . import native, stitching as bst
from .native import Context
from .spimdata import SpimData2


def _load_views(data: SpimData2, store: bn5.N5Store, view_ids, level=0):
    return {v: store.read_volume(bn5.bdv_dataset(v[1], v[0], level)) for v in view_ids}


def stitching(xml_path, ctx: Context, downsampling=(2, 2, 1), peaks_to_check=5, disable_subpixel=False,
              min_r=0.3, max_r=1.0, max_shift_xyz=None, max_shift_total=None, dry_run=False):
    """`./stitching -x dataset.xml [-ds 2,2,1] [-p 5] ...`: phase-correlate every overlapping tile
    pair and store the filtered results in the XML's <StitchingResults>.  Returns all raw results."""
    data = SpimData2.load(xml_path)
    fmt, n5_path = data.image_loader()
    if fmt != "bdv.n5":
        raise NotImplementedError(f"ImageLoader format {fmt}")
    store = bn5.N5Store(n5_path)
    pairs = data.stitching_pairs()
    needed = sorted({v for p in pairs for v in p})
    tiles = _load_views(data, store, needed)
    models = {v: data.model(*v) for v in needed}
    params = bst.PairwiseStitchingParameters(peaks_to_check=peaks_to_check, do_subpixel=not disable_subpixel)
    raw = bst.stitch_pairs(pairs, tiles, models, params, downsampling, ctx)
    for r in raw:
        if r is not None:
            a, b = r.pair
            r.hash = SpimData2.transform_hash(data.registrations[a], data.registrations[b])
    kept = bst.filter_results(raw, min_r, max_r, max_shift_xyz, max_shift_total)
    data.set_stitching_results([dict(pair=r.pair, shift=r.transform, r=r.r, hash=r.hash, bbox_min=r.bbox_min,
                                     bbox_max=r.bbox_max) for r in kept])
    if not dry_run:
        data.save(xml_path)
    return raw


def create_fusion_container(xml_path, out_path, block_size=(128, 128, 128), dtype="float32", min_intensity=None,
                            max_intensity=None, preserve_anisotropy=False, anisotropy_factor=float("nan"),
                            storage=None, downsamplings=()):
    """`./create-fusion-container -x dataset.xml -o fused.n5 -s N5 -d FLOAT32 --blockSize ...`:
    bounding box of all views (Import.getBoundingBox, J/CreateFusionContainer.java:184-211) + container."""
    data = SpimData2.load(xml_path)
    lo = np.full(3, np.inf)
    hi = np.full(3, -np.inf)
    af = anisotropy_factor if preserve_anisotropy else float("nan")
    regs = bf.adjust_all_transforms({v: data.model(*v) for v in data.view_ids()}, af)
    for v, M in regs.items():
        bmin, bmax = bf.transformed_bounding_box(data.setups[v[1]].size, M)
        lo = np.minimum(lo, bmin)
        hi = np.maximum(hi, bmax)
    # storage guessed from the extension like the reference (J/SparkAffineFusion.java:206-225); the
    # reference's default is OME-ZARR (J/CreateFusionContainer.java:67-69)
    if storage is None:
        storage = "ZARR" if out_path.rstrip("/").lower().endswith(".zarr") else "N5"
    kw = dict(anisotropy_factor=af if preserve_anisotropy else None)
    if storage.upper() == "ZARR":
        if downsamplings:
            raise NotImplementedError("multi-resolution OME-ZARR containers (N5 supports --multiRes here)")
        return bzarr.create_fusion_container_zarr(out_path, os.path.abspath(xml_path), lo.astype(np.int64),
                                                  hi.astype(np.int64), block_size, dtype, min_intensity, max_intensity, **kw)
    return bn5.create_fusion_container(out_path, os.path.abspath(xml_path), lo.astype(np.int64), hi.astype(np.int64),
                                       block_size, dtype, min_intensity, max_intensity, downsamplings=downsamplings, **kw)


def affine_fusion(out_path, ctx: Context, fusion_type="AVG_BLEND", block_scale=(2, 2, 1), channel=0, timepoint=0,
                  retries=5):
    """`./affine-fusion -o fused.n5 [-f AVG_BLEND] [--blockScale 2,2,1]`: read the container
    metadata, fuse every super-block on the device and write it with N5Utils.saveBlock semantics."""
    is_zarr = os.path.exists(os.path.join(out_path, ".zgroup"))
    store, meta = (bzarr.read_fusion_container_zarr if is_zarr else bn5.read_fusion_container)(out_path)
    data = SpimData2.load(meta["input_xml"])
    fmt, n5_in = data.image_loader()
    src = bn5.N5Store(n5_in)
    view_ids = [v for v in data.view_ids() if v[0] == data.timepoints[timepoint]]
    images = _load_views(data, src, view_ids)
    regs = bf.adjust_all_transforms({v: data.model(*v) for v in view_ids},
                                    meta["anisotropy_factor"] if meta["preserve_anisotropy"] else float("nan"))
    bb_min, bb_max = meta["bb_min"], meta["bb_max"]
    dims = [bb_max[d] - bb_min[d] + 1 for d in range(3)]
    supplier = bf.BlkAffineFusion.init(ctx, images, regs, fusion_type, 1, (bb_min, bb_max), meta["dtype"],
                                       meta["min_intensity"], meta["max_intensity"])
    dataset = meta["mr_infos"][0 if is_zarr else channel + timepoint * meta["num_channels"]][0]["dataset"]

    def sink(grid_block, block):
        if is_zarr:   # 5-D grid offset {gx, gy, gz, c, t} (J/SparkAffineFusion.java:630-643)
            store.save_block(dataset, block, tuple(grid_block[2]) + (channel, timepoint))
        else:
            store.save_block(dataset, block, grid_block[2])

    levels = meta["mr_infos"][0 if is_zarr else channel + timepoint * meta["num_channels"]]
    if len(levels) == 1:
        bf.fuse_volume(supplier, dims, meta["block_size"], block_scale, retries, sink)
    else:
        # multi-resolution container: every super-block is fused into a resident volume and its lower
        # levels are derived on the device (bs_downsample) before anything is downloaded, instead of
        # re-reading s(l-1) from the container per level (J/SparkAffineFusion.java:703-782).  Requires
        # super-blocks whose extent is divisible by the absolute downsampling of the last level.
        bs = meta["block_size"]
        compute = tuple(bs[d] * block_scale[d] for d in range(3))
        np_dt = native._BS2NP[supplier.out_dtype]
        for (off, size, gpos) in bf.grid_create(dims, compute, bs):
            h = supplier.copy_to_volume(off, tuple(off[d] + size[d] - 1 for d in range(3)))
            cur_h, cur_size, cur_off = h, list(size), list(off)
            store.save_block(levels[0]["dataset"], ctx.volume_download(cur_h, cur_size, np_dt), gpos)
            for lv in levels[1:]:
                rel = lv["relativeDownsampling"]
                nh = ctx.downsample(cur_h, rel)
                ctx.volume_free(cur_h)
                cur_h = nh
                cur_size = [cur_size[d] // rel[d] for d in range(3)]
                cur_off = [cur_off[d] // rel[d] for d in range(3)]
                blk = ctx.volume_download(cur_h, cur_size, np_dt)
                # a lower-level piece generally covers only part of a storage block (and may not start
                # on a block boundary): read-modify-write of full-extent blocks
                _write_region(store, lv["dataset"], blk, cur_off)
            ctx.volume_free(cur_h)
    for h in supplier.handles.values():
        ctx.volume_free(h)
    return dataset


def _write_region(store, dataset, block, off_xyz):
    """Write a [z,y,x] region at an arbitrary voxel offset (read-modify-write of the touched blocks)."""
    a = store.dataset_attributes(dataset)
    bs, dims = a["blockSize"], a["dimensions"]
    z, y, x = block.shape
    lo = [off_xyz[0], off_xyz[1], off_xyz[2]]
    hi = [min(lo[0] + x, dims[0]), min(lo[1] + y, dims[1]), min(lo[2] + z, dims[2])]
    for gz in range(lo[2] // bs[2], -(-hi[2] // bs[2])):
        for gy in range(lo[1] // bs[1], -(-hi[1] // bs[1])):
            for gx in range(lo[0] // bs[0], -(-hi[0] // bs[0])):
                b0 = [gx * bs[0], gy * bs[1], gz * bs[2]]
                ext = [min(bs[d], dims[d] - b0[d]) for d in range(3)]
                cur = store.read_block(dataset, (gx, gy, gz))
                if cur is None or list(cur.shape[::-1]) != ext:
                    cur = np.zeros(ext[::-1], dtype=block.dtype)
                s0 = [max(lo[d], b0[d]) for d in range(3)]
                s1 = [min(hi[d], b0[d] + ext[d]) for d in range(3)]
                cur[s0[2] - b0[2]:s1[2] - b0[2], s0[1] - b0[1]:s1[1] - b0[1], s0[0] - b0[0]:s1[0] - b0[0]] = \
                    block[s0[2] - lo[2]:s1[2] - lo[2], s0[1] - lo[1]:s1[1] - lo[1], s0[0] - lo[0]:s1[0] - lo[0]]
                store.write_block(dataset, (gx, gy, gz), cur)
