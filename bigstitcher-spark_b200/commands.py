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
                            max_intensity=None, preserve_anisotropy=False, anisotropy_factor=float("nan")):
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
    return bn5.create_fusion_container(out_path, os.path.abspath(xml_path), lo.astype(np.int64), hi.astype(np.int64),
                                       block_size, dtype, min_intensity, max_intensity,
                                       anisotropy_factor=af if preserve_anisotropy else None)


def affine_fusion(out_path, ctx: Context, fusion_type="AVG_BLEND", block_scale=(2, 2, 1), channel=0, timepoint=0,
                  retries=5):
    """`./affine-fusion -o fused.n5 [-f AVG_BLEND] [--blockScale 2,2,1]`: read the container
    metadata, fuse every super-block on the device and write it with N5Utils.saveBlock semantics."""
    store, meta = bn5.read_fusion_container(out_path)
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
    dataset = meta["mr_infos"][channel + timepoint * meta["num_channels"]][0]["dataset"]

    def sink(grid_block, block):
        store.save_block(dataset, block, grid_block[2])

    bf.fuse_volume(supplier, dims, meta["block_size"], block_scale, retries, sink)
    for h in supplier.handles.values():
        ctx.volume_free(h)
    return dataset
