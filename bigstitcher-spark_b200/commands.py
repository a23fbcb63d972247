"""The `call()` bodies of the reference's `stitching`, `create-fusion-container` and
`affine-fusion` commands with the Spark RDD collapsed to a plain host work queue over one
`native.Context` (BASELINE north_star), SpimData2 XML in, N5 / OME-Zarr blocks out.

  stitching               J/SparkPairwiseStitching.java:110-392
  create_fusion_container J/CreateFusionContainer.java:122-519   (N5 or OME-ZARR, all channels / timepoints, pyramid)
  affine_fusion           J/SparkAffineFusion.java:179-800       (s0 + multi-resolution pyramid)

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
              min_r=0.3, max_r=1.0, max_shift_xyz=None, max_shift_total=None, dry_run=False,
              channel_combine="AVERAGE", illum_combine="PICK_BRIGHTEST", shard=(0, 1), allgather=None, view_selection=None):
    """`./stitching -x dataset.xml [-ds 2,2,1] [-p 5] [--channelCombine AVERAGE] [--illumCombine PICK_BRIGHTEST] ...`:
    phase-correlate every overlapping pair of tile GROUPS (a tile's channels / illuminations are combined,
    J/SparkPairwiseStitching.java:103-107,141-165,204-208) and store the filtered results in the XML's
    <StitchingResults>.  Returns all raw results.  ``view_selection``: keyword arguments of SpimData2.select_views
    (`--angleId`, `--tileId`, `--channelId`, `--illuminationId`, `--timepointId` or `-vi`, default all views).

    Multi-GPU: pairs are independent, so rank r of w takes pairs[r::w] (``shard``) with no data-path
    collective; ``allgather(obj) -> [obj of every rank]`` (e.g. torch.distributed.all_gather_object) merges the
    20-doubles-per-pair results and rank 0 writes the XML."""
    data = SpimData2.load(xml_path)
    fmt, n5_path = data.image_loader()
    if fmt != "bdv.n5":
        raise NotImplementedError(f"ImageLoader format {fmt}")
    store = bn5.N5Store(n5_path)
    selected = data.select_views(**view_selection) if view_selection else None
    all_pairs = data.stitching_groups(selected)
    rank, world = shard
    pairs = all_pairs[rank::world]
    needed = sorted({v for p in pairs for g in p for v in g})
    tiles = _load_views(data, store, needed)
    models = {v: data.model(*v) for v in needed}
    attributes = {v: data.setups[v[1]].attributes for v in needed}
    params = bst.PairwiseStitchingParameters(peaks_to_check=peaks_to_check, do_subpixel=not disable_subpixel)
    raw = bst.stitch_pairs(pairs, tiles, models, params, downsampling, ctx, attributes=attributes,
                           channel_combine=channel_combine, illum_combine=illum_combine)
    for (ga, gb), r in zip(pairs, raw):
        if r is not None:   # hash of the FIRST views' registrations (J/SparkPairwiseStitching.java:287-289)
            r.hash = SpimData2.transform_hash(data.registrations[ga[0]], data.registrations[gb[0]])
    if world > 1:
        parts = allgather(raw)
        merged = [None] * len(all_pairs)
        for r_, part in enumerate(parts):
            merged[r_::world] = part
        raw, pairs = merged, all_pairs
        if rank != 0:
            return raw
    # every compared pair loses its stored result (a->b and b->a), found or not (:323-325); then the new ones go in
    data.remove_stitching_results(pairs)
    kept = bst.filter_results(raw, min_r, max_r, max_shift_xyz, max_shift_total)
    data.set_stitching_results([dict(pair=r.pair, shift=r.transform, r=r.r, hash=r.hash, bbox_min=r.bbox_min,
                                     bbox_max=r.bbox_max) for r in kept])
    if not dry_run:
        data.save(xml_path)
    return raw


def estimate_multires_pyramid(dims_xyz, anisotropy_factor=float("nan"), min_size=64):
    """`--multiRes`: ExportN5Api.estimateMultiResPyramid (mvrecon; recalled, PARITY_GAPS): keep halving the axes
    whose accumulated voxel size is the smallest (so anisotropic z catches up) until no axis could be halved without
    dropping below ``min_size`` voxels.  Returns RELATIVE steps after s0."""
    cur = [int(v) for v in dims_xyz]
    voxel = [1.0, 1.0, 1.0 if math.isnan(anisotropy_factor) else float(anisotropy_factor)]
    steps = []
    while True:
        smallest = min(voxel)
        rel = [2 if (voxel[d] <= smallest * 1.5 and cur[d] // 2 >= min_size) else 1 for d in range(3)]
        if rel == [1, 1, 1]:
            break
        steps.append(tuple(rel))
        cur = [cur[d] // rel[d] for d in range(3)]
        voxel = [voxel[d] * rel[d] for d in range(3)]
    return steps


def create_fusion_container(xml_path, out_path, block_size=(128, 128, 128), dtype="float32", min_intensity=None,
                            max_intensity=None, preserve_anisotropy=False, anisotropy_factor=float("nan"),
                            storage=None, downsamplings=(), multi_res=False, compression="zstd"):
    """`./create-fusion-container -x dataset.xml -o fused.zarr [-s ZARR|N5] -d FLOAT32 --blockSize ... [--multiRes |
    -ds ...] [-c Zstandard]`: bounding box of all views (Import.getBoundingBox, J/CreateFusionContainer.java:184-211),
    NumChannels / NumTimepoints from the XML (:213-216), datasets + metadata.  Defaults follow the reference: OME-ZARR
    storage unless the path ends in .n5 (:67-69), Zstandard compression (:71-76).  ``downsamplings``: relative steps."""
    data = SpimData2.load(xml_path)
    lo = np.full(3, np.inf)
    hi = np.full(3, -np.inf)
    af = anisotropy_factor if preserve_anisotropy else float("nan")
    regs = bf.adjust_all_transforms({v: data.model(*v) for v in data.view_ids()}, af)
    for v, M in regs.items():
        bmin, bmax = bf.transformed_bounding_box(data.setups[v[1]].size, M)
        lo = np.minimum(lo, bmin)
        hi = np.maximum(hi, bmax)
    if storage is None:   # guessed from the extension like the reference (J/SparkAffineFusion.java:206-225)
        storage = "N5" if out_path.rstrip("/").lower().endswith(".n5") else "ZARR"
    dims = [int(hi[d] - lo[d] + 1) for d in range(3)]
    if multi_res and not downsamplings:
        downsamplings = estimate_multires_pyramid(dims, af)
    if dtype != "float32":   # J/CreateFusionContainer.java:226-242
        top = 255.0 if dtype == "uint8" else 65535.0
        min_intensity = 0.0 if min_intensity is None else min_intensity
        max_intensity = top if max_intensity is None else max_intensity
    kw = dict(anisotropy_factor=af if preserve_anisotropy else None, num_channels=len(data.channels_ordered()),
              num_timepoints=len(data.timepoints), compression=compression, downsamplings=downsamplings)
    if storage.upper() == "ZARR":
        return bzarr.create_fusion_container_zarr(out_path, os.path.abspath(xml_path), lo.astype(np.int64),
                                                  hi.astype(np.int64), block_size, dtype, min_intensity, max_intensity, **kw)
    return bn5.create_fusion_container(out_path, os.path.abspath(xml_path), lo.astype(np.int64), hi.astype(np.int64),
                                       block_size, dtype, min_intensity, max_intensity, **kw)


# --------------------------------------------------------------------------------------------- affine-fusion
class _Sink:
    """N5Utils.saveBlock on the 3-D dataset (N5) or the 5-D view (OME-ZARR, J/SparkAffineFusion.java:630-670)."""

    def __init__(self, store, is_zarr, c, t):
        self.store, self.is_zarr, self.c, self.t = store, is_zarr, c, t

    def save(self, dataset, block, grid_pos):
        if self.is_zarr:
            self.store.save_block(dataset, block, tuple(grid_pos) + (self.c, self.t))
        else:
            self.store.save_block(dataset, block, grid_pos)

    def save_chunk(self, dataset, block, chunk_pos):
        """One storage block / chunk exactly as the device packed it (N5: already big-endian)."""
        if self.is_zarr:
            gx, gy, gz = chunk_pos
            self.store.write_chunk(dataset, (self.t, self.c, gz, gy, gx), block)
        else:
            self.store.write_block(dataset, chunk_pos, block)

    def read_region(self, dataset, mn, size):
        if self.is_zarr:
            return self.store.read_region(dataset, mn, size, self.c, self.t)
        return self.store.read_region(dataset, mn, size)

    def write_region(self, dataset, block, off_xyz, dims, bs):
        """[z,y,x] region at a block-aligned voxel offset (read-modify-write when it covers blocks only partly)."""
        z, y, x = block.shape
        for gz in range(off_xyz[2] // bs[2], -(-(off_xyz[2] + z) // bs[2])):
            for gy in range(off_xyz[1] // bs[1], -(-(off_xyz[1] + y) // bs[1])):
                for gx in range(off_xyz[0] // bs[0], -(-(off_xyz[0] + x) // bs[0])):
                    b0 = [gx * bs[0], gy * bs[1], gz * bs[2]]
                    ext = [min(bs[d], dims[d] - b0[d]) for d in range(3)]
                    if any(e <= 0 for e in ext):
                        continue
                    s0 = [max(off_xyz[d], b0[d]) for d in range(3)]
                    s1 = [min(off_xyz[0] + x, b0[0] + ext[0]), min(off_xyz[1] + y, b0[1] + ext[1]), min(off_xyz[2] + z, b0[2] + ext[2])]
                    part = block[s0[2] - off_xyz[2]:s1[2] - off_xyz[2], s0[1] - off_xyz[1]:s1[1] - off_xyz[1],
                                 s0[0] - off_xyz[0]:s1[0] - off_xyz[0]]
                    if list(part.shape[::-1]) == ext:
                        cur = part
                    else:
                        cur = self.read_region(dataset, b0, ext)
                        cur[s0[2] - b0[2]:s1[2] - b0[2], s0[1] - b0[1]:s1[1] - b0[1], s0[0] - b0[0]:s1[0] - b0[0]] = part
                    self.save(dataset, np.ascontiguousarray(cur), (gx, gy, gz))


def _mipmap_info(src: bn5.N5Store, setup: int):
    """downsamplingFactors of a BDV-N5 setup and the default mipmap transforms (scale f, shift (f - 1) / 2)."""
    a = src.get_attributes(f"setup{setup}")
    factors = [tuple(int(v) for v in f) for f in a.get("downsamplingFactors", [[1, 1, 1]])]
    return factors, [bzarr.mipmap_transform_default(f) for f in factors]


def _compose(reg, mt):
    R = np.vstack([np.asarray(reg, dtype=np.float64).reshape(3, 4), [0, 0, 0, 1]])
    M = np.vstack([np.asarray(mt, dtype=np.float64).reshape(3, 4), [0, 0, 0, 1]])
    return (R @ M)[:3]


def _source_window(src_to_world, level_dims, wmin, wmax, margin=3):
    """Source-pixel interval of a level volume that the world box [wmin, wmax] can sample (n-linear taps + margin),
    clipped to the volume; x is widened to multiples of 8 voxels (16-byte TMA rows).  None when empty."""
    inv = np.linalg.inv(np.vstack([src_to_world, [0, 0, 0, 1]]))[:3]
    c = np.array([[x, y, z] for x in (wmin[0], wmax[0]) for y in (wmin[1], wmax[1]) for z in (wmin[2], wmax[2])], dtype=np.float64)
    s = c @ inv[:, :3].T + inv[:, 3]
    lo = np.floor(s.min(axis=0)).astype(np.int64) - margin
    hi = np.ceil(s.max(axis=0)).astype(np.int64) + margin
    dims = np.asarray(level_dims, dtype=np.int64)
    lo = np.maximum(lo, 0)
    hi = np.minimum(hi, dims - 1)
    if np.any(hi < lo):
        return None
    lo[0] = (lo[0] // 8) * 8
    hi[0] = min(dims[0] - 1, (hi[0] // 8) * 8 + 7)
    return lo, hi - lo + 1


def affine_fusion(out_path, ctx: Context, fusion_type="AVG_BLEND", block_scale=(2, 2, 1), channel=None, timepoint=None,
                  retries=5, blocks_per_call=16, interpolation=1, shard=(0, 1), barrier=None, masks=False,
                  mask_offset=(0.0, 0.0, 0.0), view_selection=None):
    """`./affine-fusion -o fused.zarr [-f AVG_BLEND] [--blockScale 2,2,1] [-c channelIndex] [-t timepointIndex]
    [--masks [--maskOffset x,y,z]]`:
    read the container metadata and, for every (channel, timepoint) volume (J/SparkAffineFusion.java:425-440), fuse
    its views super-block by super-block on the device, write the blocks with N5Utils.saveBlock semantics and build
    the multi-resolution pyramid (:703-782).  Returns the list of s0 datasets written.

    Source staging is block-wise (OverlappingBlocks / ViewUtil.findOverlappingBlocks, J/fusion/OverlappingBlocks.java:
    133-161): the super-block grid is walked in z-slabs; for every slab only the source WINDOW of each overlapping
    view is read from its container -- at the mipmap level ViewUtil's best-resolution rule picks
    (J/util/ViewUtil.java:425-493) -- uploaded as a windowed view and freed after the slab.

    ``view_selection``: keyword arguments of SpimData2.select_views (the AbstractSelectableViews flags); every
    (channel, timepoint) volume fuses its views out of that selection (J/SparkAffineFusion.java:425-440).

    ``masks``: save only the coverage masks (J/SparkAffineFusion.java:564-578, GenerateComputeBlockMasks): no image
    data is read, a voxel is 255 / 65535 / 1.0 where any view's pixel grid (grown by ``mask_offset`` source pixels)
    covers it; the pyramid is then built from that s0 as usual.

    Multi-GPU ("one N5 block-grid slab per device"): rank r of w (``shard``) fuses a contiguous run of
    z-slabs, no data-path collective; ``barrier()`` separates s0 from the pyramid levels that re-read it."""
    is_zarr = os.path.exists(os.path.join(out_path, ".zgroup"))
    store, meta = (bzarr.read_fusion_container_zarr if is_zarr else bn5.read_fusion_container)(out_path)
    data = SpimData2.load(meta["input_xml"])
    fmt, n5_in = data.image_loader()
    if fmt != "bdv.n5":
        raise NotImplementedError(f"ImageLoader format {fmt}")
    src = bn5.N5Store(n5_in)
    nc, nt = int(meta["num_channels"]), int(meta["num_timepoints"])
    if nc != len(data.channels_ordered()) or nt != len(data.timepoints):
        raise ValueError(f"container says {nc} channel(s) / {nt} timepoint(s), the XML has "
                         f"{len(data.channels_ordered())} / {len(data.timepoints)}")
    written = []
    done = set()
    selected = data.select_views(**view_selection) if view_selection else None
    for c in range(nc):
        for t in range(nt):
            ci = c if channel is None else int(channel)
            ti = t if timepoint is None else int(timepoint)
            if (ci, ti) in done:
                continue
            done.add((ci, ti))
            levels = meta["mr_infos"][0 if is_zarr else ci + ti * nc]
            vol_views = data.views_of(ci, ti, selected)
            if not vol_views:
                continue                                             # nothing selected for this volume
            _fuse_volume_blockwise(ctx, data, src, _Sink(store, is_zarr, ci, ti), meta, levels, vol_views,
                                   fusion_type, interpolation, block_scale, retries, blocks_per_call, shard, barrier,
                                   masks, mask_offset)
            written.append(levels[0]["dataset"])
    return written


def _fuse_volume_blockwise(ctx, data, src, sink, meta, levels, view_ids, fusion_type, interpolation, block_scale,
                           retries, blocks_per_call, shard=(0, 1), barrier=None, masks=False, mask_offset=(0.0, 0.0, 0.0)):
    af = meta["anisotropy_factor"] if meta["preserve_anisotropy"] else float("nan")
    regs = bf.adjust_all_transforms({v: data.model(*v) for v in view_ids}, af)
    bb_min = np.asarray(meta["bb_min"], dtype=np.int64)
    dims = [int(meta["bb_max"][d] - meta["bb_min"][d] + 1) for d in range(3)]
    bs = [int(v) for v in meta["block_size"]]
    compute = tuple(bs[d] * int(block_scale[d]) for d in range(3))
    ft = native.FUSION_TYPES[fusion_type] if isinstance(fusion_type, str) else int(fusion_type)
    od = {"float32": native.DTYPE_F32, "uint16": native.DTYPE_U16, "uint8": native.DTYPE_U8}[meta["dtype"]]
    params = ctx.fuse_params(ft, interpolation, od, 0, float(meta["min_intensity"] or 0.0), float(meta["max_intensity"] or 65535.0))
    np_dt = native._BS2NP[od]
    s0 = levels[0]["dataset"]
    # blocks that go straight to the container leave the device as STORAGE chunks in the container's byte order
    # (N5 payloads are big-endian): the host neither re-strides nor swaps them
    chunk_params = ctx.fuse_params(ft, interpolation, od, 0, float(meta["min_intensity"] or 0.0),
                                   float(meta["max_intensity"] or 65535.0), out_big_endian=not sink.is_zarr)
    # the fusion kernels work on 64 x 16 x 8 output tiles anchored at the block origin: storage blocks that are whole
    # tiles (the usual 128^3 / 64^3) are packed by the device, smaller ones would leave most of every tile empty
    pack_on_device = bs[0] % 64 == 0 and bs[1] % 16 == 0 and bs[2] % 8 == 0

    # ---- per view: mipmap level by the reference's rule, level volume size, source -> world of that level
    info = {}
    for v in view_ids:
        factors, mts = _mipmap_info(src, v[1])
        lvl = bf.best_mipmap_level(regs[v], factors, mts)
        lvl_dims = src.dataset_attributes(bn5.bdv_dataset(v[1], v[0], lvl))["dimensions"]
        m = _compose(regs[v], mts[lvl])
        info[v] = dict(level=lvl, dims=tuple(int(d) for d in lvl_dims), model=m, blending=bf.adjust_blending(m),
                       dtype=src.dataset_attributes(bn5.bdv_dataset(v[1], v[0], lvl))["dataType"])
    vdims = {v: info[v]["dims"] for v in view_ids}
    vregs = {v: info[v]["model"] for v in view_ids}
    windowed_ok = (ft in (native.FUSE_AVG, native.FUSE_AVG_BLEND) and interpolation == 1 and
                   all(info[v]["dtype"] == "uint16" and info[v]["dims"][0] % 8 == 0 for v in view_ids))
    content = ft in (native.FUSE_AVG_CONTENT, native.FUSE_AVG_BLEND_CONTENT)

    # pyramid straight from the resident fused block when every super-block maps onto whole voxels of every level
    abs_last = [int(v) for v in levels[-1]["absoluteDownsampling"][:3]]
    fast_pyramid = len(levels) > 1 and all(compute[d] % abs_last[d] == 0 for d in range(3)) and not masks

    grid = bf.grid_create(dims, compute, bs)
    slabs = {}
    for gb in grid:
        slabs.setdefault(gb[0][2], []).append(gb)
    rank, world = shard
    zs = sorted(slabs)
    per = -(-len(zs) // world)
    my_z = zs[rank * per:(rank + 1) * per]           # contiguous run of z-slabs per device
    whole = {}      # fallback residency (content weights, winner types, float sources ...): whole views, kept
    if masks:
        # geometry only: full-resolution registrations and view sizes (GenerateComputeBlockMasks.java:119-128)
        fdims = {v: tuple(int(d) for d in data.setups[v[1]].size) for v in view_ids}
        for z0 in my_z:
            todo, attempt = list(slabs[z0]), 0
            while todo:
                attempt += 1
                if attempt > retries:
                    raise RuntimeError(f"masks: {len(todo)} block(s) still failing after {retries} attempts")
                failed = []
                for gb in todo:
                    off, size, gpos = gb
                    mn = tuple(int(v) for v in bb_min + np.asarray(off, dtype=np.int64))
                    sz = tuple(int(v) for v in size)
                    # the views of THIS block (OverlappingViews on the block expanded by 2, like the fusing path): a
                    # mask offset never pulls in a view the block does not overlap
                    vids = bf.find_overlapping_views(fdims, regs, np.asarray(mn), np.asarray(mn) + np.asarray(sz) - 1, sorted(view_ids))
                    gviews = [dict(src_to_world=regs[v], vol_handle=0, full_dims=fdims[v]) for v in vids]
                    try:
                        blk = ctx.mask_blocks(gviews, [mn], [sz], mask_offset, od)[0]
                    except native.BsError:
                        failed.append(gb)
                        continue
                    sink.save(s0, blk, gpos)
                todo = failed
        my_z = []
    try:
        for z0 in my_z:
            blocks = slabs[z0]
            lo = bb_min + np.array([0, 0, z0])
            hi = bb_min + np.array([dims[0] - 1, dims[1] - 1, min(dims[2], z0 + compute[2]) - 1])
            vids = bf.find_overlapping_views(vdims, vregs, lo, hi, view_ids)
            staged, views = {}, {}
            try:
                for v in vids:
                    border, rng = info[v]["blending"]
                    ds_name = bn5.bdv_dataset(v[1], v[0], info[v]["level"])
                    if windowed_ok:
                        w = _source_window(info[v]["model"], info[v]["dims"], lo - bf.AFFINE_EXPANSION, hi + bf.AFFINE_EXPANSION)
                        if w is None:
                            continue
                        wmin, wsize = w
                        staged[v] = ctx.volume_upload(src.read_region(ds_name, wmin, wsize))
                        views[v] = dict(src_to_world=info[v]["model"], vol_handle=staged[v], blend_border=border, blend_range=rng,
                                        full_dims=info[v]["dims"], window_min=tuple(int(x) for x in wmin))
                    else:
                        if v not in whole:
                            h = ctx.volume_upload(src.read_volume(ds_name))
                            whole[v] = (h, ctx.content_weights(h) if content else 0)
                        views[v] = dict(src_to_world=info[v]["model"], vol_handle=whole[v][0], content_handle=whole[v][1],
                                        blend_border=border, blend_range=rng)
                # ---- the slab's super-blocks, `blocks_per_call` per launch, RetryTracker policy (<= 5 attempts)
                todo, attempt = list(blocks), 0
                while todo:
                    attempt += 1
                    if attempt > retries:
                        raise RuntimeError(f"fusion: {len(todo)} block(s) still failing after {retries} attempts")
                    failed = []
                    if fast_pyramid:
                        # every super-block is fused into a RESIDENT volume and its lower levels are derived on the
                        # device (bs_downsample) before anything is downloaded -- level l-1 is never re-read from
                        # the container (the reference re-reads it, J/SparkAffineFusion.java:703-782)
                        for gb in todo:
                            try:
                                _fuse_block_with_pyramid(ctx, gb, bb_min, vdims, vregs, views, params, np_dt, sink, levels, bs)
                            except native.BsError:
                                failed.append(gb)
                    else:
                        for c0 in range(0, len(todo), blocks_per_call):
                            chunk = todo[c0:c0 + blocks_per_call]
                            if not pack_on_device:       # small storage blocks: fuse super-blocks, split on the host
                                try:
                                    outs = _fuse_chunk(ctx, chunk, bb_min, vdims, vregs, views, params, np_dt)
                                except native.BsError:
                                    failed.extend(chunk)
                                    continue
                                for (off, size, gpos), blk in zip(chunk, outs):
                                    sink.save(s0, blk, gpos)
                                continue
                            cells = [cell for gb in chunk for cell in _storage_cells(gb, bs)]
                            try:
                                outs = _fuse_chunk(ctx, cells, bb_min, vdims, vregs, views, chunk_params, np_dt)
                            except native.BsError:
                                failed.extend(chunk)
                                continue
                            for (off, size, gpos), blk in zip(cells, outs):
                                sink.save_chunk(s0, blk, gpos)
                    todo = failed
            finally:
                for h in staged.values():
                    ctx.volume_free(h)
    finally:
        for h, ch in whole.values():
            ctx.volume_free(h)
            if ch:
                ctx.volume_free(ch)

    # ---- pyramid s1 .. sN when super-blocks do not map onto whole voxels of every level: every block of level l is
    # the 2x average of its region of level l-1 (N5ApiTools.writeDownsampledBlock[5dOMEZARR]), read back from the
    # container and averaged on the device (bs_downsample)
    for li in range(1, 1 if fast_pyramid else len(levels)):
        if barrier is not None:
            barrier()                                # level l-1 is complete on every rank
        prev, cur = levels[li - 1], levels[li]
        rel = [int(v) for v in cur["relativeDownsampling"][:3]]
        cdims = [int(v) for v in cur["dimensions"][:3]]
        todo, attempt = bf.grid_create(cdims, compute, bs)[rank::world], 0
        while todo:
            attempt += 1
            if attempt > retries:
                raise RuntimeError(f"pyramid s{li}: {len(todo)} block(s) still failing after {retries} attempts")
            failed = []
            for gb in todo:
                off, size, gpos = gb
                h = h2 = None
                try:
                    srcblk = sink.read_region(prev["dataset"], [off[d] * rel[d] for d in range(3)], [size[d] * rel[d] for d in range(3)])
                    h = ctx.volume_upload(np.ascontiguousarray(srcblk))
                    h2 = ctx.downsample(h, rel)
                    sink.save(cur["dataset"], ctx.volume_download(h2, size, np_dt), gpos)
                except native.BsError:
                    failed.append(gb)
                finally:
                    for hh in (h, h2):
                        if hh is not None:
                            ctx.volume_free(hh)
            todo = failed


def _fuse_block_with_pyramid(ctx, gb, bb_min, vdims, vregs, views, params, np_dt, sink, levels, bs):
    off, size, gpos = gb
    wmin = bb_min + np.asarray(off, dtype=np.int64)
    vids = bf.find_overlapping_views(vdims, vregs, wmin, wmin + np.asarray(size) - 1, sorted(views))
    cur_h = ctx.fuse_block_to_volume([views[v] for v in vids], tuple(int(v) for v in wmin), tuple(int(v) for v in size), params)
    try:
        cur_size, cur_off = [int(v) for v in size], [int(v) for v in off]
        sink.save(levels[0]["dataset"], ctx.volume_download(cur_h, cur_size, np_dt), gpos)
        for lv in levels[1:]:
            rel = [int(v) for v in lv["relativeDownsampling"][:3]]
            nxt = [cur_size[d] // rel[d] for d in range(3)]
            if min(nxt) < 1:
                break        # an edge block thinner than the step: no voxel of this (or any deeper) level
            nh = ctx.downsample(cur_h, rel)
            ctx.volume_free(cur_h)
            cur_h, cur_size, cur_off = nh, nxt, [cur_off[d] // rel[d] for d in range(3)]
            sink.write_region(lv["dataset"], ctx.volume_download(cur_h, cur_size, np_dt), cur_off,
                              [int(v) for v in lv["dimensions"][:3]], bs)
    finally:
        ctx.volume_free(cur_h)


def _storage_cells(gb, bs):
    """The storage blocks of one super-block: [(offset, size, grid position)] in the container's block grid."""
    off, size, gpos = gb
    cells = []
    for kz in range(-(-size[2] // bs[2])):
        for ky in range(-(-size[1] // bs[1])):
            for kx in range(-(-size[0] // bs[0])):
                k = (kx, ky, kz)
                cells.append((tuple(int(off[d] + k[d] * bs[d]) for d in range(3)),
                              tuple(int(min(bs[d], size[d] - k[d] * bs[d])) for d in range(3)),
                              tuple(int(gpos[d] + k[d]) for d in range(3))))
    return cells


def _fuse_chunk(ctx, chunk, bb_min, vdims, vregs, views, params, np_dt):
    mins, sizes = [], []
    lo = np.array([np.inf] * 3)
    hi = np.array([-np.inf] * 3)
    for (off, size, _) in chunk:
        wmin = bb_min + np.asarray(off, dtype=np.int64)
        mins.append(tuple(int(v) for v in wmin))
        sizes.append(tuple(int(v) for v in size))
        lo = np.minimum(lo, wmin)
        hi = np.maximum(hi, wmin + np.asarray(size) - 1)
    vids = [v for v in bf.find_overlapping_views(vdims, vregs, lo.astype(np.int64), hi.astype(np.int64), sorted(views))]
    return ctx.fuse_blocks([views[v] for v in vids], mins, sizes, params)
