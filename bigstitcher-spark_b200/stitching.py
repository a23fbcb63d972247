"""Host-side mirror of the reference's pairwise-stitching operator interface.

Mirrors, name for name, what ``SparkPairwiseStitching``'s per-pair task calls
(src/main/java/net/preibisch/bigstitcher/spark/SparkPairwiseStitching.java:194-303):

    PairwiseStitchingParameters          (:200-202)
    TransformationTools.computeStitching (:247-255)   -> compute_stitching
    PairwiseStitching.getShift (upstream BigStitcher 2.5.0) -> get_shift
    result record / filters              (:284-301, :347-380) -> PairwiseStitchingResult, filter_results

Only geometry and bookkeeping happen here (overlap boxes, crops, sign conventions, result
records); all arithmetic on voxels runs in libbsgpu.so through ``native.Context``.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

from .native import Context


@dataclass
class PairwiseStitchingParameters:
    """net.preibisch.stitcher.algorithm.PairwiseStitchingParameters (defaults recalled:
    minOverlap 0.25, peaksToCheck 5, doSubpixel true, interpolateCrossCorrelation false)."""
    min_overlap: float = 0.25
    peaks_to_check: int = 5
    do_subpixel: bool = True
    interpolate_cross_correlation: bool = False
    extension: tuple = (10, 10, 10)


@dataclass
class PairwiseStitchingResult:
    """Spark.SerializablePairwiseStitchingResult (J/util/Spark.java:201-233): 3x4 double
    affine, r, bounding box min/max; ``hash`` is computed by the Java side
    (PairwiseStitchingResult.calculateHash, J/SparkPairwiseStitching.java:287-289)."""
    pair: tuple
    transform: np.ndarray          # 3x4 row-packed, global coordinates
    r: float
    bbox_min: tuple
    bbox_max: tuple
    hash: float = 0.0
    shift_px: tuple = (0.0, 0.0, 0.0)  # correction of B in (downsampled) pixel units, diagnostic


def _overlap(min1, max1, min2, max2):
    lo = np.maximum(min1, min2)
    hi = np.minimum(max1, max2)
    if np.any(hi < lo):
        return None
    return lo, hi


def local_raster_overlaps(dims1_xyz, dims2_xyz, t1, t2):
    """TransformTools.applyTranslation / getOverlap / getLocalOverlap / getLocalRasterOverlap:
    returns (interval1_min, interval2_min, size, sub1, sub2) in xyz or None.  Raster interval
    = ceil(min) .. floor(max) of the real local overlap; sub* = interval.min - localOverlap.min."""
    t1 = np.asarray(t1, dtype=np.float64)
    t2 = np.asarray(t2, dtype=np.float64)
    d1 = np.asarray(dims1_xyz, dtype=np.float64)
    d2 = np.asarray(dims2_xyz, dtype=np.float64)
    ov = _overlap(t1, t1 + d1 - 1, t2, t2 + d2 - 1)
    if ov is None:
        return None
    lo, hi = ov
    l1lo, l1hi = lo - t1, hi - t1
    l2lo, l2hi = lo - t2, hi - t2
    i1lo, i1hi = np.ceil(l1lo - 1e-9), np.floor(l1hi + 1e-9)
    i2lo, i2hi = np.ceil(l2lo - 1e-9), np.floor(l2hi + 1e-9)
    s1 = i1hi - i1lo + 1
    s2 = i2hi - i2lo + 1
    if np.any(s1 <= 0) or np.any(s2 <= 0) or np.any(s1 != s2):
        return None  # "0-sized or unequal overlap" -> null
    return (i1lo.astype(np.int64), i2lo.astype(np.int64), s1.astype(np.int64), i1lo - l1lo, i2lo - l2lo)


def get_shift(img1, img2, t1, t2, params: PairwiseStitchingParameters, ctx: Context):
    """PairwiseStitching.getShift: phase-correlate the overlapping parts of two images
    ([z,y,x] numpy arrays or device tensors) positioned at translations t1, t2 (xyz, pixel units
    of the images).  Returns (shift_xyz, r) -- the correction of image 2's position relative to
    image 1 in pixels (planted error (+3,-2,+1) is recovered as (3,-2,1))
    -- or None ("no shift found")."""
    dims1 = tuple(img1.shape)[::-1]
    dims2 = tuple(img2.shape)[::-1]
    ro = local_raster_overlaps(dims1, dims2, t1, t2)
    if ro is None:
        return None
    a1, a2, size, sub1, sub2 = ro
    c1 = img1[a1[2]:a1[2] + size[2], a1[1]:a1[1] + size[1], a1[0]:a1[0] + size[0]]
    c2 = img2[a2[2]:a2[2] + size[2], a2[1]:a2[1] + size[1], a2[0]:a2[0] + size[0]]
    if isinstance(c1, np.ndarray):
        c1 = np.ascontiguousarray(c1)
        c2 = np.ascontiguousarray(c2)
    else:
        c1 = c1.contiguous()
        c2 = c2.contiguous()
    p = ctx.pcm_params(params.peaks_to_check, params.do_subpixel, params.min_overlap, params.extension)
    res = ctx.pcm_pair(c1, c2, p)
    if not res.found or math.isinf(res.r):
        return None
    s = np.asarray(res.shift_sub if params.do_subpixel else res.shift_int, dtype=np.float64)
    # correct for the int/real coordinate difference of the two raster crops
    shift = s - (np.asarray(sub2) - np.asarray(sub1))
    return shift, res.r


def non_translations_equal(m1, m2, eps=1e-9):
    """TransformTools.nonTranslationsEqual: the 3x3 parts agree."""
    a = np.asarray(m1, dtype=np.float64).reshape(3, 4)[:, :3]
    b = np.asarray(m2, dtype=np.float64).reshape(3, 4)[:, :3]
    return bool(np.all(np.abs(a - b) < eps))


def downsample_avg(img: np.ndarray, factors_xyz):
    """2^k block averaging per axis (GroupedViewAggregator's remaining downsampling steps);
    host plumbing, output float32 when any factor > 1."""
    fx, fy, fz = (int(f) for f in factors_xyz)
    if (fx, fy, fz) == (1, 1, 1):
        return img
    z, y, x = img.shape
    z2, y2, x2 = z // fz, y // fy, x // fx
    v = img[:z2 * fz, :y2 * fy, :x2 * fx].astype(np.float32)
    v = v.reshape(z2, fz, y2, fy, x2, fx).mean(axis=(1, 3, 5), dtype=np.float32)
    return v


def aggregate_group(images, action="AVERAGE"):
    """GroupedViewAggregator for one attribute (--channelCombine / --illumCombine,
    J/SparkPairwiseStitching.java:103-107,204-208): AVERAGE -> float32 mean of the group's images,
    PICK_BRIGHTEST -> the image with the highest mean intensity.  Single-view groups pass through."""
    images = list(images)
    if len(images) == 1:
        return images[0]
    if action == "AVERAGE":
        acc = np.zeros(images[0].shape, dtype=np.float32)
        for im in images:
            acc += im.astype(np.float32)
        return acc / np.float32(len(images))
    if action == "PICK_BRIGHTEST":
        return images[int(np.argmax([float(im.mean(dtype=np.float64)) for im in images]))]
    raise ValueError(f"unknown ActionType {action}")


def _world_box(m, dims_xyz):
    c = np.array([[x, y, z] for x in (0, dims_xyz[0] - 1) for y in (0, dims_xyz[1] - 1)
                  for z in (0, dims_xyz[2] - 1)], dtype=np.float64)
    w = c @ m[:, :3].T + m[:, 3]
    return w.min(axis=0), w.max(axis=0)


def compute_stitching(img_a, img_b, model_a, model_b, params: PairwiseStitchingParameters,
                      downsample_factors=(1, 1, 1), ctx: Context | None = None):
    """TransformationTools.computeStitching for single-view groups whose registrations differ
    only by translation (the ``nonTranslationsEqual`` branch, J/SparkPairwiseStitching.java:216-255).

    model_a / model_b: row-packed 3x4 view models (source pixel -> world).  Returns
    ((resTransform 3x4, r), (bbox_min, bbox_max)) or None.  resTransform = M_B * T(shift*ds) *
    M_B^-1 in global coordinates."""
    ma = np.asarray(model_a, dtype=np.float64).reshape(3, 4)
    mb = np.asarray(model_b, dtype=np.float64).reshape(3, 4)
    if not non_translations_equal(ma, mb):
        return compute_stitching_non_equal_transformations(img_a, img_b, ma, mb, params, downsample_factors, ctx)
    ds = np.asarray(downsample_factors, dtype=np.float64)
    a = downsample_avg(img_a, downsample_factors) if isinstance(img_a, np.ndarray) else img_a
    b = downsample_avg(img_b, downsample_factors) if isinstance(img_b, np.ndarray) else img_b
    # TransformTools.getInitialTransforms: translation part expressed in (downsampled) pixel units
    lin_inv = np.linalg.inv(ma[:, :3])
    t1 = (lin_inv @ ma[:, 3]) / ds
    t2 = (lin_inv @ mb[:, 3]) / ds
    # world-space overlap bounding box of the two views (BoundingBoxMaximalGroupOverlap)
    la, ha = _world_box(ma, tuple(img_a.shape)[::-1])
    lb, hb = _world_box(mb, tuple(img_b.shape)[::-1])
    ov = _overlap(la, ha, lb, hb)
    if ov is None:
        return None
    res = get_shift(a, b, t1, t2, params, ctx)
    if res is None:
        return None
    shift, r = res
    T = np.eye(4)
    T[:3, 3] = shift * ds
    Mb = np.vstack([mb, [0, 0, 0, 1]])
    R = Mb @ T @ np.linalg.inv(Mb)
    return (R[:3, :].copy(), float(r)), (tuple(ov[0]), tuple(ov[1]))


def compute_stitching_non_equal_transformations(img_a, img_b, model_a, model_b, params: PairwiseStitchingParameters,
                                                downsample_factors=(1, 1, 1), ctx: Context | None = None):
    """TransformationTools.computeStitchingNonEqualTransformations (J/SparkPairwiseStitching.java:259-267;
    SURVEY.md 8a row a3'): when the non-translation parts differ, both views are virtually fused
    (n-linear resampling, no blending) into their common world-space overlap box on the
    ``downsample_factors`` grid -- here with the same fusion kernel that serves `affine-fusion` --
    and the two rendered volumes are phase-correlated with zero initial translations.  The shift is
    already a world-space translation: resTransform = T(shift * ds)."""
    from . import native
    ma = np.asarray(model_a, dtype=np.float64).reshape(3, 4)
    mb = np.asarray(model_b, dtype=np.float64).reshape(3, 4)
    ds = np.asarray(downsample_factors, dtype=np.float64)
    la, ha = _world_box(ma, tuple(img_a.shape)[::-1])
    lb, hb = _world_box(mb, tuple(img_b.shape)[::-1])
    ov = _overlap(la, ha, lb, hb)
    if ov is None:
        return None
    lo = np.ceil(ov[0] / ds).astype(np.int64)          # overlap box on the downsampled world grid
    hi = np.floor(ov[1] / ds).astype(np.int64)
    size = hi - lo + 1
    if np.any(size <= 0):
        return None
    S = np.diag(np.concatenate([1.0 / ds, [1.0]]))       # world -> downsampled world
    p = ctx.fuse_params("AVG", 1, native.DTYPE_F32)
    rendered = []
    for img, m in ((img_a, ma), (img_b, mb)):
        h = ctx.volume_upload(np.ascontiguousarray(img))
        try:
            m_ds = (S @ np.vstack([m, [0, 0, 0, 1]]))[:3]
            rendered.append(ctx.fuse_block([dict(src_to_world=m_ds, vol_handle=h)], lo, size, p))
        finally:
            ctx.volume_free(h)
    pp = ctx.pcm_params(params.peaks_to_check, params.do_subpixel, params.min_overlap, params.extension)
    res = ctx.pcm_pair(rendered[0], rendered[1], pp)
    if not res.found or math.isinf(res.r):
        return None
    shift = np.asarray(res.shift_sub if params.do_subpixel else res.shift_int, dtype=np.float64)
    R = np.eye(4)
    R[:3, 3] = shift * ds
    return (R[:3, :].copy(), float(res.r)), (tuple(ov[0]), tuple(ov[1]))


def filter_results(results, min_r=0.3, max_r=1.0, max_shift_xyz=None, max_shift_total=None):
    """FilteredStitchingResults with CorrelationFilter / AbsoluteShiftFilter / ShiftMagnitudeFilter
    (J/SparkPairwiseStitching.java:347-380; defaults --minR 0.3 --maxR 1.0 :85-89)."""
    kept = []
    for res in results:
        if res is None:
            continue
        if not (min_r <= res.r <= max_r):
            continue
        t = np.asarray(res.transform).reshape(3, 4)[:, 3]
        if max_shift_xyz is not None and np.any(np.abs(t) > np.asarray(max_shift_xyz)):
            continue
        if max_shift_total is not None and float(np.linalg.norm(t)) > max_shift_total:
            continue
        kept.append(res)
    return kept


def aggregate_views(images: dict, attributes: dict, channel_combine="AVERAGE", illum_combine="PICK_BRIGHTEST"):
    """GroupedViewAggregator with the reference's two actions in its order (J/SparkPairwiseStitching.java:204-208):
    first the illuminations of every channel are combined (default PICK_BRIGHTEST), then the channels (default
    AVERAGE).  images: ViewId -> [z,y,x] array of one group; attributes: ViewId -> {"channel": c, "illumination": i}."""
    by_channel = {}
    for vid in sorted(images):
        by_channel.setdefault(attributes.get(vid, {}).get("channel", 0), []).append(images[vid])
    per_channel = [aggregate_group(by_channel[c], illum_combine) for c in sorted(by_channel)]
    return aggregate_group(per_channel, channel_combine)


def _is_group(x):
    return isinstance(x, (list, tuple)) and len(x) > 0 and isinstance(x[0], (list, tuple))


def stitch_pairs(pairs, tiles, models, params=None, downsample_factors=(1, 1, 1), ctx: Context | None = None,
                 attributes: dict | None = None, channel_combine="AVERAGE", illum_combine="PICK_BRIGHTEST",
                 max_resident_bytes=64 << 30):
    """The collapsed RDD (J/SparkPairwiseStitching.java:192-312): a host work queue over pairs of view GROUPS.
    ``pairs``: [(A, B)] where A / B is a ViewId or a list of ViewIds (one tile's channels / illuminations);
    ``tiles``: ViewId -> [z,y,x] array; ``models``: ViewId -> 3x4; ``attributes``: ViewId -> channel / illumination.

    Every group is aggregated + downsampled ONCE, uploaded ONCE (async, from pinned memory when available) and
    stays resident while its pairs run; the overlap crops are cut on the device (bs_pcm_volumes_batch).  Pairs whose
    non-translation parts differ go through the virtual-fusion branch one by one."""
    params = params or PairwiseStitchingParameters()
    attributes = attributes or {}
    ds = np.asarray(downsample_factors, dtype=np.float64)
    groups = [(tuple(a) if _is_group(a) else (tuple(a),), tuple(b) if _is_group(b) else (tuple(b),)) for a, b in pairs]

    def first(g):
        return g[0]

    def pair_id(g):
        return g[0] if len(g) == 1 else g

    cache = {}

    def group_image(g):
        if g not in cache:
            img = tiles[g[0]] if len(g) == 1 else aggregate_views({v: tiles[v] for v in g}, attributes, channel_combine, illum_combine)
            cache[g] = downsample_avg(img, downsample_factors) if isinstance(img, np.ndarray) else img
        return cache[g]

    out = [None] * len(groups)
    fast = []      # (index, gA, gB, interval1, interval2, size, sub1, sub2)
    for i, (ga, gb) in enumerate(groups):
        ma = np.asarray(models[first(ga)], dtype=np.float64).reshape(3, 4)
        mb = np.asarray(models[first(gb)], dtype=np.float64).reshape(3, 4)
        if not non_translations_equal(ma, mb):
            r = compute_stitching_non_equal_transformations(group_image(ga) if len(ga) > 1 else tiles[ga[0]],
                                                            group_image(gb) if len(gb) > 1 else tiles[gb[0]],
                                                            ma, mb, params, downsample_factors, ctx)
            if r is not None:
                (tr, cc), (bmin, bmax) = r
                out[i] = PairwiseStitchingResult((pair_id(ga), pair_id(gb)), tr, cc, bmin, bmax, shift_px=tuple(tr[:, 3]))
            continue
        dims_a = tuple(tiles[first(ga)].shape)[::-1]
        dims_b = tuple(tiles[first(gb)].shape)[::-1]
        la, ha = _world_box(ma, dims_a)
        lb, hb = _world_box(mb, dims_b)
        ov = _overlap(la, ha, lb, hb)
        if ov is None:
            continue
        lin_inv = np.linalg.inv(ma[:, :3])
        t1 = (lin_inv @ ma[:, 3]) / ds
        t2 = (lin_inv @ mb[:, 3]) / ds
        da = tuple(int(dims_a[d] // int(ds[d])) for d in range(3))
        db = tuple(int(dims_b[d] // int(ds[d])) for d in range(3))
        ro = local_raster_overlaps(da, db, t1, t2)
        if ro is None:
            continue
        fast.append((i, ga, gb, ro, ov, mb))

    # resident-tile batches: greedily fill the device budget, upload, correlate, free
    p = ctx.pcm_params(params.peaks_to_check, params.do_subpixel, params.min_overlap, params.extension) if fast else None
    pos = 0
    while pos < len(fast):
        need, nbytes, end = {}, 0, pos
        while end < len(fast):
            extra = [g for g in (fast[end][1], fast[end][2]) if g not in need]
            add = sum(group_image(g).nbytes for g in set(extra))
            if need and nbytes + add > max_resident_bytes:
                break
            for g in extra:
                need[g] = True
            nbytes += add
            end += 1
        handles = {}
        try:
            for g in need:
                img = np.ascontiguousarray(group_image(g))
                cache[g] = img
                handles[g] = ctx.volume_upload(img)
            by_dtype = {}
            for item in fast[pos:end]:
                by_dtype.setdefault((group_image(item[1]).dtype, group_image(item[2]).dtype), []).append(item)
            for (dta, dtb), items in by_dtype.items():
                if dta != dtb:   # one side averaged (float32), the other a single uint16 view: host crops
                    for (i, ga, gb, ro, ov, mb) in items:
                        a1, a2, size, sub1, sub2 = ro
                        A, B = group_image(ga), group_image(gb)
                        c1 = np.ascontiguousarray(A[a1[2]:a1[2] + size[2], a1[1]:a1[1] + size[1], a1[0]:a1[0] + size[0]], dtype=np.float32)
                        c2 = np.ascontiguousarray(B[a2[2]:a2[2] + size[2], a2[1]:a2[1] + size[1], a2[0]:a2[0] + size[0]], dtype=np.float32)
                        res = [ctx.pcm_pair(c1, c2, p)]
                        _finish_fast(out, [(i, ga, gb, ro, ov, mb)], res, params, ds, pair_id)
                    continue
                jobs = [(handles[ga], handles[gb], tuple(int(v) for v in ro[0]), tuple(int(v) for v in ro[1]),
                         tuple(int(v) for v in ro[2])) for (_, ga, gb, ro, _, _) in items]
                _finish_fast(out, items, ctx.pcm_volumes_batch(jobs, p), params, ds, pair_id)
        finally:
            for h in handles.values():
                ctx.volume_free(h)
        pos = end
    return out


def _finish_fast(out, items, results, params, ds, pair_id):
    for (i, ga, gb, ro, ov, mb), res in zip(items, results):
        if not res.found or math.isinf(res.r):
            continue
        s = np.asarray(res.shift_sub if params.do_subpixel else res.shift_int, dtype=np.float64)
        shift = s - (np.asarray(ro[4]) - np.asarray(ro[3]))     # int / real coordinate difference of the raster crops
        T = np.eye(4)
        T[:3, 3] = shift * ds
        Mb = np.vstack([mb, [0, 0, 0, 1]])
        R = (Mb @ T @ np.linalg.inv(Mb))[:3, :].copy()
        out[i] = PairwiseStitchingResult((pair_id(ga), pair_id(gb)), R, float(res.r), tuple(ov[0]), tuple(ov[1]),
                                         shift_px=tuple(R[:, 3]))
