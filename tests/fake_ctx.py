"""Oracle-backed stand-in for bsgpu.native.Context -- TESTS ONLY.  It lets the CPU suite drive the
host-side logic (stitching / fusion / commands: geometry, work queues, XML + N5/Zarr I/O, retry)
end to end without a GPU.  The product has no such fallback (tests/test_abi.py)."""
import numpy as np

from oracle import fusion_oracle as fo
from oracle import pcm_oracle as po

_FT = {"AVG": fo.AVG, "AVG_BLEND": fo.AVG_BLEND, "AVG_CONTENT": fo.AVG_CONTENT, "AVG_BLEND_CONTENT": fo.AVG_BLEND_CONTENT,
       "MAX_INTENSITY": fo.MAX_INTENSITY, "LOWEST_VIEWID_WINS": fo.LOWEST_VIEWID_WINS,
       "HIGHEST_VIEWID_WINS": fo.HIGHEST_VIEWID_WINS, "CLOSEST_PIXEL_WINS": fo.CLOSEST_PIXEL_WINS}
_DT = {0: "uint16", 1: "float32", 2: "uint8"}


class _FuseParams:
    def __init__(self, fusion_type, interpolation, out_dtype, blend_lut_n, min_intensity, max_intensity):
        self.fusion_type = _FT[fusion_type] if isinstance(fusion_type, str) else int(fusion_type)
        self.interpolation, self.out_dtype, self.blend_lut_n = interpolation, out_dtype, blend_lut_n
        self.min_intensity, self.max_intensity = min_intensity, max_intensity


class FakeContext:
    device = 0

    def __init__(self):
        self.vols, self.next = {}, 1
        self.calls = {"pcm": 0, "fuse": 0, "downsample": 0}

    # -- hot path 1
    @staticmethod
    def pcm_params(peaks_to_check=5, do_subpixel=True, min_overlap_frac=0.25, extension=(10, 10, 10)):
        return dict(peaks_to_check=peaks_to_check, do_subpixel=do_subpixel, min_overlap_frac=min_overlap_frac,
                    extension=extension)

    def pcm_pair(self, a, b, p=None):
        self.calls["pcm"] += 1
        return po.pcm_shift(np.asarray(a), np.asarray(b), **(p or self.pcm_params()))

    def pcm_volumes_batch(self, jobs, p=None):
        out = []
        for (v1, v2, m1, m2, d) in jobs:
            a = self.vols[v1][m1[2]:m1[2] + d[2], m1[1]:m1[1] + d[1], m1[0]:m1[0] + d[0]]
            b = self.vols[v2][m2[2]:m2[2] + d[2], m2[1]:m2[1] + d[1], m2[0]:m2[0] + d[0]]
            out.append(self.pcm_pair(np.ascontiguousarray(a), np.ascontiguousarray(b), p))
        return out

    # -- volumes
    def volume_upload(self, vol):
        h = self.next
        self.next += 1
        self.vols[h] = np.ascontiguousarray(vol)
        return h

    def volume_free(self, h):
        del self.vols[h]

    def volume_download(self, h, dims_xyz=None, dtype=None):
        v = self.vols[h]
        if dims_xyz is not None:
            assert tuple(v.shape[::-1]) == tuple(int(d) for d in dims_xyz)
        if dtype is not None:
            assert v.dtype == np.dtype(dtype)
        return v.copy()

    def content_weights(self, h, sigma1=20.0, sigma2=40.0):
        return self.volume_upload(fo.content_weights(self.vols[h], sigma1, sigma2))

    def downsample(self, h, factors):
        self.calls["downsample"] += 1
        return self.volume_upload(fo.downsample2x(self.vols[h], tuple(int(f) for f in factors)))

    # -- hot path 2
    @staticmethod
    def fuse_params(fusion_type="AVG_BLEND", interpolation=1, out_dtype=1, blend_lut_n=0, min_intensity=0.0,
                    max_intensity=65535.0, out_big_endian=False):
        p = _FuseParams(fusion_type, interpolation, out_dtype, blend_lut_n, min_intensity, max_intensity)
        p.out_big_endian = bool(out_big_endian)
        return p

    def _views(self, views):
        out = []
        for v in views:
            vol = self.vols[v["vol_handle"]]
            fd = tuple(int(x) for x in v.get("full_dims", (0, 0, 0)))
            if fd[0] > 0:   # windowed view: the resident array is the sub-interval [window_min, +shape) of the view
                w = tuple(int(x) for x in v.get("window_min", (0, 0, 0)))
                full = np.zeros(fd[::-1], dtype=vol.dtype)
                z, y, x = vol.shape
                full[w[2]:w[2] + z, w[1]:w[1] + y, w[0]:w[0] + x] = vol
                vol = full
            c = v.get("content_handle", 0)
            out.append(fo.View(vol, np.asarray(v["src_to_world"], dtype=np.float64).reshape(3, 4),
                               v.get("blend_border", (0, 0, 0)), v.get("blend_range", (40, 40, 40)),
                               self.vols[c] if c else None))
        return out

    def fuse_block(self, views, block_min, block_size, params=None, out=None):
        self.calls["fuse"] += 1
        p = params or self.fuse_params()
        res = fo.fuse_block(self._views(views), tuple(int(v) for v in block_min), tuple(int(v) for v in block_size),
                            p.fusion_type, p.interpolation, _DT[p.out_dtype], p.min_intensity, p.max_intensity, p.blend_lut_n)
        if out is not None:
            out[...] = res
            return out
        return res

    def fuse_blocks(self, views, block_mins, block_sizes, params=None, outs=None):
        res = [self.fuse_block(views, mn, sz, params) for mn, sz in zip(block_mins, block_sizes)]
        if getattr(params, "out_big_endian", False):      # the device hands back big-endian payloads
            res = [r.astype(r.dtype.newbyteorder(">")) if r.dtype.itemsize > 1 else r for r in res]
        if outs is not None:
            for o, r in zip(outs, res):
                o[...] = r
            return outs
        return res

    def mask_blocks(self, views, block_mins, block_sizes, mask_offset=(0.0, 0.0, 0.0), out_dtype=2, out_big_endian=False):
        geom = []
        for v in views:
            fd = tuple(int(x) for x in v.get("full_dims", (0, 0, 0)))
            if fd[0] <= 0:
                fd = tuple(self.vols[v["vol_handle"]].shape[::-1])
            geom.append((np.asarray(v["src_to_world"], dtype=np.float64).reshape(3, 4), fd))
        return [fo.mask_block(geom, mn, sz, mask_offset, _DT[out_dtype]) for mn, sz in zip(block_mins, block_sizes)]

    def fuse_block_to_volume(self, views, block_min, block_size, params=None):
        return self.volume_upload(self.fuse_block(views, block_min, block_size, params))
