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

    # -- volumes
    def volume_upload(self, vol):
        h = self.next
        self.next += 1
        self.vols[h] = np.ascontiguousarray(vol)
        return h

    def volume_free(self, h):
        del self.vols[h]

    def volume_download(self, h, dims_xyz, dtype=np.float32):
        v = self.vols[h]
        assert tuple(v.shape[::-1]) == tuple(int(d) for d in dims_xyz) and v.dtype == np.dtype(dtype)
        return v.copy()

    def downsample(self, h, factors):
        self.calls["downsample"] += 1
        return self.volume_upload(fo.downsample2x(self.vols[h], tuple(int(f) for f in factors)))

    # -- hot path 2
    @staticmethod
    def fuse_params(fusion_type="AVG_BLEND", interpolation=1, out_dtype=1, blend_lut_n=0, min_intensity=0.0,
                    max_intensity=65535.0):
        return _FuseParams(fusion_type, interpolation, out_dtype, blend_lut_n, min_intensity, max_intensity)

    def _views(self, views):
        out = []
        for v in views:
            out.append(fo.View(self.vols[v["vol_handle"]], np.asarray(v["src_to_world"], dtype=np.float64).reshape(3, 4),
                               v.get("blend_border", (0, 0, 0)), v.get("blend_range", (40, 40, 40))))
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

    def fuse_block_to_volume(self, views, block_min, block_size, params=None):
        return self.volume_upload(self.fuse_block(views, block_min, block_size, params))
