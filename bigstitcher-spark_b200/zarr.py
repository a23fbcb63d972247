"""Minimal OME-Zarr (Zarr v2) writer / reader for the fusion output -- the reference's DEFAULT
container (J/CreateFusionContainer.java:67-69,331-389): one 5-D array per resolution level, N5-API
axis order x,y,z,c,t == Zarr array order t,c,z,y,x, chunks {1,1,bz,by,bx}, levels named "0","1",...
(:346), `multiscales` v0.4 metadata (:374-388), grid offsets {gx,gy,gz,c,t} at write time
(J/SparkAffineFusion.java:630-643).  Little-endian C-order chunks, always full chunk shape (edge
chunks padded with fill_value 0), dimension_separator "/"; compressor null (raw) or gzip -- zstd,
the reference default, is not available in this image.  Host-side plumbing only.
"""
from __future__ import annotations

import gzip
import json
import os

import numpy as np

_ZDT = {"uint8": "|u1", "uint16": "<u2", "float32": "<f4"}
_NPDT = {"|u1": np.uint8, "<u2": np.uint16, "<f4": np.float32}


class ZarrStore:
    def __init__(self, root: str, create: bool = False):
        self.root = root
        if create:
            os.makedirs(root, exist_ok=True)
            self._write_json("", ".zgroup", {"zarr_format": 2})
        elif not os.path.isdir(root):
            raise FileNotFoundError(root)

    def _write_json(self, group, name, obj):
        d = os.path.join(self.root, group.strip("/"))
        os.makedirs(d, exist_ok=True)
        with open(os.path.join(d, name), "w") as f:
            json.dump(obj, f)

    def _read_json(self, group, name):
        p = os.path.join(self.root, group.strip("/"), name)
        if not os.path.exists(p):
            return {}
        with open(p) as f:
            return json.load(f)

    def get_attributes(self, group=""):
        return self._read_json(group, ".zattrs")

    def set_attributes(self, group, attrs: dict):
        cur = self.get_attributes(group)
        cur.update(attrs)
        self._write_json(group, ".zattrs", cur)

    def create_array(self, path, shape_tczyx, chunks_tczyx, dtype: str, compression="raw"):
        comp = None if compression == "raw" else {"id": "gzip", "level": 1}
        if compression not in ("raw", "gzip"):
            raise NotImplementedError(f"compression {compression} (not available in this image)")
        self._write_json(path, ".zarray", {"zarr_format": 2, "shape": [int(v) for v in shape_tczyx],
                                           "chunks": [int(v) for v in chunks_tczyx], "dtype": _ZDT[dtype],
                                           "compressor": comp, "fill_value": 0, "order": "C", "filters": None,
                                           "dimension_separator": "/"})

    def array_meta(self, path):
        m = self._read_json(path, ".zarray")
        if not m:
            raise KeyError(f"{path} is not a Zarr array")
        return m

    def _chunk_path(self, path, idx_tczyx):
        return os.path.join(self.root, path.strip("/"), *[str(int(i)) for i in idx_tczyx])

    def write_chunk(self, path, idx_tczyx, block_zyx: np.ndarray):
        """block_zyx: the valid [z,y,x] part of one chunk; padded to the full chunk shape."""
        m = self.array_meta(path)
        cz, cy, cx = m["chunks"][2:]
        dt = np.dtype(_NPDT[m["dtype"]])
        full = np.zeros((cz, cy, cx), dtype=dt)
        z, y, x = block_zyx.shape
        full[:z, :y, :x] = block_zyx
        payload = full.astype(dt.newbyteorder("<"), copy=False).tobytes()
        if m["compressor"] is not None:
            payload = gzip.compress(payload, compresslevel=m["compressor"].get("level", 1))
        p = self._chunk_path(path, idx_tczyx)
        os.makedirs(os.path.dirname(p), exist_ok=True)
        with open(p, "wb") as f:
            f.write(payload)

    def read_chunk(self, path, idx_tczyx):
        m = self.array_meta(path)
        p = self._chunk_path(path, idx_tczyx)
        cz, cy, cx = m["chunks"][2:]
        dt = np.dtype(_NPDT[m["dtype"]])
        if not os.path.exists(p):
            return np.zeros((cz, cy, cx), dtype=dt)
        with open(p, "rb") as f:
            payload = f.read()
        if m["compressor"] is not None:
            payload = gzip.decompress(payload)
        return np.frombuffer(payload, dtype=dt.newbyteorder("<")).astype(dt).reshape(cz, cy, cx)

    def save_block(self, path, volume_zyx: np.ndarray, grid_offset_xyzct):
        """N5Utils.saveBlock on the 5-D view (J/SparkAffineFusion.java:630-643,670): split a
        [z,y,x] super-block into chunks starting at chunk index {gx,gy,gz} of channel c, timepoint t."""
        m = self.array_meta(path)
        gx, gy, gz, c, t = (int(v) for v in grid_offset_xyzct)
        sz, sy, sx = m["shape"][2:]
        cz, cy, cx = m["chunks"][2:]
        vz, vy, vx = volume_zyx.shape
        for kz in range(-(-vz // cz)):
            for ky in range(-(-vy // cy)):
                for kx in range(-(-vx // cx)):
                    iz, iy, ix = gz + kz, gy + ky, gx + kx
                    if iz * cz >= sz or iy * cy >= sy or ix * cx >= sx:
                        continue
                    blk = volume_zyx[kz * cz:(kz + 1) * cz, ky * cy:(ky + 1) * cy, kx * cx:(kx + 1) * cx]
                    blk = blk[:sz - iz * cz, :sy - iy * cy, :sx - ix * cx]
                    self.write_chunk(path, (t, c, iz, iy, ix), blk)

    def read_volume(self, path, c=0, t=0):
        m = self.array_meta(path)
        sz, sy, sx = m["shape"][2:]
        cz, cy, cx = m["chunks"][2:]
        out = np.zeros((sz, sy, sx), dtype=_NPDT[m["dtype"]])
        for iz in range(-(-sz // cz)):
            for iy in range(-(-sy // cy)):
                for ix in range(-(-sx // cx)):
                    ch = self.read_chunk(path, (t, c, iz, iy, ix))
                    z0, y0, x0 = iz * cz, iy * cy, ix * cx
                    out[z0:z0 + cz, y0:y0 + cy, x0:x0 + cx] = ch[:sz - z0, :sy - y0, :sx - x0]
        return out


def create_fusion_container_zarr(root, input_xml, bb_min, bb_max, block_size=(128, 128, 128), dtype="float32",
                                 min_intensity=None, max_intensity=None, num_timepoints=1, num_channels=1,
                                 anisotropy_factor=None, compression="raw", voxel_size=(1.0, 1.0, 1.0)):
    """`create-fusion-container -s ZARR` (J/CreateFusionContainer.java:331-389): 5-D array "0" with the
    `multiscales` attribute and the `Bigstitcher-Spark/*` root attributes."""
    st = ZarrStore(root, create=True)
    dims = [int(bb_max[d] - bb_min[d] + 1) for d in range(3)]
    st.create_array("0", (num_timepoints, num_channels, dims[2], dims[1], dims[0]),
                    (1, 1, block_size[2], block_size[1], block_size[0]), dtype, compression)
    multiscales = [{"version": "0.4", "name": "/",
                    "axes": [{"name": "t", "type": "time", "unit": "second"}, {"name": "c", "type": "channel"},
                             {"name": "z", "type": "space", "unit": "micrometer"},
                             {"name": "y", "type": "space", "unit": "micrometer"},
                             {"name": "x", "type": "space", "unit": "micrometer"}],
                    "datasets": [{"path": "0", "coordinateTransformations": [
                        {"type": "scale", "scale": [1.0, 1.0, voxel_size[2], voxel_size[1], voxel_size[0]]},
                        {"type": "translation", "translation": [0.0, 0.0, 0.0, 0.0, 0.0]}]}]}]
    mr = [[{"dataset": "0", "dimensions": dims + [num_channels, num_timepoints],
            "blockSize": list(block_size) + [1, 1], "relativeDownsampling": [1, 1, 1],
            "absoluteDownsampling": [1, 1, 1], "dataType": dtype}]]
    attrs = {"multiscales": multiscales, "Bigstitcher-Spark/FusionFormat": "OME-ZARR",
             "Bigstitcher-Spark/InputXML": input_xml, "Bigstitcher-Spark/NumTimepoints": num_timepoints,
             "Bigstitcher-Spark/NumChannels": num_channels,
             "Bigstitcher-Spark/Boundingbox_min": [int(v) for v in bb_min],
             "Bigstitcher-Spark/Boundingbox_max": [int(v) for v in bb_max],
             "Bigstitcher-Spark/PreserveAnisotropy": anisotropy_factor is not None,
             "Bigstitcher-Spark/DataType": dtype.upper(), "Bigstitcher-Spark/BlockSize": list(block_size),
             "Bigstitcher-Spark/MultiResolutionInfos": mr}
    if anisotropy_factor is not None:
        attrs["Bigstitcher-Spark/AnisotropyFactor"] = float(anisotropy_factor)
    if dtype != "float32":
        attrs["Bigstitcher-Spark/MinIntensity"] = float(min_intensity)
        attrs["Bigstitcher-Spark/MaxIntensity"] = float(max_intensity)
    st.set_attributes("", attrs)
    return st


def read_fusion_container_zarr(root):
    st = ZarrStore(root)
    a = st.get_attributes("")
    g = lambda k, d=None: a.get("Bigstitcher-Spark/" + k, d)  # noqa: E731
    if g("FusionFormat") != "OME-ZARR":
        raise KeyError("not a BigStitcher-Spark OME-ZARR fusion container")
    return st, {
        "format": "OME-ZARR", "input_xml": g("InputXML"), "num_timepoints": g("NumTimepoints", 1),
        "num_channels": g("NumChannels", 1), "bb_min": g("Boundingbox_min"), "bb_max": g("Boundingbox_max"),
        "preserve_anisotropy": g("PreserveAnisotropy", False), "anisotropy_factor": g("AnisotropyFactor", float("nan")),
        "dtype": g("DataType", "FLOAT32").lower(), "block_size": g("BlockSize"),
        "min_intensity": g("MinIntensity", 0.0), "max_intensity": g("MaxIntensity", 65535.0),
        "mr_infos": g("MultiResolutionInfos"),
    }
