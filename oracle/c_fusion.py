"""ctypes loader of oracle/c/libfusion_oracle.so (C / OpenMP restatement of fuse_block for AVG and
AVG_BLEND).  TEST INFRASTRUCTURE ONLY: used by tests/ and by bench.py's cpu_baseline leg."""
import ctypes as C
import os
import subprocess

import numpy as np

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "c")
_LIB = os.path.join(_DIR, "libfusion_oracle.so")
_lib = None
_DT = {np.dtype(np.uint16): 0, np.dtype(np.float32): 1, np.dtype(np.uint8): 2}


def build():
    subprocess.check_call(["make", "-C", _DIR])


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            build()
        _lib = C.CDLL(_LIB)
        _lib.fo_num_threads.restype = C.c_int
    return _lib


def set_num_threads(n: int):
    load().fo_set_num_threads(int(n))


def physical_cores() -> int:
    """Physical cores visible to this process (psutil when present, else the logical count)."""
    try:
        import psutil
        n = psutil.cpu_count(logical=False) or 0
    except Exception:
        n = 0
    avail = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    return max(1, min(n, avail)) if n else avail


def num_threads():
    return load().fo_num_threads()


def fuse_block(views, block_min_xyz, block_size_xyz, fusion_type):
    """views: oracle.fusion_oracle.View list; fusion_type 0 (AVG) or 1 (AVG_BLEND)."""
    lib = load()
    n = len(views)
    imgs = (C.c_void_p * max(n, 1))()
    dts = (C.c_int * max(n, 1))()
    dims = (C.c_longlong * (3 * max(n, 1)))()
    s2w = (C.c_double * (12 * max(n, 1)))()
    bo = (C.c_float * (3 * max(n, 1)))()
    rg = (C.c_float * (3 * max(n, 1)))()
    keep = []
    for i, v in enumerate(views):
        img = np.ascontiguousarray(v.img)
        keep.append(img)
        imgs[i] = img.ctypes.data
        dts[i] = _DT[img.dtype]
        dims[3 * i:3 * i + 3] = [int(d) for d in img.shape[::-1]]
        s2w[12 * i:12 * i + 12] = [float(x) for x in np.asarray(v.src_to_world, dtype=np.float64).ravel()]
        bo[3 * i:3 * i + 3] = [float(x) for x in (v.blend_border if v.blend_border is not None else (0, 0, 0))]
        rg[3 * i:3 * i + 3] = [float(x) for x in (v.blend_range if v.blend_range is not None else (40, 40, 40))]
    bmin = (C.c_longlong * 3)(*[int(x) for x in block_min_xyz])
    bsz = (C.c_longlong * 3)(*[int(x) for x in block_size_xyz])
    out = np.empty(tuple(int(x) for x in block_size_xyz)[::-1], dtype=np.float32)
    rc = lib.fo_fuse_block(n, imgs, dts, dims, s2w, bo, rg, bmin, bsz, int(fusion_type), C.c_void_p(out.ctypes.data))
    if rc != 0:
        raise RuntimeError(f"fo_fuse_block failed: {rc}")
    return out
