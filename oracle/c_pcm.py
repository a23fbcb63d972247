"""ctypes loader of oracle/c/libpcm_oracle.so (C / OpenMP restatement of pcm_oracle.pcm_shift for uint16 crops).
TEST INFRASTRUCTURE ONLY: used by tests/ and by bench.py's CPU arms (cpu_baseline, --impl reference)."""
import ctypes as C
import os
import subprocess

import numpy as np

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "c")
_LIB = os.path.join(_DIR, "libpcm_oracle.so")
_lib = None


def build():
    subprocess.check_call(["make", "-C", _DIR])


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            build()
        _lib = C.CDLL(_LIB)
        _lib.po_num_threads.restype = C.c_int
    return _lib


def set_num_threads(n: int):
    load().po_set_num_threads(int(n))


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
    return load().po_num_threads()


def pcm_shift(img1, img2, peaks_to_check=5, do_subpixel=True, min_overlap_frac=0.25, extension=(10, 10, 10)):
    """Same result record as oracle.pcm_oracle.pcm_shift (without the candidate diagnostics)."""
    from .pcm_oracle import PcmResult
    a = np.ascontiguousarray(img1, dtype=np.uint16)
    b = np.ascontiguousarray(img2, dtype=np.uint16)
    if a.shape != b.shape:
        raise ValueError("crops must have equal size")
    dims = (C.c_longlong * 3)(*a.shape[::-1])
    ext = (C.c_int * 3)(*[int(e) for e in extension])
    out = (C.c_double * 16)()
    rc = load().po_pcm_shift(C.c_void_p(a.ctypes.data), C.c_void_p(b.ctypes.data), dims, int(peaks_to_check), 1 if do_subpixel else 0,
                             C.c_double(min_overlap_frac), ext, out)
    if rc != 0:
        raise RuntimeError(f"po_pcm_shift failed: {rc}")
    pad = tuple(int(out[13 + i]) for i in range(3))
    if not out[0]:
        return PcmResult(False, pad=pad)
    return PcmResult(True, shift_int=tuple(int(out[1 + i]) for i in range(3)), shift_sub=tuple(float(out[4 + i]) for i in range(3)),
                     r=float(out[7]), n_overlap_px=int(out[8]), peak_index=tuple(int(out[9 + i]) for i in range(3)),
                     pcm_value=float(out[12]), pad=pad)
