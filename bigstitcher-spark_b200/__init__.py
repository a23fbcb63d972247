"""bsgpu -- B200-native (sm_100a) phase-correlation stitching and affine fusion behind the
BigStitcher-Spark `stitching` / `affine-fusion` operator interface.

The directory name carries a hyphen (``bigstitcher-spark_b200``); import it through the
top-level alias module ``bsgpu`` (``import bsgpu``).  The compute path is
``libbsgpu.so`` (hand-written CUDA behind the C ABI in ``include/bsgpu.h``); nothing here
falls back to the CPU.
"""
from . import native, stitching, fusion, parallel, n5, zarr, spimdata, commands  # noqa: F401
from .native import BsError, Context, load_library, good_fft_size  # noqa: F401
