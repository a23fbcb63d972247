"""Zstandard block codec for the N5 / Zarr writers (the reference's DEFAULT compression: `-c Zstandard`, level 3,
J/CreateFusionContainer.java:71-76, J/util/N5Util.java:91-92).

Two back ends, same frame format (RFC 8878):

  * libzstd.so.1 through ctypes when the shared library is on the machine (it is in this image, 1.5.5; there is no
    zstd.h, so nothing can be compiled against it): real compression at the reference's level 3 and a full decoder,
    i.e. containers written by the reference with its defaults can be read and vice versa.
  * a pure-Python fallback: the WRITER emits valid frames made of Raw_Block / RLE_Block only (every zstd decoder reads
    them; no entropy coding), the READER decodes frames that consist of raw / RLE blocks and raises for
    entropy-coded (Compressed_Block) input.

Host-side plumbing only.
"""
from __future__ import annotations

import ctypes as C
import struct

MAGIC = 0xFD2FB528
BLOCK_MAX = 128 * 1024
DEFAULT_LEVEL = 3

_lib = None
_lib_tried = False


def _load():
    global _lib, _lib_tried
    if _lib_tried:
        return _lib
    _lib_tried = True
    for name in ("libzstd.so.1", "libzstd.so"):
        try:
            lib = C.CDLL(name)
        except OSError:
            continue
        lib.ZSTD_compressBound.restype = C.c_size_t
        lib.ZSTD_compressBound.argtypes = [C.c_size_t]
        lib.ZSTD_compress.restype = C.c_size_t
        lib.ZSTD_compress.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int]
        lib.ZSTD_decompress.restype = C.c_size_t
        lib.ZSTD_decompress.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
        lib.ZSTD_getFrameContentSize.restype = C.c_ulonglong
        lib.ZSTD_getFrameContentSize.argtypes = [C.c_void_p, C.c_size_t]
        lib.ZSTD_isError.restype = C.c_uint
        lib.ZSTD_isError.argtypes = [C.c_size_t]
        lib.ZSTD_getErrorName.restype = C.c_char_p
        lib.ZSTD_getErrorName.argtypes = [C.c_size_t]
        _lib = lib
        break
    return _lib


def have_library() -> bool:
    return _load() is not None


# ------------------------------------------------------------------------------------------ pure-Python frames
def compress_store(data: bytes) -> bytes:
    """A valid Zstandard frame without entropy coding: Single_Segment frame header with the content size, then
    Raw_Block / RLE_Block (a block whose bytes are all equal) of at most 128 KiB, no checksum."""
    data = bytes(data)
    n = len(data)
    # Frame_Header_Descriptor: FCS field size flag (bits 7-6), Single_Segment (bit 5), no checksum, no dictionary
    if n <= 255:
        fhd, fcs = (0 << 6) | (1 << 5), struct.pack("<B", n)
    elif n <= 65535 + 256:
        fhd, fcs = (1 << 6) | (1 << 5), struct.pack("<H", n - 256)
    elif n <= 0xFFFFFFFF:
        fhd, fcs = (2 << 6) | (1 << 5), struct.pack("<I", n)
    else:
        fhd, fcs = (3 << 6) | (1 << 5), struct.pack("<Q", n)
    out = [struct.pack("<I", MAGIC), bytes([fhd]), fcs]
    if n == 0:
        out.append(struct.pack("<I", 1)[:3])            # last block, Raw_Block, size 0
        return b"".join(out)
    pos = 0
    while pos < n:
        chunk = data[pos:pos + BLOCK_MAX]
        pos += len(chunk)
        last = 1 if pos >= n else 0
        if len(chunk) > 1 and chunk.count(chunk[:1]) == len(chunk):
            hdr = last | (1 << 1) | (len(chunk) << 3)   # RLE_Block: Block_Size = regenerated size, one byte follows
            out.append(struct.pack("<I", hdr)[:3] + chunk[:1])
        else:
            hdr = last | (0 << 1) | (len(chunk) << 3)   # Raw_Block
            out.append(struct.pack("<I", hdr)[:3] + chunk)
    return b"".join(out)


def decompress_store(buf: bytes) -> bytes:
    """Decode frames made of raw / RLE blocks (what compress_store writes); skippable frames are skipped."""
    buf = bytes(buf)
    pos, out = 0, []
    while pos < len(buf):
        (magic,) = struct.unpack_from("<I", buf, pos)
        pos += 4
        if 0x184D2A50 <= magic <= 0x184D2A5F:           # skippable frame
            (sz,) = struct.unpack_from("<I", buf, pos)
            pos += 4 + sz
            continue
        if magic != MAGIC:
            raise ValueError("not a Zstandard frame")
        fhd = buf[pos]
        pos += 1
        fcs_flag, single, checksum, did = fhd >> 6, (fhd >> 5) & 1, (fhd >> 2) & 1, fhd & 3
        if not single:
            pos += 1                                     # Window_Descriptor
        pos += (0, 1, 2, 4)[did]
        fcs_size = (1 if single else 0, 2, 4, 8)[fcs_flag]
        pos += fcs_size
        while True:
            hdr = buf[pos] | (buf[pos + 1] << 8) | (buf[pos + 2] << 16)
            pos += 3
            last, btype, bsize = hdr & 1, (hdr >> 1) & 3, hdr >> 3
            if btype == 0:
                out.append(buf[pos:pos + bsize])
                pos += bsize
            elif btype == 1:
                out.append(buf[pos:pos + 1] * bsize)
                pos += 1
            else:
                raise NotImplementedError("entropy-coded Zstandard block: needs libzstd.so.1 (not found on this machine)")
            if last:
                break
        if checksum:
            pos += 4
    return b"".join(out)


# ------------------------------------------------------------------------------------------ public API
def compress(data: bytes, level: int = DEFAULT_LEVEL) -> bytes:
    lib = _load()
    if lib is None:
        return compress_store(data)
    data = bytes(data)
    cap = lib.ZSTD_compressBound(len(data))
    dst = C.create_string_buffer(cap)
    n = lib.ZSTD_compress(dst, cap, data, len(data), int(level))
    if lib.ZSTD_isError(n):
        raise RuntimeError("ZSTD_compress: " + lib.ZSTD_getErrorName(n).decode())
    return dst.raw[:n]


def decompress(buf: bytes) -> bytes:
    lib = _load()
    if lib is None:
        return decompress_store(buf)
    buf = bytes(buf)
    size = lib.ZSTD_getFrameContentSize(buf, len(buf))
    if size in (0xFFFFFFFFFFFFFFFF, 0xFFFFFFFFFFFFFFFE):   # unknown / error: the store decoder handles our own frames
        return decompress_store(buf)
    dst = C.create_string_buffer(max(1, int(size)))
    n = lib.ZSTD_decompress(dst, int(size), buf, len(buf))
    if lib.ZSTD_isError(n):
        raise RuntimeError("ZSTD_decompress: " + lib.ZSTD_getErrorName(n).decode())
    return dst.raw[:n]
