"""Seeded synthetic volumes shared by the tests (SURVEY.md 8d generators, small sizes)."""
import numpy as np
from scipy.ndimage import gaussian_filter


def field(shape_zyx, seed=1234, sigma=1.0, mean=1000.0, std=300.0):
    rng = np.random.default_rng(seed)
    g = gaussian_filter(rng.standard_normal(shape_zyx).astype(np.float32), sigma)
    return ((g - g.mean()) / g.std() * std + mean).astype(np.float32)


def tile_from(G, off_zyx, shape_zyx, seed, noise=30.0, dtype=np.uint16):
    z, y, x = off_zyx
    t = G[z:z + shape_zyx[0], y:y + shape_zyx[1], x:x + shape_zyx[2]].astype(np.float32)
    if noise > 0:
        t = t + np.random.default_rng(seed).standard_normal(shape_zyx).astype(np.float32) * noise
    if dtype == np.float32:
        return t.astype(np.float32)
    info = np.iinfo(dtype)
    return np.clip(np.rint(t), info.min, info.max).astype(dtype)


def shifted_pair(shape_zyx, shift_xyz, seed=0, margin=24, sigma=1.0, noise=30.0, dtype=np.uint16):
    """img2(p) = img1(p + shift) on a common field; expected PCM shift == shift_xyz."""
    big = tuple(s + 2 * margin for s in shape_zyx)
    G = field(big, seed=seed, sigma=sigma)
    a = tile_from(G, (margin, margin, margin), shape_zyx, 1000 + seed, noise, dtype)
    sx, sy, sz = shift_xyz
    b = tile_from(G, (margin + sz, margin + sy, margin + sx), shape_zyx, 2000 + seed, noise, dtype)
    return a, b


def rot_z(deg, center_xyz=(0, 0, 0)):
    """3x4 rotation about z through ``center``."""
    a = np.deg2rad(deg)
    R = np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]], dtype=np.float64)
    c = np.asarray(center_xyz, dtype=np.float64)
    t = c - R @ c
    return np.hstack([R, t[:, None]])


def translation(t_xyz):
    M = np.hstack([np.eye(3), np.asarray(t_xyz, dtype=np.float64)[:, None]])
    return M


def subpixel_pair(shape_zyx, shift_xyz, seed=0, margin=24, sigma=1.5, noise=20.0, dtype=np.uint16, workers=-1):
    """img2(p) = img1(p + shift) with a REAL-valued shift: the common field is translated in the
    Fourier domain ("additional Fourier-domain sub-pixel shift"), then both tiles are
    cropped at the same offset."""
    from scipy import fft as sfft
    big = tuple(s + 2 * margin for s in shape_zyx)
    G = field(big, seed=seed, sigma=sigma)
    F = sfft.rfftn(G, workers=workers)
    sx, sy, sz = shift_xyz
    kz = sfft.fftfreq(big[0])[:, None, None]
    ky = sfft.fftfreq(big[1])[None, :, None]
    kx = sfft.rfftfreq(big[2])[None, None, :]
    F *= np.exp(2j * np.pi * (kz * sz + ky * sy + kx * sx)).astype(np.complex64)  # G2(p) = G(p + s)
    G2 = sfft.irfftn(F, s=big, workers=workers).astype(np.float32)
    a = tile_from(G, (margin,) * 3, shape_zyx, 1000 + seed, noise, dtype)
    b = tile_from(G2, (margin,) * 3, shape_zyx, 2000 + seed, noise, dtype)
    return a, b


def rot_x(deg, center_xyz=(0, 0, 0)):
    """3x4 rotation about x through ``center`` (couples y and z: no xy-affine structure)."""
    a = np.deg2rad(deg)
    R = np.array([[1, 0, 0], [0, np.cos(a), -np.sin(a)], [0, np.sin(a), np.cos(a)]], dtype=np.float64)
    c = np.asarray(center_xyz, dtype=np.float64)
    t = c - R @ c
    return np.hstack([R, t[:, None]])
