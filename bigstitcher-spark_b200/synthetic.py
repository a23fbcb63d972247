"""Synthetic workloads of BASELINE.json / SURVEY.md 8(d), generated on the device with torch
(plumbing only: random fields, Gaussian smoothing, crops -- none of it is on the timed path).

Band-limited noise field G: white noise, Gaussian sigma = 2 px, rescaled to mean 1000 / std 300,
plus per-tile i.i.d. noise sigma = 30, clipped to the uint16 range.
"""
from __future__ import annotations

import numpy as np


def _gauss_kernel(sigma, device):
    import torch
    r = int(3 * sigma + 0.5)
    x = torch.arange(-r, r + 1, device=device, dtype=torch.float32)
    k = torch.exp(-0.5 * (x / sigma) ** 2)
    return k / k.sum()


def smooth_field(shape_zyx, seed, device, sigma=2.0, mean=1000.0, std=300.0):
    """float32 field on ``device``: smoothed white noise, normalised."""
    import torch
    import torch.nn.functional as F
    g = torch.Generator(device=device)
    g.manual_seed(int(seed))
    v = torch.randn(shape_zyx, generator=g, device=device, dtype=torch.float32)[None, None]
    k = _gauss_kernel(sigma, device)
    r = (k.numel() - 1) // 2
    for ax in range(3):
        shape = [1, 1, 1, 1, 1]
        shape[2 + ax] = k.numel()
        pad = [0, 0, 0, 0, 0, 0]
        pad[2 * (2 - ax)] = pad[2 * (2 - ax) + 1] = r
        v = F.conv3d(F.pad(v, pad, mode="circular"), k.view(shape))
    v = v[0, 0]
    v = (v - v.mean()) / v.std() * std + mean
    return v


def tile_from_field(field, off_zyx, shape_zyx, seed, noise=30.0):
    """uint16 tile (stored as torch.int16 bit pattern; values stay < 32768) cut from ``field``."""
    import torch
    z, y, x = off_zyx
    t = field[z:z + shape_zyx[0], y:y + shape_zyx[1], x:x + shape_zyx[2]]
    g = torch.Generator(device=field.device)
    g.manual_seed(int(seed))
    t = t + torch.randn(t.shape, generator=g, device=field.device, dtype=torch.float32) * noise
    return torch.clamp(torch.round(t), 0, 32767).to(torch.int16).contiguous()


def fourier_shift(field, shift_zyx):
    """field translated by a real-valued shift: out(p) = field(p + shift) (periodic), via the Fourier shift theorem."""
    import torch
    nz, ny, nx = field.shape
    F = torch.fft.rfftn(field)
    kz = torch.fft.fftfreq(nz, device=field.device)[:, None, None]
    ky = torch.fft.fftfreq(ny, device=field.device)[None, :, None]
    kx = torch.fft.rfftfreq(nx, device=field.device)[None, None, :]
    ph = 2.0 * np.pi * (kz * shift_zyx[0] + ky * shift_zyx[1] + kx * shift_zyx[2])
    F = F * torch.polar(torch.ones_like(ph), ph)
    return torch.fft.irfftn(F, s=(nz, ny, nx))


def make_pcm_workload(n_pairs, n=512, device="cuda", seed=42, n_fields=4, max_shift=20, subpixel_every=2):
    """BASELINE config 2: ``n_pairs`` overlap-crop pairs of n^3 uint16 with planted integer shifts drawn
    uniformly from [-max_shift, max_shift]^3 (numpy default_rng(seed)); every ``subpixel_every``-th pair (half of
    them by default) carries an additional Fourier-domain sub-pixel shift in [-0.5, 0.5)^3.  Returns
    (imgs1, imgs2, shifts_xyz) with real-valued shifts.  Pair i is cut from field i % n_fields (field seed 2000 + f),
    tile noise seeds 100000 + 2i / 2i + 1."""
    rng = np.random.default_rng(seed)
    shifts = rng.integers(-max_shift, max_shift + 1, size=(n_pairs, 3)).astype(np.float64)
    fracs = rng.uniform(-0.5, 0.5, size=(n_pairs, 3))
    m = max_shift + 4
    big = (n + 2 * m,) * 3
    imgs1, imgs2, planted = [], [], []
    fields = {}
    for i in range(n_pairs):
        f = i % n_fields
        if f not in fields:
            fields[f] = smooth_field(big, 2000 + f, device)
        sx, sy, sz = (int(v) for v in shifts[i])
        imgs1.append(tile_from_field(fields[f], (m, m, m), (n, n, n), 100000 + 2 * i))
        src = fields[f]
        total = shifts[i].copy()
        if subpixel_every and i % subpixel_every == 1:
            fx, fy, fz = fracs[i]
            src = fourier_shift(fields[f], (fz, fy, fx))
            total = total + fracs[i]
        imgs2.append(tile_from_field(src, (m + sz, m + sy, m + sx), (n, n, n), 100001 + 2 * i))
        planted.append(tuple(float(v) for v in total))
        del src
    del fields
    return imgs1, imgs2, planted


def grid_pairs_4x4x2():
    """The 112 overlapping pairs of a 4x4x2 tile grid (BASELINE configs[1]): 64 face neighbours plus the 48
    xz / yz edge neighbours."""
    def tid(i, j, k):
        return (k * 4 + j) * 4 + i
    pairs = []
    for k in range(2):
        for j in range(4):
            for i in range(4):
                if i < 3:
                    pairs.append((tid(i, j, k), tid(i + 1, j, k)))
                if j < 3:
                    pairs.append((tid(i, j, k), tid(i, j + 1, k)))
                if k < 1:
                    pairs.append((tid(i, j, k), tid(i, j, k + 1)))
                if i < 3 and k < 1:
                    pairs.append((tid(i, j, k), tid(i + 1, j, k + 1)))
                    pairs.append((tid(i + 1, j, k), tid(i, j, k + 1)))
                if j < 3 and k < 1:
                    pairs.append((tid(i, j, k), tid(i, j + 1, k + 1)))
                    pairs.append((tid(i, j + 1, k), tid(i, j, k + 1)))
    assert len(pairs) == 112
    return pairs


def make_pcm_grid_workload(n=512, device="cuda", seed=43, max_shift=10, n_tiles=32):
    """End-to-end form of config 2: 32 tiles of n^3 uint16 (a 4x4x2 grid), every tile a crop of ONE field at its
    own planted offset s_t, so that any pair (a, b) has the true shift s_b - s_a.  Returns (tiles, offsets_xyz)."""
    rng = np.random.default_rng(seed)
    offs = rng.integers(-max_shift, max_shift + 1, size=(n_tiles, 3))
    m = max_shift + 2
    fld = smooth_field((n + 2 * m,) * 3, 3000, device)
    tiles = [tile_from_field(fld, (m + int(o[2]), m + int(o[1]), m + int(o[0])), (n, n, n), 200000 + t)
             for t, o in enumerate(offs)]
    return tiles, [tuple(int(v) for v in o) for o in offs]


def make_fusion_workload(grid=(4, 4, 4), tile=576, stride=491, device="cuda", n_distinct=4, seed=7,
                         rot_deg=0.0):
    """BASELINE config 3: grid of tile^3 uint16 tiles, registration = translation(stride*(i,j,k)
    + jitter in [-2,2]^3 (default_rng(seed))) (optionally times a small z rotation).  To bound
    set-up time only ``n_distinct`` distinct tile volumes are generated and assigned cyclically;
    every view still has its own registration.  Returns (tiles, models, dims_xyz)."""
    rng = np.random.default_rng(seed)
    vols = [tile_from_field(smooth_field((tile,) * 3, 9000 + d, device), (0, 0, 0), (tile,) * 3, 9100 + d)
            for d in range(n_distinct)]
    tiles, models = [], []
    a = np.deg2rad(rot_deg)
    R = np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]])
    idx = 0
    for k in range(grid[2]):
        for j in range(grid[1]):
            for i in range(grid[0]):
                t = stride * np.array([i, j, k], dtype=np.float64) + rng.uniform(-2, 2, 3)
                c = np.array([tile / 2, tile / 2, 0.0])
                M = np.hstack([R, (t + c - R @ c)[:, None]])
                tiles.append(vols[idx % n_distinct])
                models.append(M)
                idx += 1
    return tiles, models, (tile, tile, tile)
