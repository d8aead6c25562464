"""Synthetic quantized-coefficient generator (SURVEY.md 8d recipe).

The reference ships no test images, so every parity test and the benchmark use
this generator: integer-only pixels (smooth field + hard edges + hash noise)
-> exact integer 8x8 FDCT -> division by scaled JPEG Annex-K tables with
round-half-away.  Everything is integer arithmetic on counter-based hashes, so
the output is bit-identical on every machine and independent of chunking.
"""
from __future__ import annotations

import numpy as np

from .image import CoefImage, Component, JCS_GRAYSCALE, JCS_YCbCr, blocks_for

ANNEX_K_LUMA = np.array([
    16, 11, 10, 16, 24, 40, 51, 61, 12, 12, 14, 19, 26, 58, 60, 55,
    14, 13, 16, 24, 40, 57, 69, 56, 14, 17, 22, 29, 51, 87, 80, 62,
    18, 22, 37, 56, 68, 109, 103, 77, 24, 35, 55, 64, 81, 104, 113, 92,
    49, 64, 78, 87, 103, 121, 120, 101, 72, 92, 95, 98, 112, 100, 103, 99], dtype=np.int64)
ANNEX_K_CHROMA = np.array([
    17, 18, 24, 47, 99, 99, 99, 99, 18, 21, 26, 66, 99, 99, 99, 99,
    24, 26, 56, 99, 99, 99, 99, 99, 47, 66, 99, 99, 99, 99, 99, 99] + [99] * 32, dtype=np.int64)


def quality_table(base: np.ndarray, quality: int) -> np.ndarray:
    """libjpeg's jpeg_quality_scaling + jpeg_add_quant_table (baseline clamp 1..255)."""
    quality = min(max(int(quality), 1), 100)
    scale = 5000 // quality if quality < 50 else 200 - 2 * quality
    return np.clip((base * scale + 50) // 100, 1, 255).astype(np.uint16)


def _fdct_matrix(bits: int = 14) -> np.ndarray:
    u = np.arange(8)[:, None]
    x = np.arange(8)[None, :]
    m = 0.5 * np.cos((2 * x + 1) * u * np.pi / 16)
    m[0, :] *= np.sqrt(0.5)
    return np.rint(m * (1 << bits)).astype(np.float64)   # exact small integers


_M = _fdct_matrix()
_MBITS = 28


def _hash_noise(X, Y, c, seed):
    h = (X.astype(np.uint32) * np.uint32(0x9E3779B1)) ^ (Y.astype(np.uint32) * np.uint32(0x85EBCA77))
    h = h ^ np.uint32((c * 0xC2B2AE3D + seed * 0x27D4EB2F) & 0xFFFFFFFF)
    h ^= h >> np.uint32(15)
    h *= np.uint32(0x2C1B3C6D)
    h ^= h >> np.uint32(12)
    h *= np.uint32(0x297A2D39)
    h ^= h >> np.uint32(15)
    return h


def _tri(t, period, amp):
    return (np.abs(2 * (t % period) - period) * (2 * amp)) // period - amp


def pixels(c: int, y0: int, y1: int, width: int, sx: int, sy: int, seed: int, noise: int = 4) -> np.ndarray:
    """Pixels of component c, rows [y0,y1) x [0,width), as int64 in [0,255].
    (sx, sy) = sub-sampling of this component relative to luma coordinates."""
    Y = (np.arange(y0, y1, dtype=np.int64) * sy)[:, None]
    X = (np.arange(width, dtype=np.int64) * sx)[None, :]
    amp = 1 if c == 0 else 2            # chroma has half the contrast
    v = _tri(X + 61 * c, 483, 30) + _tri(Y + 50 * c, 369, 30) + _tri(X + 2 * Y, 23, 8)
    v = v + (((X // 37) + (Y // 53)) & 1) * 50 - 25
    v = v // amp
    if noise:
        h = _hash_noise(np.broadcast_to(X, v.shape), np.broadcast_to(Y, v.shape), c, seed)
        v = v + (h % np.uint32(2 * noise + 1)).astype(np.int64) - noise
    return np.clip(128 + v, 0, 255)


def quantize_blocks(px: np.ndarray, q: np.ndarray) -> np.ndarray:
    """px int64 [8*hb, 8*wb] -> int16 [hb, wb, 64] quantized coefficients."""
    hb, wb = px.shape[0] // 8, px.shape[1] // 8
    b = (px - 128).astype(np.float64).reshape(hb, 8, wb, 8).transpose(0, 2, 1, 3)
    f = _M @ b @ _M.T                     # exact: |values| < 2^39
    f = f.astype(np.int64).reshape(hb, wb, 64)
    qs = q.astype(np.int64)[None, None, :] << _MBITS
    mag = (np.abs(f) + (qs >> 1)) // qs
    return (np.sign(f) * mag).astype(np.int16)


def make_image(width: int, height: int, subsampling: str = "420", quality: int = 50,
               seed: int = 12345, noise: int = 4, chunk_rows: int = 64, mcu_rows=None) -> CoefImage:
    """subsampling: 'gray', '444', '422', '420', '440'.
    mcu_rows=(m0, m1): generate only the slab of MCU rows [m0, m1) of the image (each
    component then holds block rows [m0*v_samp, min(m1*v_samp, hblk))); the values are
    identical to the corresponding rows of the full image."""
    if subsampling == "gray":
        samp = [(1, 1)]
        cs = JCS_GRAYSCALE
    else:
        hv = {"444": (1, 1), "422": (2, 1), "420": (2, 2), "440": (1, 2)}[subsampling]
        samp = [hv, (1, 1), (1, 1)]
        cs = JCS_YCbCr
    max_h = max(s[0] for s in samp)
    max_v = max(s[1] for s in samp)
    tables = [quality_table(ANNEX_K_LUMA, quality), quality_table(ANNEX_K_CHROMA, quality)]
    comps = []
    for c, (hs, vs) in enumerate(samp):
        wb = blocks_for(width, hs, max_h)
        hb = blocks_for(height, vs, max_v)
        q = tables[0 if c == 0 else 1]
        b0, b1 = 0, hb
        if mcu_rows is not None:
            b0, b1 = min(mcu_rows[0] * vs, hb), min(mcu_rows[1] * vs, hb)
        coef = np.empty((b1 - b0, wb, 64), dtype=np.int16)
        for r0 in range(b0, b1, chunk_rows):
            r1 = min(b1, r0 + chunk_rows)
            px = pixels(c, r0 * 8, r1 * 8, wb * 8, max_h // hs, max_v // vs, seed, noise)
            coef[r0 - b0:r1 - b0] = quantize_blocks(px, q)
        comps.append(Component(coef=coef, quant=q.copy(), h_samp=hs, v_samp=vs,
                               quant_tbl_no=0 if c == 0 else 1))
    return CoefImage(width=width, height=height, colorspace=cs, comps=comps)


# ---- the same generator on a torch device (big configurations are generated on the GPU) ----
def _tri_t(t, period, amp):
    return ((2 * (t % period) - period).abs() * (2 * amp)) // period - amp


def pixels_torch(c, y0, y1, width, sx, sy, seed, noise, device):
    import torch
    Y = (torch.arange(y0, y1, dtype=torch.int64, device=device) * sy)[:, None]
    X = (torch.arange(width, dtype=torch.int64, device=device) * sx)[None, :]
    amp = 1 if c == 0 else 2
    v = _tri_t(X + 61 * c, 483, 30) + _tri_t(Y + 50 * c, 369, 30) + _tri_t(X + 2 * Y, 23, 8)
    v = v + (((X // 37) + (Y // 53)) & 1) * 50 - 25
    v = torch.div(v, amp, rounding_mode="floor")
    if noise:
        M = 0xFFFFFFFF
        h = ((X * 0x9E3779B1) & M) ^ ((Y * 0x85EBCA77) & M)
        h = h ^ ((c * 0xC2B2AE3D + seed * 0x27D4EB2F) & M)
        h = h ^ (h >> 15)
        h = (h * 0x2C1B3C6D) & M
        h = h ^ (h >> 12)
        h = (h * 0x297A2D39) & M
        h = h ^ (h >> 15)
        v = v + (h % (2 * noise + 1)) - noise
    return (128 + v).clamp_(0, 255)


def quantize_blocks_torch(px, q):
    import torch
    hb, wb = px.shape[0] // 8, px.shape[1] // 8
    M = torch.from_numpy(_M).to(px.device)
    b = (px - 128).to(torch.float64).reshape(hb, 8, wb, 8).permute(0, 2, 1, 3)
    f = (M @ b @ M.T).to(torch.int64).reshape(hb, wb, 64)        # exact: |values| < 2^39
    qs = torch.from_numpy(q.astype(np.int64)).to(px.device)[None, None, :] << _MBITS
    mag = (f.abs() + (qs >> 1)) // qs
    return (torch.sign(f) * mag).to(torch.int16)


def make_image_torch(width, height, subsampling="420", quality=50, seed=12345, noise=4,
                     chunk_rows=32, mcu_rows=None, device="cuda"):
    """Same values as make_image, but the coefficient arrays are torch int16 tensors on
    `device` (Component.coef holds the tensor)."""
    import torch
    if subsampling == "gray":
        samp, cs = [(1, 1)], JCS_GRAYSCALE
    else:
        hv = {"444": (1, 1), "422": (2, 1), "420": (2, 2), "440": (1, 2)}[subsampling]
        samp, cs = [hv, (1, 1), (1, 1)], JCS_YCbCr
    max_h = max(s[0] for s in samp)
    max_v = max(s[1] for s in samp)
    tables = [quality_table(ANNEX_K_LUMA, quality), quality_table(ANNEX_K_CHROMA, quality)]
    comps = []
    for c, (hs, vs) in enumerate(samp):
        wb = blocks_for(width, hs, max_h)
        hb = blocks_for(height, vs, max_v)
        q = tables[0 if c == 0 else 1]
        b0, b1 = 0, hb
        if mcu_rows is not None:
            b0, b1 = min(mcu_rows[0] * vs, hb), min(mcu_rows[1] * vs, hb)
        coef = torch.empty((b1 - b0, wb, 64), dtype=torch.int16, device=device)
        for r0 in range(b0, b1, chunk_rows):
            r1 = min(b1, r0 + chunk_rows)
            px = pixels_torch(c, r0 * 8, r1 * 8, wb * 8, max_h // hs, max_v // vs, seed, noise, device)
            coef[r0 - b0:r1 - b0] = quantize_blocks_torch(px, q)
        comps.append(Component(coef=coef, quant=q.copy(), h_samp=hs, v_samp=vs,
                               quant_tbl_no=0 if c == 0 else 1))
    return CoefImage(width=width, height=height, colorspace=cs, comps=comps)
