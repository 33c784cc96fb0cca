"""Synthetic inputs of the shapes BASELINE.json names (there are no data files in the
reference). Host-side numpy only; used by tests/ and bench.py.

Bit matrices are exchanged in the reference's word layout (src/binmat.h:114-116,
src/binmat.cpp:140-149): row-major uint64 words, ceil(cols/64) per row, bit j of a row at
MSB >> (j % 64); pad bits zero.
"""
from __future__ import annotations

import numpy as np


def wpr64(cols: int) -> int:
    return (cols + 63) // 64


def pack_rows(bits: np.ndarray) -> np.ndarray:
    """dense {0,1} array (rows, cols) -> (rows, ceil(cols/64)) uint64 in reference layout."""
    bits = np.asarray(bits, dtype=np.uint8)
    rows, cols = bits.shape
    w = wpr64(cols)
    by = np.packbits(bits, axis=1, bitorder="big")
    pad = w * 8 - by.shape[1]
    if pad:
        by = np.concatenate([by, np.zeros((rows, pad), np.uint8)], axis=1)
    return np.ascontiguousarray(by).view(">u8").astype(np.uint64).reshape(rows, w)


def unpack_rows(words: np.ndarray, cols: int) -> np.ndarray:
    """inverse of pack_rows -> dense uint8 (rows, cols)."""
    words = np.ascontiguousarray(words, dtype=np.uint64)
    rows = words.shape[0]
    by = words.astype(">u8").view(np.uint8).reshape(rows, -1)
    return np.unpackbits(by, axis=1, bitorder="big")[:, :cols]


def pbm_bytes(bits: np.ndarray) -> np.ndarray:
    """dense (rows, cols) -> P4 payload, rows padded to whole bytes (src/pbm.cpp:54-77)."""
    return np.packbits(np.asarray(bits, np.uint8), axis=1, bitorder="big")


def structured_page(rows: int = 3508, cols: int = 2480, seed: int = 7, salt: float = 0.003) -> np.ndarray:
    """Text-like binary page: an alphabet of 40 stroke glyphs (24 high x 16 wide, a few 3-px
    bars each) stamped on a 32 x 16 px grid inside margins, plus salt noise. ~14 % ink.
    Pure Bernoulli noise is a degenerate input for bsvd (2 iterations, no atom ever changes),
    so every config uses this generator."""
    rng = np.random.default_rng(seed)
    gh, gw = 24, 16
    glyphs = np.zeros((40, gh, gw), np.uint8)
    for g in range(40):
        for _ in range(int(rng.integers(2, 5))):
            if rng.random() < 0.5:  # horizontal bar
                y = int(rng.integers(0, gh - 3))
                x0 = int(rng.integers(0, gw // 2))
                x1 = int(rng.integers(x0 + 4, gw + 1))
                glyphs[g, y:y + 3, x0:x1] = 1
            else:  # vertical bar
                x = int(rng.integers(0, gw - 3))
                y0 = int(rng.integers(0, gh // 2))
                y1 = int(rng.integers(y0 + 4, gh + 1))
                glyphs[g, y0:y1, x:x + 3] = 1
    page = np.zeros((rows, cols), np.uint8)
    margin = min(100, rows // 8, cols // 8)
    cell_h, cell_w = 32, 16
    ny = max(0, (rows - 2 * margin - gh) // cell_h + 1) if rows - 2 * margin >= gh else 0
    nx = max(0, (cols - 2 * margin - gw) // cell_w + 1) if cols - 2 * margin >= gw else 0
    if ny and nx:
        which = rng.integers(0, 40, size=(ny, nx))
        on = rng.random((ny, nx)) < 0.85
        for iy in range(ny):
            y = margin + iy * cell_h
            for ix in range(nx):
                if on[iy, ix]:
                    x = margin + ix * cell_w
                    page[y:y + gh, x:x + gw] |= glyphs[which[iy, ix]]
    if salt > 0:
        page ^= (rng.random((rows, cols)) < salt).astype(np.uint8)
    return page


def structured_canvas(rows: int, cols: int, seed: int = 4) -> np.ndarray:
    """Large raster made by tiling A4 pages from structured_page (config 4)."""
    th, tw = 3508, 2480
    out = np.zeros((rows, cols), np.uint8)
    s = seed * 1000
    for y in range(0, rows, th):
        for x in range(0, cols, tw):
            h, w = min(th, rows - y), min(tw, cols - x)
            out[y:y + h, x:x + w] = structured_page(th, tw, seed=s)[:h, :w]
            s += 1
    return out


_COS_TABLE = None


def _cos_table() -> np.ndarray:
    global _COS_TABLE
    if _COS_TABLE is None:
        k = np.arange(65536, dtype=np.float64)
        _COS_TABLE = np.round(32767.0 * np.cos(2.0 * np.pi * k / 65536.0)).astype(np.int64)
    return _COS_TABLE


def _pgm16_params(seed: int):
    rng = np.random.default_rng(seed)
    u = rng.integers(0, 49, size=64).astype(np.int64)      # cycles per 65536 rows
    v = rng.integers(0, 49, size=64).astype(np.int64)
    w = rng.integers(0, 65536, size=64).astype(np.int64)   # phase
    a = rng.integers(50, 256, size=64).astype(np.int64)    # amplitude
    return u, v, w, a


def smooth_pgm16(rows: int = 8192, cols: int = 8192, seed: int = 2, noise_bits: int = 6,
                 y0: int = 0, x0: int = 0, device=None):
    """16-bit grey image (config 2): a smooth field (sum of 64 low-frequency cosines) with hashed
    uniform noise in the low `noise_bits` bits, so planes run from structured (MSBs) to ~50 %
    density (LSBs). All integer arithmetic (cosine through a 65536-entry table, noise through an
    integer hash of the pixel coordinates), so any crop (y0, x0, rows, cols) of the infinite image
    is bit-identical whether it is produced by numpy (device=None) or by torch on a GPU."""
    u, v, w, a = _pgm16_params(seed)
    T = _cos_table()
    if device is None:
        y = (np.arange(rows, dtype=np.int64) + y0)[:, None]
        x = (np.arange(cols, dtype=np.int64) + x0)[None, :]
        acc = np.zeros((rows, cols), np.int64)
        for i in range(64):
            acc += a[i] * T[(u[i] * y * 8 + v[i] * x * 8 + w[i]) & 0xFFFF]
        S = int(a.sum()) * 32767
        img = np.clip(((3 * acc + S) * 65535) // (2 * S), 0, 65535)  # gain 3: the sum of 64 random phases rarely leaves +-S/3
        h = (y * 73856093) ^ (x * 19349663) ^ (seed * 83492791)
        h = ((h ^ (h >> 13)) * 1540483477) & 0xFFFFFFFF
        h = (h ^ (h >> 15)) & ((1 << noise_bits) - 1)
        img = (img & ~((1 << noise_bits) - 1)) | h
        return img.astype(np.uint16)
    import torch
    y = (torch.arange(rows, dtype=torch.int64, device=device) + y0)[:, None]
    x = (torch.arange(cols, dtype=torch.int64, device=device) + x0)[None, :]
    Tt = torch.from_numpy(T).to(device)
    acc = torch.zeros((rows, cols), dtype=torch.int64, device=device)
    for i in range(64):
        acc += int(a[i]) * Tt[(int(u[i]) * 8 * y + int(v[i]) * 8 * x + int(w[i])) & 0xFFFF]
    S = int(a.sum()) * 32767
    img = torch.clamp(torch.div((3 * acc + S) * 65535, 2 * S, rounding_mode="floor"), 0, 65535)
    h = (y * 73856093) ^ (x * 19349663) ^ (seed * 83492791)
    h = ((h ^ (h >> 13)) * 1540483477) & 0xFFFFFFFF
    h = (h ^ (h >> 15)) & ((1 << noise_bits) - 1)
    img = (img & ~((1 << noise_bits) - 1)) | h
    return img  # int64 tensor with values in [0, 65535]


def bitplane(img, b: int):
    """plane b (0 = LSB) exactly as bitplane_tool does it (src/bitplane_tool.cpp:24-39)."""
    if isinstance(img, np.ndarray):
        return ((img >> np.uint16(b)) & np.uint16(1)).astype(np.uint8)
    return ((img >> b) & 1).to(dtype=__import__("torch").uint8)


def pbm_bytes_torch(bits):
    """torch twin of pbm_bytes: dense {0,1} uint8 (rows, cols) -> P4 payload rows (uint8)"""
    import torch
    rows, cols = bits.shape
    pad = (-cols) % 8
    if pad:
        bits = torch.nn.functional.pad(bits, (0, pad))
    w = torch.tensor([128, 64, 32, 16, 8, 4, 2, 1], dtype=torch.int32, device=bits.device)
    return (bits.reshape(rows, -1, 8).to(torch.int32) * w).sum(dim=2).to(torch.uint8)


def bernoulli_bits(nbits: int, rho: float, seed: int = 5) -> np.ndarray:
    """i.i.d. Bernoulli(rho) bit array as a 1 x nbits dense matrix (config 5)."""
    rng = np.random.default_rng(seed)
    return (rng.random((1, nbits)) < rho).astype(np.uint8)
