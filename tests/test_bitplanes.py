"""Bit planes of a grey image (src/bitplane_tool.cpp:24-39, SURVEY 8f row 4 / BASELINE.json configs[1]'s first step).
CPU: the oracle against the reference's own bitplane_tool binary (oracle/_ref/bitplane_tool) on small P5 files.
GPU: bic_split_bitplanes through the C ABI against the oracle, incl. widths that are not multiples of 8 / 32 / 256."""
import subprocess
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
TOOL = ROOT / "oracle" / "_ref" / "bitplane_tool"


def p5_payload(img, maxval):
    img = np.asarray(img)
    if maxval < 256:
        return img.astype(np.uint8).reshape(-1)
    out = np.empty(img.size * 2, np.uint8)
    out[0::2] = (img.reshape(-1) >> 8) & 0xFF     # high byte first, src/pnm.cpp:119-121
    out[1::2] = img.reshape(-1) & 0xFF
    return out


def read_p4(path):
    data = path.read_bytes()
    tok, pos = [], 0
    while len(tok) < 3:                            # magic, cols, rows
        while data[pos:pos + 1].isspace():
            pos += 1
        s = pos
        while not data[pos:pos + 1].isspace():
            pos += 1
        tok.append(data[s:pos])
    pos += 1
    cols, rows = int(tok[1]), int(tok[2])
    return np.frombuffer(data, np.uint8, rows * ((cols + 7) // 8), pos).reshape(rows, -1), rows, cols


@pytest.mark.parametrize("rows,cols,maxval", [(13, 37, 65535), (20, 64, 255), (9, 100, 1023), (5, 8, 15)])
def test_oracle_vs_reference_bitplane_tool(tmp_path, oracle, synth, rows, cols, maxval):
    if not TOOL.exists():
        pytest.skip("oracle/_ref/bitplane_tool not built (reference sources not mounted)")
    rng = np.random.default_rng(rows * cols)
    img = rng.integers(0, maxval + 1, size=(rows, cols))
    pay = p5_payload(img, maxval)
    (tmp_path / "in.pgm").write_bytes(f"P5\n{cols} {rows}\n{maxval}\n".encode() + pay.tobytes())
    r = subprocess.run([str(TOOL), "in.pgm"], cwd=tmp_path, capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, r.stderr
    planes = oracle.split_bitplanes(pay, rows, cols, maxval)
    files = sorted(tmp_path.glob("plane_*.pbm"))
    assert len(files) == planes.shape[0]
    for bi, f in enumerate(files):
        bits, fr, fc = read_p4(f)
        assert (fr, fc) == (rows, cols)
        want = synth.pbm_bytes(synth.unpack_rows(planes[bi], cols))
        assert np.array_equal(bits, want), f"plane {bi}"
        assert np.array_equal(synth.unpack_rows(planes[bi], cols), ((img >> bi) & 1).astype(np.uint8))


@pytest.mark.gpu
@pytest.mark.parametrize("rows,cols,maxval", [(64, 512, 65535), (33, 264, 65535), (17, 300, 65535), (7, 8, 65535), (40, 1000, 65535),
                                              (19, 77, 65535), (25, 256, 255), (11, 45, 1023), (3, 2048, 65535)])
def test_split_bitplanes_vs_oracle(oracle, rows, cols, maxval):
    import importlib
    bic = importlib.import_module("binary-image-compression_b200")
    ctx = bic.Context(0)
    rng = np.random.default_rng(rows + cols)
    img = rng.integers(0, maxval + 1, size=(rows, cols))
    pay = p5_payload(img, maxval)
    want = oracle.split_bitplanes(pay, rows, cols, maxval)
    planes = ctx.split_bitplanes(pay, rows, cols, maxval)
    assert len(planes) == want.shape[0]
    for bi, M in enumerate(planes):
        assert np.array_equal(M.download(), want[bi]), f"plane {bi}"
        M.destroy()
    ctx.close()


@pytest.mark.gpu
def test_split_bitplanes_full_size_matches_the_bench_generator(synth):
    """the bench's planes (bitplane() of the synthetic 16-bit image) are what the kernel makes of the image's P5 payload"""
    import importlib
    import torch
    bic = importlib.import_module("binary-image-compression_b200")
    ctx = bic.Context(0)
    S = 2048
    img = synth.smooth_pgm16(S, S, seed=2, device="cuda:0")
    pay = torch.stack(((img >> 8) & 0xFF, img & 0xFF), dim=-1).to(torch.uint8).reshape(-1).cpu().numpy()
    planes = ctx.split_bitplanes(pay, S, S, 65535)
    for b in (0, 5, 11, 15):
        want = synth.pbm_bytes_torch(synth.bitplane(img, b)).cpu().numpy()
        assert np.array_equal(planes[b].download_pbm(), want)
    ctx.close()


@pytest.mark.gpu
def test_pgm_to_containers_equals_plane_by_plane(synth):
    """configs[1] end to end: P5 payload -> bic_split_bitplanes -> bic_encode_raster_resident gives, plane for plane, the
    container bic_encode_raster makes of that plane's P4 payload, and the containers decode back to the planes"""
    import importlib
    bic = importlib.import_module("binary-image-compression_b200")
    ctx = bic.Context(0)
    S = 512
    img = synth.smooth_pgm16(S, S, seed=2)
    pay = p5_payload(img.astype(np.int64), 65535)
    planes = ctx.split_bitplanes(pay, S, S, 65535)
    assert len(planes) == 16
    for b in (0, 7, 12, 15):
        p4 = synth.pbm_bytes(synth.bitplane(img, b))
        if not p4.any():
            continue  # an all-zero plane has no pivot (the reference's initialiser never returns on it)
        c1, i1 = ctx.encode_raster(p4, S, S, 8, 32)
        c2, i2 = ctx.encode_raster_resident(planes[b], 8, 32)
        assert i1.container_bytes == i2.container_bytes and np.array_equal(np.array(c1), np.array(c2))
        back, r, c_ = ctx.decode_raster(np.array(c2))
        assert (r, c_) == (S, S) and np.array_equal(back, p4)
    ctx.close()
