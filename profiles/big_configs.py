"""Full-size runs of BASELINE.json configs[0] (A4 page, 8x8/32) and configs[3] (one 65536 x 65536 raster,
32x32 patches, 1024 atoms) on ONE B200, with the size-independent checks of the test-suite:
learner fixed point, E == A*D xor X, Golomb round trip of E. Prints one JSON line per config.
Usage (GPU box): python profiles/big_configs.py [side=65536]"""
import importlib
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402

bic = importlib.import_module("binary-image-compression_b200")
synth = bic.synth
ctx = bic.Context(0)


def run(name, payload, rows, cols, W, K):
    R = ctx.matrix(rows, cols)
    R.upload_pbm(payload)
    ctx.sync()
    t0 = time.perf_counter()
    ctx.timer_start()
    X = ctx.extract_patches(R, W)
    n, m = X.rows, W * W
    D, A, E = ctx.matrix(K, m), ctx.matrix(n, K), ctx.matrix(n, m)
    ctx.initialize_model_neighbor(X, D, A, ctx.rand48(34503498))
    iters, tr = ctx.learn_model_traditional(X, E, D, A)
    streams = [ctx.golomb_encode(M) for M in (D, A, E)]
    ms = ctx.timer_stop()
    wall = time.perf_counter() - t0
    # checks
    fixed = ctx.update_coefficients(E, D, A) == 0 and ctx.update_dictionary(E, D, A) == 0
    E2 = ctx.matrix(n, m)
    ctx.residual(X, A, D, E2)
    L = ctx.L
    import ctypes as C
    d = C.c_uint64(0)
    ctx._ck(L.bic_mat_dist(ctx.h, E.h, E2.h, C.byref(d)))
    E3 = ctx.matrix(n, m)
    ctx.golomb_decode(streams[2], E3)
    d2 = C.c_uint64(0)
    ctx._ck(L.bic_mat_dist(ctx.h, E.h, E3.h, C.byref(d2)))
    rec = {"config": name, "rows": rows, "cols": cols, "W": W, "K": K, "n": n, "iterations": iters,
           "changed_atoms_per_iteration": [int(x) for x in tr[:, 1]][:12], "device_ms": ms, "wall_s": wall,
           "Mpixel_per_s": rows * cols / 1e6 / (ms / 1e3), "weight_X": X.weight(), "weight_E": E.weight(), "weight_A": A.weight(),
           "golomb_bits": [int(s.info.bitcount) for s in streams], "raw_bits": rows * cols,
           "fixed_point": bool(fixed), "E_equals_AD_xor_X": int(d.value) == 0, "golomb_roundtrip_E": int(d2.value) == 0,
           "launches": ctx.launches}
    print(json.dumps(rec), flush=True)
    for mm in (R, X, D, A, E, E2, E3):
        mm.destroy()
    for s in streams:
        s.destroy()


# configs[0]: A4 @ 300 dpi, 8x8 patches, 32 atoms
page = synth.structured_page(3508, 2480, seed=7)
run("configs[0] A4 2480x3508, 8x8/32", synth.pbm_bytes(page), 3508, 2480, 8, 32)
run("A4 2480x3508, 16x16/256", synth.pbm_bytes(page), 3508, 2480, 16, 256)

# configs[3]: one big raster = bitplane 12 of the synthetic 16-bit field (structured contours), built in bands on the GPU
side = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
payload = np.empty((side, side // 8), np.uint8)
band = 2048
for y0 in range(0, side, band):
    img = synth.smooth_pgm16(band, side, seed=4, y0=y0, device="cuda:0")
    payload[y0:y0 + band] = synth.pbm_bytes_torch(synth.bitplane(img, 12)).cpu().numpy()
    del img
torch.cuda.empty_cache()
run(f"configs[3] {side}x{side}, 32x32/1024", payload, side, side, 32, 1024)
ctx.close()
