"""bic_split_bitplanes (src/bitplane_tool.cpp:24-39) on the bench image: one 16-bit 8192 x 8192 P5 payload -> 16 rasters.
Times the kernel alone (payload resident in HBM; per-launch CUDA events from the library, bic_prof_*) and the call from a
pinned host buffer. Algorithmic bytes: 2 B in + 16 bits out per pixel = 4 B/pixel. Usage: python profiles/bitplane_bench.py"""
import ctypes as C
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
S = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
ctx = bic.Context(0)
img = synth.smooth_pgm16(S, S, seed=2, device="cuda:0")
pay_t = torch.stack(((img >> 8) & 0xFF, img & 0xFF), dim=-1).to(torch.uint8).reshape(-1)
host = ctx.pinned(pay_t.numel())
host[:] = pay_t.cpu().numpy()
del img, pay_t
planes = [ctx.matrix(S, S) for _ in range(16)]
ctx.split_bitplanes(host, S, S, 65535, planes)          # warm-up
ctx.prof_reset()
ctx.prof_enable(True)
reps = 20
t0 = time.perf_counter()
for _ in range(reps):
    ctx.split_bitplanes(host, S, S, 65535, planes)
wall = (time.perf_counter() - t0) / reps
ctx.prof_enable(False)
prof = ctx.prof_stats()
n, ms = prof["k_bitplanes"]
peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {}
peak = float(peaks.get("hbm_gbs", 6650.0))
alg = 4.0 * S * S
gbs = alg / (ms / n / 1e3) / 1e9
print(json.dumps({"kernel": "k_bitplanes16", "image": f"{S}x{S} 16-bit P5 payload", "launches": n, "avg_kernel_us": round(1e3 * ms / n, 2),
                  "algorithmic_bytes": alg, "achieved_gbs": round(gbs, 1), "hbm_peak_gbs": peak, "frac_of_hbm_peak": round(gbs / peak, 3),
                  "from_pinned_host_ms": round(1e3 * wall, 3), "from_pinned_host_gpixel_s": round(S * S / wall / 1e9, 2)}))
ctx.close()
