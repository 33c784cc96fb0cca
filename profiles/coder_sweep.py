"""BASELINE.json configs[4]: Golomb residual-coding-only sweep over densities 0.1 % - 50 %.
Device encode + decode GB/s (of packed input) vs the reference's serial GolombCoder on the host
(oracle/_ref, bit counting only -- the reference writes no bits). i.i.d. Bernoulli(rho) bit arrays.
Usage (GPU box): python profiles/coder_sweep.py [log2_bits=31] [rho,rho,...] > gpurun_out/coder_sweep.json"""
import importlib
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import torch  # noqa: E402
from oracle_bindings import load_reference  # noqa: E402

bic = importlib.import_module("binary-image-compression_b200")
synth = bic.synth
LOG2 = int(sys.argv[1]) if len(sys.argv) > 1 else 31
N = 1 << LOG2
cols = 1 << 15
rows = N // cols
ctx = bic.Context(0)
import os  # noqa: E402
for kv in filter(None, os.environ.get("BIC_SWEEP_OPTS", "").split(",")):   # e.g. BIC_SWEEP_OPTS=gol_list=0,gol_scan=0
    k, v = kv.split("=")
    ctx.set_option(k, int(v))
ref = load_reference()
out = []
RHOS = [float(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0.001, 0.003, 0.01, 0.03, 0.1, 0.2, 0.5]
for rho in RHOS:
    g = torch.Generator(device="cuda").manual_seed(5)
    M = ctx.matrix(rows, cols)
    chunk_rows = 4096
    host = np.empty((rows, cols // 8), np.uint8)
    for r0 in range(0, rows, chunk_rows):
        bits = (torch.rand((chunk_rows, cols), device="cuda", generator=g) < rho).to(torch.uint8)
        host[r0:r0 + chunk_rows] = synth.pbm_bytes_torch(bits).cpu().numpy()
    M.upload_pbm(host)
    s = ctx.stream()
    ctx.golomb_encode(M, out=s)          # warm-up (sizes the stream buffers)
    M2 = ctx.matrix(rows, cols)
    ctx.golomb_decode(s, M2)
    reps = 5
    ctx.timer_start()
    for _ in range(reps):
        ctx.golomb_encode(M, out=s)
    ms_enc = ctx.timer_stop() / reps
    ctx.timer_start()
    for _ in range(reps):
        ctx.golomb_decode(s, M2)
    ms_dec = ctx.timer_stop() / reps
    ok = bool(np.array_equal(M2.download_pbm(), host))
    # per-kernel device times of one more encode + decode (event pairs around every launch)
    ctx.prof_reset()
    ctx.prof_enable(True)
    ctx.golomb_encode(M, out=s)
    ctx.golomb_decode(s, M2)
    ctx.prof_enable(False)
    kernels = {k: round(v[1], 4) for k, v in ctx.prof_stats().items()}
    info = s.info
    # serial reference coder on a 2^26-bit prefix (single core, bit counting only)
    sample_rows = (1 << 26) // cols
    words = synth.pack_rows(np.unpackbits(host[:sample_rows], axis=1))
    t0 = time.perf_counter()
    ref_bits = ref.golomb_matrix(words, cols) if ref else None
    t_ref = time.perf_counter() - t0
    gb = N / 8 / 1e9
    rec = {"rho": rho, "input_bits": N, "bitcount": int(info.bitcount), "ratio": info.bitcount / N, "nsamples": int(info.nsamples),
           "encode_ms": ms_enc, "decode_ms": ms_dec, "encode_GBps_in": gb / (ms_enc / 1e3), "decode_GBps_out": gb / (ms_dec / 1e3),
           "encode_GBps_in_plus_out": (gb + info.bitcount / 8e9) / (ms_enc / 1e3), "roundtrip_ok": ok,
           "ref_serial_GBps_in": ((1 << 26) / 8 / 1e9) / t_ref if ref else None, "ref_sample_bits": 1 << 26, "kernel_ms": kernels}
    out.append(rec)
    print(json.dumps(rec), flush=True)
    for x in (M, M2, s):
        x.destroy()
ctx.close()
