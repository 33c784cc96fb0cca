"""Small ncu target: ONE bitplane of the bench workload (8192x8192, 8x8 patches, 32 atoms) through
the C ABI on one stream: extract -> init -> learn -> Golomb(D, A, E). Run under
`ncu --set full -k regex:...` (see profiles/README.md)."""
import importlib
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
bic = importlib.import_module("binary-image-compression_b200")
synth = bic.synth

plane = int(sys.argv[1]) if len(sys.argv) > 1 else 10
S, W, K = 8192, 8, 32
import torch  # noqa: E402
img = synth.smooth_pgm16(S, S, seed=2, device="cuda:0")
pay = synth.pbm_bytes_torch(synth.bitplane(img, plane)).cpu().numpy()
del img
ctx = bic.Context(0)
R = ctx.matrix(S, S)
R.upload_pbm(pay)
X = ctx.extract_patches(R, W)
n, m = X.rows, W * W
D, A, E = ctx.matrix(K, m), ctx.matrix(n, K), ctx.matrix(n, m)
ctx.initialize_model_neighbor(X, D, A, ctx.rand48(34503498))
it, tr = ctx.learn_model_traditional(X, E, D, A)
bits = [ctx.golomb_encode(M).info.bitcount for M in (D, A, E)]
print("plane", plane, "iterations", it, "trace", tr.tolist(), "golomb bits", bits, "launches", ctx.launches)
ctx.close()
