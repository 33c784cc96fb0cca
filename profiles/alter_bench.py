"""SURVEY 8f row 4 measured: the role-switched learners (src/bsvd.cpp:1245-1434) on one synthetic A4 page, 8x8 patches,
32 atoms, through the C ABI on one B200 and through the compiled reference (oracle/_ref, all host threads) on the same
inputs; results compared bit for bit. Usage (GPU box): python profiles/alter_bench.py [K=32]"""
import importlib
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from oracle_bindings import Oracle, load_reference  # noqa: E402  (the checker, timed as the CPU baseline)
from test_oracle_mdl_cpu import valid_bits  # noqa: E402

bic = importlib.import_module("binary-image-compression_b200")
synth = bic.synth
K = int(sys.argv[1]) if len(sys.argv) > 1 else 32
rows, cols, W, seed = 3508, 2480, 8, 34503498
m = W * W
oracle, ref = Oracle(), load_reference()
page = synth.structured_page(rows, cols, seed=7)
Xo = oracle.extract_patches(synth.pack_rows(page), rows, cols, W)
n = Xo.shape[0]
D0, A0, _ = oracle.init_neighbor(Xo, m, K, seed)
ctx = bic.Context(0)
for variant in (1, 2, 3):
    for warm in (True, False):
        X, E, D, A = ctx.matrix(n, m, Xo), ctx.matrix(n, m), ctx.matrix(K, m, D0), ctx.matrix(n, K, A0)
        l0 = ctx.launches
        ctx.sync()
        t0 = time.perf_counter()
        it = ctx.learn_model_alter(variant, X, E, D, A)
        ctx.sync()
        t_gpu = time.perf_counter() - t0
    rec = {"learner": f"learn_model_alter{variant}", "page": f"A4 {cols}x{rows}, {W}x{W} patches, {n} patches, {K} atoms", "iterations": it,
           "weight_E": E.weight(), "b200_s": round(t_gpu, 4), "b200_launches": ctx.launches - l0}
    if ref is not None and getattr(ref, "has_alter", False):
        Dr, Ar = D0.copy(), A0.copy()
        t0 = time.perf_counter()
        Er, itr = ref.learn_alter(variant, Xo, Dr, Ar, m, K)
        t_ref = time.perf_counter() - t0
        same = (itr == it and np.array_equal(valid_bits(Er, m), E.download()) and np.array_equal(valid_bits(Dr, m), D.download())
                and np.array_equal(valid_bits(Ar, K), A.download()))
        rec.update({"reference_s": round(t_ref, 3), "reference_threads": ref.max_threads(), "speedup": round(t_ref / t_gpu, 1),
                    "identical_to_reference": bool(same)})
    print(json.dumps(rec), flush=True)
ctx.close()
