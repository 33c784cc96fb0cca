"""SURVEY 8f row 3 measured: the MDL learners (src/bsvd.cpp:1463-1660) on one synthetic A4 page, 8x8 patches,
through the C ABI on one B200 and through the compiled reference (oracle/_ref, all host threads) on the same inputs;
results compared bit for bit. Usage (GPU box): python profiles/mdl_bench.py [K=16]"""
import ctypes as C
import importlib
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from oracle_bindings import Oracle, load_reference, _p64, u64, wpr  # noqa: E402  (the checker, timed as the CPU baseline)

bic = importlib.import_module("binary-image-compression_b200")
synth = bic.synth
K = int(sys.argv[1]) if len(sys.argv) > 1 else 16
rows, cols, W, seed = 3508, 2480, 8, 34503498
m = W * W
oracle, ref = Oracle(), load_reference()
page = synth.structured_page(rows, cols, seed=7)
Xo = oracle.extract_patches(synth.pack_rows(page), rows, cols, W)
n = Xo.shape[0]
ctx = bic.Context(0)
for lm, name in ((4, "forward selection"), (5, "backward selection")):
    for warm in (True, False):   # the first pass loads the kernels and sizes the scratch areas; the second is timed
        X, E, D, A = ctx.matrix(n, m, Xo), ctx.matrix(n, m), ctx.matrix(K, m), ctx.matrix(n, K)
        rng = ctx.rand48(seed)
        ctx.initialize_model_neighbor(X, D, A, rng)
        if warm:
            ctx.learn_model_mdl(lm, X, E, D, A, rng)
    l0 = ctx.launches
    ctx.sync()
    t0 = time.perf_counter()
    bestL, Dn, An = ctx.learn_model_mdl(lm, X, E, D, A, rng)
    ctx.sync()
    t_gpu = time.perf_counter() - t0
    rec = {"learner": name, "page": f"A4 {cols}x{rows}, {W}x{W} patches, {n} patches", "K_start": K, "K_selected": Dn.rows if Dn is not None else 0,
           "best_codelength_bits": bestL, "raw_bits": rows * cols, "b200_s": round(t_gpu, 4), "b200_launches": ctx.launches - l0}
    if ref is not None and ref.has_mdl:
        Dr, Ar, _ = ref.init_neighbor(Xo, m, K, seed)
        Er = np.zeros_like(Xo)
        t0 = time.perf_counter()
        h = ref.lib.ref_learn_mdl(lm, _p64(Xo), _p64(Er), _p64(Dr), _p64(Ar), n, m, K, 0, 0)
        t_ref = time.perf_counter() - t0
        pk, L = u64(0), u64(0)
        ref.lib.ref_mdl_result_info(h, C.byref(pk), C.byref(L))
        pk = int(pk.value)
        Dref, Aref = np.zeros((pk, wpr(m)), np.uint64), np.zeros((n, wpr(pk) if pk else 0), np.uint64)
        if pk:
            ref.lib.ref_mdl_result_copy(h, _p64(Dref), _p64(Aref))
        ref.lib.ref_mdl_result_free(h)
        same = (pk == rec["K_selected"] and int(L.value) == bestL and np.array_equal(Er, E.download())
                and (pk == 0 or np.array_equal(Dref, Dn.download())))
        rec.update({"reference_s": round(t_ref, 3), "reference_threads": ref.max_threads(), "speedup": round(t_ref / t_gpu, 1),
                    "identical_to_reference": bool(same)})
    print(json.dumps(rec), flush=True)
ctx.close()
