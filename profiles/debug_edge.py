import importlib, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
from oracle_bindings import Oracle
bic = importlib.import_module("binary-image-compression_b200"); synth = bic.synth
o = Oracle(); ctx = bic.Context(0)
for (n, m, p) in [(1, 1, 1), (1, 64, 1), (3, 5, 2), (33, 31, 1), (2, 64, 64)]:
    rng = np.random.default_rng(n * 100 + m + p)
    bits = (rng.random((n, m)) < 0.6).astype(np.uint8); bits[0, 0] = 1
    Xw = synth.pack_rows(bits)
    piv, _ = o.draw_pivots(Xw, m, p, o.rng(3))
    X = ctx.matrix(n, m, Xw); D, A, E = ctx.matrix(p, m), ctx.matrix(n, p), ctx.matrix(n, m)
    ctx.initialize_model_neighbor_pivots(X, piv, D, A)
    for name, fn in (("residual", lambda: ctx.residual(X, A, D, E)), ("coef", lambda: ctx.update_coefficients(E, D, A)),
                     ("dict", lambda: ctx.update_dictionary(E, D, A)), ("dict0", lambda: (ctx.set_option("dict_algo", 0), ctx.update_dictionary(E, D, A), ctx.set_option("dict_algo", 1)))):
        try:
            fn(); ctx.sync(); print((n, m, p), name, "ok")
        except Exception as ex:
            print((n, m, p), name, "FAILED", ex)
