"""Generates tests/golden/mdl_*.npz by running the COMPILED REFERENCE (oracle/_ref/libbic_ref.so, built by
oracle/Makefile from /root/reference/src) on the seeded inputs of tests/test_oracle_mdl_cpu.py::MDL_CASES:
learn_model_mdl_forward_selection / _backward_selection / _full_search (src/bsvd.cpp:1463-1717) after
initialize_model_neighbor with the same seed. Run here (the reference is not on the GPU box):
    python tests/golden/make_golden_mdl.py"""
import ctypes as C
import importlib
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from oracle_bindings import Oracle, load_reference, _p64, u64, wpr  # noqa: E402
from test_oracle_mdl_cpu import MDL_CASES, mdl_inputs, valid_bits  # noqa: E402

synth = importlib.import_module("binary-image-compression_b200.synth")
oracle, ref = Oracle(), load_reference()
assert ref is not None and ref.has_mdl
for name, rows, cols, W, K, seed, lm in MDL_CASES:
    m = W * W
    X, D, A, _ = mdl_inputs(oracle, synth, rows, cols, W, K, seed)
    Dr, Ar, _ = ref.init_neighbor(X, m, K, seed)
    n = X.shape[0]
    E = np.zeros_like(X)
    h = ref.lib.ref_learn_mdl(lm, _p64(X), _p64(E), _p64(Dr), _p64(Ar), n, m, K, 0, 0)
    pk, L = u64(0), u64(0)
    ref.lib.ref_mdl_result_info(h, C.byref(pk), C.byref(L))
    pk = int(pk.value)
    Do, Ao = np.zeros((pk, wpr(m)), np.uint64), np.zeros((n, wpr(pk) if pk else 0), np.uint64)
    if pk:
        ref.lib.ref_mdl_result_copy(h, _p64(Do), _p64(Ao))
    ref.lib.ref_mdl_result_free(h)
    np.savez_compressed(ROOT / "tests" / "golden" / f"mdl_{name}.npz", X=X, D=Do, A=valid_bits(Ao, pk), E=E,
                        p=np.uint64(pk), bestL=np.uint64(int(L.value)))
    print(name, "p", pk, "bestL", int(L.value))
