"""CPU tests of the boundary: the C-ABI library loads, exports every symbol include/bic_b200.h
declares, and refuses to run without a GPU (no CPU fallback). No compute calls here."""
import ctypes
import importlib
import subprocess

import numpy as np
import pytest

bic = importlib.import_module("binary-image-compression_b200")


@pytest.fixture(scope="module")
def lib():
    if not bic.LIB_PATH.exists():
        bic.build()
    return bic.lib()


def test_library_exports_every_declared_symbol(lib):
    names = bic.exported_symbols_declared_in_header()
    assert len(names) >= 40
    out = subprocess.run(["nm", "-D", "--defined-only", str(bic.LIB_PATH)], capture_output=True, text=True).stdout
    exported = {line.split()[-1] for line in out.splitlines() if line.strip()}
    missing = [n for n in names if n not in exported]
    assert not missing, f"declared in bic_b200.h but not exported: {missing}"


def test_library_is_sm100a_only():
    out = subprocess.run(["cuobjdump", "--list-elf", str(bic.LIB_PATH)], capture_output=True, text=True).stdout
    archs = {tok for line in out.splitlines() for tok in line.replace(".", " ").split() if tok.startswith("sm_")}
    assert archs == {"sm_100a"}, archs


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    h = ctypes.c_void_p()
    st = lib.bic_ctx_create(0, ctypes.byref(h))
    assert st == 5  # BIC_ERR_NO_DEVICE
    with pytest.raises(bic.BicError):
        bic.Context(0)


def test_rand48_host_helper_matches_oracle(lib, oracle):
    """the pivot RNG lives on the host side of the C ABI, so it can be checked without a GPU"""
    for seed in (34503498, 1, 0, 77):
        s = ctypes.c_uint64(0)
        lib.bic_rand48_seed(ctypes.byref(s), seed)
        r = oracle.rng(seed)
        for n in (1000, 7, 136090, 2 ** 22):
            for _ in range(200):
                assert lib.bic_rand48_uniform_int(ctypes.byref(s), n) == oracle.uniform_int(r, n)


def test_product_does_not_link_or_load_the_oracle():
    out = subprocess.run(["ldd", str(bic.LIB_PATH)], capture_output=True, text=True).stdout
    assert "oracle" not in out and "bic_ref" not in out
    src = (bic.PKG_DIR / "__init__.py").read_text()
    assert "oracle" not in src.replace("no CPU fallback", "")


def test_synth_pack_roundtrip(synth):
    rng = np.random.default_rng(0)
    bits = (rng.random((37, 131)) < 0.3).astype(np.uint8)
    w = synth.pack_rows(bits)
    assert w.shape == (37, 3)
    assert np.array_equal(synth.unpack_rows(w, 131), bits)
    assert (w[:, 2] & np.uint64((1 << 61) - 1)).sum() == 0  # pad bits zero
