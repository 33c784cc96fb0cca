"""The C++ shim that keeps the reference's names (binary-image-compression_b200/host/): compile-time
API coverage on CPU, behaviour on the GPU -- including our bsvd_test driver against the reference's
own bsvd_test binary (oracle/_ref/bsvd_test) on the same PBM."""
import hashlib
import importlib
import subprocess
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
HOST = ROOT / "binary-image-compression_b200" / "host"


def _build():
    bic = importlib.import_module("binary-image-compression_b200")
    if not bic.LIB_PATH.exists():
        bic.build()
    subprocess.run(["make", "-C", str(HOST)], check=True, capture_output=True)


def test_shim_builds_and_keeps_the_reference_names():
    _build()
    out = subprocess.run(["nm", "-DC", "--defined-only", str(ROOT / "binary-image-compression_b200" / "libbic_host.so")],
                         capture_output=True, text=True).stdout
    for sym in ["binary_matrix::weight() const", "binary_matrix::copy_submatrix_to", "binary_matrix::set_vectorized",
                "add(binary_matrix const&, binary_matrix const&, binary_matrix&)", "dist(binary_matrix const&, binary_matrix const&)",
                "mul(binary_matrix const&, bool, binary_matrix const&, bool, binary_matrix&)", "learn_model_setup(int, int, int, int, int)",
                "initialize_model_neighbor", "update_coefficients_omp", "update_dictionary_steepest", "learn_model_traditional",
                "initialize_model", "update_coefficients", "update_dictionary", "learn_model", "random_seed", "set_grid_width",
                "learn_model_mdl_forward_selection", "learn_model_mdl_backward_selection", "learn_model_mdl_full_search",
                "learn_model_alter1", "learn_model_alter2", "learn_model_alter3", "update_dictionary_proximus",
                "model_codelength(binary_matrix const&, binary_matrix const&, binary_matrix const&)", "universal_codelength"]:
        assert sym in out, sym


def _write_pbm(path, bits):
    rows, cols = bits.shape
    with open(path, "wb") as f:
        f.write(f"P4\n{cols} {rows}\n".encode())
        f.write(np.packbits(bits, axis=1, bitorder="big").tobytes())


@pytest.mark.gpu
def test_shim_selftest_binary():
    _build()
    r = subprocess.run([str(HOST / "shim_selftest")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "shim selftest ok" in r.stdout


@pytest.mark.gpu
def test_driver_proximus_matches_reference_driver(tmp_path, synth):
    """-d 1 (update_dictionary_proximus inside the traditional learner): same output files as the reference's driver"""
    ref_bin = ROOT / "oracle" / "_ref" / "bsvd_test"
    if not ref_bin.exists():
        pytest.skip("oracle/_ref/bsvd_test not built")
    _build()
    page = synth.structured_page(200, 168, seed=5, salt=0.01)
    pbm = tmp_path / "in.pbm"
    _write_pbm(pbm, page)
    flags = ["-I", "1", "-k", "12", "-r", "777", "-m", "0", "-M", "0", "-d", "1", "-w", "8"]
    outs = {}
    for name, exe in (("ref", ref_bin), ("b200", HOST / "bsvd_test_b200")):
        d = tmp_path / name
        d.mkdir()
        r = subprocess.run([str(exe)] + flags + [str(pbm)], cwd=d, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout[-1000:] + r.stderr[-1000:]
        outs[name] = {f: hashlib.md5((d / f).read_bytes()).hexdigest() for f in ("dictionary.pbm", "coefficients.pbm", "residual.pbm")}
    assert outs["ref"] == outs["b200"]


@pytest.mark.gpu
@pytest.mark.parametrize("lm,K", [(4, 6), (5, 10)])
def test_driver_mdl_with_proximus_matches_reference_driver(tmp_path, synth, lm, K):
    """-d 1 -l 4/5: the MDL learners' inner fit goes through the global update pointers (src/bsvd.cpp:1229,1235), so the
    reference runs PROXIMUS inside the search; so must we"""
    ref_bin = ROOT / "oracle" / "_ref" / "bsvd_test"
    if not ref_bin.exists():
        pytest.skip("oracle/_ref/bsvd_test not built")
    _build()
    page = synth.structured_page(200, 168, seed=5, salt=0.01)
    pbm = tmp_path / "in.pbm"
    _write_pbm(pbm, page)
    flags = ["-I", "1", "-k", str(K), "-r", "777", "-m", "0", "-M", "0", "-d", "1", "-l", str(lm), "-w", "8"]
    outs = {}
    for name, exe in (("ref", ref_bin), ("b200", HOST / "bsvd_test_b200")):
        d = tmp_path / name
        d.mkdir()
        r = subprocess.run([str(exe)] + flags + [str(pbm)], cwd=d, capture_output=True, text=True, timeout=900)
        assert r.returncode == 0, r.stdout[-1000:] + r.stderr[-1000:]
        outs[name] = {f: hashlib.md5((d / f).read_bytes()).hexdigest() for f in ("dictionary.pbm", "coefficients.pbm", "residual.pbm")}
    assert outs["ref"] == outs["b200"]


@pytest.mark.gpu
@pytest.mark.parametrize("mode,W,K,rows,cols,lm", [(1, 8, 32, 400, 328, 0), (1, 16, 24, 300, 260, 0), (0, 0, 12, 200, 150, 0),
                                                    (1, 8, 6, 200, 168, 4), (1, 8, 10, 200, 168, 5), (1, 8, 45, 160, 128, 6),
                                                    (1, 8, 12, 200, 168, 1), (1, 8, 12, 200, 168, 2), (1, 16, 10, 160, 160, 3)])
def test_driver_matches_reference_driver(tmp_path, synth, mode, W, K, rows, cols, lm):
    """same flags, same PBM -> same dictionary.pbm / coefficients.pbm / residual.pbm bytes (-l 1/2/3: the role-switched learners; -l 4/5/6: the MDL learners,
    which change the number of atoms)"""
    ref_bin = ROOT / "oracle" / "_ref" / "bsvd_test"
    if not ref_bin.exists():
        pytest.skip("oracle/_ref/bsvd_test not built")
    _build()
    page = synth.structured_page(rows, cols, seed=5, salt=0.01)
    pbm = tmp_path / "in.pbm"
    _write_pbm(pbm, page)
    flags = ["-I", str(mode), "-k", str(K), "-r", "777", "-m", "0", "-M", "0", "-l", str(lm)] + (["-w", str(W)] if mode else [])
    outs = {}
    for name, exe in (("ref", ref_bin), ("b200", HOST / "bsvd_test_b200")):
        d = tmp_path / name
        d.mkdir()
        r = subprocess.run([str(exe)] + flags + [str(pbm)], cwd=d, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout[-1000:] + r.stderr[-1000:]
        outs[name] = {f: hashlib.md5((d / f).read_bytes()).hexdigest() for f in ("dictionary.pbm", "coefficients.pbm", "residual.pbm")}
        outs[name]["E"] = [ln for ln in r.stdout.splitlines() if ln.startswith("|E|")][-1]
    assert outs["ref"] == outs["b200"]
