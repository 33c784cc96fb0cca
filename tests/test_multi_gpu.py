"""Row-sharded fit over several GPUs through the C ABI (csrc/dist.cu, NCCL inside the library)."""
import importlib
import socket
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def test_sharded_path_with_one_rank(oracle, synth):
    """the sharded driver with a 1-rank communicator must equal the plain path (and the oracle)"""
    bic = importlib.import_module("binary-image-compression_b200")
    ctx = bic.Context(0)
    comm = ctx.comm_create(0, 1, ctx.comm_unique_id())
    rows, cols, W, K = 400, 320, 8, 32
    page = synth.structured_page(rows, cols, seed=9, salt=0.01)
    Xw = oracle.extract_patches(synth.pack_rows(page), rows, cols, W)
    m, n = W * W, Xw.shape[0]
    X = ctx.matrix(n, m, Xw)
    D, A, E = ctx.matrix(K, m), ctx.matrix(n, K), ctx.matrix(n, m)
    ctx.dist_initialize_model_neighbor(comm, X, D, A, ctx.rand48(42))
    Do, Ao, _ = oracle.init_neighbor(Xw, m, K, 42)
    assert np.array_equal(D.download(), Do)
    it, tr = ctx.dist_learn_model_traditional(comm, X, E, D, A)
    Eo, ito, tro = oracle.learn_traditional(Xw, Do, Ao, m, K)
    assert it == ito and np.array_equal(tr, tro)
    assert np.array_equal(D.download(), Do) and np.array_equal(A.download(), Ao) and np.array_equal(E.download(), Eo)
    # sharded coder with one shard == the plain coder
    s_plain = ctx.golomb_encode(E)
    s_shard, si = ctx.dist_golomb_encode(comm, E)
    assert si.code_bit_offset == 0 and si.global_bitcount == s_plain.info.bitcount == s_shard.info.bitcount
    assert si.global_nsamples == s_plain.info.nsamples
    b0, i0 = s_plain.download()
    b1, i1 = s_shard.download()
    assert np.array_equal(b0, b1) and np.array_equal(i0, i1)
    ctx.comm_destroy(comm)
    ctx.close()


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_fit_matches_oracle(world):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                        "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
                        str(ROOT / "tests" / "dist_gpu_worker.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert f"dist ok world={world}" in r.stdout
