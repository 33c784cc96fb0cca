"""CPU, world_size 2, gloo: the row-sharded protocol (tests/dist_protocol.py, the numpy restatement of
csrc/dist.cu's message flow) reproduces the oracle's single-process init and dictionary update bit for
bit on every shard."""
import importlib
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, split):
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "tests"))
    import torch
    import torch.distributed as dist
    from dist_protocol import sharded_init, sharded_update_dictionary
    from oracle_bindings import Oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    synth = importlib.import_module("binary-image-compression_b200.synth")
    oracle = Oracle()

    def allreduce(a):
        t = torch.from_numpy(np.ascontiguousarray(a, np.int64).copy())
        dist.all_reduce(t)
        return t.numpy()

    def allgather(obj):
        out = [None] * world
        dist.all_gather_object(out, obj)
        return out
    allgather.rank = rank

    for (rows, cols, W, K, seed) in [(256, 192, 8, 16, 1), (200, 256, 16, 24, 2)]:
        page = synth.structured_page(rows, cols, seed=seed, salt=0.02)
        Xw = oracle.extract_patches(synth.pack_rows(page), rows, cols, W)
        m = W * W
        X = synth.unpack_rows(Xw, m)
        n = X.shape[0]
        cut = int(n * split)
        lo, hi = (0, cut) if rank == 0 else (cut, n)
        # ---- init
        r = oracle.rng(99 + seed)
        D0 = sharded_init(X[lo:hi], K, lambda nn: oracle.uniform_int(r, nn), allgather, allreduce)
        Do, Ao, _ = oracle.init_neighbor(Xw, m, K, 99 + seed)
        assert np.array_equal(D0, synth.unpack_rows(Do, m)), "sharded init differs from the oracle"
        # ---- two iterations: coefficient update is local (rows independent), dictionary update sharded
        Eo = oracle.residual(Xw, Ao, Do, m, K)
        for it in range(2):
            oracle.update_coefficients(Eo, Do, Ao, m, K)
            El = synth.unpack_rows(Eo, m)[lo:hi].copy()
            Al = synth.unpack_rows(Ao, K)[lo:hi].copy()
            Dl = synth.unpack_rows(Do, m)
            newD, changed, ncoll = sharded_update_dictionary(El, Dl, Al, allreduce)
            ca = oracle.update_dictionary(Eo, Do, Ao, m, K)
            assert changed == ca, (it, changed, ca)
            assert ncoll == 1 + ca                       # one allreduce + one per CHANGED atom
            assert np.array_equal(newD, synth.unpack_rows(Do, m))
            assert np.array_equal(El, synth.unpack_rows(Eo, m)[lo:hi])
    dist.destroy_process_group()


@pytest.mark.parametrize("split", [0.5, 0.23])
def test_sharded_protocol_matches_oracle_world2_gloo(split):
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(2, _free_port(), split), nprocs=2, join=True)


def _golomb_worker(rank, world, port, split):
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "tests"))
    import torch.distributed as dist
    from dist_protocol import sharded_golomb
    from oracle_bindings import Oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    synth = importlib.import_module("binary-image-compression_b200.synth")
    oracle = Oracle()

    def allgather(obj):
        out = [None] * world
        dist.all_gather_object(out, obj)
        return out

    cases = [(300, 64, 0.07, None), (257, 100, 0.002, None), (128, 32, 0.5, None), (90, 333, 0.0, None), (200, 64, 0.05, "empty_tail")]
    for rows, cols, rho, special in cases:
        rng = np.random.default_rng(rows + cols)
        bits = (rng.random((rows, cols)) < rho).astype(np.uint8)
        cut = int(rows * split)
        if special == "empty_tail":
            bits[cut:] = 0                              # the last shard has no ones: it only writes the closing run
        whole, nbits, ns = oracle.golomb_encode(synth.pack_rows(bits), cols)
        lo, hi = (0, cut) if rank == 0 else (cut, rows)
        Ml = synth.pack_rows(bits[lo:hi])

        def enc(t0, pos0, prev0, closing, total):
            return oracle.golomb_encode_shard(Ml, cols, t0, pos0, prev0, closing, total)

        by, local_bits, code0, gbits, gns = sharded_golomb(bits[lo:hi], rank, world, allgather, enc)
        assert (gbits, gns) == (nbits, ns), (gbits, gns, nbits, ns)
        # every rank checks the concatenation: its bits sit at [code0, code0 + local_bits) of the single-stream code
        parts = allgather((code0, local_bits, np.unpackbits(by)[:local_bits]))
        cat = np.zeros(nbits, np.uint8)
        for c0, lb, b in parts:
            assert not cat[c0: c0 + lb].any()
            cat[c0: c0 + lb] = b
        assert np.array_equal(cat, np.unpackbits(whole)[:nbits])
    dist.destroy_process_group()


@pytest.mark.parametrize("split", [0.5, 0.31])
def test_sharded_golomb_prefix_state_world2_gloo(split):
    """the two small all-gathers that give a row shard its coder state and code offset (k_g2_shard_base1/2): the ranks' codes
    are exact substrings of the single-stream code, also when a shard has no ones at all"""
    import torch.multiprocessing as mp
    mp.spawn(_golomb_worker, args=(2, _free_port(), split), nprocs=2, join=True)
