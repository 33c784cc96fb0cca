"""bic_pipeline (csrc/pipeline.cu): rasters queued on one host thread come out byte for byte as bic_encode_raster produces
them (and therefore as the oracle says: test_gpu_parity.py pins that path), whatever the batch sizes, the number of slots and
the mix of shapes; the device-side pivot draw equals the host replay; failures are reported per job."""
import importlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
SEED = 34503498


@pytest.fixture(scope="module")
def bic():
    return importlib.import_module("binary-image-compression_b200")


@pytest.fixture(scope="module")
def ctx(bic):
    c = bic.Context(0)
    yield c
    c.close()


def _pages(synth, n, rows, cols, noise_every=4):
    out = []
    for s in range(n):
        if noise_every and s % noise_every == noise_every - 1:
            bits = (np.random.default_rng(s).random((rows, cols)) < 0.5).astype(np.uint8)   # noise: stops after 2 iterations
        else:
            bits = synth.structured_page(rows, cols, seed=100 + s, salt=0.002 + 0.004 * s)
        out.append(bits)
    return out


@pytest.mark.parametrize("first,nxt,slots", [(2, 2, 4), (1, 1, 3), (5, 3, 8), (64, 1, 2), (2, 2, 5)])
def test_pipeline_containers_equal_the_synchronous_encoder(bic, ctx, synth, oracle, first, nxt, slots):
    rows, cols, W, K = 328, 264, 8, 32
    pages = _pages(synth, 10, rows, cols)
    want = []
    for bits in pages:
        cont, info = ctx.encode_raster(synth.pbm_bytes(bits), rows, cols, W, K, seed=SEED)
        want.append((np.array(cont, copy=True), int(info.iterations), (int(info.bits_D), int(info.bits_A), int(info.bits_E))))
    assert len({w[1] for w in want}) > 1          # pages that need different numbers of iterations share the pool
    # one page against the oracle directly, so this test does not lean on the synchronous path alone
    Xo = oracle.extract_patches(synth.pack_rows(pages[0]), rows, cols, W)
    Do, Ao, _ = oracle.init_neighbor(Xo, W * W, K, SEED)
    Eo, ito, _ = oracle.learn_traditional(Xo, Do, Ao, W * W, K)
    assert want[0][1] == ito and want[0][2][2] == oracle.golomb_encode(Eo, W * W)[1]
    pipe = bic.Pipeline(0, slots)
    pipe.set_option("first_batch", first)
    pipe.set_option("next_batch", nxt)
    if slots == 5:      # the coder's other routes inside the pipeline: lists for sparse tiles, tile scans as their own launches
        pipe.set_option("gol_list", 2)
        pipe.set_option("gol_scan", 2)
    try:
        bufs, jobs = [], []
        for bits in pages:
            pay = ctx.pinned(rows * ((cols + 7) // 8))
            pay[:] = synth.pbm_bytes(bits).reshape(-1)
            out = ctx.pinned(2 * pay.size + (1 << 16))
            out[:] = 0xAB
            job, info = pipe.submit(pay, rows, cols, W, K, seed=SEED, out=out)
            bufs.append((pay, out, info))
            jobs.append(job)
        pipe.wait()
        for (pay, out, info), job, (cont, iters, bits3) in zip(bufs, jobs, want):
            done, st, msg = pipe.status(job)
            assert done and st == 0, msg
            assert int(info.iterations) == iters and (int(info.bits_D), int(info.bits_A), int(info.bits_E)) == bits3
            assert int(info.container_bytes) == len(cont)
            assert np.array_equal(out[: len(cont)], cont), job
        assert pipe.stats()["sync_fallbacks"] == 0
    finally:
        pipe.close()


def test_pipeline_mixed_shapes_resident_and_decode(bic, ctx, synth):
    """different rasters and patch sizes through the same slots (workspaces are re-made on a shape change), rasters that are
    already on the device, and a dictionary too large for the chain (the synchronous path inside the slot)"""
    pipe = bic.Pipeline(0, 3)
    try:
        cases = [(200, 168, 8, 12), (256, 256, 16, 32), (200, 168, 8, 12), (130, 128, 24, 6), (260, 200, 16, 256), (97, 45, 5, 4)]
        items = []
        for i, (rows, cols, W, K) in enumerate(cases):
            bits = synth.structured_page(rows, cols, seed=40 + i, salt=0.01)
            pay = synth.pbm_bytes(bits)
            out = ctx.pinned(4 * pay.size + (1 << 18))
            if i % 2:
                R = ctx.matrix(rows, cols)
                R.upload_pbm(pay)
                ctx.sync()
                job, info = pipe.submit_resident(R, W, K, seed=SEED + i, out=out, producer=ctx)
            else:
                hp = ctx.pinned(pay.size)
                hp[:] = pay.reshape(-1)
                job, info = pipe.submit(hp, rows, cols, W, K, seed=SEED + i, out=out)
            items.append((job, info, out, pay, rows, cols, W, K, i))
        pipe.wait()
        for job, info, out, pay, rows, cols, W, K, i in items:
            done, st, msg = pipe.status(job)
            assert done and st == 0, (i, msg)
            cont, info2 = ctx.encode_raster(pay, rows, cols, W, K, seed=SEED + i)
            assert np.array_equal(out[: len(cont)], cont), i
            back, r, c = ctx.decode_raster(out[: int(info.container_bytes)])
            assert np.array_equal(back, pay)
        assert pipe.stats()["sync_fallbacks"] == 1     # 16x16 / 256 atoms: histograms do not fit the chain's shared memory
    finally:
        pipe.close()


def test_pipeline_reports_failures_per_job(bic, ctx, synth):
    pipe = bic.Pipeline(0, 2)
    try:
        rows, cols, W, K = 64, 64, 8, 4
        zero = ctx.pinned(rows * cols // 8)
        zero[:] = 0
        good = ctx.pinned(rows * cols // 8)
        good[:] = synth.pbm_bytes(synth.structured_page(rows, cols, seed=1, salt=0.05)).reshape(-1)
        small = ctx.pinned(64)
        big = ctx.pinned(1 << 16)
        j1, _ = pipe.submit(zero, rows, cols, W, K, out=big)       # all-zero raster: the reference's draw loop never ends
        j2, i2 = pipe.submit(good, rows, cols, W, K, out=small)    # container does not fit
        j3, i3 = pipe.submit(good, rows, cols, W, K, out=ctx.pinned(1 << 16))
        j4, i4 = pipe.submit(good, rows, cols, W, K, out=None)     # sizes only
        pipe.wait()
        assert pipe.status(j1)[1] == 1
        assert pipe.status(j2)[1] == 4 and int(i2.container_bytes) > 64
        assert pipe.status(j3)[1] == 0 and pipe.status(j4)[1] == 0
        assert int(i4.container_bytes) == int(i3.container_bytes) == int(i2.container_bytes)
        with pytest.raises(bic.BicError):
            pipe.result(j1)
        pipe.forget_finished()
        j5, i5 = pipe.submit(good, rows, cols, W, K, out=None)
        assert int(pipe.result(j5).container_bytes) == int(i3.container_bytes)
    finally:
        pipe.close()


@pytest.mark.parametrize("n,p,zero_frac,seed", [(1000, 32, 0.0, 1), (5000, 64, 0.9, 2), (100000, 32, 0.999, 3), (77, 200, 0.5, 4), (33, 5, 0.97, 5)])
def test_device_pivot_draw_equals_the_host_replay(ctx, oracle, synth, n, p, zero_frac, seed):
    """k_draw_pivots (a warp replaying rand48 with jump-ahead, 32 draws per round) against the oracle's serial draw: same
    pivots, same number of draws, same generator state afterwards -- with most rows all zero, so rounds hold many rejections"""
    rng = np.random.default_rng(seed)
    bits = (rng.random((n, 64)) < 0.3).astype(np.uint8)
    bits[rng.random(n) < zero_frac] = 0
    bits[n // 2, 3] = 1
    Xo = synth.pack_rows(bits)
    r = oracle.rng(777 + seed)
    piv_o, nd_o = oracle.draw_pivots(Xo, 64, p, r)
    X = ctx.matrix(n, 64, Xo)
    st = ctx.rand48(777 + seed)
    piv, nd = ctx.draw_pivots(X, p, st)
    assert nd == nd_o and np.array_equal(piv, piv_o)
    # the generators continue identically
    L = ctx.L
    import ctypes as C
    for _ in range(5):
        assert int(L.bic_rand48_uniform_int(C.byref(st), 1000)) == oracle.uniform_int(r, 1000)
    X.destroy()


def test_sharded_pipeline_with_one_rank(bic, ctx, synth, oracle):
    """the sharded mode of the pipeline on a one-rank communicator per slot: the whole asynchronous sharded path (device-side
    draw over the gathered bitmap, the hooked cluster chain, the device-resident Golomb shard base, the shard container) must
    reproduce the oracle's fit and the serial coder's streams; tests/dist_gpu_worker.py runs the same over several GPUs"""
    pipe = bic.Pipeline(0, 3)
    pipe.make_sharded(0, 1, lambda uid: uid)
    try:
        W, K = 8, 16
        jobs = []
        for s_ in range(7):
            rows, cols = 96 + 8 * (s_ % 3), 128
            if s_ % 3 == 2:
                page = (np.random.default_rng(100 + s_).random((rows, cols)) < 0.45).astype(np.uint8)
            else:
                page = synth.structured_page(rows, cols, seed=60 + s_, salt=0.01)
            pay = ctx.pinned(rows * cols // 8)
            pay[:] = synth.pbm_bytes(page).reshape(-1)
            out = ctx.pinned(1 << 18)
            job, info = pipe.submit(pay, rows, cols, W, K, seed=900, out=out)
            jobs.append((job, info, out, page, rows, cols))
        pipe.wait()
        m = W * W
        for job, info, out, page, rows, cols in jobs:
            done, st, msg = pipe.status(job)
            assert done and st == 0, msg
            Xw = oracle.extract_patches(synth.pack_rows(page), rows, cols, W)
            Do, Ao, _ = oracle.init_neighbor(Xw, m, K, 900)
            Eo, ito, _ = oracle.learn_traditional(Xw, Do, Ao, m, K)
            sc = bic.Pipeline.parse_shard_container(out[: int(info.container_bytes)])
            assert sc["iterations"] == ito == int(info.iterations)
            for name, Mo, cols_ in (("D", Do, m), ("A", Ao, K), ("E", Eo, m)):
                so, nbits, ns = oracle.golomb_encode(Mo, cols_)
                stt = sc["streams"][name]
                assert stt["local_bits"] == nbits and stt["local_samples"] == ns, (name, stt["local_bits"], nbits)
                assert np.array_equal(stt["bytes"], so), name
                if name != "D":
                    assert stt["global_bitcount"] == nbits and stt["code_bit_offset"] == 0
    finally:
        pipe.close()
