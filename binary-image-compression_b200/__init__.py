"""binary-image-compression_b200: B200 (sm_100a) implementation of the encoder hot path of
nacho-pancho/binary-image-compression -- bsvd fit (patch extraction, neighbour init, greedy
coefficient update, ordered majority-vote dictionary update) + Golomb/EG coding.

The product is the C-ABI shared library `libbic_b200.so` (CUDA kernels, include/bic_b200.h) and
the C++ shim in host/ that keeps the reference's names. This Python module is only a ctypes
binding over that C ABI for tests/, bench.py and __graft_entry__.py. There is no CPU fallback:
loading fails loudly when the library is missing and every call fails when no GPU is present.

The directory name has a hyphen, so import it with
    importlib.import_module("binary-image-compression_b200")      (or `import bic_b200`)
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

from . import synth  # noqa: F401

PKG_DIR = Path(__file__).resolve().parent
# BIC_B200_LIB selects another build of the same library (e.g. libbic_b200_dbg.so: device-side bounds checks compiled in)
LIB_PATH = Path(os.environ["BIC_B200_LIB"]).resolve() if os.environ.get("BIC_B200_LIB") else PKG_DIR / "libbic_b200.so"

BIC_OK = 0
STATUS_NAMES = {0: "ok", 1: "invalid", 2: "cuda", 3: "nomem", 4: "capacity", 5: "no_device", 6: "corrupt", 7: "unsupported"}
CODER_GOLOMB, CODER_EG = 1, 2

_u64 = C.c_uint64
_u64p = C.POINTER(C.c_uint64)
_u8p = C.POINTER(C.c_uint8)
_vp = C.c_void_p


class BicError(RuntimeError):
    def __init__(self, status, msg=""):
        super().__init__(f"libbic_b200: {STATUS_NAMES.get(status, status)} {msg}")
        self.status = status


class StreamInfo(C.Structure):
    _fields_ = [("coder", C.c_uint32), ("chunk_samples", C.c_uint32), ("rows", _u64), ("cols", _u64),
                ("bitcount", _u64), ("nsamples", _u64), ("nchunks", _u64)]


class ShardInfo(C.Structure):
    _fields_ = [(n, _u64) for n in ("global_bitcount", "global_nsamples", "code_bit_offset", "local_code_bits",
                                    "first_chunk", "local_chunks")]


class MatchTotals(C.Structure):
    _fields_ = [("matches", _u64), ("weight_sum", _u64), ("bits_match", _u64), ("bits_nomatch", _u64), ("L", C.c_double)]


MATCH_DTYPE = np.dtype([(n, np.uint64) for n in ("besti", "bestj", "bestd", "weight", "match_len", "nomatch_len", "use_match")])


class EncodeInfo(C.Structure):
    _fields_ = [(n, _u64) for n in ("rows", "cols", "W", "K", "n", "m", "iterations", "weight_E", "weight_A",
                                    "weight_D", "bits_D", "bits_A", "bits_E", "container_bytes")]


def build(verbose: bool = False) -> Path:
    """Compile libbic_b200.so in-tree with nvcc for sm_100a (no GPU needed)."""
    r = subprocess.run(["make", "-C", str(PKG_DIR / "csrc"), "-j8", "all", "debug"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("building libbic_b200.so failed:\n" + r.stdout[-4000:] + r.stderr[-4000:])
    if verbose:
        print(r.stdout[-2000:])
    return LIB_PATH


_lib = None


def lib() -> C.CDLL:
    """Load the C-ABI library. Raises if it has not been built -- no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise ImportError(f"{LIB_PATH} is missing: run __graft_entry__.build() (nvcc, sm_100a). "
                          "There is no CPU fallback for this path.")
    L = C.CDLL(str(LIB_PATH))
    sig = {
        "bic_ctx_create": [C.c_int, C.POINTER(_vp)],
        "bic_ctx_create_on_stream": [C.c_int, _vp, C.POINTER(_vp)],
        "bic_ctx_destroy": [_vp],
        "bic_ctx_sync": [_vp],
        "bic_timer_start": [_vp],
        "bic_timer_stop": [_vp, C.POINTER(C.c_float)],
        "bic_ctx_set_option": [_vp, C.c_char_p, C.c_int64],
        "bic_ctx_wait_ctx": [_vp, _vp],
        "bic_prof_enable": [_vp, C.c_int],
        "bic_prof_reset": [_vp],
        "bic_prof_get": [_vp, C.c_int, C.POINTER(C.c_char_p), _u64p, C.POINTER(C.c_double)],
        "bic_ctx_read_counter": [_vp, C.c_char_p, _u64p],
        "bic_host_alloc": [C.c_size_t, C.POINTER(_vp)],
        "bic_host_free": [_vp],
        "bic_mat_create": [_vp, _u64, _u64, C.POINTER(_vp)],
        "bic_mat_destroy": [_vp, _vp],
        "bic_mat_upload_words64": [_vp, _vp, _u64p],
        "bic_mat_download_words64": [_vp, _vp, _u64p],
        "bic_mat_upload_pbm": [_vp, _vp, _u8p],
        "bic_mat_download_pbm": [_vp, _vp, _u8p],
        "bic_mat_clear": [_vp, _vp],
        "bic_mat_copy": [_vp, _vp, _vp],
        "bic_mat_copy_rows": [_vp, _vp, C.c_uint64, C.c_uint64, _vp, C.c_uint64],
        "bic_mat_weight": [_vp, _vp, _u64p],
        "bic_mat_dist": [_vp, _vp, _vp, _u64p],
        "bic_mat_xor": [_vp, _vp, _vp, _vp],
        "bic_extract_patches": [_vp, _vp, _u64, _vp],
        "bic_assemble_patches": [_vp, _vp, _u64, _vp],
        "bic_draw_pivots": [_vp, _vp, _u64, _u64p, _u64p, _u64p],
        "bic_initialize_model_neighbor_pivots": [_vp, _vp, _u64p, _u64, _vp, _vp],
        "bic_initialize_model_neighbor": [_vp, _vp, _vp, _vp, _u64p],
        "bic_update_coefficients": [_vp, _vp, _vp, _vp, _u64p],
        "bic_update_dictionary_steepest": [_vp, _vp, _vp, _vp, _u64p],
        "bic_residual": [_vp, _vp, _vp, _vp, _vp],
        "bic_learn_model_traditional": [_vp, _vp, _vp, _vp, _vp, _u64p, _u64p, _u64],
        "bic_model_codelength": [_vp, _vp, _vp, _vp, _u64p],
        "bic_mat_transpose": [_vp, _vp, _vp],
        "bic_update_dictionary_proximus": [_vp, _vp, _vp, _vp, _u64p],
        "bic_learn_model_alter": [_vp, C.c_int, _vp, _vp, _vp, _vp, _u64p],
        "bic_split_bitplanes": [_vp, _u8p, C.c_uint64, C.c_uint64, C.c_uint32, C.POINTER(_vp), C.c_uint32],
        "bic_learn_model_mdl": [_vp, C.c_int, _vp, _vp, C.POINTER(_vp), C.POINTER(_vp), _u64p, _u64p],
        "bic_comm_unique_id": [_u8p],
        "bic_comm_create": [_vp, C.c_int, C.c_int, _u8p, C.POINTER(_vp)],
        "bic_comm_destroy": [_vp, _vp],
        "bic_dist_initialize_model_neighbor": [_vp, _vp, _vp, _vp, _vp, _u64p],
        "bic_dist_update_dictionary_steepest": [_vp, _vp, _vp, _vp, _vp, _u64p],
        "bic_dist_learn_model_traditional": [_vp, _vp, _vp, _vp, _vp, _vp, _u64p, _u64p, _u64],
        "bic_learn_model_traditional_batched": [_vp, C.c_uint32, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp), _u64p],
        "bic_dist_golomb_encode": [_vp, _vp, _vp, C.c_uint32, _vp, C.POINTER(ShardInfo)],
        "bic_golomb_encode_shard": [_vp, _vp, C.c_uint32, _u64, _u64, C.c_int64, _u64, C.c_int, _u64, _vp, C.POINTER(ShardInfo)],
        "bic_match_patches_v1": [_vp, _vp, _u64, _vp, C.POINTER(MatchTotals)],
        "bic_match_patches_v4": [_vp, _vp, _u64, _u64, _u64, _vp, C.POINTER(MatchTotals)],
        "bic_pipeline_create": [C.c_int, C.c_int, C.POINTER(_vp)],
        "bic_pipeline_destroy": [_vp],
        "bic_pipeline_set_option": [_vp, C.c_char_p, C.c_int64],
        "bic_pipeline_submit": [_vp, _u8p, _u64, _u64, _u64, _u64, C.c_ulong, _u8p, _u64, C.POINTER(EncodeInfo), _u64p],
        "bic_pipeline_submit_resident": [_vp, _vp, _vp, _u64, _u64, C.c_ulong, _u8p, _u64, C.POINTER(EncodeInfo), _u64p],
        "bic_pipeline_poll": [_vp, _u64p],
        "bic_pipeline_wait": [_vp, _u64],
        "bic_pipeline_job_status": [_vp, _u64, C.POINTER(C.c_int), C.POINTER(C.c_char_p)],
        "bic_pipeline_forget_finished": [_vp],
        "bic_pipeline_stats": [_vp, _u64p, _u64p, _u64p, _u64p, _u64p],
        "bic_pipeline_attach_comms": [_vp, C.POINTER(_vp), C.c_int],
        "bic_merge_shard_containers": [C.POINTER(_u8p), C.POINTER(_u64), C.c_int, _u8p, _u64, C.POINTER(_u64)],
        "bic_pipeline_wait_ctx": [_vp, _vp],
        "bic_ctx_wait_pipeline": [_vp, _vp],
        "bic_stream_create": [_vp, C.POINTER(_vp)],
        "bic_stream_destroy": [_vp, _vp],
        "bic_stream_get_info": [_vp, C.POINTER(StreamInfo)],
        "bic_stream_download": [_vp, _vp, _u8p, _u64, _u64p, _u64],
        "bic_stream_upload": [_vp, _vp, C.POINTER(StreamInfo), _u8p, _u64p],
        "bic_golomb_encode": [_vp, _vp, C.c_uint32, _vp],
        "bic_golomb_bitcount": [_vp, _vp, _u64p, _u64p],
        "bic_golomb_decode": [_vp, _vp, _vp],
        "bic_eg_encode": [_vp, _vp, _vp],
        "bic_eg_decode": [_vp, _vp, _vp],
        "bic_encode_raster": [_vp, _u8p, _u64, _u64, _u64, _u64, C.c_ulong, _u8p, _u64, C.POINTER(EncodeInfo)],
        "bic_encode_raster_resident": [_vp, _vp, _u64, _u64, C.c_ulong, _u8p, _u64, C.POINTER(EncodeInfo)],
        "bic_decode_raster": [_vp, _u8p, _u64, _u8p, _u64, _u64p, _u64p],
    }
    for name, args in sig.items():
        f = getattr(L, name)
        f.argtypes = args
        f.restype = C.c_int
    L.bic_pipeline_slot_ctx.argtypes = [_vp, C.c_int]
    L.bic_pipeline_slot_ctx.restype = _vp
    L.bic_enumL.argtypes = [_u64, _u64]
    L.bic_enumL.restype = C.c_double
    L.bic_prof_kernel_count.argtypes = []
    L.bic_prof_kernel_count.restype = C.c_int
    L.bic_comm_collective_count.argtypes = [_vp]
    L.bic_comm_collective_count.restype = _u64
    L.bic_ctx_last_error.argtypes = [_vp]
    L.bic_ctx_last_error.restype = C.c_char_p
    L.bic_status_string.argtypes = [C.c_int]
    L.bic_status_string.restype = C.c_char_p
    L.bic_ctx_cuda_stream.argtypes = [_vp]
    L.bic_ctx_cuda_stream.restype = _vp
    L.bic_ctx_sm_count.argtypes = [_vp]
    L.bic_ctx_sm_count.restype = C.c_int
    L.bic_ctx_launch_count.argtypes = [_vp]
    L.bic_ctx_launch_count.restype = _u64
    for n in ("bic_mat_rows", "bic_mat_cols", "bic_mat_stride_words32"):
        getattr(L, n).argtypes = [_vp]
        getattr(L, n).restype = _u64
    L.bic_mat_device_ptr.argtypes = [_vp]
    L.bic_mat_device_ptr.restype = _vp
    L.bic_rand48_seed.argtypes = [_u64p, C.c_ulong]
    L.bic_rand48_seed.restype = None
    L.bic_rand48_uniform_int.argtypes = [_u64p, _u64]
    L.bic_rand48_uniform_int.restype = _u64
    _lib = L
    return L


def exported_symbols_declared_in_header():
    """names declared in include/bic_b200.h (used by the CPU test that checks the .so exports them)"""
    import re
    text = (PKG_DIR.parent / "include" / "bic_b200.h").read_text()
    return sorted(set(re.findall(r"\b(bic_[a-z0-9_]+)\s*\(", text)))


def _wpr64(cols):
    return (cols + 63) // 64


class Matrix:
    """Device bit matrix handle (stands in for binary_matrix, src/binmat.h:29)."""

    def __init__(self, ctx: "Context", rows: int, cols: int):
        self.ctx, self.rows, self.cols = ctx, int(rows), int(cols)
        h = _vp()
        ctx._ck(ctx.L.bic_mat_create(ctx.h, self.rows, self.cols, C.byref(h)))
        self.h = h

    def upload(self, words64: np.ndarray) -> "Matrix":
        w = np.ascontiguousarray(words64, np.uint64)
        assert w.size == self.rows * _wpr64(self.cols), (w.shape, self.rows, self.cols)
        self.ctx._ck(self.ctx.L.bic_mat_upload_words64(self.ctx.h, self.h, w.ctypes.data_as(_u64p)))
        self.ctx.sync()  # w may be a temporary
        return self

    def download(self) -> np.ndarray:
        out = np.zeros((self.rows, _wpr64(self.cols)), np.uint64)
        self.ctx._ck(self.ctx.L.bic_mat_download_words64(self.ctx.h, self.h, out.ctypes.data_as(_u64p)))
        return out

    def upload_pbm(self, payload: np.ndarray) -> "Matrix":
        b = np.ascontiguousarray(payload, np.uint8)
        assert b.size == self.rows * ((self.cols + 7) // 8)
        self.ctx._ck(self.ctx.L.bic_mat_upload_pbm(self.ctx.h, self.h, b.ctypes.data_as(_u8p)))
        self.ctx.sync()
        return self

    def download_pbm(self) -> np.ndarray:
        out = np.zeros((self.rows, (self.cols + 7) // 8), np.uint8)
        self.ctx._ck(self.ctx.L.bic_mat_download_pbm(self.ctx.h, self.h, out.ctypes.data_as(_u8p)))
        return out

    def clear(self):
        self.ctx._ck(self.ctx.L.bic_mat_clear(self.ctx.h, self.h))

    def copy_from(self, other: "Matrix"):
        self.ctx._ck(self.ctx.L.bic_mat_copy(self.ctx.h, other.h, self.h))

    def copy_rows_from(self, other: "Matrix", src_row0: int, nrows: int, dst_row0: int):
        self.ctx._ck(self.ctx.L.bic_mat_copy_rows(self.ctx.h, other.h, src_row0, nrows, self.h, dst_row0))

    def weight(self) -> int:
        w = _u64(0)
        self.ctx._ck(self.ctx.L.bic_mat_weight(self.ctx.h, self.h, C.byref(w)))
        return int(w.value)

    def destroy(self):
        if self.h:
            self.ctx.L.bic_mat_destroy(self.ctx.h, self.h)
            self.h = None


class Stream:
    def __init__(self, ctx: "Context"):
        self.ctx = ctx
        h = _vp()
        ctx._ck(ctx.L.bic_stream_create(ctx.h, C.byref(h)))
        self.h = h

    @property
    def info(self) -> StreamInfo:
        si = StreamInfo()
        self.ctx._ck(self.ctx.L.bic_stream_get_info(self.h, C.byref(si)))
        return si

    def download(self):
        si = self.info
        nb = (si.bitcount + 7) // 8
        by = np.zeros(max(nb, 1), np.uint8)
        idx = np.zeros(max(2 * si.nchunks, 2), np.uint64)
        self.ctx._ck(self.ctx.L.bic_stream_download(self.ctx.h, self.h, by.ctypes.data_as(_u8p), by.size,
                                                    idx.ctypes.data_as(_u64p), idx.size // 2))
        return by[:nb], idx[: 2 * si.nchunks].reshape(-1, 2)

    def upload(self, info: StreamInfo, by: np.ndarray, idx: np.ndarray):
        by = np.ascontiguousarray(by, np.uint8)
        idx = np.ascontiguousarray(idx, np.uint64)
        self.ctx._ck(self.ctx.L.bic_stream_upload(self.ctx.h, self.h, C.byref(info), by.ctypes.data_as(_u8p),
                                                  idx.ctypes.data_as(_u64p)))
        self.ctx.sync()

    def destroy(self):
        if self.h:
            self.ctx.L.bic_stream_destroy(self.ctx.h, self.h)
            self.h = None


class Pipeline:
    """bic_pipeline: a pool of encoder slots driven by the calling thread; rasters are queued and polled, nothing blocks on the
    device (see include/bic_b200.h)."""

    def __init__(self, device: int = 0, nslots: int = 16):
        self.L = lib()
        h = _vp()
        st = self.L.bic_pipeline_create(device, nslots, C.byref(h))
        if st != 0:
            raise BicError(st, "bic_pipeline_create")
        self.h = h
        self.nslots = nslots
        self._keep = {}

    def _ck(self, st, what=""):
        if st != 0:
            raise BicError(st, what)

    def set_option(self, name: str, value: int):
        self._ck(self.L.bic_pipeline_set_option(self.h, name.encode(), int(value)), name)

    def submit(self, payload: np.ndarray, rows: int, cols: int, W: int, K: int, seed: int = 34503498, out: np.ndarray | None = None):
        """queue a host P4 payload; returns (job id, EncodeInfo that is filled when the job is done)"""
        info = EncodeInfo()
        job = _u64(0)
        outp = out.ctypes.data_as(_u8p) if out is not None else None
        self._ck(self.L.bic_pipeline_submit(self.h, payload.ctypes.data_as(_u8p), rows, cols, W, K, seed, outp,
                                            out.size if out is not None else 0, C.byref(info), C.byref(job)))
        self._keep[int(job.value)] = (payload, out, info)
        return int(job.value), info

    def submit_resident(self, raster: "Matrix", W: int, K: int, seed: int = 34503498, out: np.ndarray | None = None, producer: "Context | None" = None):
        info = EncodeInfo()
        job = _u64(0)
        outp = out.ctypes.data_as(_u8p) if out is not None else None
        self._ck(self.L.bic_pipeline_submit_resident(self.h, raster.h, producer.h if producer is not None else None, W, K, seed, outp,
                                                     out.size if out is not None else 0, C.byref(info), C.byref(job)))
        self._keep[int(job.value)] = (raster, out, info)
        return int(job.value), info

    def poll(self) -> int:
        left = _u64(0)
        self._ck(self.L.bic_pipeline_poll(self.h, C.byref(left)))
        return int(left.value)

    def wait(self, job: int = 0):
        self._ck(self.L.bic_pipeline_wait(self.h, job))

    def status(self, job: int):
        """(done, status code, message)"""
        done = C.c_int(0)
        msg = C.c_char_p()
        st = self.L.bic_pipeline_job_status(self.h, job, C.byref(done), C.byref(msg))
        return bool(done.value), st, (msg.value or b"").decode()

    def result(self, job: int):
        """wait for the job; raises on failure; returns its EncodeInfo"""
        self.wait(job)
        done, st, msg = self.status(job)
        if st != 0:
            raise BicError(st, msg)
        return self._keep[job][2]

    def forget_finished(self):
        self._ck(self.L.bic_pipeline_forget_finished(self.h))
        self._keep = {j: v for j, v in self._keep.items() if not self.status_safe(j)}

    def status_safe(self, job):
        try:
            return self.status(job)[0]
        except BicError:
            return True

    def stats(self) -> dict:
        v = [_u64(0) for _ in range(5)]
        self._ck(self.L.bic_pipeline_stats(self.h, *[C.byref(x) for x in v]))
        return dict(zip(("launches", "polls", "batches", "sync_fallbacks", "recodes"), (int(x.value) for x in v)))

    def wait_for(self, ctx: "Context"):
        self._ck(self.L.bic_pipeline_wait_ctx(self.h, ctx.h))

    # ---- sharded mode (several GPUs): one communicator per slot, the same slot index on every rank
    def make_sharded(self, rank: int, world: int, share_id):
        """share_id(id_or_None) -> id: the caller's plumbing that takes the 128-byte id generated on rank 0 (None elsewhere) to every
        rank (e.g. a torch.distributed broadcast). Creates one communicator per slot and switches the pipeline to sharded jobs."""
        Context._preload_nccl()
        self._comms = []
        for i in range(self.nslots):
            uid = None
            if rank == 0:
                buf = np.zeros(128, np.uint8)
                self._ck(self.L.bic_comm_unique_id(buf.ctypes.data_as(_u8p)), "comm_unique_id")
                uid = buf
            uid = np.ascontiguousarray(share_id(uid), np.uint8)
            comm = _vp()
            self._ck(self.L.bic_comm_create(self.L.bic_pipeline_slot_ctx(self.h, i), rank, world, uid.ctypes.data_as(_u8p), C.byref(comm)),
                     "comm_create")
            self._comms.append(comm)
        arr = (_vp * self.nslots)(*self._comms)
        self._ck(self.L.bic_pipeline_attach_comms(self.h, arr, self.nslots), "attach_comms")
        self.rank, self.world = rank, world

    @staticmethod
    def parse_shard_container(buf: np.ndarray) -> dict:
        """the fields and streams of a sharded job's output (layout: csrc/pipeline.cu)"""
        h = buf[:48 * 8].view(np.uint64)
        assert int(h[0]) == 0x0044524853434942, "not a shard container"
        out = {"rank": int(h[2]), "ranks": int(h[3]), "rows": int(h[4]), "cols": int(h[5]), "W": int(h[6]), "K": int(h[7]), "n": int(h[8]),
               "m": int(h[9]), "iterations": int(h[10]), "seed": int(h[11]), "streams": {}}
        off = 48 * 8
        for i, name in enumerate(("D", "A", "E")):
            f = [int(x) for x in h[12 + 7 * i: 19 + 7 * i]]
            nb, ni = (f[4] + 7) // 8, f[6]
            nbp = (nb + 7) & ~7
            st = {"chunk_samples": f[1], "rows": f[2], "cols": f[3], "local_bits": f[4], "local_samples": f[5], "local_chunks": f[6],
                  "bytes": buf[off: off + nb], "index": buf[off + nbp: off + nbp + ni * 16].view(np.uint64).reshape(-1, 2)}
            if i:
                sh = [int(x) for x in h[33 + 6 * (i - 1): 39 + 6 * (i - 1)]]
                st.update(dict(zip(("global_bitcount", "global_nsamples", "code_bit_offset", "local_code_bits", "first_chunk", "local_chunks_sh"), sh)))
            out["streams"][name] = st
            off += nbp + ni * 16
        out["bytes"] = off
        return out

    @staticmethod
    def merge_shard_containers(shards) -> np.ndarray:
        """the N shard containers of one sharded job (one per rank, any order) -> the ordinary container of the whole raster,
        byte for byte what encode_raster gives for the concatenated bands (bic_merge_shard_containers: host code, no device)"""
        L = lib()
        bufs = []
        for b in shards:
            a = np.zeros((len(b) + 7) // 8 * 8, np.uint8)            # 8-byte aligned copies (numpy allocations are)
            a[: len(b)] = np.frombuffer(bytes(b), np.uint8) if not isinstance(b, np.ndarray) else b
            bufs.append(a)
        n = len(bufs)
        ptrs = (_u8p * n)(*[a.ctypes.data_as(_u8p) for a in bufs])
        sizes = (_u64 * n)(*[len(b) for b in shards])
        need = _u64(0)
        st = L.bic_merge_shard_containers(ptrs, sizes, n, None, 0, C.byref(need))
        if st != BIC_OK:
            raise BicError(st, L.bic_status_string(st).decode())
        out = np.zeros(int(need.value), np.uint8)
        st = L.bic_merge_shard_containers(ptrs, sizes, n, out.ctypes.data_as(_u8p), out.size, C.byref(need))
        if st != BIC_OK:
            raise BicError(st, L.bic_status_string(st).decode())
        return out

    def close(self):
        if self.h:
            for i, comm in enumerate(getattr(self, "_comms", [])):
                self.L.bic_comm_destroy(self.L.bic_pipeline_slot_ctx(self.h, i), comm)
            self._comms = []
            self.L.bic_pipeline_destroy(self.h)
            self.h = None


class Context:
    """One device context (one CUDA stream). Method names follow the reference's plug points."""

    def __init__(self, device: int = 0, cuda_stream: int | None = None):
        self.L = lib()
        h = _vp()
        if cuda_stream is None:
            st = self.L.bic_ctx_create(device, C.byref(h))
        else:
            st = self.L.bic_ctx_create_on_stream(device, _vp(cuda_stream), C.byref(h))
        if st != BIC_OK:
            raise BicError(st, self.L.bic_status_string(st).decode())
        self.h = h

    def _ck(self, st):
        if st != BIC_OK:
            raise BicError(st, self.L.bic_ctx_last_error(self.h).decode())

    def close(self):
        if self.h:
            self.L.bic_ctx_destroy(self.h)
            self.h = None

    def sync(self):
        self._ck(self.L.bic_ctx_sync(self.h))

    @property
    def launches(self) -> int:
        return int(self.L.bic_ctx_launch_count(self.h))

    @property
    def sm_count(self) -> int:
        return int(self.L.bic_ctx_sm_count(self.h))

    def timer_start(self):
        self._ck(self.L.bic_timer_start(self.h))

    def timer_stop(self) -> float:
        ms = C.c_float(0)
        self._ck(self.L.bic_timer_stop(self.h, C.byref(ms)))
        return float(ms.value)

    def wait_for_pipeline(self, pipe: "Pipeline"):
        self._ck(self.L.bic_ctx_wait_pipeline(self.h, pipe.h))

    def wait_for(self, other: "Context"):
        """order this context's stream after everything queued so far on `other`'s stream"""
        self._ck(self.L.bic_ctx_wait_ctx(self.h, other.h))

    def set_option(self, name: str, value: int):
        self._ck(self.L.bic_ctx_set_option(self.h, name.encode(), int(value)))

    def read_counter(self, name: str) -> int:
        v = _u64(0)
        self._ck(self.L.bic_ctx_read_counter(self.h, name.encode(), C.byref(v)))
        return int(v.value)

    def prof_enable(self, on: bool):
        self._ck(self.L.bic_prof_enable(self.h, 1 if on else 0))

    def prof_reset(self):
        self._ck(self.L.bic_prof_reset(self.h))

    def prof_stats(self) -> dict:
        """{kernel name: (launches, total device ms)} for kernels launched while profiling was on"""
        out = {}
        for k in range(self.L.bic_prof_kernel_count()):
            name, n, ms = C.c_char_p(), _u64(0), C.c_double(0)
            self._ck(self.L.bic_prof_get(self.h, k, C.byref(name), C.byref(n), C.byref(ms)))
            if n.value:
                key = name.value.decode()
                pn, pms = out.get(key, (0, 0.0))
                out[key] = (pn + int(n.value), pms + float(ms.value))
        return out

    def pinned(self, nbytes: int) -> np.ndarray:
        """uint8 array backed by page-locked host memory (kept alive by the context)"""
        p = _vp()
        st = self.L.bic_host_alloc(nbytes, C.byref(p))
        if st != BIC_OK:
            raise BicError(st, "pinned allocation")
        buf = (C.c_uint8 * max(nbytes, 1)).from_address(p.value)
        arr = np.frombuffer(buf, dtype=np.uint8, count=nbytes)
        self._pinned = getattr(self, "_pinned", [])
        self._pinned.append((p, buf))
        return arr

    # ---- matrices
    def matrix(self, rows, cols, words64=None) -> Matrix:
        m = Matrix(self, rows, cols)
        if words64 is not None:
            m.upload(words64)
        return m

    def stream(self) -> Stream:
        return Stream(self)

    # ---- bsvd path (names as in src/bsvd.h / src/bsvd_test.cpp)
    def extract_patches(self, raster: Matrix, W: int, out: "Matrix | None" = None) -> Matrix:
        n = ((raster.rows + W - 1) // W) * ((raster.cols + W - 1) // W)
        X = out if out is not None else Matrix(self, n, W * W)
        self._ck(self.L.bic_extract_patches(self.h, raster.h, W, X.h))
        return X

    def assemble_patches(self, X: Matrix, W: int, rows: int, cols: int) -> Matrix:
        R = Matrix(self, rows, cols)
        self._ck(self.L.bic_assemble_patches(self.h, X.h, W, R.h))
        return R

    def rand48(self, seed: int):
        s = _u64(0)
        self.L.bic_rand48_seed(C.byref(s), seed)
        return s

    def draw_pivots(self, X: Matrix, p: int, rng_state) -> tuple[np.ndarray, int]:
        piv = np.zeros(max(p, 1), np.uint64)
        nd = _u64(0)
        self._ck(self.L.bic_draw_pivots(self.h, X.h, p, C.byref(rng_state), piv.ctypes.data_as(_u64p), C.byref(nd)))
        return piv[:p], int(nd.value)

    def initialize_model_neighbor_pivots(self, X: Matrix, pivots, D: Matrix, A: Matrix):
        piv = np.ascontiguousarray(pivots, np.uint64)
        self._ck(self.L.bic_initialize_model_neighbor_pivots(self.h, X.h, piv.ctypes.data_as(_u64p), piv.size, D.h, A.h))
        self.sync()

    def initialize_model_neighbor(self, X: Matrix, D: Matrix, A: Matrix, rng_state):
        self._ck(self.L.bic_initialize_model_neighbor(self.h, X.h, D.h, A.h, C.byref(rng_state)))

    def update_coefficients(self, E: Matrix, D: Matrix, A: Matrix) -> int:
        ch = _u64(0)
        self._ck(self.L.bic_update_coefficients(self.h, E.h, D.h, A.h, C.byref(ch)))
        return int(ch.value)

    def update_dictionary(self, E: Matrix, D: Matrix, A: Matrix) -> int:
        ch = _u64(0)
        self._ck(self.L.bic_update_dictionary_steepest(self.h, E.h, D.h, A.h, C.byref(ch)))
        return int(ch.value)

    def residual(self, X: Matrix, A: Matrix, D: Matrix, E: Matrix):
        self._ck(self.L.bic_residual(self.h, X.h, A.h, D.h, E.h))

    def learn_model_traditional(self, X: Matrix, E: Matrix, D: Matrix, A: Matrix, trace_cap: int = 256):
        it = _u64(0)
        tr = np.zeros(2 * trace_cap, np.uint64)
        self._ck(self.L.bic_learn_model_traditional(self.h, X.h, E.h, D.h, A.h, C.byref(it),
                                                    tr.ctypes.data_as(_u64p), trace_cap))
        n = int(it.value)
        return n, tr[: 2 * min(n, trace_cap)].reshape(-1, 2)

    def update_dictionary_proximus(self, E: Matrix, D: Matrix, A: Matrix) -> int:
        ch = _u64(0)
        self._ck(self.L.bic_update_dictionary_proximus(self.h, E.h, D.h, A.h, C.byref(ch)))
        return int(ch.value)

    # ---- role-switched learners (src/bsvd.cpp:1245-1434)
    def transpose(self, M: Matrix) -> Matrix:
        T = Matrix(self, M.cols, M.rows)
        self._ck(self.L.bic_mat_transpose(self.h, M.h, T.h))
        return T

    def learn_model_alter(self, variant: int, X: Matrix, E: Matrix, D: Matrix, A: Matrix) -> int:
        it = _u64(0)
        self._ck(self.L.bic_learn_model_alter(self.h, variant, X.h, E.h, D.h, A.h, C.byref(it)))
        return int(it.value)

    # ---- bit planes of a grey image (src/bitplane_tool.cpp:24-39)
    def split_bitplanes(self, p5_payload: np.ndarray, rows: int, cols: int, maxval: int, planes=None, sync: bool = True):
        """P5 payload bytes (1 byte per pixel if maxval < 256 else 2, high byte first) -> list of rows x cols planes, LSB first"""
        n = 0
        b = 1
        while b < maxval:
            n += 1
            b <<= 1
        planes = planes or [Matrix(self, rows, cols) for _ in range(n)]
        pay = np.ascontiguousarray(p5_payload, np.uint8)
        assert pay.size == rows * cols * (2 if maxval >= 256 else 1)
        arr = (_vp * n)(*[m.h for m in planes])
        self._ck(self.L.bic_split_bitplanes(self.h, pay.ctypes.data_as(_u8p), rows, cols, maxval, arr, n))
        if sync:
            self.sync()  # pay may be a temporary
        return planes

    # ---- MDL model selection (src/bsvd.cpp:1438-1717)
    def model_codelength(self, E: Matrix, D: "Matrix | None", A: "Matrix | None") -> int:
        L = _u64(0)
        self._ck(self.L.bic_model_codelength(self.h, E.h, D.h if D is not None else None, A.h if A is not None else None, C.byref(L)))
        return int(L.value)

    def learn_model_mdl(self, lm: int, X: Matrix, E: Matrix, D: Matrix, A: Matrix, rng_state=None):
        """lm = 4 forward selection, 5 backward selection, 6 full search. D and A are consumed: their handles are
        replaced by the selected model's matrices (None, None for the empty model). Returns (bestL, D, A)."""
        L = _u64(0)
        dh, ah = _vp(D.h.value), _vp(A.h.value)
        D.h = A.h = None  # ownership moves into the call
        self._ck(self.L.bic_learn_model_mdl(self.h, lm, X.h, E.h, C.byref(dh), C.byref(ah),
                                            C.byref(rng_state) if rng_state is not None else None, C.byref(L)))
        if not dh.value:
            return int(L.value), None, None
        Dn, An = Matrix.__new__(Matrix), Matrix.__new__(Matrix)
        for M, h in ((Dn, dh), (An, ah)):
            M.ctx, M.h = self, h
            M.rows, M.cols = int(self.L.bic_mat_rows(h)), int(self.L.bic_mat_cols(h))
        return int(L.value), Dn, An

    def learn_model_traditional_batched(self, Xs, Es, Ds, As) -> list[int]:
        """independent fits of identical shape, every kernel launched once for the whole batch"""
        n = len(Xs)
        arr = lambda ms: (_vp * n)(*[m.h for m in ms])  # noqa: E731
        its = np.zeros(n, np.uint64)
        self._ck(self.L.bic_learn_model_traditional_batched(self.h, n, arr(Xs), arr(Es), arr(Ds), arr(As),
                                                            its.ctypes.data_as(_u64p)))
        return [int(v) for v in its]

    # ---- several GPUs (rows sharded, D replicated)
    @staticmethod
    def _preload_nccl():
        """Map the NCCL build torch ships (if any) before the library looks for one: a process must
        not end up with two different libnccl.so.2."""
        import importlib.util
        import os
        if os.environ.get("BIC_NCCL_LIB"):
            return
        try:
            spec = importlib.util.find_spec("nvidia.nccl")
            for loc in (spec.submodule_search_locations or []) if spec else []:
                cand = os.path.join(loc, "lib", "libnccl.so.2")
                if os.path.exists(cand):
                    C.CDLL(cand, mode=C.RTLD_GLOBAL)
                    os.environ["BIC_NCCL_LIB"] = cand
                    return
        except Exception:
            pass

    def comm_unique_id(self) -> np.ndarray:
        self._preload_nccl()
        uid = np.zeros(128, np.uint8)
        self._ck(self.L.bic_comm_unique_id(uid.ctypes.data_as(_u8p)))
        return uid

    def comm_create(self, rank: int, nranks: int, uid: np.ndarray):
        self._preload_nccl()
        h = _vp()
        u = np.ascontiguousarray(uid, np.uint8)
        self._ck(self.L.bic_comm_create(self.h, rank, nranks, u.ctypes.data_as(_u8p), C.byref(h)))
        return h

    def comm_destroy(self, comm):
        self._ck(self.L.bic_comm_destroy(self.h, comm))

    def comm_collectives(self, comm) -> int:
        return int(self.L.bic_comm_collective_count(comm))

    def dist_initialize_model_neighbor(self, comm, X: Matrix, D: Matrix, A: Matrix, rng_state):
        self._ck(self.L.bic_dist_initialize_model_neighbor(self.h, comm, X.h, D.h, A.h, C.byref(rng_state)))

    def dist_update_dictionary(self, comm, E: Matrix, D: Matrix, A: Matrix) -> int:
        ch = _u64(0)
        self._ck(self.L.bic_dist_update_dictionary_steepest(self.h, comm, E.h, D.h, A.h, C.byref(ch)))
        return int(ch.value)

    def dist_learn_model_traditional(self, comm, X: Matrix, E: Matrix, D: Matrix, A: Matrix, trace_cap: int = 256):
        it = _u64(0)
        tr = np.zeros(2 * trace_cap, np.uint64)
        self._ck(self.L.bic_dist_learn_model_traditional(self.h, comm, X.h, E.h, D.h, A.h, C.byref(it),
                                                         tr.ctypes.data_as(_u64p), trace_cap))
        n = int(it.value)
        return n, tr[: 2 * min(n, trace_cap)].reshape(-1, 2)

    def dist_golomb_encode(self, comm, M: Matrix, out: "Stream | None" = None, chunk_samples: int = 256):
        """this rank's rows coded as a substring of the one global stream; returns (Stream, ShardInfo)"""
        out = out or Stream(self)
        si = ShardInfo()
        self._ck(self.L.bic_dist_golomb_encode(self.h, comm, M.h, chunk_samples, out.h, C.byref(si)))
        return out, si

    def golomb_encode_shard(self, M: Matrix, ones_before: int, bits_before: int, last_one_before: int, code_bits_before: int,
                            closing: bool, total_bits: int, out: "Stream | None" = None, chunk_samples: int = 256, lengths_only=False):
        """M's rows coded as a substring of the one global stream, the prefix state given explicitly; returns (Stream, ShardInfo)"""
        if not lengths_only:
            out = out or Stream(self)
        si = ShardInfo()
        self._ck(self.L.bic_golomb_encode_shard(self.h, M.h, chunk_samples, ones_before, bits_before, last_one_before, code_bits_before,
                                                int(closing), total_bits, None if lengths_only else out.h, C.byref(si)))
        return out, si

    # ---- template matching (compress*_test)
    def match_patches(self, raster: Matrix, W: int, version: int = 1, T: int = 0, R: int = 10000):
        """version 1: compress_test.cpp's search over the unmodified image; 4: compress4_test.cpp's windowed search with the
        in-place residual replacement (raster is rewritten). Returns (records structured array, MatchTotals)."""
        n = ((raster.rows + W - 1) // W) * ((raster.cols + W - 1) // W)
        recs = np.zeros(n, MATCH_DTYPE)
        tot = MatchTotals()
        if version == 1:
            self._ck(self.L.bic_match_patches_v1(self.h, raster.h, W, recs.ctypes.data_as(_vp), C.byref(tot)))
        else:
            self._ck(self.L.bic_match_patches_v4(self.h, raster.h, W, T, R, recs.ctypes.data_as(_vp), C.byref(tot)))
        return recs, tot

    # ---- coding
    def golomb_encode(self, M: Matrix, out: Stream | None = None, chunk_samples: int = 256) -> Stream:
        out = out or Stream(self)
        self._ck(self.L.bic_golomb_encode(self.h, M.h, chunk_samples, out.h))
        return out

    def golomb_bitcount(self, M: Matrix):
        b, s = _u64(0), _u64(0)
        self._ck(self.L.bic_golomb_bitcount(self.h, M.h, C.byref(b), C.byref(s)))
        return int(b.value), int(s.value)

    def golomb_decode(self, s: Stream, M: Matrix):
        self._ck(self.L.bic_golomb_decode(self.h, s.h, M.h))

    def eg_encode(self, M: Matrix, out: Stream | None = None) -> Stream:
        out = out or Stream(self)
        self._ck(self.L.bic_eg_encode(self.h, M.h, out.h))
        return out

    def eg_decode(self, s: Stream, M: Matrix):
        self._ck(self.L.bic_eg_decode(self.h, s.h, M.h))

    # ---- whole encoder
    def encode_raster(self, payload: np.ndarray, rows: int, cols: int, W: int, K: int, seed: int = 34503498,
                      out: np.ndarray | None = None):
        """payload: P4 rows (uint8). Returns (container bytes ndarray, EncodeInfo)."""
        b = np.ascontiguousarray(payload, np.uint8)
        info = EncodeInfo()
        if out is None:
            out = np.zeros(max(64 * 1024, 2 * b.size + 64 * 1024), np.uint8)
        st = self.L.bic_encode_raster(self.h, b.ctypes.data_as(_u8p), rows, cols, W, K, seed,
                                      out.ctypes.data_as(_u8p), out.size, C.byref(info))
        if st == 4:  # capacity: retry once with the size the library reported
            out = np.zeros(int(info.container_bytes) + 64, np.uint8)
            st = self.L.bic_encode_raster(self.h, b.ctypes.data_as(_u8p), rows, cols, W, K, seed,
                                          out.ctypes.data_as(_u8p), out.size, C.byref(info))
        self._ck(st)
        return out[: int(info.container_bytes)], info

    def encode_raster_resident(self, raster: Matrix, W: int, K: int, seed: int = 34503498, out: np.ndarray | None = None):
        """the same container from a raster already in device memory (e.g. a plane of split_bitplanes)"""
        info = EncodeInfo()
        if out is None:
            out = np.zeros(max(64 * 1024, 2 * raster.rows * ((raster.cols + 7) // 8) + 64 * 1024), np.uint8)
        st = self.L.bic_encode_raster_resident(self.h, raster.h, W, K, seed, out.ctypes.data_as(_u8p), out.size, C.byref(info))
        if st == 4:  # capacity: retry once with the size the library reported, as encode_raster does
            out = np.zeros(int(info.container_bytes) + 64, np.uint8)
            st = self.L.bic_encode_raster_resident(self.h, raster.h, W, K, seed, out.ctypes.data_as(_u8p), out.size, C.byref(info))
        self._ck(st)
        return out[: int(info.container_bytes)], info

    def decode_raster(self, container: np.ndarray):
        cbuf = np.ascontiguousarray(container, np.uint8)
        r, c_ = _u64(0), _u64(0)
        self._ck(self.L.bic_decode_raster(self.h, cbuf.ctypes.data_as(_u8p), cbuf.size, None, 0, C.byref(r), C.byref(c_)))
        rows, cols = int(r.value), int(c_.value)
        out = np.zeros((rows, (cols + 7) // 8), np.uint8)
        self._ck(self.L.bic_decode_raster(self.h, cbuf.ctypes.data_as(_u8p), cbuf.size, out.ctypes.data_as(_u8p),
                                          out.size, C.byref(r), C.byref(c_)))
        return out, rows, cols
