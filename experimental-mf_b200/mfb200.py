"""ctypes binding of the C ABI in include/mf_b200.h (libmf_b200.so, hand-written sm_100a CUDA).

Python here is plumbing for tests and bench.py only; the host side of the product is the C++
mirror of the reference's MF / DPMF / AdaptRegMF classes (csrc/model.h, the `mf` binary).
There is no CPU fallback: if the library is missing or no CUDA device is usable, calls raise.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libmf_b200.so")

THETA, PHI, BU, BV, THETA_OLD, PHI_OLD, BU_OLD, BV_OLD, UR, VR, LAMBDA_U, LAMBDA_V = range(12)
MODE_HOGWILD, MODE_ORDERED, MODE_ATOMIC = 0, 1, 2

f32p = C.POINTER(C.c_float)
i32p = C.POINTER(C.c_int32)
i64p = C.POINTER(C.c_int64)

# every symbol include/mf_b200.h declares (tests/test_abi.py checks the header against this)
SIGNATURES = {
    "mfb_last_error": (C.c_char_p, []),
    "mfb_version": (C.c_char_p, []),
    "mfb_padding": (C.c_int, [C.c_int]),
    "mfb_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_int]),
    "mfb_destroy": (None, [C.c_void_p]),
    "mfb_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "mfb_sync": (C.c_int, [C.c_void_p]),
    "mfb_set_option": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int]),
    "mfb_enable": (C.c_int, [C.c_void_p, C.c_int]),
    "mfb_upload": (C.c_int, [C.c_void_p, C.c_int, f32p, C.c_int64, C.c_int64, C.c_int64]),
    "mfb_download": (C.c_int, [C.c_void_p, C.c_int, f32p, C.c_int64, C.c_int64, C.c_int64]),
    "mfb_init_normal": (C.c_int, [C.c_void_p, C.c_uint64, C.c_float]),
    "mfb_snapshot_old": (C.c_int, [C.c_void_p]),
    "mfb_device_ptr": (C.c_void_p, [C.c_void_p, C.c_int]),
    "mfb_dataset_create": (C.c_int, [C.c_void_p, C.POINTER(C.c_int)]),
    "mfb_dataset_append_block": (C.c_int, [C.c_void_p, C.c_int, C.c_int32, i32p, i32p, i32p, f32p]),
    "mfb_dataset_load_file": (C.c_int, [C.c_void_p, C.c_int, C.c_char_p]),
    "mfb_dataset_finalize": (C.c_int, [C.c_void_p, C.c_int]),
    "mfb_dataset_free": (C.c_int, [C.c_void_p, C.c_int]),
    "mfb_dataset_num_ratings": (C.c_int64, [C.c_void_p, C.c_int]),
    "mfb_dataset_num_runs": (C.c_int64, [C.c_void_p, C.c_int]),
    "mfb_blocks_read": (C.c_int, [C.c_char_p, C.POINTER(C.c_void_p)]),
    "mfb_blocks_from_arrays": (C.c_int, [C.c_int64, i64p, C.c_int64, i32p, i32p, i32p, f32p,
                                         C.POINTER(C.c_void_p)]),
    "mfb_blocks_write": (C.c_int, [C.c_void_p, C.c_char_p]),
    "mfb_blocks_free": (None, [C.c_void_p]),
    "mfb_blocks_num_blocks": (C.c_int64, [C.c_void_p]),
    "mfb_blocks_num_runs": (C.c_int64, [C.c_void_p]),
    "mfb_blocks_num_ratings": (C.c_int64, [C.c_void_p]),
    "mfb_blocks_block_off": (i64p, [C.c_void_p]),
    "mfb_blocks_run_uid": (i32p, [C.c_void_p]),
    "mfb_blocks_run_off": (i32p, [C.c_void_p]),
    "mfb_blocks_vid": (i32p, [C.c_void_p]),
    "mfb_blocks_rating": (f32p, [C.c_void_p]),
    "mfb_dataset_append_blocks": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "mfb_gen_defaults": (None, [C.c_void_p, C.c_int32, C.c_int32, C.c_int64]),
    "mfb_generate": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                               C.POINTER(C.c_void_p)]),
    "mfb_sgd_epoch": (C.c_int, [C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_float, C.c_int]),
    "mfb_sgd_epoch_blocks": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.c_float, C.c_float, C.c_float, C.c_int]),
    "mfb_dataset_num_blocks": (C.c_int64, [C.c_void_p, C.c_int]),
    "mfb_sgd_epoch_from_host": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_float, C.c_float,
                                          C.c_float, C.c_int, C.c_int64]),
    "mfb_sgd_epoch_from_file": (C.c_int, [C.c_void_p, C.c_char_p, C.c_float, C.c_float, C.c_float, C.c_int, C.c_int64,
                                          C.POINTER(C.c_int64)]),
    "mfb_dataset_refresh_from_host": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "mfb_blocks_pin": (C.c_int, [C.c_void_p]),
    "mfb_blocks_unpin": (C.c_int, [C.c_void_p]),
    "mfb_sse": (C.c_int, [C.c_void_p, C.c_int, C.c_float, C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
    "mfb_sse_link": (C.c_int, [C.c_void_p, C.c_int, C.c_float, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
    "mfb_seteta": (C.c_float, [C.c_float, C.c_int, C.c_float]),
    "mfb_seteta_cutoff": (C.c_float, [C.c_float, C.c_int, C.c_float, C.c_float]),
    "mfb_dp_weights": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_int32)]),
    "mfb_dp_bound": (C.c_float, [C.c_float, C.c_int, C.c_int]),
    "mfb_sgld_epoch": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_float, C.c_int]),
    "mfb_sgld_flush_noise": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "mfb_col_sqnorms": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double),
                                  C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "mfb_set_noise_table": (C.c_int, [C.c_void_p, f32p, C.c_int64]),
    "mfb_admf_set_validation": (C.c_int, [C.c_void_p, C.c_int64, i32p, i32p, f32p]),
    "mfb_admf_set_draws": (C.c_int, [C.c_void_p, C.c_int64, i32p]),
    "mfb_admf_set_lams": (C.c_int, [C.c_void_p, f32p]),
    "mfb_admf_get_lams": (C.c_int, [C.c_void_p, f32p]),
    "mfb_admf_epoch": (C.c_int, [C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_int, C.c_float, C.c_int]),
    "mfb_blocks_split_by_item": (C.c_int, [C.c_void_p, C.c_int, i32p, C.POINTER(C.c_void_p)]),
    "mfb_blocks_merge_runs": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]),
    "mfb_blocks_regroup": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "mfb_dataset_ingest_file": (C.c_int, [C.c_void_p, C.c_int, C.c_char_p, C.c_int, C.c_float, C.c_float, C.c_float, C.c_int,
                                          C.c_int64, C.POINTER(C.c_int64)]),
    "mfb_wire_index_file": (C.c_int, [C.c_char_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "mfb_comm_unique_id": (C.c_int, [C.c_void_p]),
    "mfb_comm_init": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "mfb_comm_destroy": (C.c_int, [C.c_void_p]),
    "mfb_comm_ipc_export": (C.c_int, [C.c_void_p, C.c_void_p]),
    "mfb_comm_ipc_import": (C.c_int, [C.c_void_p, C.c_void_p]),
    "mfb_comm_ipc_close": (C.c_int, [C.c_void_p]),
    "mfb_dsgd_epoch": (C.c_int, [C.c_void_p, C.POINTER(C.c_int), i32p, C.c_float, C.c_float, C.c_float, C.c_int]),
    "mfb_dsgd_epoch_ex": (C.c_int, [C.c_void_p, C.POINTER(C.c_int), i32p, C.c_int, C.c_int, C.c_float, C.c_float,
                                    C.c_float, C.c_int]),
    "mfb_dsgd_timeline": (C.c_int, [C.c_void_p, f32p, C.c_int]),
    "mfb_placement_report": (C.c_int, [C.c_void_p, C.c_int, f32p, C.c_int, C.POINTER(C.c_int)]),
    "mfb_comm_allgather_items": (C.c_int, [C.c_void_p, i32p]),
    "mfb_comm_allreduce_sse": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
    "mfb_probe_arm": (C.c_int, [C.c_void_p, C.c_int]),
    "mfb_probe_read": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64)]),
    "mfb_last_kernel_ms": (C.c_float, [C.c_void_p]),
    "mfb_launch_count": (C.c_int64, [C.c_void_p]),
    "mfb_h2d_bytes": (C.c_int64, [C.c_void_p]),
    "mfb_last_launch": (C.c_int, [C.c_void_p, C.POINTER(C.c_int)]),
}

_lib = None


class MfbError(RuntimeError):
    pass


def _prefer_bundled_nccl():
    """One NCCL per process: libmf_b200.so binds NCCL at run time (dlopen "libnccl.so.2", MFB_NCCL_LIB overrides).
    In a Python process that later imports torch, a system libnccl loaded first would be the one torch then gets
    (same SONAME) - and torch needs its own, newer one.  Point the library at the pip-installed copy torch uses."""
    if os.environ.get("MFB_NCCL_LIB"):
        return
    try:
        import importlib.util
        spec = importlib.util.find_spec("nvidia.nccl")
        for d in (spec.submodule_search_locations if spec else []):
            cand = os.path.join(d, "lib", "libnccl.so.2")
            if os.path.exists(cand):
                os.environ["MFB_NCCL_LIB"] = cand
                return
    except Exception:
        pass


def lib():
    """Load libmf_b200.so (fails loudly if it has not been built: run __graft_entry__.build())."""
    global _lib
    if _lib is None:
        _prefer_bundled_nccl()
        if not os.path.exists(LIB_PATH):
            raise MfbError("%s not built (make -C experimental-mf_b200)" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def _check(rc):
    if rc != 0:
        raise MfbError("mf_b200 error %d: %s" % (rc, lib().mfb_last_error().decode()))


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


class GenParams(C.Structure):
    """mfb_gen_params (include/mf_b200.h)."""
    _fields_ = [("nu", C.c_int32), ("nv", C.c_int32), ("nnz", C.c_int64), ("rank", C.c_int32),
                ("gb", C.c_float), ("noise_sd", C.c_float), ("degree_sigma", C.c_float),
                ("zipf_s", C.c_float), ("test_frac", C.c_float), ("valid_frac", C.c_float),
                ("split", C.c_int32), ("users_per_block", C.c_int32), ("seed", C.c_uint64),
                ("user_begin", C.c_int32), ("user_end", C.c_int32), ("threads", C.c_int32)]


class SgldParams(C.Structure):
    """mfb_sgld_params (include/mf_b200.h)."""
    _fields_ = [("eta", C.c_float), ("temp", C.c_float), ("bound", C.c_float), ("ntrain", C.c_int32),
                ("lambda_r", C.c_float), ("lambda_ub", C.c_float), ("lambda_vb", C.c_float),
                ("seed", C.c_uint64), ("round", C.c_uint32), ("use_table", C.c_int32),
                ("table_offset", C.c_int32)]


class Blocks:
    """A parsed rating file on the host (mfb_blocks): numpy views of its flat arrays."""

    def __init__(self, handle):
        self.h = C.c_void_p(handle) if not isinstance(handle, C.c_void_p) else handle

    @staticmethod
    def read(path):
        h = C.c_void_p()
        _check(lib().mfb_blocks_read(path.encode(), C.byref(h)))
        return Blocks(h)

    @staticmethod
    def from_arrays(block_off, run_uid, run_off, vid, rating):
        bo = np.ascontiguousarray(block_off, np.int64)
        ru = np.ascontiguousarray(run_uid, np.int32)
        ro = np.ascontiguousarray(run_off, np.int32)
        vi = np.ascontiguousarray(vid, np.int32)
        ra = _f32(rating)
        h = C.c_void_p()
        _check(lib().mfb_blocks_from_arrays(len(bo) - 1, bo.ctypes.data_as(i64p), len(ru),
                                            ru.ctypes.data_as(i32p), ro.ctypes.data_as(i32p),
                                            vi.ctypes.data_as(i32p), ra.ctypes.data_as(f32p), C.byref(h)))
        return Blocks(h)

    def write(self, path):
        _check(lib().mfb_blocks_write(self.h, path.encode()))
        return path

    def pin(self):
        _check(lib().mfb_blocks_pin(self.h))

    def unpin(self):
        _check(lib().mfb_blocks_unpin(self.h))

    def close(self):
        if self.h:
            lib().mfb_blocks_unpin(self.h)
            lib().mfb_blocks_free(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def split_by_item(self, bounds):
        b = np.ascontiguousarray(bounds, np.int32)
        n = len(b) - 1
        out = (C.c_void_p * n)()
        _check(lib().mfb_blocks_split_by_item(self.h, n, b.ctypes.data_as(i32p), out))
        return [Blocks(C.c_void_p(x)) for x in out]

    def merge_runs(self, users_per_block=500):
        out = C.c_void_p()
        _check(lib().mfb_blocks_merge_runs(self.h, users_per_block, C.byref(out)))
        return Blocks(out)

    def regroup(self, merge_users=True, longest_first=True, users_per_block=500):
        out = C.c_void_p()
        _check(lib().mfb_blocks_regroup(self.h, int(merge_users), int(longest_first), users_per_block, C.byref(out)))
        return Blocks(out)

    @property
    def nblocks(self):
        return lib().mfb_blocks_num_blocks(self.h)

    @property
    def nruns(self):
        return lib().mfb_blocks_num_runs(self.h)

    @property
    def nratings(self):
        return lib().mfb_blocks_num_ratings(self.h)

    def _view(self, fn, n, dtype):
        if n == 0:
            return np.zeros(0, dtype)
        return np.ctypeslib.as_array(fn(self.h), (n,))

    @property
    def block_off(self):
        return self._view(lib().mfb_blocks_block_off, self.nblocks + 1, np.int64)

    @property
    def run_uid(self):
        return self._view(lib().mfb_blocks_run_uid, self.nruns, np.int32)

    @property
    def run_off(self):
        return self._view(lib().mfb_blocks_run_off, self.nruns + 1, np.int32)

    @property
    def vid(self):
        return self._view(lib().mfb_blocks_vid, self.nratings, np.int32)

    @property
    def rating(self):
        return self._view(lib().mfb_blocks_rating, self.nratings, np.float32)


def gen_params(nu, nv, nnz, **kw):
    p = GenParams()
    lib().mfb_gen_defaults(C.byref(p), nu, nv, nnz)
    for k, v in kw.items():
        assert hasattr(p, k), k
        setattr(p, k, v)
    return p


def generate(params):
    """-> (train, test, valid) Blocks of the synthetic data set described by `params`."""
    t, s, v = C.c_void_p(), C.c_void_p(), C.c_void_p()
    _check(lib().mfb_generate(C.byref(params), C.byref(t), C.byref(s), C.byref(v)))
    return Blocks(t), Blocks(s), Blocks(v)


class Context:
    """One model (theta, phi, bu, bv [+ optional arrays]) resident in the HBM of one GPU."""

    def __init__(self, nu, nv, dim, device=0):
        self.nu, self.nv, self.dim = nu, nv, dim
        self.h = C.c_void_p()
        _check(lib().mfb_create(C.byref(self.h), device, nu, nv, dim))

    def close(self):
        if self.h:
            lib().mfb_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- plumbing
    def set_stream(self, cuda_stream):
        _check(lib().mfb_set_stream(self.h, C.c_void_p(cuda_stream)))

    def sync(self):
        _check(lib().mfb_sync(self.h))

    def set_option(self, name, value):
        _check(lib().mfb_set_option(self.h, name.encode(), int(value)))

    def enable(self, group):
        _check(lib().mfb_enable(self.h, group))

    # -- factors
    def rows(self, which):
        return {THETA: self.nu, THETA_OLD: self.nu, BU: self.nu, BU_OLD: self.nu, UR: self.nu,
                PHI: self.nv, PHI_OLD: self.nv, BV: self.nv, BV_OLD: self.nv, VR: self.nv,
                LAMBDA_U: self.dim, LAMBDA_V: self.dim}[which]

    def cols(self, which):
        return self.dim if which in (THETA, PHI, THETA_OLD, PHI_OLD) else 1

    def upload(self, which, arr, row0=0):
        a = _f32(arr)
        a2 = a.reshape(a.shape[0], -1)
        _check(lib().mfb_upload(self.h, which, a2.ctypes.data_as(f32p), row0, a2.shape[0], a2.shape[1]))

    def download(self, which, row0=0, nrows=None):
        nrows = self.rows(which) - row0 if nrows is None else nrows
        c = self.cols(which)
        out = np.empty((nrows, c), np.float32)
        _check(lib().mfb_download(self.h, which, out.ctypes.data_as(f32p), row0, nrows, c))
        return out if c > 1 else out.reshape(-1)

    def set_factors(self, theta, phi, bu, bv):
        self.upload(THETA, theta)
        self.upload(PHI, phi)
        self.upload(BU, bu)
        self.upload(BV, bv)

    def get_factors(self):
        return self.download(THETA), self.download(PHI), self.download(BU), self.download(BV)

    def init_normal(self, seed, scale=1e-2):
        _check(lib().mfb_init_normal(self.h, seed, scale))

    def snapshot_old(self):
        _check(lib().mfb_snapshot_old(self.h))

    def device_ptr(self, which):
        return lib().mfb_device_ptr(self.h, which)

    # -- datasets
    def dataset_from_arrays(self, block_off, run_uid, run_off, vid, rating):
        """File-order arrays (blocks -> user-runs -> records) -> one finalized dataset id."""
        ds = C.c_int()
        _check(lib().mfb_dataset_create(self.h, C.byref(ds)))
        run_uid = np.ascontiguousarray(run_uid, np.int32)
        vid = np.ascontiguousarray(vid, np.int32)
        rating = _f32(rating)
        run_off = np.asarray(run_off, np.int64)
        block_off = np.asarray(block_off, np.int64)
        for b in range(len(block_off) - 1):
            r0, r1 = int(block_off[b]), int(block_off[b + 1])
            base = int(run_off[r0])
            rec_off = np.ascontiguousarray(run_off[r0:r1 + 1] - base, np.int32)
            _check(lib().mfb_dataset_append_block(
                self.h, ds, r1 - r0, run_uid[r0:r1].ctypes.data_as(i32p), rec_off.ctypes.data_as(i32p),
                vid[base:].ctypes.data_as(i32p), rating[base:].ctypes.data_as(f32p)))
        _check(lib().mfb_dataset_finalize(self.h, ds))
        return ds.value

    def dataset_from_blocks(self, blocks):
        ds = C.c_int()
        _check(lib().mfb_dataset_create(self.h, C.byref(ds)))
        _check(lib().mfb_dataset_append_blocks(self.h, ds, blocks.h))
        _check(lib().mfb_dataset_finalize(self.h, ds))
        return ds.value

    def dataset_from_file(self, path):
        ds = C.c_int()
        _check(lib().mfb_dataset_create(self.h, C.byref(ds)))
        _check(lib().mfb_dataset_load_file(self.h, ds, path.encode()))
        _check(lib().mfb_dataset_finalize(self.h, ds))
        return ds.value

    def dataset_ingest_file(self, path, with_epoch=False, eta=0.0, lam=0.0, gb=0.0, mode=MODE_ATOMIC, tile_ratings=0):
        """streaming ingest: a finalized dataset from the file in one pass (records decoded on the GPU), optionally with
        the SGD epoch run on every chunk as it lands; returns (dataset id, records)"""
        ds = C.c_int()
        _check(lib().mfb_dataset_create(self.h, C.byref(ds)))
        n = C.c_int64()
        _check(lib().mfb_dataset_ingest_file(self.h, ds.value, path.encode(), int(with_epoch), eta, lam, gb, mode, tile_ratings,
                                             C.byref(n)))
        return ds.value, n.value

    def dataset_free(self, ds):
        _check(lib().mfb_dataset_free(self.h, ds))

    def num_ratings(self, ds):
        return lib().mfb_dataset_num_ratings(self.h, ds)

    def num_runs(self, ds):
        return lib().mfb_dataset_num_runs(self.h, ds)

    # -- hot path
    def sgd_epoch(self, ds, eta, lam, gb, mode=MODE_HOGWILD):
        _check(lib().mfb_sgd_epoch(self.h, ds, eta, lam, gb, mode))

    def sgd_epoch_blocks(self, ds, block_begin, block_end, eta, lam, gb, mode=MODE_ATOMIC):
        _check(lib().mfb_sgd_epoch_blocks(self.h, ds, block_begin, block_end, eta, lam, gb, mode))

    def num_blocks(self, ds):
        return lib().mfb_dataset_num_blocks(self.h, ds)

    def dataset_refresh_from_host(self, ds, blocks):
        _check(lib().mfb_dataset_refresh_from_host(self.h, ds, blocks.h))

    def sgd_epoch_from_file(self, path, eta, lam, gb, mode=MODE_ATOMIC, tile_ratings=0):
        """one out-of-core epoch over the rating file at `path`; returns the number of records processed"""
        n = C.c_int64()
        _check(lib().mfb_sgd_epoch_from_file(self.h, path.encode(), eta, lam, gb, mode, tile_ratings, C.byref(n)))
        return n.value

    def sgd_epoch_from_host(self, ds, blocks, eta, lam, gb, mode=MODE_HOGWILD, chunk_ratings=0):
        _check(lib().mfb_sgd_epoch_from_host(self.h, ds, blocks.h, eta, lam, gb, mode, chunk_ratings))

    # -- dpmf
    def dp_weights(self, ds):
        n = C.c_int32()
        _check(lib().mfb_dp_weights(self.h, ds, C.byref(n)))
        return n.value

    def sgld_epoch(self, ds, params, gb, mode=MODE_HOGWILD):
        _check(lib().mfb_sgld_epoch(self.h, ds, C.byref(params), gb, mode))

    def sgld_flush_noise(self, ds, params):
        _check(lib().mfb_sgld_flush_noise(self.h, ds, C.byref(params)))

    def col_sqnorms(self):
        nu_, nv_ = np.zeros(self.dim), np.zeros(self.dim)
        bu2, bv2 = C.c_double(), C.c_double()
        dp = C.POINTER(C.c_double)
        _check(lib().mfb_col_sqnorms(self.h, nu_.ctypes.data_as(dp), nv_.ctypes.data_as(dp),
                                     C.byref(bu2), C.byref(bv2)))
        return nu_, nv_, bu2.value, bv2.value

    def set_noise_table(self, table):
        t = _f32(table)
        _check(lib().mfb_set_noise_table(self.h, t.ctypes.data_as(f32p), len(t)))

    # -- admf
    def admf_set_validation(self, u, v, r):
        u, v, r = np.ascontiguousarray(u, np.int32), np.ascontiguousarray(v, np.int32), _f32(r)
        _check(lib().mfb_admf_set_validation(self.h, len(u), u.ctypes.data_as(i32p), v.ctypes.data_as(i32p),
                                             r.ctypes.data_as(f32p)))

    def admf_set_draws(self, draws):
        d = np.ascontiguousarray(draws, np.int32)
        _check(lib().mfb_admf_set_draws(self.h, len(d), d.ctypes.data_as(i32p)))

    def admf_set_lams(self, lams):
        a = _f32(lams)
        _check(lib().mfb_admf_set_lams(self.h, a.ctypes.data_as(f32p)))

    def admf_get_lams(self):
        a = np.zeros(4, np.float32)
        _check(lib().mfb_admf_get_lams(self.h, a.ctypes.data_as(f32p)))
        return a

    def admf_epoch(self, ds, eta, eta_reg, loss, gb, mode=MODE_ATOMIC):
        _check(lib().mfb_admf_epoch(self.h, ds, eta, eta_reg, loss, gb, mode))

    # -- multi-GPU (DSGD ring)
    def comm_init(self, rank, world, unique_id):
        buf = C.create_string_buffer(bytes(unique_id), 128)
        _check(lib().mfb_comm_init(self.h, rank, world, buf))

    def comm_ipc_export(self):
        buf = C.create_string_buffer(208)
        _check(lib().mfb_comm_ipc_export(self.h, buf))
        return buf.raw

    def comm_ipc_import(self, handles208):
        assert len(handles208) == 208
        _check(lib().mfb_comm_ipc_import(self.h, C.create_string_buffer(handles208, 208)))

    def comm_ipc_close(self):
        _check(lib().mfb_comm_ipc_close(self.h))

    def dsgd_epoch(self, datasets, item_bounds, eta, lam, gb, mode=MODE_ATOMIC, halves=1, rotations=1):
        ds = (C.c_int * len(datasets))(*datasets)
        b = np.ascontiguousarray(item_bounds, np.int32)
        _check(lib().mfb_dsgd_epoch_ex(self.h, ds, b.ctypes.data_as(i32p), halves, rotations, eta, lam, gb, mode))

    def placement_report(self, which=None):
        """(calibration ms of every candidate placement, index of the one kept) for the plane-layout
        working copy (which=1) or the rows of the item matrix (which=0); None = whichever search ran"""
        if which is None:
            a = self.placement_report(1)
            return a if a[0] else self.placement_report(0)
        ms, best = np.zeros(64, np.float32), C.c_int(-1)
        n = lib().mfb_placement_report(self.h, which, ms.ctypes.data_as(f32p), 64, C.byref(best))
        return ms[:max(n, 0)].tolist(), best.value

    def dsgd_timeline(self, steps):
        """(wait ms, kernel ms) alternating, for the first `steps` cell kernels of the most recent DSGD epoch"""
        out = np.zeros(2 * steps, np.float32)
        n = lib().mfb_dsgd_timeline(self.h, out.ctypes.data_as(f32p), len(out))
        if n < 0:
            _check(n)
        return out[:n]

    def allgather_items(self, item_bounds):
        b = np.ascontiguousarray(item_bounds, np.int32)
        _check(lib().mfb_comm_allgather_items(self.h, b.ctypes.data_as(i32p)))

    def allreduce_sse(self, sse, n):
        s, k = C.c_double(sse), C.c_int64(n)
        _check(lib().mfb_comm_allreduce_sse(self.h, C.byref(s), C.byref(k)))
        return s.value, k.value

    def sse(self, ds, gb, link=0):
        s, n = C.c_double(), C.c_int64()
        _check(lib().mfb_sse_link(self.h, ds, gb, link, C.byref(s), C.byref(n)))
        return s.value, n.value

    def rmse(self, ds, gb):
        s, n = self.sse(ds, gb)
        return float(np.sqrt(s / max(n, 1)))

    def last_kernel_ms(self):
        return lib().mfb_last_kernel_ms(self.h)

    def probe_arm(self, item):
        _check(lib().mfb_probe_arm(self.h, int(item)))

    def probe_read(self):
        """(mean stale updates per update over all items, mean for the probed item, updates)"""
        out = (C.c_uint64 * 4)()
        _check(lib().mfb_probe_read(self.h, out))
        s, n, hs, hn = [int(x) for x in out]
        return s / max(n, 1), hs / max(hn, 1), n

    def launch_count(self):
        return lib().mfb_launch_count(self.h)

    def h2d_bytes(self):
        return lib().mfb_h2d_bytes(self.h)

    def last_launch(self):
        out = (C.c_int * 4)()
        _check(lib().mfb_last_launch(self.h, out))
        return {"kernel": out[0], "grid": out[1], "threads": out[2], "ring": out[3]}


def seeded_model(nu, nv, dim, seed, scale=1e-2):
    """(theta[nu][dim], phi[nv][dim], bu[nu], bv[nv]) ~ N(0,1)*scale from numpy's seeded generator: the initial
    model of every full-size parity run (tests/golden/make_fullsize_golden.py feeds the same arrays to the
    reference, whose own init is clock-seeded, model.cc:3-5); equals tests/oraclelib.Model(nu, nv, dim, seed)."""
    rng = np.random.default_rng(seed)
    theta = rng.standard_normal((nu, dim), dtype=np.float32) * np.float32(scale)
    phi = rng.standard_normal((nv, dim), dtype=np.float32) * np.float32(scale)
    bu = rng.standard_normal(nu, dtype=np.float32) * np.float32(scale)
    bv = rng.standard_normal(nv, dtype=np.float32) * np.float32(scale)
    return theta, phi, bu, bv


def wire_index_file(path):
    """(frames, serialized users, bytes inside them) of a [u32][mf.Block] file: the host half of the device decoder"""
    a, b, c = C.c_int64(), C.c_int64(), C.c_int64()
    _check(lib().mfb_wire_index_file(path.encode(), C.byref(a), C.byref(b), C.byref(c)))
    return a.value, b.value, c.value


def comm_unique_id():
    buf = C.create_string_buffer(128)
    _check(lib().mfb_comm_unique_id(buf))
    return buf.raw


def seteta(eta0, rnd, gam):
    return lib().mfb_seteta(eta0, rnd, gam)
