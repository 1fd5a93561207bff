"""ctypes bindings for the TEST-ONLY oracle libraries (oracle/libmf_oracle.so and, when it has
been built, oracle/_ref/libmf_ref.so) plus small synthetic-data helpers used by the tests.

Nothing in the product imports this module."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "libmf_oracle.so")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "libmf_ref.so")
REF_BIN = os.path.join(ORACLE_DIR, "_ref", "mf_ref")

f32p = C.POINTER(C.c_float)
i32p = C.POINTER(C.c_int32)
i64p = C.POINTER(C.c_int64)
u64p = C.POINTER(C.c_uint64)


def _p(a, t):
    return a.ctypes.data_as(t)


class MfoModel(C.Structure):
    _fields_ = [("nu", C.c_int32), ("nv", C.c_int32), ("dim", C.c_int32), ("stride", C.c_int32),
                ("theta", f32p), ("phi", f32p), ("bu", f32p), ("bv", f32p)]


class MfoData(C.Structure):
    _fields_ = [("nblocks", C.c_int64), ("block_off", i64p), ("nruns", C.c_int64),
                ("run_uid", i32p), ("run_off", i64p), ("vid", i32p), ("rating", f32p)]


class MfoFile(C.Structure):
    _fields_ = [("d", MfoData), ("nratings", C.c_int64)]


class MfoDpState(C.Structure):
    _fields_ = [("eta", C.c_float), ("temp", C.c_float), ("bound", C.c_float),
                ("ntrain", C.c_int32), ("lambda_r", C.c_float), ("lambda_ub", C.c_float),
                ("lambda_vb", C.c_float), ("lambda_u", f32p), ("lambda_v", f32p), ("ur", f32p),
                ("vr", f32p), ("gcount", C.c_uint64), ("gcountu", u64p), ("gcountv", u64p)]


class MfoAdState(C.Structure):
    _fields_ = [("eta", C.c_float), ("eta_reg", C.c_float), ("loss", C.c_int32),
                ("lam_u", C.c_float), ("lam_v", C.c_float), ("lam_bu", C.c_float),
                ("lam_bv", C.c_float), ("theta_old", f32p), ("phi_old", f32p), ("bu_old", f32p),
                ("bv_old", f32p), ("nvalid", C.c_int64), ("val_u", i32p), ("val_v", i32p),
                ("val_r", f32p), ("draws", i32p), ("draw_pos", C.c_int64)]


class MfoNoiseTable(C.Structure):
    _fields_ = [("table", f32p), ("size", C.c_int64), ("offset", C.c_int32)]


class MfoNoisePhilox(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("round", C.c_uint32)]


NOISE_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_int, C.c_int32, C.c_int64, C.c_int32, C.c_int32, f32p)


def build_oracle():
    """(Re)build the oracle libraries with oracle/Makefile; _ref only where /root/reference exists."""
    subprocess.run(["make", "-s", "-C", ORACLE_DIR, "all"], check=True)


_oracle = None


def oracle():
    global _oracle
    if _oracle is None:
        if not os.path.exists(ORACLE_SO):
            build_oracle()
        L = C.CDLL(ORACLE_SO)
        L.mfo_padding.restype = C.c_int
        L.mfo_seteta.restype = C.c_float
        L.mfo_seteta.argtypes = [C.c_float, C.c_int, C.c_float]
        L.mfo_seteta_cutoff.restype = C.c_float
        L.mfo_seteta_cutoff.argtypes = [C.c_float, C.c_int, C.c_float, C.c_float]
        L.mfo_read_blocks.restype = C.POINTER(MfoFile)
        L.mfo_read_blocks.argtypes = [C.c_char_p]
        L.mfo_free_file.argtypes = [C.POINTER(MfoFile)]
        L.mfo_write_blocks.argtypes = [C.c_char_p, C.POINTER(MfoData)]
        L.mfo_sgd_epoch.argtypes = [C.POINTER(MfoModel), C.POINTER(MfoData), C.c_float, C.c_float,
                                    C.c_float]
        L.mfo_sse.restype = C.c_float
        L.mfo_sse.argtypes = [C.POINTER(MfoModel), C.POINTER(MfoData), C.c_float, i64p]
        L.mfo_sse_link.restype = C.c_float
        L.mfo_sse_link.argtypes = [C.POINTER(MfoModel), C.POINTER(MfoData), C.c_float, C.c_int, i64p]
        L.mfo_dp_bound.restype = C.c_float
        L.mfo_dp_bound.argtypes = [C.c_float, C.c_int]
        L.mfo_dp_weights.restype = C.c_int32
        L.mfo_dp_weights.argtypes = [C.POINTER(MfoData), C.c_int, C.c_int, f32p, f32p]
        L.mfo_sgld_epoch.argtypes = [C.POINTER(MfoModel), C.POINTER(MfoData),
                                     C.POINTER(MfoDpState), C.c_float, C.c_void_p, C.c_void_p]
        L.mfo_finish_noise.argtypes = [C.POINTER(MfoModel), C.POINTER(MfoDpState), C.c_void_p,
                                       C.c_void_p]
        L.mfo_sample_hyper.argtypes = [C.POINTER(MfoModel), C.POINTER(MfoDpState), C.c_float,
                                       C.c_float, C.c_float]
        L.mfo_sample_gamma.restype = C.c_float
        L.mfo_sample_gamma.argtypes = [C.c_float, C.c_float]
        L.mfo_shuffle_valid.argtypes = [C.c_int64, i32p, i32p, f32p]
        L.mfo_admf_epoch.argtypes = [C.POINTER(MfoModel), C.POINTER(MfoData),
                                     C.POINTER(MfoAdState), C.c_float]
        L.mfo_srand.argtypes = [C.c_uint]
        L.mfo_rand_draws.argtypes = [C.c_int64, C.c_int64, i32p]
        L.mfo_philox4x32_10.argtypes = [C.POINTER(C.c_uint32), C.POINTER(C.c_uint32),
                                        C.POINTER(C.c_uint32)]
        L.mfo_philox_normal4.argtypes = [C.c_uint64, C.c_uint32, C.c_int, C.c_int32, C.c_int64,
                                         C.c_uint32, f32p]
        L.mfo_philox_bias_normal.restype = C.c_float
        L.mfo_philox_bias_normal.argtypes = [C.c_uint64, C.c_uint32, C.c_int, C.c_int32, C.c_int64]
        _oracle = L
    return _oracle


def have_ref():
    return os.path.exists(REF_SO)


_ref = None


def ref():
    global _ref
    if _ref is None:
        L = C.CDLL(REF_SO)
        L.ref_create_mf.restype = C.c_void_p
        L.ref_create_mf.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_float, C.c_float,
                                    C.c_float, C.c_float, C.c_int, C.c_int]
        L.ref_create_dpmf.restype = C.c_void_p
        L.ref_create_dpmf.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_float, C.c_float,
                                      C.c_float, C.c_float, C.c_int, C.c_int, C.c_float, C.c_float,
                                      C.c_float, C.c_int, C.c_int, C.c_float, C.c_float]
        L.ref_create_admf.restype = C.c_void_p
        L.ref_create_admf.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_int, C.c_float,
                                      C.c_float, C.c_float, C.c_float, C.c_int, C.c_int, C.c_int,
                                      C.c_float]
        for name in ("ref_set_factors", "ref_get_factors", "ref_get_old"):
            getattr(L, name).argtypes = [C.c_void_p, f32p, f32p, f32p, f32p]
        L.ref_seteta.argtypes = [C.c_void_p, C.c_int]
        L.ref_get_eta.restype = C.c_float
        L.ref_get_eta.argtypes = [C.c_void_p]
        L.ref_get_etareg.restype = C.c_float
        L.ref_get_etareg.argtypes = [C.c_void_p]
        L.ref_epoch.argtypes = [C.c_void_p]
        L.ref_calc_mse.restype = C.c_float
        L.ref_calc_mse.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int)]
        L.ref_dpmf_set_noise.argtypes = [C.c_void_p, f32p, C.c_int]
        L.ref_dpmf_set_offset.argtypes = [C.c_void_p, C.c_int]
        L.ref_dpmf_finish_noise.argtypes = [C.c_void_p]
        L.ref_dpmf_sample_hyper.argtypes = [C.c_void_p, C.c_float]
        L.ref_dpmf_get_hyper.argtypes = [C.c_void_p, f32p]
        L.ref_dpmf_set_hyper.argtypes = [C.c_void_p, f32p]
        L.ref_dpmf_get_weights.argtypes = [C.c_void_p, f32p, f32p]
        L.ref_dpmf_info.argtypes = [C.c_void_p, C.POINTER(C.c_int), f32p, C.POINTER(C.c_int)]
        L.ref_admf_get_lams.argtypes = [C.c_void_p, f32p]
        L.ref_num_valid.argtypes = [C.c_void_p]
        L.ref_get_valid.argtypes = [C.c_void_p, i32p, i32p, f32p]
        L.ref_set_io_paths.argtypes = [C.c_char_p, C.c_char_p]
        L.ref_save_model.argtypes = [C.c_void_p, C.c_int]
        L.ref_read_model.argtypes = [C.c_void_p]
        L.ref_dpmf_read_hyper.argtypes = [C.c_void_p]
        L.ref_get_lambda.restype = C.c_float
        L.ref_get_lambda.argtypes = [C.c_void_p]
        L.ref_srand.argtypes = [C.c_uint]
        L.ref_seed_generator.argtypes = [C.c_uint]
        L.ref_padding.argtypes = [C.c_int]
        _ref = L
    return _ref


class Dataset:
    """A rating file in file order (blocks -> user-runs -> records), as numpy arrays."""

    def __init__(self, block_off, run_uid, run_off, vid, rating):
        # always own the memory: the inputs may be views into an mfb_blocks that is freed later
        self.block_off = np.array(block_off, dtype=np.int64, order="C", copy=True)
        self.run_uid = np.array(run_uid, dtype=np.int32, order="C", copy=True)
        self.run_off = np.array(run_off, dtype=np.int64, order="C", copy=True)
        self.vid = np.array(vid, dtype=np.int32, order="C", copy=True)
        self.rating = np.array(rating, dtype=np.float32, order="C", copy=True)

    @property
    def nratings(self):
        return int(self.vid.shape[0])

    @property
    def nruns(self):
        return int(self.run_uid.shape[0])

    @property
    def nblocks(self):
        return int(self.block_off.shape[0] - 1)

    def as_mfo(self):
        return MfoData(self.nblocks, _p(self.block_off, i64p), self.nruns, _p(self.run_uid, i32p),
                       _p(self.run_off, i64p), _p(self.vid, i32p), _p(self.rating, f32p))

    def block_range(self, k0, k1):
        """Blocks [k0, k1) of the file as a Dataset of their own (a slice of an epoch)."""
        r0, r1 = int(self.block_off[k0]), int(self.block_off[k1])
        o0, o1 = int(self.run_off[r0]), int(self.run_off[r1])
        return Dataset(self.block_off[k0:k1 + 1] - r0, self.run_uid[r0:r1], self.run_off[r0:r1 + 1] - o0,
                       self.vid[o0:o1], self.rating[o0:o1])

    def uid_per_rating(self):
        return np.repeat(self.run_uid, np.diff(self.run_off))

    def write(self, path):
        d = self.as_mfo()
        rc = oracle().mfo_write_blocks(path.encode(), C.byref(d))
        assert rc == 0
        return path

    @staticmethod
    def read(path):
        L = oracle()
        fp = L.mfo_read_blocks(path.encode())
        assert fp, "cannot read %s" % path
        f = fp.contents
        d = f.d
        nb, nr, n = d.nblocks, d.nruns, f.nratings
        out = Dataset(np.ctypeslib.as_array(d.block_off, (nb + 1,)).copy(),
                      np.ctypeslib.as_array(d.run_uid, (nr,)).copy() if nr else np.zeros(0, np.int32),
                      np.ctypeslib.as_array(d.run_off, (nr + 1,)).copy(),
                      np.ctypeslib.as_array(d.vid, (n,)).copy() if n else np.zeros(0, np.int32),
                      np.ctypeslib.as_array(d.rating, (n,)).copy() if n else np.zeros(0, np.float32))
        L.mfo_free_file(fp)
        return out

    @staticmethod
    def from_triples(u, v, r, users_per_block=50):
        """Group consecutive equal-uid triples into runs, `users_per_block` runs per block."""
        u = np.asarray(u, np.int32)
        n = len(u)
        if n == 0:
            return Dataset([0], [], [0], [], [])
        starts = np.flatnonzero(np.r_[True, u[1:] != u[:-1]])
        run_off = np.r_[starts, n].astype(np.int64)
        run_uid = u[starts]
        nruns = len(run_uid)
        block_off = np.r_[np.arange(0, nruns, users_per_block), nruns].astype(np.int64)
        return Dataset(block_off, run_uid, run_off, v, r)


class Model:
    """Dense fp32 model with the reference's row stride padding(dim) (util.h:163-165)."""

    def __init__(self, nu, nv, dim, seed=0, scale=1e-2):
        self.nu, self.nv, self.dim = nu, nv, dim
        self.stride = oracle().mfo_padding(dim)
        rng = np.random.default_rng(seed)
        self.theta = np.zeros((nu, self.stride), np.float32)
        self.phi = np.zeros((nv, self.stride), np.float32)
        self.theta[:, :dim] = rng.standard_normal((nu, dim), dtype=np.float32) * scale
        self.phi[:, :dim] = rng.standard_normal((nv, dim), dtype=np.float32) * scale
        self.bu = (rng.standard_normal(nu, dtype=np.float32) * scale).astype(np.float32)
        self.bv = (rng.standard_normal(nv, dtype=np.float32) * scale).astype(np.float32)

    def copy(self):
        m = Model.__new__(Model)
        m.nu, m.nv, m.dim, m.stride = self.nu, self.nv, self.dim, self.stride
        m.theta, m.phi, m.bu, m.bv = self.theta.copy(), self.phi.copy(), self.bu.copy(), self.bv.copy()
        return m

    def as_mfo(self):
        return MfoModel(self.nu, self.nv, self.dim, self.stride, _p(self.theta, f32p),
                        _p(self.phi, f32p), _p(self.bu, f32p), _p(self.bv, f32p))

    def dense(self):
        """(theta[nu][dim], phi[nv][dim]) without the padding columns, C-contiguous."""
        return (np.ascontiguousarray(self.theta[:, :self.dim]),
                np.ascontiguousarray(self.phi[:, :self.dim]))

    def set_dense(self, theta, phi, bu, bv):
        self.theta[:, :self.dim] = theta
        self.phi[:, :self.dim] = phi
        self.bu[:] = bu
        self.bv[:] = bv


def make_ratings(nu, nv, nnz, seed=1, split=4, users_per_block=50, gb=2.76, test_frac=0.1,
                 valid_frac=0.0, rank=8):
    """Small numpy re-creation of the SURVEY 8d recipe (planted low-rank model, skewed degrees and
    popularity, no duplicate (u,i), getdata.cc-style `--split` user grouping).  Returns
    (train, test, valid) Datasets; valid is None when valid_frac == 0."""
    rng = np.random.default_rng(seed)
    U = rng.standard_normal((nu, rank)).astype(np.float32) * 0.5
    V = rng.standard_normal((nv, rank)).astype(np.float32) * 0.5
    deg = rng.lognormal(0.0, 1.0, nu)
    deg = np.clip(np.round(deg / deg.sum() * nnz), 1, max(1, nv // 2)).astype(np.int64)
    pop = 1.0 / np.arange(1, nv + 1)
    pop /= pop.sum()
    perm = rng.permutation(nv)
    us, vs = [], []
    for u in range(nu):
        items = perm[rng.choice(nv, size=int(deg[u]), replace=False, p=pop)]
        us.append(np.full(len(items), u, np.int32))
        vs.append(items.astype(np.int32))
    u = np.concatenate(us)
    v = np.concatenate(vs)
    r = gb + np.einsum("ij,ij->i", U[u], V[v]) + rng.normal(0, 0.5, len(u))
    r = np.clip(np.round(r), 1, 5).astype(np.float32)
    order = rng.permutation(len(u))  # getdata.cc:33-36 global shuffle
    u, v, r = u[order], v[order], r[order]
    n = len(u)
    n_test = int(n * test_frac)
    n_valid = int(n * valid_frac)

    def grouped(uu, vv, rr, nsplit):
        # getdata.cc:53-80: cut into `nsplit` chunks, group each chunk by user
        outs = []
        for ch in np.array_split(np.arange(len(uu)), nsplit):
            if len(ch) == 0:
                continue
            users = rng.permutation(np.unique(uu[ch]))
            rank_of = np.zeros(nu, np.int64)
            rank_of[users] = np.arange(len(users))
            o = ch[np.argsort(rank_of[uu[ch]], kind="stable")]
            outs.append(o)
        o = np.concatenate(outs) if outs else np.zeros(0, np.int64)
        return Dataset.from_triples(uu[o], vv[o], rr[o], users_per_block)

    test = grouped(u[:n_test], v[:n_test], r[:n_test], 1)
    valid = grouped(u[n_test:n_test + n_valid], v[n_test:n_test + n_valid],
                    r[n_test:n_test + n_valid], 1) if n_valid else None
    s = n_test + n_valid
    train = grouped(u[s:], v[s:], r[s:], split)
    return train, test, valid


class Ref:
    """One reference model object (MF / DPMF / AdaptRegMF) behind oracle/_ref/libmf_ref.so."""

    def __init__(self, handle, nu, nv, dim):
        self.h, self.nu, self.nv, self.dim = handle, nu, nv, dim

    def set_factors(self, theta, phi, bu, bv):
        a = [np.ascontiguousarray(x, np.float32) for x in (theta, phi, bu, bv)]
        ref().ref_set_factors(self.h, *[_p(x, f32p) for x in a])

    def get_factors(self):
        theta = np.zeros((self.nu, self.dim), np.float32)
        phi = np.zeros((self.nv, self.dim), np.float32)
        bu = np.zeros(self.nu, np.float32)
        bv = np.zeros(self.nv, np.float32)
        ref().ref_get_factors(self.h, _p(theta, f32p), _p(phi, f32p), _p(bu, f32p), _p(bv, f32p))
        return theta, phi, bu, bv

    def get_old(self):
        theta = np.zeros((self.nu, self.dim), np.float32)
        phi = np.zeros((self.nv, self.dim), np.float32)
        bu = np.zeros(self.nu, np.float32)
        bv = np.zeros(self.nv, np.float32)
        ref().ref_get_old(self.h, _p(theta, f32p), _p(phi, f32p), _p(bu, f32p), _p(bv, f32p))
        return theta, phi, bu, bv

    def seteta(self, rnd):
        ref().ref_seteta(self.h, rnd)

    @property
    def eta(self):
        return ref().ref_get_eta(self.h)

    def epoch(self):
        ref().ref_epoch(self.h)

    def sse(self, which):
        n = C.c_int(0)
        s = ref().ref_calc_mse(self.h, which, C.byref(n))
        return s, n.value
