"""Live pin of the C oracle against the reference's own sources (oracle/_ref/, built in place from
/root/reference by oracle/Makefile).  Skipped where /root/reference was never available (the GPU
box runs test_oracle_golden.py against the committed outputs of this same reference build)."""
import ctypes as C
import re
import struct
import subprocess

import numpy as np
import pytest

import oraclelib as ol
from oraclelib import (Dataset, MfoAdState, MfoDpState, MfoNoiseTable, Model, _p, f32p, i32p, u64p)

pytestmark = pytest.mark.skipif(not ol.have_ref(), reason="oracle/_ref not built (no /root/reference)")

GB = 2.76


def files(tmp_path, nu, nv, nnz, seed, valid_frac=0.0):
    train, test, valid = ol.make_ratings(nu, nv, nnz, seed=seed, valid_frac=valid_frac)
    tp, sp = train.write(str(tmp_path / "train")), test.write(str(tmp_path / "test"))
    vp = valid.write(str(tmp_path / "valid")) if valid is not None else None
    return train, test, valid, tp, sp, vp


def maxdiff(ref_factors, m):
    d = m.dim
    return max(float(np.abs(a - b).max()) for a, b in
               zip(ref_factors, (m.theta[:, :d], m.phi[:, :d], m.bu, m.bv)))


@pytest.mark.parametrize("nu,nv,dim,seed", [(90, 40, 16, 1), (150, 80, 33, 2), (64, 300, 128, 3)])
def test_sgd_bit_exact(oracle_lib, tmp_path, nu, nv, dim, seed):
    L, R = oracle_lib, ol.ref()
    train, test, _, tp, sp, _ = files(tmp_path, nu, nv, 3000, seed)
    m = Model(nu, nv, dim, seed=seed)
    eta0, gam, lam = 3e-2, 0.8, 1e-2
    r = ol.Ref(R.ref_create_mf(tp.encode(), sp.encode(), dim, eta0, gam, lam, GB, nu, nv), nu, nv, dim)
    r.set_factors(*m.dense(), m.bu, m.bv)
    mm, dd, tt = m.as_mfo(), train.as_mfo(), test.as_mfo()
    for ep in (1, 2):
        r.seteta(ep)
        eta = L.mfo_seteta(eta0, ep, gam)
        assert np.float32(eta) == np.float32(r.eta)
        r.epoch()
        L.mfo_sgd_epoch(C.byref(mm), C.byref(dd), eta, lam, GB)
        assert maxdiff(r.get_factors(), m) == 0.0
        n = C.c_int64()
        s = L.mfo_sse(C.byref(mm), C.byref(tt), GB, C.byref(n))
        assert (np.float32(s), n.value) == (np.float32(r.sse(1)[0]), r.sse(1)[1])


@pytest.mark.parametrize("eps", [0.0, 0.7])
def test_sgld_bit_exact(oracle_lib, tmp_path, eps):
    L, R = oracle_lib, ol.ref()
    nu, nv, dim = 100, 50, 24
    train, test, _, tp, sp, _ = files(tmp_path, nu, nv, 3000, 11)
    m = Model(nu, nv, dim, seed=4)
    ntr = train.nratings
    eta0, gam, temp, mineta, ha, hb = np.float32(3e-2 / ntr), 0.5, 0.5, 1e-13, 1.0, 100.0
    noise_size = nv * (dim + 1) + 15000
    h = R.ref_create_dpmf(tp.encode(), sp.encode(), dim, eta0, gam, 5e-3, GB, nu, nv, ha, hb, eps, 0,
                          noise_size, temp, mineta)
    r = ol.Ref(h, nu, nv, dim)
    r.set_factors(*m.dense(), m.bu, m.bv)
    table = np.random.default_rng(1).standard_normal(noise_size).astype(np.float32)
    off = 4242
    R.ref_dpmf_set_noise(h, _p(table, f32p), noise_size)
    R.ref_dpmf_set_offset(h, off)
    mm, dd = m.as_mfo(), train.as_mfo()
    ur, vr = np.zeros(nu, np.float32), np.zeros(nv, np.float32)
    ntrain = L.mfo_dp_weights(C.byref(dd), nu, nv, _p(ur, f32p), _p(vr, f32p))
    bound = L.mfo_dp_bound(eps, nv)
    lu, lv = np.full(dim, 1e2, np.float32), np.full(dim, 1e2, np.float32)
    gcu, gcv = np.zeros(nu, np.uint64), np.zeros(nv, np.uint64)
    st = MfoDpState(eta0, temp, bound, ntrain, 1.0, 1e2, 1e2, _p(lu, f32p), _p(lv, f32p),
                    _p(ur, f32p), _p(vr, f32p), 0, _p(gcu, u64p), _p(gcv, u64p))
    ntab = MfoNoiseTable(_p(table, f32p), noise_size, off)
    fn = C.cast(L.mfo_noise_from_table, C.c_void_p)
    for ep in (1, 2):
        r.epoch()
        L.mfo_sgld_epoch(C.byref(mm), C.byref(dd), C.byref(st), GB, fn, C.byref(ntab))
        assert maxdiff(r.get_factors(), m) == 0.0
        R.ref_dpmf_finish_noise(h)
        L.mfo_finish_noise(C.byref(mm), C.byref(st), fn, C.byref(ntab))
        assert maxdiff(r.get_factors(), m) == 0.0
        s_tr, _ = r.sse(0)
        R.ref_srand(100 + ep)
        R.ref_dpmf_sample_hyper(h, s_tr)
        L.mfo_srand(100 + ep)
        L.mfo_sample_hyper(C.byref(mm), C.byref(st), ha, hb, s_tr)
        hyp = np.zeros(3 + 2 * dim, np.float32)
        R.ref_dpmf_get_hyper(h, _p(hyp, f32p))
        np.testing.assert_array_equal(hyp, np.r_[st.lambda_r, st.lambda_ub, st.lambda_vb, lu, lv].astype(np.float32))
        r.seteta(ep + 1)
        st.eta = L.mfo_seteta_cutoff(eta0, ep + 1, gam, mineta)
        assert np.float32(st.eta) == np.float32(r.eta)


@pytest.mark.parametrize("loss", [0, 1])
def test_admf_bit_exact(oracle_lib, tmp_path, loss):
    L, R = oracle_lib, ol.ref()
    nu, nv, dim = 80, 70, 16
    train, test, valid, tp, sp, vp = files(tmp_path, nu, nv, 3000, 21, valid_frac=0.05)
    m = Model(nu, nv, dim, seed=8)
    eta0, gam, lam, eta_reg0 = 2e-2, 1.0, 5e-3, 5e-2
    R.ref_srand(9)
    h = R.ref_create_admf(tp.encode(), sp.encode(), vp.encode(), dim, eta0, gam, lam, GB, nu, nv, loss,
                          eta_reg0)
    r = ol.Ref(h, nu, nv, dim)
    r.set_factors(*m.dense(), m.bu, m.bv)
    want = []
    for ep in (1, 2):
        r.seteta(ep)
        r.epoch()
        l4 = np.zeros(4, np.float32)
        R.ref_admf_get_lams(h, _p(l4, f32p))
        want.append((l4, r.get_factors()))
    vu, vv, vr = valid.uid_per_rating().copy(), valid.vid.copy(), valid.rating.copy()
    L.mfo_srand(9)
    L.mfo_shuffle_valid(len(vu), _p(vu, i32p), _p(vv, i32p), _p(vr, f32p))
    tho, pho, buo, bvo = m.theta.copy(), m.phi.copy(), m.bu.copy(), m.bv.copy()
    st = MfoAdState(eta0, eta_reg0, loss, lam, lam, lam, lam, _p(tho, f32p), _p(pho, f32p),
                    _p(buo, f32p), _p(bvo, f32p), len(vu), _p(vu, i32p), _p(vv, i32p), _p(vr, f32p),
                    None, 0)
    mm, dd = m.as_mfo(), train.as_mfo()
    for ep in (1, 2):
        st.eta, st.eta_reg = L.mfo_seteta(eta0, ep, gam), L.mfo_seteta(eta_reg0, ep, gam)
        L.mfo_admf_epoch(C.byref(mm), C.byref(dd), C.byref(st), GB)
        l4, fac = want[ep - 1]
        np.testing.assert_array_equal(l4, np.array([st.lam_u, st.lam_v, st.lam_bu, st.lam_bv], np.float32))
        assert maxdiff(fac, m) == 0.0


def write_model_file(path, m, lam):
    """MF::save_model layout, model.cc:98-122: nv nu dim | lambda | bv | phi | bu | theta."""
    with open(path, "wb") as f:
        f.write(struct.pack("<iiif", m.nv, m.nu, m.dim, lam))
        f.write(m.bv.tobytes())
        f.write(np.ascontiguousarray(m.phi[:, :m.dim]).tobytes())
        f.write(m.bu.tobytes())
        f.write(np.ascontiguousarray(m.theta[:, :m.dim]).tobytes())


def test_reference_binary_prints_oracle_rmse(oracle_lib, tmp_path):
    """The reference's own main() (main.cc:95-164 -> run(MF&), --fly 1, --model for a seeded
    start) prints the tRMSE the oracle computes, epoch by epoch."""
    L = oracle_lib
    nu, nv, dim = 90, 40, 16
    train, test, _, tp, sp, _ = files(tmp_path, nu, nv, 3000, 5)
    m = Model(nu, nv, dim, seed=2)
    eta0, gam, lam = 2e-2, 1.0, 5e-3
    mp = str(tmp_path / "model")
    write_model_file(mp, m, lam)
    out = subprocess.run([ol.REF_BIN, "--alg", "mf", "--train", tp, "--test", sp, "--nu", str(nu),
                          "--nv", str(nv), "--dim", str(dim), "--iter", "3", "--fly", "1", "--eta",
                          str(eta0), "--lambda", str(lam), "--gam", str(gam), "--bias", str(GB),
                          "--model", mp], capture_output=True, text=True, check=True).stdout
    got = [float(x) for x in re.findall(r"tRMSE=([0-9.]+)", out)]
    assert len(got) == 3
    mm, dd, tt = m.as_mfo(), train.as_mfo(), test.as_mfo()
    for ep in (1, 2, 3):
        L.mfo_sgd_epoch(C.byref(mm), C.byref(dd), L.mfo_seteta(eta0, ep, gam), np.float32(lam), GB)
        n = C.c_int64()
        s = L.mfo_sse(C.byref(mm), C.byref(tt), GB, C.byref(n))
        assert abs(np.sqrt(s * 1.0 / n.value) - got[ep - 1]) < 2e-6
