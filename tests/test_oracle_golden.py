"""The plain-C oracle (oracle/mf_oracle.c) against outputs of the reference itself
(tests/golden/ref_small.npz, made by tests/golden/make_golden.py from /root/reference's own
sources).  Bit-exact: both sides are sequential fp32 with -ffp-contract=off."""
import ctypes as C

import numpy as np

import oraclelib as ol
from oraclelib import (Dataset, MfoAdState, MfoDpState, MfoNoiseTable, Model, _p, f32p, i32p, i64p,
                       u64p)


def load_ds(g, prefix):
    return Dataset(g[prefix + "_block_off"], g[prefix + "_run_uid"], g[prefix + "_run_off"],
                   g[prefix + "_vid"], g[prefix + "_rating"])


def init_model(g):
    nu, nv, dim = [int(x) for x in g["shape"]]
    m = Model(nu, nv, dim)
    m.theta[:] = 0
    m.phi[:] = 0
    m.set_dense(g["init_theta"], g["init_phi"], g["init_bu"], g["init_bv"])
    return m


def assert_model_equal(m, g, key, ep):
    d = m.dim
    np.testing.assert_array_equal(m.theta[:, :d], g["%stheta_%d" % (key, ep)])
    np.testing.assert_array_equal(m.phi[:, :d], g["%sphi_%d" % (key, ep)])
    np.testing.assert_array_equal(m.bu, g["%sbu_%d" % (key, ep)])
    np.testing.assert_array_equal(m.bv, g["%sbv_%d" % (key, ep)])
    assert not m.theta[:, d:].any() and not m.phi[:, d:].any()  # padding columns stay zero


def test_padding_kat(oracle_lib):
    # util.h:163-165 (SURVEY 8c KAT)
    assert [oracle_lib.mfo_padding(d) for d in (16, 32, 64, 100, 128, 1, 17, 20)] == \
        [16, 32, 64, 112, 128, 16, 32, 32]


def test_sgd_epochs_match_reference(oracle_lib, golden):
    g = golden
    m, train, test = init_model(g), load_ds(g, "train"), load_ds(g, "test")
    eta0, gam, lam = [float(x) for x in g["sgd_params"]]
    gb = float(g["gb"])
    mm, dd, tt = m.as_mfo(), train.as_mfo(), test.as_mfo()
    for ep in (1, 2, 3):
        eta = oracle_lib.mfo_seteta(eta0, ep, gam)
        assert np.float32(eta) == g["sgd_eta_%d" % ep]
        oracle_lib.mfo_sgd_epoch(C.byref(mm), C.byref(dd), eta, lam, gb)
        assert_model_equal(m, g, "sgd_", ep)
        n = C.c_int64()
        s = oracle_lib.mfo_sse(C.byref(mm), C.byref(tt), gb, C.byref(n))
        assert np.float32(s) == g["sgd_test_sse_%d" % ep] and n.value == int(g["sgd_test_n_%d" % ep])


def test_sgld_dp_epochs_match_reference(oracle_lib, golden):
    g = golden
    L = oracle_lib
    m, train, test = init_model(g), load_ds(g, "train"), load_ds(g, "test")
    nu, nv, dim = m.nu, m.nv, m.dim
    eta0, gam, eps, temp, mineta, ha, hb = [np.float32(x) for x in g["dp_params"]]
    tau, off, seed = [int(x) for x in g["dp_tau_off_seed"]]
    gb = float(g["gb"])
    mm, dd, tt = m.as_mfo(), train.as_mfo(), test.as_mfo()
    ur, vr = np.zeros(nu, np.float32), np.zeros(nv, np.float32)
    ntrain = L.mfo_dp_weights(C.byref(dd), nu, nv, _p(ur, f32p), _p(vr, f32p))
    assert ntrain == int(g["dp_ntrain_tau"][0])
    np.testing.assert_array_equal(ur, g["dp_ur"])
    np.testing.assert_array_equal(vr, g["dp_vr"])
    tau_eff = int(g["dp_ntrain_tau"][1])
    assert tau_eff == (tau if tau > 0 else nv)  # model.cc:239
    bound = L.mfo_dp_bound(eps, tau_eff)
    assert np.float32(bound) == g["dp_bound"]
    lu, lv = np.full(dim, 1e2, np.float32), np.full(dim, 1e2, np.float32)  # model.cc:226
    gcu, gcv = np.zeros(nu, np.uint64), np.zeros(nv, np.uint64)
    st = MfoDpState(eta0, temp, bound, ntrain, 1.0, 1e2, 1e2, _p(lu, f32p), _p(lv, f32p),
                    _p(ur, f32p), _p(vr, f32p), 0, _p(gcu, u64p), _p(gcv, u64p))
    table = np.ascontiguousarray(g["dp_table"])
    ntab = MfoNoiseTable(_p(table, f32p), len(table), off)
    fn = C.cast(L.mfo_noise_from_table, C.c_void_p)
    L.mfo_srand(seed)
    for ep in (1, 2, 3):
        assert np.float32(st.eta) == g["dp_eta_%d" % ep]
        L.mfo_sgld_epoch(C.byref(mm), C.byref(dd), C.byref(st), gb, fn, C.byref(ntab))
        assert_model_equal(m, g, "dp_pre_", ep)
        L.mfo_finish_noise(C.byref(mm), C.byref(st), fn, C.byref(ntab))
        assert_model_equal(m, g, "dp_", ep)
        assert st.gcount == 0 and not gcu.any() and not gcv.any()
        n = C.c_int64()
        s_tr = L.mfo_sse(C.byref(mm), C.byref(dd), gb, C.byref(n))
        s_te = L.mfo_sse(C.byref(mm), C.byref(tt), gb, C.byref(n))
        assert np.float32(s_tr) == g["dp_train_sse_%d" % ep]
        assert np.float32(s_te) == g["dp_test_sse_%d" % ep]
        L.mfo_sample_hyper(C.byref(mm), C.byref(st), ha, hb, s_tr)
        hyp = np.r_[st.lambda_r, st.lambda_ub, st.lambda_vb, lu, lv].astype(np.float32)
        np.testing.assert_array_equal(hyp, g["dp_hyper_%d" % ep])
        st.eta = L.mfo_seteta_cutoff(eta0, ep + 1, gam, mineta)


def test_admf_epochs_and_lambda_trajectory_match_reference(oracle_lib, golden):
    g = golden
    L = oracle_lib
    train, valid = load_ds(g, "train"), load_ds(g, "valid")
    eta0, gam, lam, eta_reg0 = [np.float32(x) for x in g["ad_params"]]
    gb = float(g["gb"])
    # AdaptRegMF::plain_read_valid (model.cc:390-415): flatten in file order, random_shuffle
    vu, vv, vr = valid.uid_per_rating().copy(), valid.vid.copy(), valid.rating.copy()
    L.mfo_srand(int(g["ad_seed"]))
    L.mfo_shuffle_valid(len(vu), _p(vu, i32p), _p(vv, i32p), _p(vr, f32p))
    np.testing.assert_array_equal(vu, g["ad_valid_u"])
    np.testing.assert_array_equal(vv, g["ad_valid_v"])
    np.testing.assert_array_equal(vr, g["ad_valid_r"])
    for loss in (0, 1):
        m = init_model(g)
        L.mfo_srand(int(g["ad_seed"]))
        tu, tv, tr = vu.copy(), vv.copy(), vr.copy()
        L.mfo_shuffle_valid(len(tu), _p(tu, i32p), _p(tv, i32p), _p(tr, f32p))  # same rand() draws
        tho, pho, buo, bvo = m.theta.copy(), m.phi.copy(), m.bu.copy(), m.bv.copy()
        st = MfoAdState(eta0, eta_reg0, loss, lam, lam, lam, lam, _p(tho, f32p), _p(pho, f32p),
                        _p(buo, f32p), _p(bvo, f32p), len(vu), _p(vu, i32p), _p(vv, i32p),
                        _p(vr, f32p), None, 0)
        mm, dd = m.as_mfo(), train.as_mfo()
        k = "ad%d_" % loss
        for ep in (1, 2, 3):
            st.eta = L.mfo_seteta(eta0, ep, gam)
            st.eta_reg = L.mfo_seteta(eta_reg0, ep, gam)  # model.cc:386-388
            L.mfo_admf_epoch(C.byref(mm), C.byref(dd), C.byref(st), gb)
            lams = np.array([st.lam_u, st.lam_v, st.lam_bu, st.lam_bv], np.float32)
            np.testing.assert_array_equal(lams, g[k + "lams_%d" % ep])
            assert_model_equal(m, g, k, ep)
            d = m.dim
            np.testing.assert_array_equal(tho[:, :d], g[k + "theta_old_%d" % ep])
            np.testing.assert_array_equal(pho[:, :d], g[k + "phi_old_%d" % ep])
            np.testing.assert_array_equal(buo, g[k + "bu_old_%d" % ep])
            np.testing.assert_array_equal(bvo, g[k + "bv_old_%d" % ep])


def test_wire_reader_matches_reference_bytes(oracle_lib, golden, tmp_path):
    # the oracle's reader on the exact bytes the reference build parsed
    p = tmp_path / "train.bin"
    golden["train_bytes"].tofile(p)
    ds = Dataset.read(str(p))
    t = load_ds(golden, "train")
    for a in ("block_off", "run_uid", "run_off", "vid", "rating"):
        np.testing.assert_array_equal(getattr(ds, a), getattr(t, a))
    # and the writer reproduces them
    q = tmp_path / "again.bin"
    ds.write(str(q))
    assert q.read_bytes() == p.read_bytes()


def test_link_aware_evaluation_matches_numpy(oracle_lib):
    """calc_mse with the logistic link of --loss 1 (util.h:90-95) applied to the prediction, vs numpy;
    link 0 equals MF::calc_mse up to the order of the additions."""
    import ctypes as C
    nu, nv, dim = 90, 40, 24
    train, test, _ = ol.make_ratings(nu, nv, 3000, seed=4)
    m = ol.Model(nu, nv, dim, seed=8, scale=0.4)
    mm, tt = m.as_mfo(), test.as_mfo()
    n = C.c_int64()
    u = test.uid_per_rating()
    pred = np.einsum("ij,ij->i", m.theta[u, :dim].astype(np.float64), m.phi[test.vid, :dim].astype(np.float64)) \
        + m.bu[u] + m.bv[test.vid] + 2.76
    r01 = (test.rating > 3).astype(np.float32)          # a 0/1 target, what the logistic link is for
    t01 = ol.Dataset(test.block_off, test.run_uid, test.run_off, test.vid, r01)
    s1 = oracle_lib.mfo_sse_link(C.byref(mm), C.byref(t01.as_mfo()), 2.76, 1, C.byref(n))
    want1 = float(((r01 - 1.0 / (1.0 + np.exp(-pred))) ** 2).sum())
    assert n.value == test.nratings and abs(s1 - want1) <= 1e-5 * want1
    s0 = oracle_lib.mfo_sse_link(C.byref(mm), C.byref(tt), 2.76, 0, C.byref(n))
    ref0 = oracle_lib.mfo_sse(C.byref(mm), C.byref(tt), 2.76, C.byref(n))
    assert abs(s0 - ref0) <= 1e-5 * ref0
