"""dpmf path (SGLD / DP) on the GPU through the C ABI, against the reference's golden outputs, the
CPU oracle, and the noise statistics the reference's lazy-noise scheme guarantees."""
import ctypes as C

import numpy as np
import pytest

import mfb200 as mb
import oraclelib as ol
from gpu_common import ctx_from_model, model_rel_err, row_rel_err, upload_ds, vec_rel_err
from oraclelib import MfoDpState, MfoNoisePhilox, MfoNoiseTable, _p, f32p, u64p
from test_oracle_golden import init_model, load_ds

pytestmark = pytest.mark.gpu
GB = 2.76


def dp_ctx(m, train):
    c = ctx_from_model(m)
    c.enable(2)  # before finalize: builds the static logical clock
    d = upload_ds(c, train)
    ntrain = c.dp_weights(d)
    return c, d, ntrain


def test_ordered_sgld_with_table_noise_reproduces_reference_golden_bit_exact(golden):
    """The reference's SgldFilter + finish_noise with its noise_ table at a fixed offset
    (tests/golden, epsilon > 0) vs the CUDA ordered mode reading the same table."""
    g = golden
    m, train, test = init_model(g), load_ds(g, "train"), load_ds(g, "test")
    dim = m.dim
    eta0, gam, eps, temp, mineta, ha, hb = [np.float32(x) for x in g["dp_params"]]
    tau, off, _ = [int(x) for x in g["dp_tau_off_seed"]]
    c, d, ntrain = dp_ctx(m, train)
    dte = upload_ds(c, test)
    assert ntrain == int(g["dp_ntrain_tau"][0])
    np.testing.assert_array_equal(c.download(mb.UR), g["dp_ur"])
    np.testing.assert_array_equal(c.download(mb.VR), g["dp_vr"])
    bound = mb.lib().mfb_dp_bound(eps, tau, m.nv)
    assert np.float32(bound) == g["dp_bound"]
    c.set_noise_table(g["dp_table"])
    hyp = np.r_[1.0, 1e2, 1e2, np.full(2 * dim, 1e2)].astype(np.float32)  # model.h:42, model.cc:226
    eta = eta0
    for ep in (1, 2, 3):
        assert np.float32(eta) == g["dp_eta_%d" % ep]
        c.upload(mb.LAMBDA_U, hyp[3:3 + dim])
        c.upload(mb.LAMBDA_V, hyp[3 + dim:])
        p = mb.SgldParams(eta, temp, bound, ntrain, hyp[0], hyp[1], hyp[2], 0, ep, 1, off)
        c.sgld_epoch(d, p, GB, mb.MODE_ORDERED)
        for which, key in ((mb.THETA, "theta"), (mb.PHI, "phi"), (mb.BU, "bu"), (mb.BV, "bv")):
            np.testing.assert_array_equal(c.download(which), g["dp_pre_%s_%d" % (key, ep)])
        c.sgld_flush_noise(d, p)
        for which, key in ((mb.THETA, "theta"), (mb.PHI, "phi"), (mb.BU, "bu"), (mb.BV, "bv")):
            np.testing.assert_array_equal(c.download(which), g["dp_%s_%d" % (key, ep)])
        s_tr, _ = c.sse(d, GB)
        s_te, _ = c.sse(dte, GB)
        assert abs(s_tr - float(g["dp_train_sse_%d" % ep])) <= 1e-5 * s_tr
        assert abs(s_te - float(g["dp_test_sse_%d" % ep])) <= 1e-5 * s_te
        # the Gibbs step is host code; feed the reference's draws so that the kernels stay comparable
        hyp = g["dp_hyper_%d" % ep]
        eta = mb.lib().mfb_seteta_cutoff(eta0, ep + 1, gam, mineta)
    # K7 reductions against numpy on the final factors
    nu_, nv_, bu2, bv2 = c.col_sqnorms()
    th, ph, bu, bv = [x.astype(np.float64) for x in c.get_factors()]
    np.testing.assert_allclose(nu_, (th ** 2).sum(0), rtol=1e-6)
    np.testing.assert_allclose(nv_, (ph ** 2).sum(0), rtol=1e-6)
    np.testing.assert_allclose([bu2, bv2], [(bu ** 2).sum(), (bv ** 2).sum()], rtol=1e-6)
    c.close()


def _sgld_pair(dim, eps, temp, lam_r, use_table, seed=0xABCDEF0123):
    """one ordered SGLD epoch + flush on the GPU and in the oracle with the same noise source"""
    L = ol.oracle()
    nu, nv = 160, 70
    train, _, _ = ol.make_ratings(nu, nv, 4000, seed=dim)
    m = ol.Model(nu, nv, dim, seed=4, scale=0.1)
    c, d, ntrain = dp_ctx(m, train)
    eta, temp = np.float32(2e-2 / ntrain), np.float32(temp)
    bound = mb.lib().mfb_dp_bound(eps, 0, nv)
    lam = (np.random.default_rng(1).uniform(0.5, 2.0, 2 * dim)).astype(np.float32)
    c.upload(mb.LAMBDA_U, lam[:dim])
    c.upload(mb.LAMBDA_V, lam[dim:])
    ur, vr = c.download(mb.UR), c.download(mb.VR)
    lu, lv = lam[:dim].copy(), lam[dim:].copy()
    gcu, gcv = np.zeros(nu, np.uint64), np.zeros(nv, np.uint64)
    st = MfoDpState(eta, temp, bound, ntrain, lam_r, 0.7, 0.9, _p(lu, f32p), _p(lv, f32p), _p(ur, f32p),
                    _p(vr, f32p), 0, _p(gcu, u64p), _p(gcv, u64p))
    mm, dd = m.as_mfo(), train.as_mfo()
    table = np.random.default_rng(2).standard_normal(nv * (dim + 1) + 5000).astype(np.float32)
    c.set_noise_table(table)
    errs = []
    for ep in (1, 2):
        if use_table:
            ctx, fn = MfoNoiseTable(_p(table, f32p), len(table), 77), C.cast(L.mfo_noise_from_table, C.c_void_p)
        else:
            ctx, fn = MfoNoisePhilox(seed, ep), C.cast(L.mfo_noise_from_philox, C.c_void_p)
        p = mb.SgldParams(eta, temp, bound, ntrain, lam_r, 0.7, 0.9, seed, ep, int(use_table), 77)
        c.sgld_epoch(d, p, GB, mb.MODE_ORDERED)
        L.mfo_sgld_epoch(C.byref(mm), C.byref(dd), C.byref(st), GB, fn, C.byref(ctx))
        errs.append(model_rel_err(c, m))
        c.sgld_flush_noise(d, p)
        L.mfo_finish_noise(C.byref(mm), C.byref(st), fn, C.byref(ctx))
        errs.append(model_rel_err(c, m))
    c.close()
    return max(errs)


@pytest.mark.parametrize("dim", [16, 32, 50, 128])
@pytest.mark.parametrize("eps", [0.0, 0.5])
def test_ordered_sgld_bit_exact_vs_oracle_with_table_noise(dim, eps):
    """full dynamics (eps = 0 -> bound = 1) and the DP-scaled ones, the reference's table noise"""
    assert _sgld_pair(dim, eps, 0.5, 1.3, True) == 0.0


@pytest.mark.parametrize("dim", [16, 50, 128])
def test_ordered_sgld_with_philox_noise_matches_oracle(dim):
    """Same counter-based noise stream on both sides (oracle: mfo_noise_from_philox).  The normals
    go through logf/sinf/cosf of two different libms (1e-7 apart).  With the drift off they stay
    1e-6 apart; with the DP-scaled drift 1e-5.  (With bound = 1 the bilinear dynamics amplifies a
    1e-7 perturbation to 1e-2 within one epoch - measured - so that case is pinned by the
    table-noise test above, which is bit-exact.)"""
    assert _sgld_pair(dim, 0.0, 0.5, 0.0, False) <= 1e-6
    assert _sgld_pair(dim, 0.5, 0.5, 1.3, False) <= 1e-5


@pytest.mark.parametrize("flat", [1, 0], ids=["sub-warp kernel", "warp-per-run kernel"])
@pytest.mark.parametrize("dim", [20, 32, 64, 100, 128])
def test_parallel_sgld_on_conflict_free_data_equals_the_serial_oracle(dim, flat):
    """Production schedule of the dpmf epoch (both kernels: sgld_flat_kernel and the warp-per-run one) on data where
    the order cannot matter - every item rated once, one run per user - against the serial oracle with the same
    Philox noise stream.  With little noise (temp = 1e-4) the rows agree to 1e-5; with the noise dominating the rows
    (temp = 0.5: every item is touched once, so its lazy noise covers the whole epoch) to 2e-4, the accuracy of the
    MUFU lg2 / sin / cos the production schedule builds its normals with (the ordered schedule uses libm).  The two
    kernels agree with each other far more closely than either does with libm (measured 2e-8)."""
    L = ol.oracle()
    rng = np.random.default_rng(dim)
    nruns, per = 700, 23
    n = nruns * per
    lens = rng.integers(1, 2 * per, nruns)
    lens[-1] += n - lens.sum() if lens.sum() < n else 0
    off = np.r_[0, np.cumsum(lens)]
    n = int(off[-1])
    train = ol.Dataset(np.r_[np.arange(0, nruns, 50), nruns], rng.permutation(nruns), off, rng.permutation(n),
                       rng.integers(1, 6, n))
    nu, nv = nruns, n
    for eps, lam_r, tmp, tol in ((0.0, 0.0, 1e-4, 1e-5), (0.5, 1.3, 1e-4, 1e-5), (0.5, 1.3, 0.5, 2e-4)):
        m = ol.Model(nu, nv, dim, seed=4, scale=0.1)
        c, d, ntrain = dp_ctx(m, train)
        c.set_option("sgld_flat", flat)
        eta, temp = np.float32(2e-2 / ntrain), np.float32(tmp)
        bound = mb.lib().mfb_dp_bound(eps, 0, nv)
        lam = (np.random.default_rng(1).uniform(0.5, 2.0, 2 * dim)).astype(np.float32)
        c.upload(mb.LAMBDA_U, lam[:dim])
        c.upload(mb.LAMBDA_V, lam[dim:])
        ur, vr = c.download(mb.UR), c.download(mb.VR)
        lu, lv = lam[:dim].copy(), lam[dim:].copy()
        gcu, gcv = np.zeros(nu, np.uint64), np.zeros(nv, np.uint64)
        st = MfoDpState(eta, temp, bound, ntrain, lam_r, 0.7, 0.9, _p(lu, f32p), _p(lv, f32p), _p(ur, f32p),
                        _p(vr, f32p), 0, _p(gcu, u64p), _p(gcv, u64p))
        mm, dd = m.as_mfo(), train.as_mfo()
        fn = C.cast(L.mfo_noise_from_philox, C.c_void_p)
        for ep in (1, 2):
            ctx = MfoNoisePhilox(99, ep)
            p = mb.SgldParams(eta, temp, bound, ntrain, lam_r, 0.7, 0.9, 99, ep, 0, 0)
            c.sgld_epoch(d, p, GB, mb.MODE_HOGWILD)
            L.mfo_sgld_epoch(C.byref(mm), C.byref(dd), C.byref(st), GB, fn, C.byref(ctx))
            assert model_rel_err(c, m) <= tol, (eps, lam_r, ep, model_rel_err(c, m))
            c.sgld_flush_noise(d, p)
            L.mfo_finish_noise(C.byref(mm), C.byref(st), fn, C.byref(ctx))
        c.close()


@pytest.mark.parametrize("mode", [mb.MODE_ORDERED, mb.MODE_HOGWILD])
def test_noise_statistics_per_epoch_invariant(mode):
    """with the drift switched off (lambda_r = lambda_* = 0) one epoch + flush adds to
    EVERY coordinate of EVERY row - touched often, once or never - independent N(0, temp*eta*ntrain)
    noise.  Checked on mean, variance, kurtosis and cross-coordinate correlation."""
    nu, nv, dim = 3000, 400, 32
    tr, _, _ = mb.generate(mb.gen_params(nu, nv, 150_000, test_frac=0.0, users_per_block=100))
    m = ol.Model(nu + 50, nv + 20, dim, seed=1, scale=0.0)  # 50 users / 20 items never rated
    c = ctx_from_model(m)
    c.enable(2)
    d = c.dataset_from_blocks(tr)
    ntrain = c.dp_weights(d)
    # ur/vr are inf for unseen rows (as in the reference); harmless here because they multiply 0
    c.upload(mb.UR, np.ones(m.nu, np.float32))
    c.upload(mb.VR, np.ones(m.nv, np.float32))
    eta, temp = np.float32(1e-6), np.float32(0.7)
    p = mb.SgldParams(eta, temp, 1.0, ntrain, 0.0, 0.0, 0.0, 2024, 1, 0, 0)
    c.sgld_epoch(d, p, GB, mode)
    c.sgld_flush_noise(d, p)
    th, ph, bu, bv = c.get_factors()
    var = float(temp * eta * ntrain)
    for name, a in (("theta", th), ("phi", ph), ("bu", bu.reshape(-1, 1)), ("bv", bv.reshape(-1, 1))):
        z = a.astype(np.float64) / np.sqrt(var)
        n = z.size
        assert abs(z.mean()) < 5 / np.sqrt(n), name
        tol = 0.02
        assert abs(z.var() - 1) < tol + 4 * np.sqrt(2 / n), (name, z.var())
        assert abs((z ** 4).mean() - 3) < 0.3 + 10 / np.sqrt(n), name
    zt = th.astype(np.float64) / np.sqrt(var)
    corr = np.corrcoef(zt.T)
    assert np.abs(corr - np.eye(dim)).max() < 0.1
    # never-rated rows get the same variance (all of it from the flush)
    assert abs(zt[nu:].var() - 1) < 0.15
    # per-row variance does not depend on how often the row was touched
    cnt = np.bincount(tr.vid, minlength=m.nv)
    zp = ph.astype(np.float64) / np.sqrt(var)
    hot, cold = zp[np.argsort(cnt)[-50:]], zp[np.argsort(cnt)[:50]]
    print("phi noise variance: hot items %.3f cold items %.3f (mode %d)" % (hot.var(), cold.var(), mode))
    assert abs(hot.var() - 1) < 0.15 and abs(cold.var() - 1) < 0.15
    c.close()


def test_hogwild_sgld_rmse_close_to_serial_oracle():
    """Posterior-sample RMSE of the Hogwild SGLD kernel vs the serial oracle on the same data with the
    same Philox noise stream (different interleaving => statistical agreement only)."""
    L = ol.oracle()
    nu, nv, dim = 6040, 3706, 32
    tr, te, _ = mb.generate(mb.gen_params(nu, nv, 1_000_000, test_frac=0.1))
    train = ol.Dataset(tr.block_off, tr.run_uid, tr.run_off, tr.vid, tr.rating)
    test = ol.Dataset(te.block_off, te.run_uid, te.run_off, te.vid, te.rating)
    m = ol.Model(nu, nv, dim, seed=11)
    c = ctx_from_model(m)
    c.enable(2)
    d, dte = c.dataset_from_blocks(tr), c.dataset_from_blocks(te)
    ntrain = c.dp_weights(d)
    eta0, temp, gam = np.float32(2e-2 / ntrain), np.float32(0.01), 0.5
    lam = np.full(dim, 2.0, np.float32)
    c.upload(mb.LAMBDA_U, lam)
    c.upload(mb.LAMBDA_V, lam)
    ur, vr = c.download(mb.UR), c.download(mb.VR)
    gcu, gcv = np.zeros(nu, np.uint64), np.zeros(nv, np.uint64)
    lu, lv = lam.copy(), lam.copy()
    st = MfoDpState(eta0, temp, 1.0, ntrain, 1.0, 2.0, 2.0, _p(lu, f32p), _p(lv, f32p), _p(ur, f32p),
                    _p(vr, f32p), 0, _p(gcu, u64p), _p(gcv, u64p))
    mm, dd, tt = m.as_mfo(), train.as_mfo(), test.as_mfo()
    fn = C.cast(L.mfo_noise_from_philox, C.c_void_p)
    got, want = [], []
    for ep in range(1, 7):
        eta = mb.lib().mfb_seteta_cutoff(eta0, ep, gam, 1e-13)
        p = mb.SgldParams(eta, temp, 1.0, ntrain, 1.0, 2.0, 2.0, 7, ep, 0, 0)
        c.sgld_epoch(d, p, GB, mb.MODE_HOGWILD)
        c.sgld_flush_noise(d, p)
        got.append(c.rmse(dte, GB))
        st.eta = eta
        ph = MfoNoisePhilox(7, ep)
        L.mfo_sgld_epoch(C.byref(mm), C.byref(dd), C.byref(st), GB, fn, C.byref(ph))
        L.mfo_finish_noise(C.byref(mm), C.byref(st), fn, C.byref(ph))
        n = C.c_int64()
        s = L.mfo_sse(C.byref(mm), C.byref(tt), GB, C.byref(n))
        want.append(float(np.sqrt(s / n.value)))
    print("sgld oracle ", ["%.4f" % x for x in want])
    print("sgld hogwild", ["%.4f" % x for x in got])
    assert want[-1] < want[0]
    assert abs(got[-1] - want[-1]) < 1e-3   # north_star's bound on the final test RMSE (measured: 1e-4)
    c.close()
