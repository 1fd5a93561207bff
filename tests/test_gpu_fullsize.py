"""BASELINE.json configs[1] at FULL size (Netflix shape: 480,189 x 17,770, 100 M ratings, k=128) through
size-independent properties - the serial oracle would need minutes per epoch here:
every record is updated exactly once; two independently written kernels agree; eta = 0 is the identity;
the evaluation pass equals a numpy evaluation of the downloaded factors; the host-streamed epoch
(compact records, chunked) equals the resident one."""
import numpy as np
import pytest

import mfb200 as mb

pytestmark = pytest.mark.gpu
GB = 2.76
NU, NV, NNZ, K = 480_189, 17_770, 100_000_000, 128


@pytest.fixture(scope="module")
def data():
    tr, te, _ = mb.generate(mb.gen_params(NU, NV, NNZ))
    return tr, te


def chunk_count(run_off, base):
    """chunks of mfb_sgd_epoch_from_host: base, then x7/4 each up to 6x base, cut at run boundaries"""
    nruns, r0, k, want = len(run_off) - 1, 0, 0, base
    while r0 < nruns:
        r1 = int(np.searchsorted(run_off, run_off[r0] + want, side="right")) - 1
        r0 = max(r1, r0 + 1)
        k += 1
        want = min(want * 7 // 4, 6 * base)
    return k


def fresh(tr, te):
    c = mb.Context(NU, NV, K)
    c.init_normal(0x4D46B200, 1e-2)
    return c, c.dataset_from_blocks(tr), c.dataset_from_blocks(te)


def test_every_record_is_updated_exactly_once_and_eta_zero_is_identity(data):
    tr, te = data
    c, dtr, dte = fresh(tr, te)
    before = c.get_factors()
    c.probe_arm(0)
    c.set_option("kernel", 3)
    c.sgd_epoch(dtr, 0.0, 5e-3, GB, mb.MODE_ATOMIC)
    _, _, updates = c.probe_read()
    assert updates == tr.nratings
    c.probe_arm(-1)
    for kernel in (3, 4):
        c.set_option("kernel", kernel)
        c.sgd_epoch(dtr, 0.0, 5e-3, GB, mb.MODE_ATOMIC)
    for a, b in zip(before, c.get_factors()):
        np.testing.assert_array_equal(a, b)
    c.close()


def test_stream_and_burst_kernels_agree_and_learn(data):
    tr, te = data
    out = {}
    for kernel in (3, 4):
        c, dtr, dte = fresh(tr, te)
        c.set_option("kernel", kernel)
        traj = []
        for ep in (1, 2, 3):
            c.sgd_epoch(dtr, mb.seteta(2e-2, ep, 1.0), 5e-3, GB, mb.MODE_ATOMIC)
            traj.append(c.rmse(dte, GB))
        out[kernel] = traj
        c.close()
    print("stream", out[3], "burst", out[4])
    assert out[3][0] > out[3][1] > out[3][2] and out[3][2] < 0.62
    # the two kernels run with different numbers of user-runs in flight, which shows in the first epoch
    #  and fades: measured 7.5e-4 after epoch 1, 2e-4 at the end
    assert abs(out[3][0] - out[4][0]) <= 1.5e-3 and abs(out[3][-1] - out[4][-1]) <= 1e-3


def test_sse_pass_equals_numpy_on_downloaded_factors(data):
    tr, te = data
    c, dtr, dte = fresh(tr, te)
    c.sgd_epoch(dtr, 0.02, 5e-3, GB, mb.MODE_ATOMIC)
    s, n = c.sse(dte, GB)
    th, ph, bu, bv = c.get_factors()
    u = np.repeat(te.run_uid, np.diff(te.run_off))
    pred = np.einsum("ij,ij->i", th[u].astype(np.float64), ph[te.vid].astype(np.float64)) + bu[u] + bv[te.vid] + GB
    want = float(((te.rating - pred) ** 2).sum())
    assert n == te.nratings and abs(s - want) <= 1e-5 * want
    c.close()


def test_streamed_epoch_equals_resident_epoch_at_full_size(data):
    tr, te = data
    tr.pin()
    got = []
    for streamed in (False, True):
        c, dtr, dte = fresh(tr, te)
        for ep in (1, 2):
            eta = mb.seteta(2e-2, ep, 1.0)
            if streamed:
                c.sgd_epoch_from_host(dtr, tr, eta, 5e-3, GB, mb.MODE_ATOMIC, 0)
            else:
                c.sgd_epoch(dtr, eta, 5e-3, GB, mb.MODE_ATOMIC)
        got.append(c.rmse(dte, GB))
        if streamed:
            assert c.h2d_bytes() == 2 * (3 * tr.nratings + 8 * tr.nruns + 4 * chunk_count(tr.run_off, 3 << 20))
        c.close()
    tr.unpin()
    assert abs(got[0] - got[1]) <= 5e-4, got


def test_admf_at_full_size_eta_zero_is_identity_and_one_epoch_learns(data):
    """BASELINE configs[3] (adaptive regulariser, k=64) at full size through properties: with eta = 0
    every record is visited and nothing moves - factors, biases and the four regularisers come back
    bit for bit (the regularisers' step is eta_reg*eta); one real epoch then lowers the test RMSE
    below the value the plain-SGD epoch 1 reaches and keeps the regularisers finite and >= 0."""
    tr, te = data
    k = 64
    c = mb.Context(NU, NV, k)
    c.init_normal(0x4D46B200, 1e-2)
    c.enable(1)
    c.snapshot_old()
    dtr, dte = c.dataset_from_blocks(tr), c.dataset_from_blocks(te)
    vu = np.repeat(te.run_uid, np.diff(te.run_off)).astype(np.int32)   # the test file doubles as validation list
    c.admf_set_validation(vu, te.vid, te.rating)
    c.admf_set_lams([5e-3] * 4)
    rng = np.random.default_rng(0)
    c.admf_set_draws(rng.integers(0, len(vu), tr.nruns).astype(np.int32))
    before, rmse0 = c.get_factors(), c.rmse(dte, GB)
    c.admf_epoch(dtr, 0.0, 2e-2, 0, GB, mb.MODE_ATOMIC)
    assert c.last_kernel_ms() > 1.0
    for a, b in zip(before, c.get_factors()):
        np.testing.assert_array_equal(a, b)
    np.testing.assert_array_equal(np.asarray(c.admf_get_lams(), np.float32), np.full(4, 5e-3, np.float32))
    c.admf_epoch(dtr, 2e-2, 2e-2, 0, GB, mb.MODE_ATOMIC)
    rmse1, lams = c.rmse(dte, GB), np.asarray(c.admf_get_lams())
    assert rmse1 < 0.66 < rmse0, (rmse0, rmse1)
    assert np.all(np.isfinite(lams)) and np.all(lams >= 0) and lams[2] > 5e-3   # lam_bu grows on this data
    c.close()


def test_sgld_noise_invariant_at_full_size(data):
    """BASELINE configs[2] (SGLD, k=128, Philox noise) at full size through the invariant of SURVEY 8a3:
    with the drift off, one epoch + flush adds independent N(0, temp*eta*ntrain) noise to EVERY
    coordinate of EVERY row, however often the row is touched - 61 M user coordinates, 2.3 M item
    coordinates, the 20 most rated items (0.47 % of the records each) against the least rated ones."""
    tr, te = data
    k = 128
    c = mb.Context(NU, NV, k)          # factors, biases and lambda_u/lambda_v start at zero
    c.enable(2)
    d = c.dataset_from_blocks(tr)
    ntrain = c.dp_weights(d)
    assert ntrain == tr.nratings
    c.upload(mb.UR, np.ones(NU, np.float32))
    c.upload(mb.VR, np.ones(NV, np.float32))
    eta, temp = np.float32(1e-9), np.float32(0.7)
    p = mb.SgldParams(eta, temp, 1.0, ntrain, 0.0, 0.0, 0.0, 2024, 1, 0, 0)
    c.sgld_epoch(d, p, GB, mb.MODE_HOGWILD)
    c.sgld_flush_noise(d, p)
    th, ph, bu, bv = c.get_factors()
    c.close()
    sd = np.sqrt(float(temp) * float(eta) * ntrain)
    for name, a in (("theta", th), ("phi", ph), ("bu", bu), ("bv", bv)):
        z = a.astype(np.float64).ravel() / sd
        n = z.size
        assert abs(z.mean()) < 5 / np.sqrt(n), (name, z.mean())
        assert abs(z.var() - 1) < 0.01 + 4 * np.sqrt(2 / n), (name, z.var())
        assert abs((z ** 4).mean() - 3) < 0.1 + 10 / np.sqrt(n), (name, (z ** 4).mean())
    cnt = np.bincount(tr.vid, minlength=NV)
    order = np.argsort(cnt)
    hot, cold = ph[order[-20:]].astype(np.float64) / sd, ph[order[:200]].astype(np.float64) / sd
    print("phi noise variance at full size: 20 most rated items %.3f, 200 least rated %.3f" % (hot.var(), cold.var()))
    assert abs(hot.var() - 1) < 0.15 and abs(cold.var() - 1) < 0.05
