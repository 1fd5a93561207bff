"""Generates tests/golden/ref_small.npz from the REFERENCE ITSELF.

Run in the build container (needs /root/reference):  python tests/golden/make_golden.py

It drives the reference's own sources, compiled in place into oracle/_ref/libmf_ref.so by
oracle/Makefile (TBB/MKL/protobuf shimmed, -ffp-contract=off), through the harness in
oracle/ref_harness.cc: SgdFilter / SgldFilter / AdRegFilter one block at a time in file order
(`--fly 1` order), MF::calc_mse, DPMF::finish_noise / sample_hyper, AdaptRegMF::updateReg.
The reference's clock-seeded init is overwritten through its public arrays with the seeded model
stored in the fixture.  Inputs and outputs are stored so that boxes without /root/reference
(the GPU box) can still check the oracle and the CUDA path against reference outputs.
"""
import ctypes as C
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oraclelib as ol  # noqa: E402
from oraclelib import _p, f32p, i32p  # noqa: E402

NU, NV, DIM = 120, 60, 20
GB = 2.76
EPOCHS = 3


def ds_arrays(prefix, ds, out):
    out[prefix + "_block_off"] = ds.block_off
    out[prefix + "_run_uid"] = ds.run_uid
    out[prefix + "_run_off"] = ds.run_off
    out[prefix + "_vid"] = ds.vid
    out[prefix + "_rating"] = ds.rating


def main():
    ol.build_oracle()
    assert ol.have_ref(), "oracle/_ref/libmf_ref.so missing: /root/reference is needed"
    R = ol.ref()
    out = {}
    train, test, valid = ol.make_ratings(NU, NV, 4000, seed=3, valid_frac=0.05)
    d = tempfile.mkdtemp()
    tp, sp, vp = train.write(d + "/train"), test.write(d + "/test"), valid.write(d + "/valid")
    ds_arrays("train", train, out)
    ds_arrays("test", test, out)
    ds_arrays("valid", valid, out)
    out["train_bytes"] = np.fromfile(tp, dtype=np.uint8)
    out["shape"] = np.array([NU, NV, DIM], np.int32)
    out["gb"] = np.float32(GB)

    m = ol.Model(NU, NV, DIM, seed=5)
    th, ph = m.dense()
    out["init_theta"], out["init_phi"], out["init_bu"], out["init_bv"] = th, ph, m.bu, m.bv

    # ---- mf: 3 epochs of SgdFilter in file order, eta from MF::seteta --------------------------
    eta0, gam, lam = 2e-2, 1.0, 5e-3
    out["sgd_params"] = np.array([eta0, gam, lam], np.float32)
    r = ol.Ref(R.ref_create_mf(tp.encode(), sp.encode(), DIM, eta0, gam, lam, GB, NU, NV), NU, NV, DIM)
    r.set_factors(th, ph, m.bu, m.bv)
    for ep in range(1, EPOCHS + 1):
        r.seteta(ep)
        out["sgd_eta_%d" % ep] = np.float32(r.eta)
        r.epoch()
        t, p, bu, bv = r.get_factors()
        out["sgd_theta_%d" % ep], out["sgd_phi_%d" % ep] = t, p
        out["sgd_bu_%d" % ep], out["sgd_bv_%d" % ep] = bu, bv
        s, n = r.sse(1)
        out["sgd_test_sse_%d" % ep] = np.float32(s)
        out["sgd_test_n_%d" % ep] = np.int64(n)

    # ---- dpmf: SgldFilter + finish_noise + sample_hyper; table noise at a fixed offset ---------
    ntr = train.nratings
    eps, tau, temp, mineta, ha, hb = 0.5, 0, 0.3, 1e-13, 1.0, 100.0
    eta0d, gamd = np.float32(2e-2 / ntr), 0.6
    noise_size = NV * (DIM + 1) + 20000
    off = 137
    table = np.random.default_rng(9).standard_normal(noise_size).astype(np.float32)
    out["dp_params"] = np.array([eta0d, gamd, eps, temp, mineta, ha, hb], np.float32)
    out["dp_tau_off_seed"] = np.array([tau, off, 77], np.int32)
    out["dp_table"] = table
    h = R.ref_create_dpmf(tp.encode(), sp.encode(), DIM, eta0d, gamd, 5e-3, GB, NU, NV, ha, hb, eps,
                          tau, noise_size, temp, mineta)
    r = ol.Ref(h, NU, NV, DIM)
    r.set_factors(th, ph, m.bu, m.bv)
    R.ref_dpmf_set_noise(h, _p(table, f32p), noise_size)
    R.ref_dpmf_set_offset(h, off)
    nt, bd, ta = C.c_int(), C.c_float(), C.c_int()
    R.ref_dpmf_info(h, C.byref(nt), C.byref(bd), C.byref(ta))
    out["dp_ntrain_tau"] = np.array([nt.value, ta.value], np.int32)
    out["dp_bound"] = np.float32(bd.value)
    ur, vr = np.zeros(NU, np.float32), np.zeros(NV, np.float32)
    R.ref_dpmf_get_weights(h, _p(ur, f32p), _p(vr, f32p))
    out["dp_ur"], out["dp_vr"] = ur, vr
    R.ref_srand(77)
    for ep in range(1, EPOCHS + 1):  # run(DPMF&), main.cc:69-72 + finish_round, model.cc:299-310
        out["dp_eta_%d" % ep] = np.float32(r.eta)
        r.epoch()
        t, p, bu, bv = r.get_factors()
        out["dp_pre_theta_%d" % ep], out["dp_pre_phi_%d" % ep] = t, p  # before finish_noise
        out["dp_pre_bu_%d" % ep], out["dp_pre_bv_%d" % ep] = bu, bv
        R.ref_dpmf_finish_noise(h)
        t, p, bu, bv = r.get_factors()
        out["dp_theta_%d" % ep], out["dp_phi_%d" % ep] = t, p
        out["dp_bu_%d" % ep], out["dp_bv_%d" % ep] = bu, bv
        s_tr, _ = r.sse(0)
        s_te, n_te = r.sse(1)
        out["dp_train_sse_%d" % ep], out["dp_test_sse_%d" % ep] = np.float32(s_tr), np.float32(s_te)
        R.ref_dpmf_sample_hyper(h, s_tr)
        hyp = np.zeros(3 + 2 * DIM, np.float32)
        R.ref_dpmf_get_hyper(h, _p(hyp, f32p))
        out["dp_hyper_%d" % ep] = hyp
        r.seteta(ep + 1)

    # ---- admf: AdRegFilter + updateReg, both link functions ------------------------------------
    eta_reg0 = 2e-3
    out["ad_params"] = np.array([eta0, gam, lam, eta_reg0], np.float32)
    out["ad_seed"] = np.int32(5)
    for loss in (0, 1):
        R.ref_srand(5)
        h = R.ref_create_admf(tp.encode(), sp.encode(), vp.encode(), DIM, eta0, gam, lam, GB, NU, NV,
                              loss, eta_reg0)
        r = ol.Ref(h, NU, NV, DIM)
        r.set_factors(th, ph, m.bu, m.bv)
        nval = R.ref_num_valid(h)
        vu, vv, vr_ = np.zeros(nval, np.int32), np.zeros(nval, np.int32), np.zeros(nval, np.float32)
        R.ref_get_valid(h, _p(vu, i32p), _p(vv, i32p), _p(vr_, f32p))
        out["ad_valid_u"], out["ad_valid_v"], out["ad_valid_r"] = vu, vv, vr_  # after the shuffle
        for ep in range(1, EPOCHS + 1):
            r.seteta(ep)
            r.epoch()
            l4 = np.zeros(4, np.float32)
            R.ref_admf_get_lams(h, _p(l4, f32p))
            k = "ad%d_" % loss
            out[k + "lams_%d" % ep] = l4
            t, p, bu, bv = r.get_factors()
            out[k + "theta_%d" % ep], out[k + "phi_%d" % ep] = t, p
            out[k + "bu_%d" % ep], out[k + "bv_%d" % ep] = bu, bv
            t, p, bu, bv = r.get_old()
            out[k + "theta_old_%d" % ep], out[k + "phi_old_%d" % ep] = t, p
            out[k + "bu_old_%d" % ep], out[k + "bv_old_%d" % ep] = bu, bv
            s, n = r.sse(1)
            out[k + "test_sse_%d" % ep] = np.float32(s)

    path = os.path.join(HERE, "ref_small.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
