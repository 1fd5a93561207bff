import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "experimental-mf_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import mfb200 as mb, oraclelib as ol
from oraclelib import MfoDpState, MfoNoisePhilox, MfoNoiseTable, _p, f32p, u64p
from gpu_common import ctx_from_model, upload_ds, row_rel_err, vec_rel_err
L = ol.oracle(); GB = 2.76
def run(dim, eps, temp, use_table, lam_r=1.3):
    nu, nv = 160, 70
    train, _, _ = ol.make_ratings(nu, nv, 4000, seed=dim)
    m = ol.Model(nu, nv, dim, seed=4, scale=0.1)
    c = ctx_from_model(m); c.enable(2); d = upload_ds(c, train); ntrain = c.dp_weights(d)
    eta = np.float32(2e-2 / ntrain); temp = np.float32(temp)
    bound = mb.lib().mfb_dp_bound(eps, 0, nv)
    lam = (np.random.default_rng(1).uniform(0.5, 2.0, 2 * dim)).astype(np.float32)
    c.upload(mb.LAMBDA_U, lam[:dim]); c.upload(mb.LAMBDA_V, lam[dim:])
    ur, vr = c.download(mb.UR), c.download(mb.VR)
    lu, lv = lam[:dim].copy(), lam[dim:].copy()
    gcu, gcv = np.zeros(nu, np.uint64), np.zeros(nv, np.uint64)
    st = MfoDpState(eta, temp, bound, ntrain, lam_r, 0.7, 0.9, _p(lu, f32p), _p(lv, f32p), _p(ur, f32p), _p(vr, f32p), 0, _p(gcu, u64p), _p(gcv, u64p))
    mm, dd = m.as_mfo(), train.as_mfo()
    table = np.random.default_rng(2).standard_normal(nv * (dim + 1) + 5000).astype(np.float32)
    c.set_noise_table(table)
    if use_table:
        ctx = MfoNoiseTable(_p(table, f32p), len(table), 77); fn = C.cast(L.mfo_noise_from_table, C.c_void_p)
    else:
        ctx = MfoNoisePhilox(0xABCDEF0123, 1); fn = C.cast(L.mfo_noise_from_philox, C.c_void_p)
    p = mb.SgldParams(eta, temp, bound, ntrain, lam_r, 0.7, 0.9, 0xABCDEF0123, 1, int(use_table), 77)
    c.sgld_epoch(d, p, GB, mb.MODE_ORDERED)
    L.mfo_sgld_epoch(C.byref(mm), C.byref(dd), C.byref(st), GB, fn, C.byref(ctx))
    th, ph, bu, bv = c.get_factors()
    print("dim %3d eps %.1f temp %.2f table %d lam_r %.1f: theta %.2e phi %.2e bu %.2e bv %.2e | inf: ur %d vr %d" % (
        dim, eps, temp, use_table, lam_r, row_rel_err(th, m.theta[:, :dim]), row_rel_err(ph, m.phi[:, :dim]),
        vec_rel_err(bu, m.bu), vec_rel_err(bv, m.bv), np.isinf(ur).sum(), np.isinf(vr).sum()), flush=True)
    c.close()
for dim in (16, 32, 50):
    for eps in (0.0, 0.5):
        run(dim, eps, 0.0, 1); run(dim, eps, 0.5, 1); run(dim, eps, 0.5, 0); run(dim, eps, 0.5, 0, lam_r=0.0)
