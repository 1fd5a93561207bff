"""GPU experiment: dpmf (SGLD / DP) and admf epochs at the Netflix shape (BASELINE configs 3 and 4):
kernel time per epoch, updates/s, test RMSE trajectory, lambda trajectory."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "experimental-mf_b200"))
import mfb200 as mb
GB = 2.76
nu, nv, nnz = 480189, 17770, 100_000_000
tr, te, va = mb.generate(mb.gen_params(nu, nv, nnz, valid_frac=0.01))
n = tr.nratings
for k, eps in ((128, 0.0), (64, 1.0)):   # config 3: pure SGLD k=128; config 4: DP eps>0 k=64
    c = mb.Context(nu, nv, k); c.init_normal(1, 1e-2); c.enable(2)
    d, dte = c.dataset_from_blocks(tr), c.dataset_from_blocks(te)
    ntrain = c.dp_weights(d)
    bound = mb.lib().mfb_dp_bound(eps, 0, nv)
    lam = np.full(k, 1e2, np.float32); c.upload(mb.LAMBDA_U, lam); c.upload(mb.LAMBDA_V, lam)
    # effective step scal = eta*ntrain*bound = 0.02 as in plain SGD; noise variance per epoch and
    # coordinate temp*eta*ntrain = 0.002 (the reference's sweep: temp 0.1, eta of order 1/ntrain, run.py:22,34)
    eta0, temp, gam = np.float32(2e-2 / ntrain / bound), np.float32(0.1 * bound), 1.0
    ms, fl, traj = [], [], []
    for ep in range(1, 5):
        eta = mb.lib().mfb_seteta_cutoff(eta0, ep, gam, 1e-13)
        p = mb.SgldParams(eta, temp, bound, ntrain, 1.0, 1e2, 1e2, 7, ep, 0, 0)
        c.sgld_epoch(d, p, GB, mb.MODE_HOGWILD); ms.append(c.last_kernel_ms())
        c.sgld_flush_noise(d, p); fl.append(c.last_kernel_ms())
        traj.append(c.rmse(dte, GB))
    print("dpmf k=%d eps=%g bound=%.3g: epoch ms %s (%.2f G upd/s)  flush ms %.2f  rmse %s" % (
        k, eps, bound, " ".join("%.1f" % x for x in ms), n / min(ms) / 1e6, min(fl), " ".join("%.4f" % x for x in traj)), flush=True)
    c.close()
k = 64
c = mb.Context(nu, nv, k); c.init_normal(1, 1e-2); c.enable(1); c.snapshot_old()
d, dte = c.dataset_from_blocks(tr), c.dataset_from_blocks(te)
vu = np.repeat(va.run_uid, np.diff(va.run_off)).astype(np.int32)
c.admf_set_validation(vu, va.vid, va.rating); c.admf_set_lams([5e-3] * 4)
rng = np.random.default_rng(0)
ms, traj, lams = [], [], []
for ep in range(1, 5):
    c.admf_set_draws(rng.integers(0, len(vu), tr.nruns).astype(np.int32))
    c.admf_epoch(d, mb.seteta(2e-2, ep, 1.0), mb.seteta(2e-2, ep, 1.0), 0, GB, mb.MODE_ATOMIC); ms.append(c.last_kernel_ms())
    traj.append(c.rmse(dte, GB)); lams.append(c.admf_get_lams())
print("admf k=%d: epoch ms %s (%.2f G upd/s) rmse %s lams %s" % (k, " ".join("%.1f" % x for x in ms), n / min(ms) / 1e6,
      " ".join("%.4f" % x for x in traj), np.round(lams[-1], 5)), flush=True)
