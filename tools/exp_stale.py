"""GPU experiment: what drives the test-RMSE gap of the parallel schedule?  Medium shape (120k x 17,770,
25M ratings, k=128); max_groups fixes the number of sub-warps W; item popularity Zipf(s) with s = 1
(hot rows) or 0 (flat); eta0 0.02 or 0.01.  The serial oracle runs on the host for each data set."""
import ctypes as C, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "experimental-mf_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import mfb200 as mb, oraclelib as ol
GB = 2.76
nu, nv, nnz, k, EPOCHS = 120000, 17770, 25_000_000, 128, 6
for zipf in (1.0, 0.0):
    p = mb.gen_params(nu, nv, nnz); p.zipf_s = zipf
    tr, te, _ = mb.generate(p)
    train = ol.Dataset(tr.block_off, tr.run_uid, tr.run_off, tr.vid, tr.rating)
    test = ol.Dataset(te.block_off, te.run_uid, te.run_off, te.vid, te.rating)
    top = np.bincount(tr.vid, minlength=nv).max() / tr.nratings
    for eta0 in (0.02, 0.01):
        m = ol.Model(nu, nv, k, seed=11); th, ph = m.dense()
        res = {}
        for W, ring in [(420, 1), (1680, 1), (1680, 4), (6720, 1), (6720, 4), (16000, 1)]:
            c = mb.Context(nu, nv, k); c.set_factors(th, ph, m.bu, m.bv)
            c.set_option("max_groups", W); c.set_option("ring", ring)
            dtr, dte = c.dataset_from_blocks(tr), c.dataset_from_blocks(te)
            traj, ms = [], []
            for ep in range(1, EPOCHS + 1):
                c.sgd_epoch(dtr, mb.seteta(eta0, ep, 1.0), 5e-3, GB, mb.MODE_ATOMIC); ms.append(c.last_kernel_ms()); traj.append(c.rmse(dte, GB))
            res[(W, ring)] = (traj, min(ms))
            c.close()
        mm, dd, tt = m.as_mfo(), train.as_mfo(), test.as_mfo()
        want = []
        for ep in range(1, EPOCHS + 1):
            ol.oracle().mfo_sgd_epoch(C.byref(mm), C.byref(dd), mb.seteta(eta0, ep, 1.0), 5e-3, GB)
            n = C.c_int64(); s = ol.oracle().mfo_sse(C.byref(mm), C.byref(tt), GB, C.byref(n)); want.append(float(np.sqrt(s / n.value)))
        print("zipf %.1f (top item share %.4f) eta0 %.2f oracle rmse %s" % (zipf, top, eta0, " ".join("%.4f" % x for x in want)), flush=True)
        for (W, ring), (traj, ms) in res.items():
            print("   W %5d ring %d: %.2f ms  d(ep1) %+.5f  d(final) %+.5f" % (W, ring, ms, traj[0] - want[0], traj[-1] - want[-1]), flush=True)
