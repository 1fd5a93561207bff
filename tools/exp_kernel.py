"""GPU experiment: kernel variant x row_concurrency on the full Netflix shape (time) and a medium
shape (accuracy vs the serial oracle)."""
import ctypes as C, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "experimental-mf_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import mfb200 as mb, oraclelib as ol
GB = 2.76
def timing():
    nu, nv, nnz, k = 480189, 17770, 100_000_000, 128
    tr, te, _ = mb.generate(mb.gen_params(nu, nv, nnz))
    c = mb.Context(nu, nv, k); c.init_normal(1, 1e-2)
    dtr, dte = c.dataset_from_blocks(tr), c.dataset_from_blocks(te)
    for kern in (1, 2):
        for rc in (8, 16, 32):
            c.set_option("kernel", kern); c.set_option("row_concurrency", rc)
            c.init_normal(1, 1e-2)
            ms = []
            for ep in range(1, 5):
                c.sgd_epoch(dtr, mb.seteta(2e-2, ep, 1.0), 5e-3, GB, mb.MODE_ATOMIC); ms.append(c.last_kernel_ms())
            print("full  kernel %d rc %2d: %.2f ms  %.2f Gupd/s  rmse(4 ep) %.4f" % (kern, rc, min(ms), tr.nratings / min(ms) / 1e6, c.rmse(dte, GB)), flush=True)
    c.close()
def accuracy():
    nu, nv, nnz, k, EPOCHS = 120000, 17770, 25_000_000, 128, 8
    tr, te, _ = mb.generate(mb.gen_params(nu, nv, nnz))
    train = ol.Dataset(tr.block_off, tr.run_uid, tr.run_off, tr.vid, tr.rating)
    test = ol.Dataset(te.block_off, te.run_uid, te.run_off, te.vid, te.rating)
    m = ol.Model(nu, nv, k, seed=11); th, ph = m.dense()
    res = {}
    for kern in (1, 2):
        for rc in (8, 12, 16, 24):
            c = mb.Context(nu, nv, k); c.set_factors(th, ph, m.bu, m.bv)
            c.set_option("kernel", kern); c.set_option("row_concurrency", rc)
            dtr, dte = c.dataset_from_blocks(tr), c.dataset_from_blocks(te)
            traj, ms = [], []
            for ep in range(1, EPOCHS + 1):
                c.sgd_epoch(dtr, mb.seteta(2e-2, ep, 1.0), 5e-3, GB, mb.MODE_ATOMIC); ms.append(c.last_kernel_ms()); traj.append(c.rmse(dte, GB))
            res[(kern, rc)] = traj
            print("med   kernel %d rc %2d: %.2f ms %.2f Gupd/s rmse %s" % (kern, rc, min(ms), tr.nratings / min(ms) / 1e6, " ".join("%.4f" % x for x in traj)), flush=True)
            c.close()
    mm, dd, tt = m.as_mfo(), train.as_mfo(), test.as_mfo()
    want = []
    for ep in range(1, EPOCHS + 1):
        ol.oracle().mfo_sgd_epoch(C.byref(mm), C.byref(dd), mb.seteta(2e-2, ep, 1.0), 5e-3, GB)
        n = C.c_int64(); s = ol.oracle().mfo_sse(C.byref(mm), C.byref(tt), GB, C.byref(n)); want.append(float(np.sqrt(s / n.value)))
    print("oracle rmse", " ".join("%.4f" % x for x in want))
    for key, traj in res.items():
        print("kernel %d rc %2d final |d rmse| = %.5f  max over epochs %.5f" % (key[0], key[1], abs(traj[-1] - want[-1]), max(abs(a - b) for a, b in zip(traj, want))))
timing()
accuracy()
