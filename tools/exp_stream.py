"""GPU experiment: streaming sub-warp kernel (kernel 3) vs the warp-per-run kernels on the Netflix
shape (time per epoch, epochs 1..4) and on a medium shape (test RMSE vs the serial oracle)."""
import ctypes as C, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "experimental-mf_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import mfb200 as mb, oraclelib as ol
GB = 2.76
def run(c, dtr, dte, n, epochs=4):
    ms = []
    for ep in range(1, epochs + 1):
        c.sgd_epoch(dtr, mb.seteta(2e-2, ep, 1.0), 5e-3, GB, mb.MODE_ATOMIC); ms.append(c.last_kernel_ms())
    return ms, c.rmse(dte, GB)
def timing(full=True):
    nu, nv, nnz, k = 480189, 17770, 100_000_000, 128
    tr, te, _ = mb.generate(mb.gen_params(nu, nv, nnz))
    c = mb.Context(nu, nv, k)
    dtr, dte = c.dataset_from_blocks(tr), c.dataset_from_blocks(te)
    cfgs = []
    for thr in (0, 1):
        for rc in (24, 48, 96):
            for ring in (1, 2):
                cfgs.append((3, ring, rc, 1, thr))
    cfgs += [(3, 1, 6, 1, 1), (3, 1, 12, 1, 1)]
    for kern, ring, rc, esc, thr in cfgs:
        c.set_option("kernel", kern); c.set_option("row_concurrency", rc); c.set_option("eta_scaling", esc); c.set_option("throttle", thr)
        c.set_option("ring", ring)
        c.init_normal(1, 1e-2)
        ms, rmse = run(c, dtr, dte, tr.nratings)
        print("full kernel %d ring %d rc %3d eta_scaling %d throttle %d: ms %s  best %.2f Gupd/s  rmse(4 ep) %.4f" % (
            kern, ring, rc, esc, thr, " ".join("%.2f" % x for x in ms), tr.nratings / min(ms) / 1e6, rmse), flush=True)
    c.close()
def accuracy():
    nu, nv, nnz, k, EPOCHS = 120000, 17770, 25_000_000, 128, 8
    tr, te, _ = mb.generate(mb.gen_params(nu, nv, nnz))
    train = ol.Dataset(tr.block_off, tr.run_uid, tr.run_off, tr.vid, tr.rating)
    test = ol.Dataset(te.block_off, te.run_uid, te.run_off, te.vid, te.rating)
    m = ol.Model(nu, nv, k, seed=11); th, ph = m.dense()
    res = {}
    for kern, ring, rc, esc in [(3, 1, 32, 1), (3, 1, 32, 2), (3, 1, 64, 2), (3, 2, 32, 2)]:
        c = mb.Context(nu, nv, k); c.set_factors(th, ph, m.bu, m.bv)
        c.set_option("kernel", kern); c.set_option("row_concurrency", rc); c.set_option("ring", ring); c.set_option("eta_scaling", esc)
        dtr, dte = c.dataset_from_blocks(tr), c.dataset_from_blocks(te)
        traj, ms = [], []
        if len(sys.argv) > 2: c.set_option("run_fraction_ppm", int(sys.argv[2]))
        for ep in range(1, EPOCHS + 1):
            c.sgd_epoch(dtr, mb.seteta(2e-2, ep, 1.0), 5e-3, GB, mb.MODE_ATOMIC); ms.append(c.last_kernel_ms()); traj.append(c.rmse(dte, GB))
        res[(kern, ring, rc, esc)] = traj
        print("med kernel %d ring %d rc %3d esc %d: ms %s rmse %s" % (kern, ring, rc, esc, " ".join("%.2f" % x for x in ms), " ".join("%.4f" % x for x in traj)), flush=True)
        c.close()
    mm, dd, tt = m.as_mfo(), train.as_mfo(), test.as_mfo()
    want = []
    for ep in range(1, EPOCHS + 1):
        ol.oracle().mfo_sgd_epoch(C.byref(mm), C.byref(dd), mb.seteta(2e-2, ep, 1.0), 5e-3, GB)
        n = C.c_int64(); s = ol.oracle().mfo_sse(C.byref(mm), C.byref(tt), GB, C.byref(n)); want.append(float(np.sqrt(s / n.value)))
    print("oracle rmse", " ".join("%.4f" % x for x in want))
    for key, traj in res.items():
        print("%s final |d rmse| = %.5f  max over epochs %.5f" % (key, abs(traj[-1] - want[-1]), max(abs(a - b) for a, b in zip(traj, want))))
if __name__ == "__main__":
    what = sys.argv[1:] or ["timing", "accuracy"]
    if "timing" in what: timing()
    if "accuracy" in what: accuracy()
