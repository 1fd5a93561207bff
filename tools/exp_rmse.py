"""GPU experiment: test-RMSE trajectories of the parallel schedules vs the serial CPU oracle as a
function of the concurrency bound, plus kernel time.  Medium Netflix-like shape so that the
oracle finishes in ~30 s."""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "experimental-mf_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import mfb200 as mb  # noqa: E402
import oraclelib as ol  # noqa: E402

nu, nv, nnz, k = [int(x) for x in os.environ.get("SHAPE", "120000,17770,25000000,128").split(",")]
EPOCHS = int(os.environ.get("EPOCHS", 8))
GB = 2.76
tr, te, _ = mb.generate(mb.gen_params(nu, nv, nnz))
train = ol.Dataset(tr.block_off, tr.run_uid, tr.run_off, tr.vid, tr.rating)
test = ol.Dataset(te.block_off, te.run_uid, te.run_off, te.vid, te.rating)
m = ol.Model(nu, nv, k, seed=11)
th, ph = m.dense()
cnt = np.bincount(tr.vid, minlength=nv)
print("shape", nu, nv, tr.nratings, k, "top item share %.4f" % (cnt.max() / tr.nratings), flush=True)

configs = []
for mode, name in ((mb.MODE_HOGWILD, "hogwild"), (mb.MODE_ATOMIC, "atomic")):
    for rc in [int(x) for x in os.environ.get("RC", "2,8,32,0").split(",")]:
        configs.append((name, mode, rc))
res = {}
for name, mode, rc in configs:
    c = mb.Context(nu, nv, k)
    c.set_factors(th, ph, m.bu, m.bv)
    c.set_option("row_concurrency", rc)
    c.set_option("memopt", int(os.environ.get("MEMOPT", 0)))
    dtr, dte = c.dataset_from_blocks(tr), c.dataset_from_blocks(te)
    traj, ms = [], []
    for ep in range(1, EPOCHS + 1):
        c.sgd_epoch(dtr, mb.seteta(2e-2, ep, 1.0), 5e-3, GB, mode)
        ms.append(c.last_kernel_ms())
        traj.append(c.rmse(dte, GB))
    res[(name, rc)] = traj
    print("%-8s rc %3d  %.2f ms/epoch %.2f Gupd/s  rmse %s" % (
        name, rc, min(ms), tr.nratings / min(ms) / 1e6, " ".join("%.4f" % x for x in traj)), flush=True)
    c.close()
t0 = time.time()
mm, dd, tt = m.as_mfo(), train.as_mfo(), test.as_mfo()
want = []
for ep in range(1, EPOCHS + 1):
    ol.oracle().mfo_sgd_epoch(C.byref(mm), C.byref(dd), mb.seteta(2e-2, ep, 1.0), 5e-3, GB)
    n = C.c_int64()
    s = ol.oracle().mfo_sse(C.byref(mm), C.byref(tt), GB, C.byref(n))
    want.append(float(np.sqrt(s / n.value)))
print("oracle   (serial, %.0fs)            rmse %s" % (time.time() - t0, " ".join("%.4f" % x for x in want)), flush=True)
for key, traj in res.items():
    print("%-8s rc %3d  final |d rmse| = %.5f" % (key[0], key[1], abs(traj[-1] - want[-1])))
