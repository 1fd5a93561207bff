"""GPU experiment: ML-1M shape (k=32): final test RMSE vs the serial oracle for the kernels / claim sizes."""
import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "experimental-mf_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import mfb200 as mb, oraclelib as ol
GB = 2.76
nu, nv, dim, epochs = 6040, 3706, 32, 10
tr, te, _ = mb.generate(mb.gen_params(nu, nv, 1_000_000, test_frac=0.1))
train = ol.Dataset(tr.block_off, tr.run_uid, tr.run_off, tr.vid, tr.rating)
test = ol.Dataset(te.block_off, te.run_uid, te.run_off, te.vid, te.rating)
m = ol.Model(nu, nv, dim, seed=11); th, ph = m.dense()
print("top item share %.4f, runs %d" % (np.bincount(tr.vid).max() / tr.nratings, tr.nruns))
res = {}
for name, opts in [("stream", {"kernel": 3}), ("burst span 32", {"kernel": 4, "span_runs": 32}), ("burst span 8", {"kernel": 4, "span_runs": 8}),
                   ("burst span 2", {"kernel": 4, "span_runs": 2}), ("burst span 8 W 42", {"kernel": 4, "span_runs": 8, "max_groups": 42}),
                   ("burst span 8 W 21", {"kernel": 4, "span_runs": 8, "max_groups": 21}),
                   ("burst batch 8", {"kernel": 4, "batch": 8}), ("burst depth 2", {"kernel": 4, "depth": 2}),
                   ("burst batch 8 depth 2", {"kernel": 4, "batch": 8, "depth": 2}), ("default", {})]:
    c = mb.Context(nu, nv, dim); c.set_factors(th, ph, m.bu, m.bv)
    for k, v in opts.items(): c.set_option(k, v)
    dtr, dte = c.dataset_from_blocks(tr), c.dataset_from_blocks(te)
    traj, ms = [], []
    for ep in range(1, epochs + 1):
        c.sgd_epoch(dtr, mb.seteta(2e-2, ep, 1.0), 5e-3, GB, mb.MODE_ATOMIC); ms.append(c.last_kernel_ms()); traj.append(c.rmse(dte, GB))
    res[name] = (traj, min(ms), c.last_launch()); c.close()
mm, dd, tt = m.as_mfo(), train.as_mfo(), test.as_mfo()
want = []
for ep in range(1, epochs + 1):
    ol.oracle().mfo_sgd_epoch(C.byref(mm), C.byref(dd), mb.seteta(2e-2, ep, 1.0), 5e-3, GB)
    n = C.c_int64(); s = ol.oracle().mfo_sse(C.byref(mm), C.byref(tt), GB, C.byref(n)); want.append(float(np.sqrt(s / n.value)))
for name, (traj, ms, ls) in res.items():
    print("%-20s %.3f ms  d(final) %+.5f  max|d| %.5f  %s" % (name, ms, traj[-1] - want[-1], max(abs(a - b) for a, b in zip(traj, want)), ls))
