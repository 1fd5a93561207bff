"""GPU experiment (one GPU): does the DSGD cell ORDER cost accuracy?  The Netflix-shaped file is cut
into P x P cells (user shard x item block) and one GPU walks the DSGD schedule serially (sub-epoch s:
cells (p, (p+s) mod P) for p = 0..P-1), with the default concurrency bounds per cell, optionally with
explicit options.  Test RMSE after each epoch for P = 1, 2, 4, 8."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "experimental-mf_b200"))
import mfb200 as mb, mfb_dsgd
GB = 2.76
nu, nv, nnz, k, EPOCHS = 480189, 17770, 100_000_000, 128, 10
opts = dict(kv.split("=") for kv in os.environ.get("MFB_OPTS", "").split(",") if kv)
Ps = [int(x) for x in (sys.argv[1:] or ["1", "2", "4", "8"])]
for P in Ps:
    c = mb.Context(nu, nv, k); c.init_normal(0x4D46B200, 1e-2)
    for name, val in opts.items(): c.set_option(name, int(val))
    bounds = mfb_dsgd.item_bounds(nv, P)
    cells, ntrain = [], 0
    te_all = None
    for p in range(P):
        u0, u1 = mfb_dsgd.user_range(nu, p, P)
        tr, te, _ = mb.generate(mb.gen_params(nu, nv, nnz, user_begin=u0, user_end=u1))
        parts = tr.split_by_item(bounds) if P > 1 else [tr]
        cells.append([c.dataset_from_blocks(b) for b in parts]); ntrain += tr.nratings
        dte = c.dataset_from_blocks(te)
        te_all = (te_all or []) + [dte]
    traj, ms = [], []
    for ep in range(1, EPOCHS + 1):
        t = 0.0
        for s in range(P):
            for p in range(P):
                c.sgd_epoch(cells[p][(p + s) % P], mb.seteta(2e-2, ep, 1.0), 5e-3, GB, mb.MODE_ATOMIC); t += c.last_kernel_ms()
        sse = n = 0
        for d in te_all:
            s_, n_ = c.sse(d, GB); sse += s_; n += n_
        traj.append(float(np.sqrt(sse / n))); ms.append(t)
    print("P %d opts %s: rmse %s | kernel ms/epoch %s" % (P, opts, " ".join("%.4f" % x for x in traj), " ".join("%.1f" % x for x in ms)), flush=True)
    c.close()
