"""GPU experiment: dpmf epoch at the Netflix shape, sub-warp kernel (sgld_flat = 1) against the warp-per-run one (0),
k = 128 (C3) and k = 64 (C4): ms per epoch, test RMSE after 4 rounds with the flush."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "experimental-mf_b200"))
import mfb200 as mb
GB = 2.76
nu, nv, nnz = 480189, 17770, 100_000_000
tr, te, _ = mb.generate(mb.gen_params(nu, nv, nnz, test_frac=0.01))
KS = [int(x) for x in os.environ.get("KS", "128,64").split(",")]
FLATS = [int(x) for x in os.environ.get("FLATS", "1,0").split(",")]
for k in KS:
    for flat in FLATS:
        c = mb.Context(nu, nv, k); c.init_normal(1, 1e-2); c.enable(2)
        c.set_option("sgld_flat", flat)
        d, dte = c.dataset_from_blocks(tr), c.dataset_from_blocks(te)
        ntrain = c.dp_weights(d)
        lam = np.full(k, 2.0, np.float32); c.upload(mb.LAMBDA_U, lam); c.upload(mb.LAMBDA_V, lam)
        eta0, temp = np.float32(2e-2 / ntrain), np.float32(0.01)
        ms, rm = [], []
        for ep in range(1, 6):
            eta = mb.lib().mfb_seteta_cutoff(eta0, ep, 0.5, 1e-13)
            p = mb.SgldParams(eta, temp, 1.0, ntrain, 1.0, 2.0, 2.0, 7, ep, 0, 0)
            c.sgld_epoch(d, p, GB, mb.MODE_HOGWILD); ms.append(c.last_kernel_ms())
            c.sgld_flush_noise(d, p); rm.append(c.rmse(dte, GB))
        print("k %d flat %d: ms/epoch %s | tRMSE %s | launch %s" % (k, flat, " ".join("%.1f" % x for x in ms), " ".join("%.4f" % x for x in rm), c.last_launch()), flush=True)
        c.close()
