"""Two dpmf (SGLD) epochs at the Netflix shape, k from argv (default 128): the command ncu captures
for profiles/r1_sgld_*.  Prints the kernel time and launch shape."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "experimental-mf_b200"))
import mfb200 as mb
GB = 2.76
k = int(sys.argv[1]) if len(sys.argv) > 1 else 128
nu, nv, nnz = 480189, 17770, 100_000_000
tr, te, _ = mb.generate(mb.gen_params(nu, nv, nnz))
c = mb.Context(nu, nv, k); c.init_normal(1, 1e-2); c.enable(2)
d, dte = c.dataset_from_blocks(tr), c.dataset_from_blocks(te)
ntrain = c.dp_weights(d)
lam = np.full(k, 1e2, np.float32); c.upload(mb.LAMBDA_U, lam); c.upload(mb.LAMBDA_V, lam)
eta0, temp = np.float32(2e-2 / ntrain), np.float32(0.1)
for ep in (1, 2):
    eta = mb.lib().mfb_seteta_cutoff(eta0, ep, 1.0, 1e-13)
    p = mb.SgldParams(eta, temp, 1.0, ntrain, 1.0, 1e2, 1e2, 7, ep, 0, 0)
    c.sgld_epoch(d, p, GB, mb.MODE_HOGWILD)
    ms = c.last_kernel_ms()
    c.sgld_flush_noise(d, p)
    print("dpmf k=%d epoch %d: %.2f ms (%.2f G upd/s) launch %s rmse %.4f" % (
        k, ep, ms, tr.nratings / ms / 1e6, c.last_launch(), c.rmse(dte, GB)), flush=True)
c.close()
