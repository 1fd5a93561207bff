"""GPU experiment: admf (adaptive regulariser) epochs at the Netflix shape, k=64, for several
hot-row weights of the admf kernel and with/without the one-record-ahead row request:
kernel time per epoch, test RMSE and lambda trajectories (the first configuration is the
production one and the yardstick for the others)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "experimental-mf_b200"))
import mfb200 as mb
GB, k, EPOCHS = 2.76, 64, int(os.environ.get("EPOCHS", "5"))
nu, nv, nnz = 480189, 17770, 100_000_000
tr, te, va = mb.generate(mb.gen_params(nu, nv, nnz, valid_frac=0.01))
vu = np.repeat(va.run_uid, np.diff(va.run_off)).astype(np.int32)
configs = [tuple(int(x) for x in a.split(",")) for a in sys.argv[1:]] or [(6, 1), (3, 1), (2, 1), (2, 0), (1, 0)]
for weight, prefetch in configs:
    c = mb.Context(nu, nv, k); c.init_normal(1, 1e-2); c.enable(1); c.snapshot_old()
    c.set_option("admf_weight", weight); c.set_option("admf_prefetch", prefetch)
    d, dte = c.dataset_from_blocks(tr), c.dataset_from_blocks(te)
    c.admf_set_validation(vu, va.vid, va.rating); c.admf_set_lams([5e-3] * 4)
    rng = np.random.default_rng(0)
    ms, traj, lams, grids = [], [], [], []
    for ep in range(1, EPOCHS + 1):
        c.admf_set_draws(rng.integers(0, len(vu), tr.nruns).astype(np.int32))
        c.admf_epoch(d, mb.seteta(2e-2, ep, 1.0), mb.seteta(2e-2, ep, 1.0), 0, GB, mb.MODE_ATOMIC)
        ms.append(c.last_kernel_ms()); sh = c.last_launch(); grids.append(sh["grid"] * sh["threads"] // 32)
        traj.append(c.rmse(dte, GB)); lams.append(c.admf_get_lams())
    print("weight %d prefetch %d: ms %s | warps %s | rmse %s | lam_u %s | lam_bu %s" % (
        weight, prefetch, " ".join("%.1f" % x for x in ms), " ".join(str(g) for g in grids),
        " ".join("%.4f" % x for x in traj), " ".join("%.2e" % l[0] for l in lams),
        " ".join("%.3f" % l[2] for l in lams)), flush=True)
    c.close()
