"""GPU experiment (one GPU): the placement search (tune_placement, option placement_trials) at the
Netflix shape.  K contexts one after the other (kept alive, so each starts from different physical
pages): calibration times of the candidates, the one kept, and the epoch times that follow; the last
context runs with the search off for comparison."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "experimental-mf_b200"))
import mfb200 as mb
GB, k, K = 2.76, int(os.environ.get("DIM", "128")), int(os.environ.get("K", "4"))
TRIALS = int(os.environ.get("TRIALS", "16"))
nu, nv, nnz = 480189, 17770, 100_000_000
tr, te, _ = mb.generate(mb.gen_params(nu, nv, nnz))
keep = []
for i in range(K + 1):
    c = mb.Context(nu, nv, k); c.init_normal(0x4D46B200, 1e-2)
    c.set_option("placement_trials", TRIALS if i < K else 0)
    d, dte = c.dataset_from_blocks(tr), c.dataset_from_blocks(te)
    ms = []
    c.sync(); t0 = time.perf_counter()
    for ep in range(1, 7):
        c.sgd_epoch(d, mb.seteta(2e-2, ep, 1.0), 5e-3, GB, mb.MODE_ATOMIC); c.sync(); ms.append(c.last_kernel_ms())
        if ep == 1:
            first_wall = 1e3 * (time.perf_counter() - t0)
    cal, best = c.placement_report()
    print("context %d (trials %d): first epoch wall %.1f ms; epochs ms %s; tRMSE %.4f" % (
        i, TRIALS if i < K else 0, first_wall, " ".join("%.2f" % x for x in ms), c.rmse(dte, GB)), flush=True)
    if cal:
        print("   calibration ms %s -> kept %d (%.2f)" % (" ".join("%.2f" % x for x in cal), best, cal[best]), flush=True)
    keep.append((c, d))
