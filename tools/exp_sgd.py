"""GPU experiment harness (not part of the product or the tests): times the SGD epoch kernel under
different occupancy / schedule / data-skew settings.  Usage: python tools/exp_sgd.py [zipf ...]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "experimental-mf_b200"))
import mfb200 as mb  # noqa: E402

nu, nv, nnz, k = 480189, 17770, int(os.environ.get("NNZ", 100_000_000)), int(os.environ.get("K", 128))
for zipf in [float(x) for x in (sys.argv[1:] or ["1.0"])]:
    t0 = time.time()
    tr, te, _ = mb.generate(mb.gen_params(nu, nv, nnz, zipf_s=zipf))
    c = mb.Context(nu, nv, k)
    c.init_normal(1, 1e-2)
    dtr, dte = c.dataset_from_blocks(tr), c.dataset_from_blocks(te)
    print("zipf %.2f: %d ratings, gen+ingest %.1fs" % (zipf, tr.nratings, time.time() - t0), flush=True)
    for mode, name in ((mb.MODE_HOGWILD, "hogwild"), (mb.MODE_ATOMIC, "atomic")):
        for threads in (256, 128):
            for ctas in (0, 2, 1):
                c.set_option("threads", threads)
                c.set_option("ctas_per_sm", ctas)
                c.sgd_epoch(dtr, 0.02, 5e-3, 2.76, mode)
                ms = []
                for _ in range(2):
                    c.sgd_epoch(dtr, 0.01, 5e-3, 2.76, mode)
                    ms.append(c.last_kernel_ms())
                print("  %-8s threads %3d ctas/sm %d : %.2f ms  %.2f Gupd/s  rmse %.4f" % (
                    name, threads, ctas, min(ms), tr.nratings / min(ms) / 1e6, c.rmse(dte, 2.76)), flush=True)
    c.close()
