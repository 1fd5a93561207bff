"""GPU experiment: (1) where does the time of rank 0's DSGD cell kernels go (P=8, Netflix shape)?  8 launches
vs ONE launch over the same records, widths, kernels; (2) which hot-row budget keeps the parallel schedule stable
once the run bound is lifted after epoch 1: ML-1M shape vs the serial oracle, and the P=8 schedule vs the golden."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "experimental-mf_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import mfb200 as mb  # noqa: E402
import mfb_dsgd  # noqa: E402
import oraclelib as ol  # noqa: E402

GB, LAM = 2.76, 5e-3
parts = sys.argv[1:] or ["T", "M", "D"]

if "T" in parts:
    NU, NV, NNZ, K, P = 480189, 17770, 100_000_000, 128, 8
    u0, u1 = mfb_dsgd.user_range(NU, 0, P)
    tr, te, _ = mb.generate(mb.gen_params(NU, NV, NNZ, user_begin=u0, user_end=u1))
    cells = tr.split_by_item(mfb_dsgd.item_bounds(NV, P))
    c = mb.Context(NU, NV, K)
    c.init_normal(1, 1e-2)
    c.set_option("placement_trials", 0)
    ds = [c.dataset_from_blocks(b) for b in cells]
    # all cells as one file (cell after cell)
    bo, ru, ro, vi, ra, rbase, obase = [0], [], [0], [], [], 0, 0
    for b in cells:
        bo += [int(x) + rbase for x in b.block_off[1:]]
        ru.append(np.asarray(b.run_uid)); ro.append(np.asarray(b.run_off[1:], np.int64) + obase)
        vi.append(np.asarray(b.vid)); ra.append(np.asarray(b.rating))
        rbase += b.nruns; obase += b.nratings
    cat = mb.Blocks.from_arrays(np.array(bo, np.int64), np.concatenate(ru), np.concatenate([np.zeros(1, np.int64)] + ro[1:]).astype(np.int32),
                                np.concatenate(vi), np.concatenate(ra))
    dcat = c.dataset_from_blocks(cat)
    dall = c.dataset_from_blocks(tr)  # the shard in file order (whole runs)
    n0 = tr.nratings
    lens = [np.diff(np.asarray(b.run_off)) for b in cells]
    print("T: rank 0 of %d: %d ratings, %d runs in the file, cells: %s runs, mean piece %.1f, longest piece %d, top item share %s" % (
        P, n0, tr.nruns, [b.nruns for b in cells], np.mean(np.concatenate(lens)), max(int(l.max()) for l in lens),
        ["%.3f" % (np.bincount(np.asarray(b.vid)).max() / b.nratings) for b in cells]), flush=True)
    eta = 0.004
    for kern, planes in ((3, 0), (3, 1), (4, 0)):
        for age, rc in ((5, 16), (5, 32), (5, 64)):
            c.set_option("kernel", kern); c.set_option("model_age", age); c.set_option("row_concurrency", rc)
            c.set_option("phi_planes", planes)
            for rep in range(2):
                per = []
                for d in ds:
                    c.sgd_epoch_blocks(d, 0, c.num_blocks(d), eta, LAM, GB, mb.MODE_ATOMIC); per.append(c.last_kernel_ms())
                shape = c.last_launch()
                c.sgd_epoch_blocks(dcat, 0, c.num_blocks(dcat), eta, LAM, GB, mb.MODE_ATOMIC); one = c.last_kernel_ms()
                c.sgd_epoch_blocks(dall, 0, c.num_blocks(dall), eta, LAM, GB, mb.MODE_ATOMIC); whole = c.last_kernel_ms()
            print("T kernel %d planes %d rc %2d: 8 cell launches %.2f ms (%s) | one launch over the same cells %.2f ms | the shard in file "
                  "order (whole runs) %.2f ms | %s" % (kern, planes, rc, sum(per), " ".join("%.2f" % x for x in per), one, whole, shape), flush=True)
    c.close()
    del tr, te, cells, cat

if "M" in parts:
    nu, nv, nnz, k = 6040, 3706, 1_000_000, 32
    tr, te, _ = mb.generate(mb.gen_params(nu, nv, nnz, test_frac=0.1))
    train = ol.Dataset(tr.block_off, tr.run_uid, tr.run_off, tr.vid, tr.rating)
    test = ol.Dataset(te.block_off, te.run_uid, te.run_off, te.vid, te.rating)
    m = ol.Model(nu, nv, k, seed=11)
    th, ph = [x.copy() for x in m.dense()]   # (dense() returns views when dim == stride; the oracle trains m in place)
    bu, bv = m.bu.copy(), m.bv.copy()
    mm, dd, tt = m.as_mfo(), train.as_mfo(), test.as_mfo()
    want = []
    for ep in range(1, 11):
        ol.oracle().mfo_sgd_epoch(C.byref(mm), C.byref(dd), mb.seteta(2e-2, ep, 1.0), LAM, GB)
        n = C.c_int64(); s = ol.oracle().mfo_sse(C.byref(mm), C.byref(tt), GB, C.byref(n)); want.append(float(np.sqrt(s / n.value)))
    print("M: ML-1M shape, oracle %s" % " ".join("%.4f" % x for x in want), flush=True)
    c = mb.Context(nu, nv, k)
    d, dte = c.dataset_from_blocks(tr), c.dataset_from_blocks(te)
    for w1, w2 in ((0, 0), (0, 168), (0, 336), (0, 672), (0, 1344), (0, 2688), (168, 168), (336, 336), (672, 672), (0, -1)):
        c.set_factors(th, ph, bu, bv)
        traj, ms = [], []
        for ep in range(1, 11):
            # explicit width in runs: w1 in epoch 1, w2 afterwards (0 = the default bounds, -1 = run bound off)
            w = w1 if ep == 1 else w2
            c.set_option("max_groups", max(w, 0)); c.set_option("run_fraction_ppm", 0 if w < 0 else 3500)
            c.sgd_epoch(d, mb.seteta(2e-2, ep, 1.0), LAM, GB, mb.MODE_ATOMIC); ms.append(c.last_kernel_ms()); traj.append(c.rmse(dte, GB))
        print("M width epoch 1: %s, later: %s: final %.5f diff %+.5f max|diff| %.5f ms/epoch %s launch %s" % (
            w1 or "default", {0: "default", -1: "row bound only"}.get(w2, w2), traj[-1], traj[-1] - want[-1],
            max(abs(a - b) for a, b in zip(traj, want)), " ".join("%.2f" % x for x in ms), c.last_launch()), flush=True)
    c.close()

if "D" in parts:
    GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "fullsize", "c2_mf_k128.json")))
    WANT = GOLD["test_rmse"]
    NU, NV, NNZ, K = 480189, 17770, 100_000_000, 128
    P, R1 = int(os.environ.get("P", "8")), 16
    TH, PH, BU, BV = mb.seeded_model(NU, NV, K, GOLD["model_seed"])
    c = mb.Context(NU, NV, K)
    c.set_option("placement_trials", 0)
    cells, tests = [], []
    for p in range(P):
        u0, u1 = mfb_dsgd.user_range(NU, p, P)
        trp, tep, _ = mb.generate(mb.gen_params(NU, NV, NNZ, user_begin=u0, user_end=u1))
        cells.append([c.dataset_from_blocks(b) for b in trp.split_by_item(mfb_dsgd.item_bounds(NV, P))])
        tests.append(c.dataset_from_blocks(tep))
        del trp, tep
    for rbe, rc in ((1000, 32), (1, 8), (1, 12), (1, 16), (1, 24)):
        c.set_option("run_bound_epochs", rbe); c.set_option("row_concurrency", rc)
        c.set_factors(TH, PH, BU, BV)
        traj, ms = [], []
        for ep in range(1, 11):
            c.set_option("model_age", ep - 1)
            eta, rot, ms0 = mb.seteta(2e-2, ep, 1.0), (R1 if ep == 1 else 1), 0.0
            for r in range(rot):
                for s in range(P):
                    for p in range(P):
                        dsx = cells[p][(p + s) % P]; nb = c.num_blocks(dsx)
                        c.sgd_epoch_blocks(dsx, nb * r // rot, nb * (r + 1) // rot, eta, LAM, GB, mb.MODE_ATOMIC)
                        if p == 0: ms0 += c.last_kernel_ms()
            ms.append(ms0)
            sse = n = 0
            for dt in tests:
                a, b = c.sse(dt, GB); sse += a; n += b
            traj.append(float(np.sqrt(sse / n)))
        print("D P%d R1=%d run bound in %s, rc %2d: final %.5f (ref %.5f, diff %+.5f) rank-0 kernel ms %s" % (
            P, R1, "every epoch" if rbe > 1 else "epoch 1 only", rc, traj[-1], WANT[-1], traj[-1] - WANT[-1], " ".join("%.2f" % x for x in ms)), flush=True)
    c.close()
