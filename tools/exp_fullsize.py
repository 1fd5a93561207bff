"""GPU experiment at the full Netflix shape (C2: 480,189 x 17,770, 100M ratings, k=128) against the committed
trajectory of the reference (tests/golden/fullsize/c2_mf_k128.json, `mf_ref --fly 1`), same data, same seeded model.

  part S: one GPU, file order, production schedule under different concurrency bounds (per epoch)
  part D: the P-GPU DSGD schedule walked by ONE GPU (cells of a sub-epoch share no user and no item, so this is
          the P-GPU result up to intra-cell Hogwild effects): test RMSE for ring turns in epoch 1 / piece counts /
          bounds, and the summed kernel time of rank 0's cells (= what one of the P GPUs would spend per epoch)
  part K: rank 0's cells only: kernel choice and width at a late epoch's step size

  python tools/exp_fullsize.py [S] [D] [K]    (env P=8)"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "experimental-mf_b200"))
import mfb200 as mb  # noqa: E402
import mfb_dsgd  # noqa: E402

GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "fullsize", "c2_mf_k128.json")))
WANT = GOLD["test_rmse"]
NU, NV, NNZ, K = GOLD["shape"]["nu"], GOLD["shape"]["nv"], GOLD["shape"]["nnz"], 128
ETA0, LAM, GAM, GB = GOLD["eta0"], GOLD["lambda"], GOLD["gam"], GOLD["gb"]
EPOCHS = int(os.environ.get("EPOCHS", len(WANT)))
parts = [a for a in sys.argv[1:]] or ["S", "D", "K"]
TH, PH, BU, BV = mb.seeded_model(NU, NV, K, GOLD["model_seed"])


def fmt(traj):
    return " ".join("%.4f" % x for x in traj)


def report(name, traj, ms):
    d = [t - w for t, w in zip(traj, WANT)]
    print("%-46s final %.5f (ref %.5f, diff %+.5f, max|diff| %.5f) ms/epoch %s" % (
        name, traj[-1], WANT[len(traj) - 1], d[-1], max(abs(x) for x in d), " ".join("%.1f" % x for x in ms)), flush=True)
    print("%-46s traj %s" % ("", fmt(traj)), flush=True)


if "S" in parts:
    t0 = time.time()
    tr, te, _ = mb.generate(mb.gen_params(NU, NV, NNZ))
    np.asarray(tr.vid[:64 << 20]).tofile("/tmp/mfb_vid.bin")  # for tools/l2_atomic_peak (real item popularity)
    c = mb.Context(NU, NV, K)
    dtr, dte = c.dataset_from_blocks(tr), c.dataset_from_blocks(te)
    print("S: generated + ingested in %.0f s; reference trajectory %s" % (time.time() - t0, fmt(WANT)), flush=True)
    DEF = {"run_fraction_ppm": 3500, "row_concurrency": 32, "eta_scaling": 1, "max_groups": 0}
    variants = [
        ("default (run bound in epoch 1 only, rc 32)", lambda ep: {}),
        ("run bound in every epoch (round 1)", lambda ep: {"run_bound_epochs": 1000}),
        ("rc 24", lambda ep: {"row_concurrency": 24}),
        ("rc 40", lambda ep: {"row_concurrency": 40}),
        ("rf 1750 in epoch 1", lambda ep: {"run_fraction_ppm": 1750}),
        ("regrouped file (merge + longest first) from epoch 2", lambda ep: {}),
    ]
    DEF["run_bound_epochs"] = 1
    dreg = c.dataset_from_blocks(tr.regroup(True, True))
    for name, opts in variants:
        c.set_factors(TH, PH, BU, BV)
        traj, ms = [], []
        for ep in range(1, EPOCHS + 1):
            for k_, v_ in {**DEF, **opts(ep)}.items():
                c.set_option(k_, v_)
            c.sgd_epoch(dreg if (name.startswith("regrouped") and ep > 1) else dtr, mb.seteta(ETA0, ep, GAM), LAM, GB, mb.MODE_ATOMIC)
            ms.append(c.last_kernel_ms())
            traj.append(c.rmse(dte, GB))
        report("S " + name, traj, ms)
    c.close()
    del tr, te

if "D" in parts or "K" in parts:
    P = int(os.environ.get("P", "8"))
    HS = [int(x) for x in os.environ.get("HALVES", "1,2").split(",")]
    c = mb.Context(NU, NV, K)
    c.set_option("placement_trials", 0)  # one placement for all variants (the search would favour the first)
    t0 = time.time()
    cells, tests = {h: [] for h in HS}, []
    for p in range(P):
        u0, u1 = mfb_dsgd.user_range(NU, p, P)
        trp, tep, _ = mb.generate(mb.gen_params(NU, NV, NNZ, user_begin=u0, user_end=u1))
        for h in HS:
            cells[h].append([c.dataset_from_blocks(b) for b in trp.split_by_item(mfb_dsgd.item_bounds(NV, P * h))])
        tests.append(c.dataset_from_blocks(tep))
        del trp, tep
    print("D: P=%d cells ingested in %.0f s" % (P, time.time() - t0), flush=True)
    DEF = {"run_fraction_ppm": 3500, "row_concurrency": 32, "eta_scaling": 1, "max_groups": 0, "kernel": 0}

    def rmse_all():
        s = n = 0
        for d in tests:
            a, b = c.sse(d, GB)
            s, n = s + a, n + b
        return float(np.sqrt(s / n))

    def walk_epoch(H, rot, eta, ranks=None):
        """the DSGD schedule of one epoch; returns kernel ms summed over the cells of rank 0"""
        ms0 = 0.0
        for r in range(rot):
            for s in range(P):
                for p in (range(P) if ranks is None else ranks):
                    for h in range(H):
                        ds = cells[H][p][((p + s) % P) * H + h]
                        nb = c.num_blocks(ds)
                        c.sgd_epoch_blocks(ds, nb * r // rot, nb * (r + 1) // rot, eta, LAM, GB, mb.MODE_ATOMIC)
                        if p == 0:
                            ms0 += c.last_kernel_ms()
        return ms0

if "D" in parts:
    variants = [
        # name, halves, turns in epoch 1, per-epoch options
        ("H1 R1=1 default bounds (round 1)", 1, 1, lambda ep: {}),
        ("H1 R1=16", 1, 16, lambda ep: {}),
        ("H2 R1=16", 2, 16, lambda ep: {}),
        ("H2 R1=32", 2, 32, lambda ep: {}),
        ("H2 R1=16, runs unbounded from epoch 2", 2, 16, lambda ep: {"run_fraction_ppm": 3500 if ep == 1 else 0}),
        ("H2 R1=16, runs unbounded always", 2, 16, lambda ep: {"run_fraction_ppm": 0}),
        ("H2 R1=16, unbounded from ep 2, rc 64", 2, 16, lambda ep: {"run_fraction_ppm": 3500 if ep == 1 else 0, "row_concurrency": 64}),
        ("H1 R1=16, runs unbounded from epoch 2", 1, 16, lambda ep: {"run_fraction_ppm": 3500 if ep == 1 else 0}),
    ]
    for name, H, R1, opts in variants:
        if H not in HS:
            continue
        c.set_factors(TH, PH, BU, BV)
        traj, ms = [], []
        for ep in range(1, EPOCHS + 1):
            for k_, v_ in {**DEF, **opts(ep)}.items():
                c.set_option(k_, v_)
            ms.append(walk_epoch(H, R1 if ep == 1 else 1, mb.seteta(ETA0, ep, GAM)))
            traj.append(rmse_all())
        report("D P%d %s [ms = rank 0's kernels]" % (P, name), traj, ms)

if "K" in parts:
    # rank 0's cells at the step size of epoch 5: which kernel, how wide
    c.set_factors(TH, PH, BU, BV)
    for k_, v_ in DEF.items():
        c.set_option(k_, v_)
    for ep in range(1, 4):  # a model that is no longer at its initialisation
        walk_epoch(HS[0], 1, mb.seteta(ETA0, ep, GAM))
    eta = mb.seteta(ETA0, 5, GAM)
    for H in HS:
        n0 = sum(c.num_ratings(d) for d in cells[H][0])
        for kern in (0, 3, 4):
            for rf in (3500, 14000, 0):
                for rc in (32, 128):
                    for k_, v_ in {**DEF, "kernel": kern, "run_fraction_ppm": rf, "row_concurrency": rc}.items():
                        c.set_option(k_, v_)
                    walk_epoch(H, 1, eta, ranks=[0])
                    ms = walk_epoch(H, 1, eta, ranks=[0])
                    ll = c.last_launch()
                    print("K P%d H%d kernel %d rf %5d rc %3d: rank 0's %d cell kernels %.2f ms = %.2f G upd/s  (last launch %s)" % (
                        P, H, kern, rf, rc, P * H, ms, n0 / ms / 1e6, ll), flush=True)
    c.close()
