"""Generates tests/golden/fullsize/*.json: the REFERENCE's own test-RMSE (and lambda) trajectories at the
BASELINE.json config sizes (Netflix-shaped, 480,189 x 17,770, 100 M ratings).

Run in the build container (needs /root/reference and ~20 GB of /dev/shm):

    python tests/golden/make_fullsize_golden.py prep          # data files + seeded model files
    python tests/golden/make_fullsize_golden.py c2            # mf k=128, reference main() --fly 1
    python tests/golden/make_fullsize_golden.py c2fly         # mf k=128, reference main() --fly 8, 3 runs
    python tests/golden/make_fullsize_golden.py c3 [seed]     # dpmf eps=0 (pure SGLD) k=128, harness
    python tests/golden/make_fullsize_golden.py c4dp [seed]   # dpmf eps=1 k=64, harness
    python tests/golden/make_fullsize_golden.py c4ad          # admf k=64, harness

What runs is the reference's own code (oracle/Makefile compiles it where it lies, behind the
TBB/MKL/protobuf shims): `c2*` its `main()` (oracle/_ref/mf_ref, prints `iter#i <secs> tRMSE=`,
mf.h:35) started from a seeded `--model` file (MF::read_model, model.cc:75-97); `c3/c4*` its
SgldFilter / AdRegFilter / finish_noise / sample_hyper / updateReg through oracle/ref_harness.cc in
file order (`--fly 1` order), with the loops of run(DPMF&) / run(AdaptRegMF&) (main.cc:55-93,
model.cc:299-310) restated below, because the reference's main() cannot seed the factors of those two
algorithms (clock-seeded init, model.cc:3-5).

The data are regenerated anywhere from the counter-based generator (mfb_generate, same parameters), the
initial factors from numpy's seeded generator (mfb200.seeded_model), so the GPU box can replay the same
experiment without /root/reference and compare with the committed trajectories."""
import ctypes as C
import json
import os
import re
import subprocess
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "experimental-mf_b200"))
sys.path.insert(0, os.path.dirname(HERE))
import mfb200 as mb  # noqa: E402  (host-side generator / block writer only: no GPU is used here)
import oraclelib as ol  # noqa: E402
from oraclelib import _p, f32p  # noqa: E402

NU, NV, NNZ = 480_189, 17_770, 100_000_000
GB, ETA0, LAMBDA, GAM = 2.76, 2e-2, 5e-3, 1.0
MODEL_SEED = 20261018
EPOCHS = int(os.environ.get("EPOCHS", "10"))
WORK = os.environ.get("MFGOLD_DIR", "/dev/shm/mfgold")
OUT = os.path.join(HERE, "fullsize")
DP_TEMP = 0.1       # run.py:22,34 sweeps temp = 1e-1
ADMF_ETA_REG = 2e-3  # main.cc:104
# --hyperb: with the reference's default (100, main.cc:99) the Gibbs step draws lambda_u ~ 2000 while the factors are
# still near their 1e-2 initialisation, eta*ntrain*lambda_u >> 1 and the REFERENCE ITSELF ends in NaN at round 7-8
# on this data (kept on record: fullsize/*_hb100.json).  1e4 keeps the prior precisions at O(10) and the run stable.
HYPERB = float(os.environ.get("HYPERB", "1e4"))


def paths(tag):
    return [os.path.join(WORK, "%s_%s.bin" % (tag, s)) for s in ("train", "test", "valid")]


def write_ref_model(path, theta, phi, bu, bv, lam):
    """MF::save_model's byte layout (model.cc:98-121): nv, nu, dim, lambda, bv, phi, bu, theta."""
    with open(path, "wb") as f:
        np.array([phi.shape[0], theta.shape[0], theta.shape[1]], np.int32).tofile(f)
        np.array([lam], np.float32).tofile(f)
        bv.astype(np.float32).tofile(f)
        np.ascontiguousarray(phi, np.float32).tofile(f)
        bu.astype(np.float32).tofile(f)
        np.ascontiguousarray(theta, np.float32).tofile(f)


def prep():
    os.makedirs(WORK, exist_ok=True)
    for tag, vf in (("nov", 0.0), ("val", 0.01)):  # the default file set (C2/C3/C4-dp) and the one with a validation hold-out (admf)
        t0 = time.time()
        tr, te, va = mb.generate(mb.gen_params(NU, NV, NNZ, valid_frac=vf))
        tp, sp, vp = paths(tag)
        tr.write(tp)
        te.write(sp)
        if vf > 0:
            va.write(vp)
        print(tag, "train", tr.nratings, "runs", tr.nruns, "test", te.nratings, "valid", va.nratings if vf > 0 else 0,
              "%.0fs" % (time.time() - t0), flush=True)
    th, ph, bu, bv = mb.seeded_model(NU, NV, 128, MODEL_SEED)
    write_ref_model(os.path.join(WORK, "model_k128.bin"), th, ph, bu, bv, LAMBDA)


def save(name, obj):
    os.makedirs(OUT, exist_ok=True)
    obj["generated_by"] = "tests/golden/make_fullsize_golden.py " + " ".join(sys.argv[1:])
    obj["shape"] = {"nu": NU, "nv": NV, "nnz": NNZ}
    obj["model_seed"] = MODEL_SEED
    with open(os.path.join(OUT, name + ".json"), "w") as f:
        json.dump(obj, f, indent=1)
    print("wrote", name, flush=True)


def run_main(fly, extra=()):
    tp, sp, _ = paths("nov")
    cmd = [ol.REF_BIN, "--alg", "mf", "--train", tp, "--test", sp, "--nu", str(NU), "--nv", str(NV), "--dim", "128",
           "--iter", str(EPOCHS), "--fly", str(fly), "--eta", str(ETA0), "--lambda", str(LAMBDA), "--gam", str(GAM),
           "--bias", str(GB), "--model", os.path.join(WORK, "model_k128.bin")] + list(extra)
    t0 = time.time()
    out = subprocess.run(cmd, capture_output=True, text=True, check=True,
                         env=dict(os.environ, OMP_NUM_THREADS=str(max(fly, 1)))).stdout
    rmse = [float(x) for x in re.findall(r"tRMSE=([0-9.]+)", out)]
    secs = [float(x) for x in re.findall(r"iter#\d+\t([0-9.]+)\t", out)]
    print("fly", fly, "rmse", rmse, "%.0fs" % (time.time() - t0), flush=True)
    return {"cmd": " ".join(os.path.basename(c) if c.startswith("/") else c for c in cmd), "test_rmse": rmse,
            "cumulative_seconds": secs}


def c2():
    r = run_main(1)
    save("c2_mf_k128", {"config": "C2 Netflix-shaped SGD MF k=128, reference main() --fly 1 (single-thread update order)",
                        "eta0": ETA0, "lambda": LAMBDA, "gam": GAM, "gb": GB, **r})


def c2fly():
    fly = int(os.environ.get("FLY", "8"))
    runs = [run_main(fly) for _ in range(3)]
    save("c2_mf_k128_fly%d" % fly, {"config": "C2, reference main() --fly %d, three runs: the reference's own parallel "
                                    "(Hogwild over %d blocks in flight) spread" % (fly, fly), "runs": runs})


def dp_run(k, eps, seed, tag):
    """run(DPMF&) (main.cc:55-75) + DPMF::finish_round (model.cc:299-310) through the harness."""
    R = ol.ref()
    tp, sp, _ = paths("nov")
    R.ref_seed_generator(seed)
    R.ref_srand(seed)
    # eta0 and temp: the effective step scal = eta*ntrain*bound*lambda_r (dpmf.h:46) starts at 0.02 as in plain SGD,
    # the injected noise variance per coordinate and epoch is temp*eta*ntrain = 0.02*DP_TEMP (tools/exp_algs.py)
    # ntrain is known only after init (block_count); take the generator's count (checked below)
    tr, _, _ = mb.generate(mb.gen_params(NU, NV, NNZ))
    ntrain_guess = tr.nratings
    del tr
    bound = float(mb.lib().mfb_dp_bound(eps, 0, NV))
    eta0 = np.float32(2e-2 / ntrain_guess / bound)
    temp = np.float32(DP_TEMP * bound)
    noise_size = 400_000_000
    t0 = time.time()
    h = R.ref_create_dpmf(tp.encode(), sp.encode(), k, eta0, GAM, LAMBDA, GB, NU, NV, 1.0, HYPERB, eps, 0, noise_size,
                          temp, 1e-13)
    r = ol.Ref(h, NU, NV, k)
    th, ph, bu, bv = mb.seeded_model(NU, NV, k, MODEL_SEED)
    r.set_factors(th, ph, bu, bv)
    nt, bd, ta = C.c_int(), C.c_float(), C.c_int()
    R.ref_dpmf_info(h, C.byref(nt), C.byref(bd), C.byref(ta))
    assert nt.value == ntrain_guess, (nt.value, ntrain_guess)
    print("dpmf k", k, "eps", eps, "ntrain", nt.value, "bound", bd.value, "tau", ta.value, "init %.0fs" % (time.time() - t0),
          flush=True)
    out = {"config": "%s: Netflix-shaped dpmf k=%d epsilon=%g, reference SgldFilter/finish_noise/sample_hyper in file "
                     "order (harness), noise from the reference's own table (noise_size %d) and generator seed %d"
                     % (tag, k, eps, noise_size, seed),
           "k": k, "epsilon": eps, "tau": ta.value, "bound": bd.value, "ntrain": nt.value, "eta0": float(eta0),
           "temp": float(temp), "gam": GAM, "gb": GB, "mineta": 1e-13, "hyper_a": 1.0, "hyper_b": HYPERB,
           "noise_seed": seed, "eta": [], "train_rmse": [], "test_rmse": [], "lambda_r": [], "lambda_ub": [],
           "lambda_vb": [], "lambda_u_mean": [], "lambda_v_mean": [], "seconds": []}
    for ep in range(1, EPOCHS + 1):
        t0 = time.time()
        out["eta"].append(float(r.eta))
        r.epoch()
        R.ref_dpmf_finish_noise(h)
        s_tr, n_tr = r.sse(0)
        s_te, n_te = r.sse(1)
        R.ref_dpmf_sample_hyper(h, s_tr)
        hyp = np.zeros(3 + 2 * k, np.float32)
        R.ref_dpmf_get_hyper(h, _p(hyp, f32p))
        r.seteta(ep + 1)
        out["train_rmse"].append(float(np.sqrt(s_tr / n_tr)))
        out["test_rmse"].append(float(np.sqrt(s_te / n_te)))
        out["lambda_r"].append(float(hyp[0]))
        out["lambda_ub"].append(float(hyp[1]))
        out["lambda_vb"].append(float(hyp[2]))
        out["lambda_u_mean"].append(float(hyp[3:3 + k].mean()))
        out["lambda_v_mean"].append(float(hyp[3 + k:].mean()))
        out["seconds"].append(time.time() - t0)
        print(tag, "round", ep, "RMSE %.5f tRMSE %.5f lambda_r %.4f %.0fs" % (
            out["train_rmse"][-1], out["test_rmse"][-1], hyp[0], out["seconds"][-1]), flush=True)
        save("%s_seed%d%s" % (tag, seed, "" if HYPERB == 1e4 else "_hb%g" % HYPERB), out)


def c4ad():
    """run(AdaptRegMF&) (main.cc:77-93) + the epoch boundary of AdRegReadFilter (admf.h:30-36)."""
    R = ol.ref()
    k = 64
    tp, sp, vp = paths("val")
    R.ref_srand(5)
    h = R.ref_create_admf(tp.encode(), sp.encode(), vp.encode(), k, ETA0, GAM, LAMBDA, GB, NU, NV, 0, ADMF_ETA_REG)
    r = ol.Ref(h, NU, NV, k)
    th, ph, bu, bv = mb.seeded_model(NU, NV, k, MODEL_SEED)
    r.set_factors(th, ph, bu, bv)
    out = {"config": "C4 admf: Netflix-shaped (1 % validation hold-out) adaptive-regulariser MF k=64 loss=0, reference "
                     "AdRegFilter + updateReg in file order (harness), srand(5)",
           "k": k, "eta0": ETA0, "gam": GAM, "lambda": LAMBDA, "gb": GB, "eta_reg0": ADMF_ETA_REG, "srand": 5,
           "nvalid": int(R.ref_num_valid(h)), "test_rmse": [], "lams": [], "seconds": []}
    for ep in range(1, EPOCHS + 1):
        t0 = time.time()
        r.seteta(ep)
        r.epoch()
        l4 = np.zeros(4, np.float32)
        R.ref_admf_get_lams(h, _p(l4, f32p))
        s, n = r.sse(1)
        out["test_rmse"].append(float(np.sqrt(s / n)))
        out["lams"].append([float(x) for x in l4])
        out["seconds"].append(time.time() - t0)
        print("admf iter", ep, "tRMSE %.5f lams %s %.0fs" % (out["test_rmse"][-1], l4, out["seconds"][-1]), flush=True)
        save("c4_admf_k64", out)


if __name__ == "__main__":
    what = sys.argv[1]
    if what == "prep":
        prep()
    elif what == "c2":
        c2()
    elif what == "c2fly":
        c2fly()
    elif what == "c3":
        dp_run(128, 0.0, int(sys.argv[2]) if len(sys.argv) > 2 else 1, "c3_sgld_k128")
    elif what == "c4dp":
        dp_run(64, 1.0, int(sys.argv[2]) if len(sys.argv) > 2 else 1, "c4_dpmf_k64_eps1")
    elif what == "c4ad":
        c4ad()
    else:
        raise SystemExit(__doc__)
