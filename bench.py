#!/usr/bin/env python
"""bench.py - the hot path of experimental-mf on B200: one "step" = one blocked-SGD matrix
factorization epoch over the whole synthetic rating file (BASELINE.json metric: rating
updates/sec per epoch; N=1 workload = configs[1], Netflix-shaped 480,189 x 17,770, 100M ratings,
k=128, plain SGD, fp32).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

prints ONE JSON line (rank 0).  Legs of the default (b200) arm, all in one run:
  parity     10 epochs from the seeded model of tests/golden/fullsize/c2_mf_k128.json (the reference's own
             `main() --fly 1` trajectory on the same data): test_rmse / test_rmse_reference / rmse_abs_diff
  value      epochs with the rating tiles resident in HBM, CUDA events, max over ranks
  roofline   the update kernel's own duration (CUDA events around the launch on its stream) against the
             measured L2 ceiling of this access pattern (profiles/r2_l2_atomic_peak.json); the algorithmic-HBM
             figure (12 + 16k bytes per update against MEASURED_PEAKS.json) is kept beside it
  e2e        the same epochs through the C ABI with the rating tiles in pinned HOST memory
             (chunked H2D copy overlapped with the kernel) + test SSE read back every step;
             e2e_from_file: epochs straight from the protobuf FILE (out-of-core path, nothing resident)
  cpu_baseline  the reference's own CPU path (oracle/_ref/mf_ref: reference sources + shim
             TBB/MKL/protobuf) on a bounded sample of the same workload, all host cores
  configs    short sub-records for the other BASELINE.json configs (C1 ML-1M-shaped mf k=32, C3 SGLD k=128,
             C4 dpmf eps>0 / admf k=64, C5 Yahoo-shaped mf k=128 on this one GPU), each with its time per epoch,
             its bound and its test RMSE next to the reference's
`--impl reference` runs only the reference's CPU path, on the whole file, as its own arm (no product code in
that process: the data come from the `getdata` binary).
"""
import argparse
import ctypes as C
import json
import os
import re
import shutil
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "experimental-mf_b200")

WORKLOADS = {
    # name: (nu, nv, nnz, k, test_frac)   shapes: BASELINE.json configs / reference src/run.py:2-3,6-7
    "ml1m": (6040, 3706, 1_000_000, 32, 0.1),
    "netflix": (480_189, 17_770, 100_000_000, 128, 0.01),
    "yahoo": (1_000_990, 624_961, 252_000_000, 128, 0.01),
}
ETA0, LAMBDA, GAM, GB = 2e-2, 5e-3, 1.0, 2.76  # reference defaults, main.cc:97-100
METRIC, UNIT = "rating updates/sec per epoch", "updates/s"
GOLD_DIR = os.path.join(ROOT, "tests", "golden", "fullsize")
RMSE_TOL = 1e-3  # BASELINE.json north_star: final test RMSE within 1e-3 absolute of the reference


def bytes_per_update(k):
    return 12 + 16 * k  # rating record + two factor rows read and written


KERNEL_NAMES = {1: "sgd_epoch_kernel (warp per run)", 3: "sgd_stream_kernel (sub-warp per run, cp.async ring)",
                4: "sgd_burst_kernel (warp per run, Gram-batched)"}


def load_json(path):
    try:
        return json.load(open(path))
    except (OSError, ValueError):
        return None


def gold(name):
    return load_json(os.path.join(GOLD_DIR, name + ".json"))


def measured_peak():
    p = load_json(os.path.join(ROOT, "MEASURED_PEAKS.json"))
    if p and "hbm_gbs" in p:
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def l2_peak():
    """profiles/r2_l2_atomic_peak.json (tools/l2_atomic_peak.cu run on a B200 of this pool): G row-updates/s that
    the L2 sustains for gather + red.add.v4.f32 of one 512-B row per update, no arithmetic"""
    return load_json(os.path.join(ROOT, "profiles", "r2_l2_atomic_peak.json"))


def profiled(name):
    """one launch of the update kernel from the committed ncu --set full capture (tools/ncu_summary.py)"""
    for f in ("r2_%s.json" % name, "r1_%s.json" % name):
        d = load_json(os.path.join(ROOT, "profiles", f))
        if d and d.get("launches"):
            return d["launches"][0], os.path.join("profiles", f)
    return None, None


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region: NVML polled from a thread
    (about a millisecond per sample, so even a 50 ms region gets tens of samples); nvidia-smi -lms
    as the fallback when the NVML binding is missing."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index
        self.nvml, self.handle, self.stop_flag, self.t = None, None, False, None
        self.source = None

    def _nvml_index(self):
        # LOCAL_RANK counts visible devices; NVML counts physical ones
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [x.strip() for x in vis.split(",") if x.strip()]
            if self.index < len(ids) and ids[self.index].isdigit():
                return int(ids[self.index])
        return self.index

    def start(self):
        if os.environ.get("MFB_NO_SAMPLER"):
            return
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(self._nvml_index())
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.source = "nvml"
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.source = "nvidia-smi"
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _poll(self):
        n = self.nvml
        bits = [(getattr(n, "nvmlClocksEventReasonHwSlowdown", 0x8), "hw_slowdown"),
                (getattr(n, "nvmlClocksEventReasonHwThermalSlowdown", 0x40), "hw_thermal_slowdown"),
                (getattr(n, "nvmlClocksEventReasonSwThermalSlowdown", 0x20), "sw_thermal_slowdown"),
                (getattr(n, "nvmlClocksEventReasonSwPowerCap", 0x4), "sw_power_cap")]
        while not self.stop_flag:
            try:
                mhz = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
                try:
                    mask = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
                except Exception:
                    mask = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
                self.rows.append((mhz, self.max_mhz, [name for bit, name in bits if mask & bit]))
            except Exception:
                pass
            time.sleep(0.001)

    def _read(self):
        for line in self.proc.stdout:
            r = [x.strip() for x in line.split(",")]
            try:
                self.rows.append((float(r[0]), float(r[1]),
                                  [name for name, v in zip(self.NAMES, r[3:7]) if v.lower().startswith("active")]))
            except (ValueError, IndexError):
                continue

    def stop(self):
        if self.t is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0, "reasons": ["no NVML and no nvidia-smi"]}
        if self.nvml:
            self.stop_flag = True
            self.t.join(timeout=2)
        else:
            self.proc.terminate()
            self.t.join(timeout=2)
        sm = sorted(r[0] for r in self.rows)
        mx = [r[1] for r in self.rows]
        reasons = set()
        for r in self.rows:
            reasons.update(r[2])
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons), "source": self.source}


# ------------------------------------------------------------------------------------ CPU arm
def scratch_dir():
    return tempfile.mkdtemp(prefix="mfbench_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)


def run_mf_ref(train_path, test_path, nu, nv, k, iters, cores, model=None):
    """the reference's own main() (oracle/_ref/mf_ref) -> (seconds per epoch, tRMSE per epoch); mf.h:35 prints
    "iter#i \\t cumulative seconds \\t tRMSE=" (the clock includes file re-read, parse and test evaluation)"""
    ref_bin = os.path.join(ROOT, "oracle", "_ref", "mf_ref")
    cmd = [ref_bin, "--alg", "mf", "--train", train_path, "--test", test_path, "--nu", str(nu), "--nv", str(nv),
           "--dim", str(k), "--iter", str(iters), "--fly", str(cores), "--eta", str(ETA0), "--lambda", str(LAMBDA),
           "--gam", str(GAM), "--bias", str(GB)] + (["--model", model] if model else [])
    out = subprocess.run(cmd, capture_output=True, text=True, check=True,
                         env=dict(os.environ, OMP_NUM_THREADS=str(cores))).stdout
    cum = [float(x) for x in re.findall(r"iter#\d+\t([0-9.]+)\t", out)]
    rmse = [float(x) for x in re.findall(r"tRMSE=([0-9.]+)", out)]
    return [b - a for a, b in zip([0.0] + cum[:-1], cum)], rmse


def cpu_reference_run(train_blocks, test_blocks, nu, nv, k, iters, sample_ratings, cores, model_arrays=None):
    """Times the reference's own CPU implementation (its main.cc/mf.h/model.cc compiled into
    oracle/_ref/mf_ref; falls back to the oracle port when that binary was never built) on the
    first `sample_ratings` records of the training file.  Returns per-epoch seconds."""
    import numpy as np

    import mfb200 as mb
    ref_bin = os.path.join(ROOT, "oracle", "_ref", "mf_ref")
    run_off = train_blocks.run_off
    nruns = int(np.searchsorted(run_off, sample_ratings, side="right") - 1)
    nruns = max(nruns, 1)
    n = int(run_off[nruns])
    block_off = train_blocks.block_off
    nb = int(np.searchsorted(block_off, nruns, side="right") - 1)
    bo = np.r_[block_off[:nb + 1], nruns] if block_off[nb] != nruns else block_off[:nb + 1]
    sample = mb.Blocks.from_arrays(bo, train_blocks.run_uid[:nruns], run_off[:nruns + 1],
                                   train_blocks.vid[:n], train_blocks.rating[:n])
    tmp = scratch_dir()
    tp, sp = os.path.join(tmp, "train.bin"), os.path.join(tmp, "test.bin")
    sample.write(tp)
    test_blocks.write(sp)
    whole = n == train_blocks.nratings
    desc = "%s %d ratings (%d user-runs) of the training file, %d epoch(s), --fly %d" % (
        "all" if whole else "first", n, nruns, iters, cores)
    try:
        if os.path.exists(ref_bin):
            kind = "reference"
            mp = None
            if model_arrays is not None:  # MF::read_model's layout (model.cc:75-97)
                th, ph, bu, bv = model_arrays
                mp = os.path.join(tmp, "model.bin")
                with open(mp, "wb") as f:
                    np.array([nv, nu, k], np.int32).tofile(f)
                    np.array([LAMBDA], np.float32).tofile(f)
                    bv.tofile(f), ph.tofile(f), bu.tofile(f), th.tofile(f)
            secs, rmse = run_mf_ref(tp, sp, nu, nv, k, iters, cores, mp)
        else:
            kind = "port"
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            import oraclelib as ol
            cores = 1
            m = ol.Model(nu, nv, k, seed=1)
            d = ol.Dataset(bo, sample.run_uid, sample.run_off, sample.vid, sample.rating)
            mm, dd = m.as_mfo(), d.as_mfo()
            secs, rmse = [], []
            for ep in range(1, iters + 1):
                t0 = time.time()
                ol.oracle().mfo_sgd_epoch(C.byref(mm), C.byref(dd), ol.oracle().mfo_seteta(ETA0, ep, GAM), LAMBDA, GB)
                secs.append(time.time() - t0)
            desc = "first %d ratings, %d epoch(s), single-thread C port" % (n, iters)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return {"kind": kind, "cores": cores, "sample": desc, "n": n, "secs": secs, "rmse": rmse}


def run_reference_arm(args, wl):
    """The reference's own CPU implementation on the WHOLE file of the workload (or, with --cpu-sample N, on the
    first Blocks holding N ratings - then the line's config says so).  This process loads nothing of the product:
    the files come from the `getdata` binary (a separate process)."""
    nu, nv, nnz, k, test_frac = WORKLOADS[wl]
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    tmp = scratch_dir()
    try:
        cmd = [os.path.join(PKG, "getdata"), "-w", os.path.join(tmp, "d"), "--method", "synth", "--nu", str(nu), "--nv", str(nv),
               "--nnz", str(nnz), "--test", str(test_frac)]
        if args.cpu_sample and args.cpu_sample < nnz:
            cmd += ["--head", str(args.cpu_sample)]
        out = subprocess.run(cmd, capture_output=True, text=True, check=True).stdout
        n = int(re.search(r"train (\d+) ratings", out).group(1))
        iters = args.warmup + args.steps
        all_secs, rmse = run_mf_ref(os.path.join(tmp, "d.train"), os.path.join(tmp, "d.test"), nu, nv, k, iters, cores)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    secs = all_secs[args.warmup:]
    t = sum(secs)
    val = n * len(secs) / t
    cfg = workload_config(wl, args.gpus)
    whole = not (args.cpu_sample and args.cpu_sample < nnz)
    if not whole:  # the label of what was run, not of the arm it stands beside
        cfg["ratings"] = n
        cfg["workload"] = "first %d ratings of the %s" % (n, cfg["workload"])
    cfg["schedule"] = "the reference's own: TBB-style pipeline, --fly %d SgdFilter calls in flight (Hogwild)" % cores
    cfg["parallelism"] = cfg["parallelism"] + " (this arm: %d host cores)" % cores if args.gpus == 1 else cfg["parallelism"]
    sample = "%s %d training ratings, epochs %d..%d timed by the reference's own clock, --fly %d" % (
        "all" if whole else "the first", n, args.warmup + 1, iters, cores)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / len(secs),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "reference", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "test_rmse": rmse[-1] if rmse else None, "epochs_run": iters, "train_ratings": n,
        "note": "reference source + shim TBB/MKL/protobuf on the host CPU; each step = one epoch over the file; "
                "time is the reference's own printed clock (includes file re-read, parse, test eval)",
    }
    print(json.dumps(line))


def workload_config(wl, n_gpus):
    nu, nv, nnz, k, _ = WORKLOADS[wl]
    return {"workload": "%s-shaped synthetic (%d users x %d items, %d ratings) SGD MF k=%d fp32" % (wl, nu, nv, nnz, k),
            "nu": nu, "nv": nv, "ratings": nnz, "k": k, "alg": "mf", "eta": ETA0, "lambda": LAMBDA,
            "schedule": "parallel user-runs, atomic (red.add.v4.f32) accumulation of item- and user-row increments, "
                        "bounded concurrency",
            "epochs": "eta = eta0/epoch as in the reference (model.cc:36-38): warm-up steps are epochs 1..W, "
                      "timed steps the epochs after them",
            "parallelism": "1 GPU" if n_gpus == 1 else "dsgd%d" % n_gpus,
            "l2": "inputs larger than L2: each epoch streams the rating tiles (8 B/rating) and all user rows"}


def first_epoch_within(traj, target, tol=RMSE_TOL):
    for i, x in enumerate(traj):
        if x <= target + tol:
            return i + 1
    return None


# ------------------------------------------------------------------------------------ other configs
def gibbs(np, rng, a, b, sum_sq, count):
    """gamma_posterior (util.h:150-154): lambda ~ Gamma(a + n/2, rate b + sum/2)"""
    return float(rng.gamma(a + 0.5 * count, 1.0 / (b + 0.5 * sum_sq)))


def config_dpmf(mb, np, tr, te, nu, nv, k, g, device):
    """C3 / C4-dp: run(DPMF&) (main.cc:55-75, finish_round model.cc:299-310) through the C ABI from the golden
    file's seeded model and hyper-parameters; the Gibbs draws are host code (numpy's gamma here)."""
    c = mb.Context(nu, nv, k, device)
    c.set_factors(*mb.seeded_model(nu, nv, k, g["model_seed"]))
    c.enable(2)
    d, dte = c.dataset_from_blocks(tr), c.dataset_from_blocks(te)
    ntrain = c.dp_weights(d)
    bound = mb.lib().mfb_dp_bound(g["epsilon"], 0, nv)
    eta0, temp = np.float32(g["eta0"]), np.float32(g["temp"])
    lam_u, lam_v = np.full(k, 1e2, np.float32), np.full(k, 1e2, np.float32)
    lam_r, lam_ub, lam_vb = 1.0, 1e2, 1e2
    rng = np.random.default_rng(7)
    traj, ms = [], []
    for rnd in range(1, len(g["test_rmse"]) + 1):
        eta = mb.lib().mfb_seteta_cutoff(eta0, rnd, g["gam"], g["mineta"])
        c.upload(mb.LAMBDA_U, lam_u)
        c.upload(mb.LAMBDA_V, lam_v)
        p = mb.SgldParams(eta, temp, bound, ntrain, lam_r, lam_ub, lam_vb, 2026, rnd, 0, 0)
        c.sgld_epoch(d, p, g["gb"], mb.MODE_HOGWILD)
        ms.append(c.last_kernel_ms())
        c.sgld_flush_noise(d, p)
        s_tr, n_tr = c.sse(d, g["gb"])
        traj.append(c.rmse(dte, g["gb"]))
        nu_, nv_, bu2, bv2 = c.col_sqnorms()
        a, b = g["hyper_a"], g["hyper_b"]
        lam_r = gibbs(np, rng, a, b, s_tr, ntrain)
        lam_ub, lam_vb = gibbs(np, rng, a, b, bu2, nu), gibbs(np, rng, a, b, bv2, nv)
        lam_u = np.array([gibbs(np, rng, a, b, x, nu) for x in nu_], np.float32)
        lam_v = np.array([gibbs(np, rng, a, b, x, nv) for x in nv_], np.float32)
    c.close()
    steady = ms[len(ms) // 2:]
    return traj, sum(steady) / len(steady), ntrain, lam_r


def config_admf(mb, np, tr, te, va, nu, nv, k, g, device):
    """C4-admf: run(AdaptRegMF&) (main.cc:77-93, admf.h:30-36) through the C ABI from the golden file's seeded model."""
    c = mb.Context(nu, nv, k, device)
    c.set_factors(*mb.seeded_model(nu, nv, k, g["model_seed"]))
    c.enable(1)
    c.snapshot_old()
    d, dte = c.dataset_from_blocks(tr), c.dataset_from_blocks(te)
    vu = np.repeat(va.run_uid, np.diff(va.run_off)).astype(np.int32)
    perm = np.random.default_rng(5).permutation(len(vu))  # model.cc:413 shuffles the list
    c.admf_set_validation(vu[perm], np.asarray(va.vid)[perm], np.asarray(va.rating)[perm])
    c.admf_set_lams([g["lambda"]] * 4)
    rng = np.random.default_rng(6)
    traj, lams, ms = [], [], []
    for ep in range(1, len(g["test_rmse"]) + 1):
        c.admf_set_draws(rng.integers(0, len(vu), tr.nruns).astype(np.int32))  # admf.h:82: one draw per user-run
        c.admf_epoch(d, mb.seteta(g["eta0"], ep, g["gam"]), mb.seteta(g["eta_reg0"], ep, g["gam"]), 0, g["gb"], mb.MODE_ATOMIC)
        ms.append(c.last_kernel_ms())
        traj.append(c.rmse(dte, g["gb"]))
        lams.append([float(x) for x in c.admf_get_lams()])
    c.close()
    steady = ms[len(ms) // 2:]
    return traj, lams, sum(steady) / len(steady)


def config_mf_small(mb, np, wl, device, steps, warmup, cores):
    """C1 (ML-1M-shaped, k=32): resident epochs + the 10-epoch test RMSE next to the reference's own main() --fly 1
    from the same seeded model (the CPU leg of this sub-record)."""
    nu, nv, nnz, k, test_frac = WORKLOADS[wl]
    tr, te, _ = mb.generate(mb.gen_params(nu, nv, nnz, test_frac=test_frac))
    model = mb.seeded_model(nu, nv, k, 20261018)
    c = mb.Context(nu, nv, k, device)
    dtr, dte = c.dataset_from_blocks(tr), c.dataset_from_blocks(te)
    c.set_factors(*model)
    traj = []
    for ep in range(1, 11):
        c.sgd_epoch(dtr, mb.seteta(ETA0, ep, GAM), LAMBDA, GB, mb.MODE_ATOMIC)
        traj.append(c.rmse(dte, GB))
    c.set_factors(*model)
    ms = []
    for ep in range(1, warmup + steps + 1):
        c.sgd_epoch(dtr, mb.seteta(ETA0, ep, GAM), LAMBDA, GB, mb.MODE_ATOMIC)
        ms.append(c.last_kernel_ms())
    shape = c.last_launch()
    c.close()
    kms = sum(ms[warmup:]) / steps
    rec = {"workload": workload_config(wl, 1)["workload"], "ms_per_epoch": kms, "updates_per_s": tr.nratings / kms * 1e3,
           "launch": shape, "kernel": KERNEL_NAMES.get(shape["kernel"], "?"),
           "hbm_algorithmic_frac": tr.nratings * bytes_per_update(k) / (kms * 1e-3) / 1e9 / measured_peak()[0],
           "bound": "launch width: %d user-runs in the file; the factor matrices (1.2 MB) are L2-resident" % tr.nruns,
           "test_rmse": traj[-1], "rmse_trajectory": traj}
    ref_one = cpu_reference_run(tr, te, nu, nv, k, 10, tr.nratings, 1, model)
    ref_all = cpu_reference_run(tr, te, nu, nv, k, 4, tr.nratings, cores)
    if ref_one["rmse"]:
        rec["test_rmse_reference"] = ref_one["rmse"][-1]
        rec["rmse_abs_diff"] = abs(traj[-1] - ref_one["rmse"][-1])
        rec["reference"] = "oracle/_ref/mf_ref --fly 1, same file, same seeded model, 10 epochs"
    secs = ref_all["secs"][1:] or ref_all["secs"]
    rec["cpu_baseline"] = {"value": ref_all["n"] * len(secs) / sum(secs), "unit": UNIT, "cores": ref_all["cores"],
                           "kind": ref_all["kind"], "sample": ref_all["sample"]}
    return rec


def other_configs(mb, np, args, tr, te, device, cores, t_start):
    """sub-records for C1, C3, C4 (and C5 on this one GPU when the time budget allows)"""
    out = {}
    nu, nv, nnz, _, test_frac = WORKLOADS["netflix"]

    def guarded(name, fn):
        if time.time() - t_start > args.time_budget:
            out[name] = {"skipped": "time budget of %d s used up" % args.time_budget}
            return
        try:
            t0 = time.time()
            out[name] = fn()
            out[name]["wall_s"] = round(time.time() - t0, 1)
        except Exception as e:  # a sub-record must not take the headline down with it
            out[name] = {"error": "%s: %s" % (type(e).__name__, e)}

    guarded("C1_ml1m_mf_k32", lambda: config_mf_small(mb, np, "ml1m", device, args.steps, args.warmup, cores))

    def dp(name, gname, k):
        g = gold(gname)
        if not g:
            return {"skipped": "no golden file %s" % gname}
        traj, kms, ntrain, lam_r = config_dpmf(mb, np, tr, te, nu, nv, k, g, device)
        ref = g["test_rmse"]
        prof, prof_src = profiled("sgld_flat")
        bound = "instruction issue (Philox4x32-10 + Box-Muller per coordinate)"
        if prof and k == 128:  # the committed capture is the k = 128 launch
            try:
                inst = float(prof["smsp__inst_executed.sum"]["value"])
                issue = float(prof["smsp__issue_active.avg.pct_of_peak_sustained_active"]["value"])
                bound += ": %.0f warp instructions per record, issue slots %.0f %% busy" % (inst / ntrain, issue)
            except (KeyError, ValueError):
                pass
        return {"workload": g["config"], "ms_per_epoch": kms, "updates_per_s": ntrain / kms * 1e3,
                "kernel": "sgld_flat_kernel (sub-warp per run, flat loop)", "bound": bound, "profile": prof_src,
                "test_rmse": traj[-1], "test_rmse_reference": ref[-1], "rmse_abs_diff": abs(traj[-1] - ref[-1]),
                "rmse_trajectory": traj, "rmse_reference_trajectory": ref, "lambda_r": lam_r, "lambda_r_reference": g["lambda_r"][-1],
                "reference": "reference SgldFilter/finish_noise/sample_hyper in file order (tests/golden/fullsize/%s.json); "
                             "noise streams differ (table vs Philox): statistical agreement" % gname}

    guarded("C3_sgld_k128", lambda: dp("C3", "c3_sgld_k128_seed1", 128))
    guarded("C4_dpmf_k64_eps1", lambda: dp("C4", "c4_dpmf_k64_eps1_seed1", 64))

    def admf():
        g = gold("c4_admf_k64")
        if not g:
            return {"skipped": "no golden file"}
        trv, tev, vav = mb.generate(mb.gen_params(nu, nv, nnz, test_frac=test_frac, valid_frac=0.01))
        traj, lams, kms = config_admf(mb, np, trv, tev, vav, nu, nv, 64, g, device)
        ref = g["test_rmse"]
        return {"workload": g["config"], "ms_per_epoch": kms, "updates_per_s": trv.nratings / kms * 1e3,
                "bound": "serial instruction stream of a run (validation step per user-run) x runs in flight",
                "test_rmse": traj[-1], "test_rmse_reference": ref[-1], "rmse_abs_diff": abs(traj[-1] - ref[-1]),
                "rmse_trajectory": traj, "rmse_reference_trajectory": ref, "lams": lams[-1], "lams_reference": g["lams"][-1],
                "reference": "reference AdRegFilter + updateReg in file order (tests/golden/fullsize/c4_admf_k64.json); the "
                             "validation draws differ (rand() vs numpy): statistical agreement"}

    guarded("C4_admf_k64", admf)

    def yahoo():
        yu, yv, ynnz, yk, ytf = WORKLOADS["yahoo"]
        try_, tey, _ = mb.generate(mb.gen_params(yu, yv, ynnz, test_frac=ytf))
        c = mb.Context(yu, yv, yk, device)
        c.init_normal(0x4D46B200, 1e-2)
        d, dte = c.dataset_from_blocks(try_), c.dataset_from_blocks(tey)
        ms = []
        for ep in range(1, args.warmup + 3):
            c.sgd_epoch(d, mb.seteta(ETA0, ep, GAM), LAMBDA, GB, mb.MODE_ATOMIC)
            ms.append(c.last_kernel_ms())
        rm = c.rmse(dte, GB)
        shape = c.last_launch()
        # end to end from pinned host memory: item ids beyond 65,535, so the packed records take 4 bytes (16 + 8 bits of item
        # id, 8 bits of rating code) instead of 3
        e2e = None
        try:
            try_.pin()
            secs, h0 = [], c.h2d_bytes()
            for ep in range(len(ms) + 1, len(ms) + 4):
                c.sync()
                t0 = time.perf_counter()
                c.sgd_epoch_from_host(d, try_, mb.seteta(ETA0, ep, GAM), LAMBDA, GB, mb.MODE_ATOMIC, args.chunk)
                c.sse(dte, GB)
                secs.append(time.perf_counter() - t0)
            h2d_y = (c.h2d_bytes() - h0) // 3
            e2e = {"value": try_.nratings / min(secs[1:]), "unit": UNIT, "ms_per_step": 1e3 * min(secs[1:]),
                   "h2d_bytes_per_step": h2d_y, "d2h_bytes_per_step": 8,
                   "record_bytes": round((h2d_y - 8 * try_.nruns) / try_.nratings, 2),
                   "what": "mfb_sgd_epoch_from_host + mfb_sse; 624,961 items do not fit 16-bit ids: the packed records carry a "
                           "third id byte (4 bytes per record; 8 when the ratings have more than 256 distinct values)"}
            try_.unpin()
        except Exception as e:
            e2e = {"error": "%s: %s" % (type(e).__name__, e)}
        c.close()
        kms = sum(ms[args.warmup:]) / len(ms[args.warmup:])
        peak = measured_peak()[0]
        ach = try_.nratings * bytes_per_update(yk) / (kms * 1e-3) / 1e9
        prof, prof_src = profiled("sgd_stream_yahoo")
        traffic = prof.get("dram_traffic_bytes") if prof else None
        return {"workload": workload_config("yahoo", 1)["workload"] + " on ONE GPU (the item matrix, 320 MB, does not fit the L2)",
                "ms_per_epoch": kms, "updates_per_s": try_.nratings / kms * 1e3, "launch": shape,
                "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                             "traffic": traffic, "traffic_source": prof_src,
                             "dram_gbs": traffic / (kms * 1e-3) / 1e9 if traffic else None,
                             "dram_frac_of_peak": traffic / (kms * 1e-3) / 1e9 / peak if traffic else None,
                             "note": "achieved = algorithmic bytes (12 + 16k per update); dram_gbs = what ncu saw move (user rows stay "
                                     "in registers for a run, part of the 320 MB item matrix is served by the 126 MB L2): this "
                                     "shape reads and writes DRAM at a fifth of its bandwidth - the L2 atomic path binds here too "
                                     "(busiest slice's atomic unit 76 % busy, profiles/r2_sgd_stream_yahoo.md)"},
                "e2e": e2e, "test_rmse_after_%d_epochs" % len(ms): rm}

    if args.yahoo:
        guarded("C5_yahoo_mf_k128_1gpu", yahoo)
    return out


# ------------------------------------------------------------------------------------ GPU arm
def run_b200_arm(args, wl):
    import numpy as np
    import torch

    sys.path.insert(0, PKG)
    import mfb200 as mb
    nu, nv, nnz, k, test_frac = WORKLOADS[wl]
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus != world:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node %d" % args.gpus)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback for the product path")
    torch.cuda.set_device(local)
    if world > 1:
        import mfb_dsgd
        return mfb_dsgd.bench(args, wl, WORKLOADS[wl], rank, world, local, workload_config(wl, world))

    t_start = time.time()
    cores = os.cpu_count() or 1
    t0 = time.time()
    tr, te, _ = mb.generate(mb.gen_params(nu, nv, nnz, test_frac=test_frac))
    gen_s = time.time() - t0
    ntrain = tr.nratings
    c = mb.Context(nu, nv, k, local)
    stream = torch.cuda.current_stream()
    c.set_stream(stream.cuda_stream)
    t0 = time.time()
    dtr, dte = c.dataset_from_blocks(tr), c.dataset_from_blocks(te)
    ingest_s = time.time() - t0
    mode = {"hogwild": mb.MODE_HOGWILD, "atomic": mb.MODE_ATOMIC}[args.schedule]
    launches0 = c.launch_count()

    # ---- leg 0: parity with the reference's own trajectory (same data, same seeded model) ------------------
    g = gold("c2_mf_k128") if wl == "netflix" else None
    seed = g["model_seed"] if g else 20261018
    model = mb.seeded_model(nu, nv, k, seed)
    parity = None
    if g and not args.no_parity:
        c.set_factors(*model)
        traj, cum_ms = [], []
        for ep in range(1, len(g["test_rmse"]) + 1):
            c.sgd_epoch(dtr, mb.seteta(ETA0, ep, GAM), LAMBDA, GB, mode)
            cum_ms.append((cum_ms[-1] if cum_ms else 0.0) + c.last_kernel_ms())
            traj.append(c.rmse(dte, GB))
        ref = g["test_rmse"]
        reach = first_epoch_within(traj, ref[-1])
        parity = {"test_rmse": traj[-1], "test_rmse_reference": ref[-1], "rmse_abs_diff": abs(traj[-1] - ref[-1]),
                  "rmse_max_abs_diff_over_epochs": max(abs(a - b) for a, b in zip(traj, ref)),
                  "rmse_trajectory": traj, "rmse_reference_trajectory": ref,
                  "epoch_ms": [b - a for a, b in zip([0.0] + cum_ms[:-1], cum_ms)],
                  "epochs_to_rmse": reach, "seconds_to_rmse": cum_ms[reach - 1] * 1e-3 if reach else None,
                  "rmse_target": "reference's tRMSE after epoch %d (%.6f) + %g" % (len(ref), ref[-1], RMSE_TOL),
                  "reference": "oracle/_ref/mf_ref --fly 1 (the reference's own main(), single-thread update order) on the "
                               "same generated file from the same seeded model: tests/golden/fullsize/c2_mf_k128.json; its "
                               "own --fly 8 runs end within 1e-5 of it (c2_mf_k128_fly8.json)"}

    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    epoch = [0]
    c.set_factors(*model)  # the timed legs start over from the same model (a new model: the run bound applies again)

    def step_resident():
        epoch[0] += 1
        c.sgd_epoch(dtr, mb.seteta(ETA0, epoch[0], GAM), LAMBDA, GB, mode)

    # ---- leg 1: tiles resident in HBM ---------------------------------------------------------
    for _ in range(args.warmup):
        step_resident()
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    sampler.start()
    launches_a = c.launch_count()
    ev[0].record(stream)
    for _ in range(args.steps):
        step_resident()
    ev[1].record(stream)
    torch.cuda.synchronize()
    launches_resident = c.launch_count() - launches_a
    total_ms = ev[0].elapsed_time(ev[1])
    # the kernel's own duration, launch by launch (events recorded by the library around the
    # kernel on the same stream): a second, identical timed pass keeps the first one unperturbed
    kern_ms, shapes = [], []
    for _ in range(args.steps):
        step_resident()
        kern_ms.append(c.last_kernel_ms())
        shapes.append(c.last_launch())
    ms_per_step = total_ms / args.steps
    value = ntrain * args.steps / (total_ms * 1e-3)
    peak, peak_src = measured_peak()
    kavg = sum(kern_ms) / len(kern_ms)
    rmse_resident = c.rmse(dte, GB)

    # ---- leg 2: end to end from pinned host memory -------------------------------------------
    tr.pin()
    sse_host = []

    def step_e2e():
        epoch[0] += 1
        c.sgd_epoch_from_host(dtr, tr, mb.seteta(ETA0, epoch[0], GAM), LAMBDA, GB, mode, args.chunk)
        sse_host.append(c.sse(dte, GB)[0])  # D2H read of the step's result (synchronises)

    for _ in range(max(1, args.warmup // 2)):
        step_e2e()
    torch.cuda.synchronize()
    h2d0 = c.h2d_bytes()
    launches_b = c.launch_count()
    t0 = time.perf_counter()
    ev[0].record(stream)
    for _ in range(args.steps):
        step_e2e()
    ev[1].record(stream)
    torch.cuda.synchronize()
    e2e_wall = time.perf_counter() - t0
    clocks = sampler.stop()  # sampled over both timed legs (resident and end to end)
    h2d = (c.h2d_bytes() - h2d0) // args.steps
    e2e_ms = max(ev[0].elapsed_time(ev[1]), 1e3 * e2e_wall)
    e2e_value = ntrain * args.steps / (e2e_ms * 1e-3)
    launches_e2e = c.launch_count() - launches_b
    launches = launches_resident + launches_e2e  # kernels launched inside the two timed regions
    final_rmse = float(np.sqrt(sse_host[-1] / te.nratings))
    tr.unpin()

    # ---- leg 2b: out of core, straight from the protobuf file: the raw bytes go to the GPU and are decoded there ----
    from_file = None
    if not args.no_file:
        tmp = scratch_dir()
        try:
            fp = tr.write(os.path.join(tmp, "train.bin"))
            fbytes = os.path.getsize(fp)

            def file_epochs(reps):
                secs, h0 = [], c.h2d_bytes()
                for _ in range(reps):
                    epoch[0] += 1
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    n = c.sgd_epoch_from_file(fp, mb.seteta(ETA0, epoch[0], GAM), LAMBDA, GB, mode, args.tile)
                    s_ = c.sse(dte, GB)[0]
                    secs.append(time.perf_counter() - t0)
                    assert n == ntrain
                return min(secs[1:]), (c.h2d_bytes() - h0) // reps, float(np.sqrt(s_ / te.nratings))

            best, h2d_f, rmse_f = file_epochs(4)
            c.set_option("file_decode", 0)
            best_host, h2d_h, _ = file_epochs(2)
            c.set_option("file_decode", 1)
            # streaming ingest: the same pass over the file that also leaves a finalized, resident dataset behind
            ingest = None
            try:
                epoch[0] += 1
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                ds_i, n_i = c.dataset_ingest_file(fp, True, mb.seteta(ETA0, epoch[0], GAM), LAMBDA, GB, mode, args.tile)
                c.sync()
                ingest = {"ms": 1e3 * (time.perf_counter() - t0), "ratings": n_i, "runs": c.num_runs(ds_i),
                          "what": "mfb_dataset_ingest_file with the epoch fused: file -> finalized resident dataset + one SGD epoch, "
                                  "one pass (the mf driver's default way in); compare ingest_s (host arrays -> HBM) + an epoch"}
                c.dataset_free(ds_i)
            except Exception as e:
                ingest = {"error": "%s: %s" % (type(e).__name__, e)}
            from_file = {"value": ntrain / best, "unit": UNIT, "ms_per_step": 1e3 * best, "file_bytes": fbytes,
                         "file_gb_per_s": fbytes / best / 1e9,
                         "h2d_bytes_per_step": h2d_f, "tile_ratings": args.tile or (8 << 20), "host_cores": cores,
                         "test_rmse": rmse_f, "streaming_ingest": ingest,
                         "host_decode": {"value": ntrain / best_host, "ms_per_step": 1e3 * best_host, "h2d_bytes_per_step": h2d_h,
                                         "what": "option file_decode = 0: frames decoded by the host cores into pinned SoA chunks "
                                                 "(the round's first version of this path)"},
                         "what": "mfb_sgd_epoch_from_file: [u32][mf.Block] file (page cache) -> pread of whole frames into pinned "
                                 "chunks, one jump per user on the host -> H2D of the RAW bytes -> varints / records decoded on the "
                                 "GPU (mfb_wire_decode.cu) -> kernel; three slots of device buffers, nothing resident; + mfb_sse. "
                                 "Wall clock, best of 3 after one warm-up"}
        except Exception as e:
            from_file = {"error": "%s: %s" % (type(e).__name__, e)}
        finally:
            shutil.rmtree(tmp, ignore_errors=True)

    # ---- leg 3: the reference's CPU path on a bounded sample ---------------------------------
    cpu = None
    if not args.no_cpu:
        r = cpu_reference_run(tr, te, nu, nv, k, 2, min(ntrain, args.cpu_sample or 20_000_000), cores)
        secs = r["secs"][1:] or r["secs"]
        cpu = {"value": r["n"] * len(secs) / sum(secs), "unit": UNIT, "cores": r["cores"], "kind": r["kind"],
               "sample": r["sample"] + "; epoch 2 timed by the reference's own clock (incl. file re-read, parse, test eval)",
               "test_rmse_on_sample": r["rmse"][-1] if r["rmse"] else None}

    # ---- roofline of the update kernel ------------------------------------------------------------------------
    row_bytes = 4 * mb.lib().mfb_padding(k)
    upd_per_s = ntrain / (kavg * 1e-3)
    hbm_alg = upd_per_s * bytes_per_update(k) / 1e9
    prof, prof_src = profiled("sgd_stream") if wl == "netflix" else (None, None)
    traffic = prof.get("dram_traffic_bytes") if prof else None
    lp = l2_peak() if (wl == "netflix" and row_bytes == 512) else None
    if lp:
        # one update = one row gathered from the L2 + one row of increments reduced into it
        l2_achieved = upd_per_s * 2 * row_bytes / 1e9
        l2_peak_gbs = lp["uniform"]["both_grows_per_s"] * 2 * row_bytes
        roofline = {"bound": "l2_atomic", "achieved": l2_achieved, "peak": l2_peak_gbs, "unit": "GB/s",
                    "frac": l2_achieved / l2_peak_gbs, "traffic": traffic, "traffic_source": prof_src,
                    "peak_source": "profiles/r2_l2_atomic_peak.json: tools/l2_atomic_peak.cu on a B200 of this pool - gather "
                                   "(ld.global.cg.v4) + reduction (red.global.add.v4.f32) of one 512-B row per update into an "
                                   "L2-resident 9.1 MB matrix, rows spread uniformly, no arithmetic",
                    "achieved_g_updates_per_s": upd_per_s / 1e9,
                    "peak_g_updates_per_s": {"uniform_rows": lp["uniform"]["both_grows_per_s"],
                                             "this_files_item_popularity": lp["training_file"]["both_grows_per_s"],
                                             "reduction_only_uniform": lp["uniform"]["red_grows_per_s"]},
                    "frac_of_this_files_pattern": upd_per_s / 1e9 / lp["training_file"]["both_grows_per_s"],
                    "dram_gbs": traffic / (kavg * 1e-3) / 1e9 if traffic else None,
                    "l2_hit_rate": (float(prof["lts__t_sector_hit_rate.pct"]["value"]) / 100.0
                                    if prof and "lts__t_sector_hit_rate.pct" in prof else None),
                    "hbm_algorithmic": {"achieved": hbm_alg, "peak": peak, "unit": "GB/s", "frac": hbm_alg / peak,
                                        "peak_source": peak_src, "bytes_per_update": bytes_per_update(k),
                                        "note": "SURVEY 8d's figure (rating record + two factor rows read and written per "
                                                "update); not a bound of this kernel: user rows stay in registers for a "
                                                "run and the item matrix lives in the L2, so DRAM moves ~28 B per update"},
                    "note": "the kernel is bound by the L2: 512 B gathered + 512 B reduced per update into a 9.1 MB matrix; "
                            "frac is against the ceiling of an even spread over the L2 slices; hot item rows (0.47 % of the "
                            "records on one row) lower that ceiling to this_files_item_popularity"}
    else:
        roofline = {"bound": "hbm", "achieved": hbm_alg, "peak": peak, "unit": "GB/s", "frac": hbm_alg / peak,
                    "traffic": traffic, "traffic_source": prof_src, "peak_source": peak_src,
                    "note": "algorithmic bytes (rating record + two factor rows read and written per update)"}
    roofline.update({"kernel": KERNEL_NAMES.get(shapes[-1]["kernel"], "?"), "launch": shapes[-1], "kernel_ms": kavg,
                     "bytes_per_update": bytes_per_update(k), "updates_per_launch": ntrain})

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(wl, 1),
        "roofline": roofline,
        "cpu_baseline": cpu,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 8,
                "ms_per_step": e2e_ms / args.steps,
                "what": "mfb_sgd_epoch_from_host (pinned host tiles, compact 3-byte records when the data allow -> "
                        "chunked H2D overlapped with the kernel, expanded on the device) + mfb_sse"},
        "e2e_from_file": from_file,
        "clocks": clocks, "gpu_launches": launches,
        "gpu_launches_detail": {"resident_leg": launches_resident, "e2e_leg": launches_e2e,
                                "whole_run": c.launch_count() - launches0},
        "placement": dict(zip(("calibration_ms", "kept"), c.placement_report())),
        "test_rmse": parity["test_rmse"] if parity else final_rmse,
        "test_rmse_reference": parity["test_rmse_reference"] if parity else None,
        "rmse_abs_diff": parity["rmse_abs_diff"] if parity else None,
        "parity": parity,
        "test_rmse_after_all_legs": final_rmse, "test_rmse_after_resident_leg": rmse_resident,
        "epochs_run": epoch[0], "train_ratings": ntrain, "gen_s": round(gen_s, 2), "ingest_s": round(ingest_s, 2),
    }
    c.close()
    if wl == "netflix" and not args.no_configs:
        line["configs"] = other_configs(mb, np, args, tr, te, local, cores, t_start)
    line["wall_s"] = round(time.time() - t_start, 1)
    print(json.dumps(line))


def main():
    # The contract is ONE JSON line on stdout.  Libraries loaded later (NCCL prints its version
    # banner to stdout when NCCL_DEBUG is set in the environment) must not be able to add to it:
    # file descriptor 1 is pointed at stderr for the whole run and the line goes to the saved one.
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = real_stdout
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=None, choices=list(WORKLOADS))
    ap.add_argument("--schedule", default="atomic", choices=["hogwild", "atomic"])
    ap.add_argument("--chunk", type=int, default=0, help="ratings per H2D chunk in the e2e leg")
    ap.add_argument("--tile", type=int, default=0, help="ratings per device tile buffer in the from-file leg")
    ap.add_argument("--cpu-sample", type=int, default=0,
                    help="ratings in the CPU sample (b200 arm: default 20M; reference arm: default the whole file)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-file", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the sub-records of the other BASELINE configs")
    ap.add_argument("--yahoo", type=int, default=1, help="include the Yahoo-shaped sub-record on this GPU (0 = skip)")
    ap.add_argument("--time-budget", type=int, default=170, help="seconds after which remaining sub-records are skipped")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3  # timing rule: W >= 3
    wl = args.workload or "netflix"
    if args.impl == "reference":
        run_reference_arm(args, wl)
    else:
        run_b200_arm(args, wl)


if __name__ == "__main__":
    main()
