"""The `mf` host driver (experimental-mf_b200/csrc/mf_main.cc + model.cc): the reference's command
line, messages, exit codes and output lines, checked against the reference's own binary
(oracle/_ref/mf_ref, reference main.cc compiled in place) where it exists, and the CPU oracle."""
import ctypes as C
import os
import re
import struct
import subprocess

import numpy as np
import pytest

import mfb200 as mb
import oraclelib as ol

MF = os.path.join(mb.HERE, "mf")
GB = 2.76


def run(binary, *args, **kw):
    return subprocess.run([binary] + [str(a) for a in args], capture_output=True, text=True, **kw)


def write_model_file(path, m, lam):
    """checkpoint layout of the reference, model.cc:98-122"""
    with open(path, "wb") as f:
        f.write(struct.pack("<iiif", m.nv, m.nu, m.dim, lam))
        f.write(m.bv.tobytes())
        f.write(np.ascontiguousarray(m.phi[:, :m.dim]).tobytes())
        f.write(m.bu.tobytes())
        f.write(np.ascontiguousarray(m.theta[:, :m.dim]).tobytes())


def read_model_file(path):
    raw = open(path, "rb").read()
    nv, nu, dim, lam = struct.unpack_from("<iiif", raw, 0)
    a = np.frombuffer(raw, np.float32, offset=16)
    bv, a = a[:nv], a[nv:]
    phi, a = a[:nv * dim].reshape(nv, dim), a[nv * dim:]
    bu, a = a[:nu], a[nu:]
    theta = a[:nu * dim].reshape(nu, dim)
    assert len(a) == nu * dim
    return nv, nu, dim, lam, bv, phi, bu, theta


def test_cli_error_paths_match_reference_binary():
    assert os.path.exists(MF), "run __graft_entry__.build()"
    cases = [[], ["--bogus", "1"], ["--train", "x", "--nu", "3"], ["--train", "x", "--nu", "3", "--nv", "3", "--alg", "foo"]]
    want_rc = [1, 1, 1, 2]
    for args, rc in zip(cases, want_rc):
        got = run(MF, *args)
        assert got.returncode == rc, (args, got.returncode, got.stdout)
        if os.path.exists(ol.REF_BIN):
            ref = run(ol.REF_BIN, *args)
            assert ref.returncode == got.returncode
            assert ref.stdout.splitlines()[0] == got.stdout.splitlines()[0]
    # every flag of the reference's help text is accepted
    helptext = run(MF).stdout
    for flag in ("--train", "--nu", "--nv", "--test", "--valid", "--result", "--model", "--alg", "--dim",
                 "--iter", "--fly", "--stride", "--eta", "--lambda", "--gam", "--bias", "--mineta",
                 "--epsilon", "--tau", "--temp", "--noise_size", "--eta_reg", "--loss", "--measure"):
        assert flag in helptext


def test_cli_fails_loudly_without_gpu(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    got = run(MF, "--train", tmp_path / "t", "--test", tmp_path / "s", "--nu", 5, "--nv", 5, "--alg", "mf")
    assert got.returncode == 3 and "mfb_create failed" in got.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("stream_ingest", ["1", "0"], ids=["ingest fused with epoch 1", "host parse + upload"])
def test_cli_mf_single_thread_order_prints_reference_numbers(tmp_path, stream_ingest):
    """./mf --alg mf --fly 1 --model <seeded> : same command line for the reference binary and for
    ours; the per-epoch tRMSE lines must agree to the printed precision, and the checkpoint written
    by save_model must hold the oracle's factors bit for bit.  Both ways of getting the training file
    into HBM: the streaming ingest fused with the first epoch (default) and MF_STREAM_INGEST=0."""
    nu, nv, dim = 300, 120, 32
    train, test, _ = ol.make_ratings(nu, nv, 9000, seed=5)
    tp, sp = train.write(str(tmp_path / "train")), test.write(str(tmp_path / "test"))
    m = ol.Model(nu, nv, dim, seed=2)
    eta0, gam, lam = 2e-2, 1.0, 5e-3
    mp = str(tmp_path / "model")
    write_model_file(mp, m, lam)
    args = ["--alg", "mf", "--train", tp, "--test", sp, "--nu", nu, "--nv", nv, "--dim", dim, "--iter", 3,
            "--fly", 1, "--eta", eta0, "--lambda", lam, "--gam", gam, "--bias", GB, "--model", mp,
            "--result", str(tmp_path / "out")]
    got = run(MF, *args, env=dict(os.environ, MF_SAVE_EVERY="3", MF_STREAM_INGEST=stream_ingest))
    assert got.returncode == 0, got.stderr
    lines = got.stdout.strip().splitlines()
    assert len(lines) == 3 and all(re.fullmatch(r"iter#\d+\t[0-9.]+\ttRMSE=[0-9.]+", l) for l in lines)
    rmse = [float(x) for x in re.findall(r"tRMSE=([0-9.]+)", got.stdout)]
    L = ol.oracle()
    mm, dd, tt = m.as_mfo(), train.as_mfo(), test.as_mfo()
    for ep in (1, 2, 3):
        L.mfo_sgd_epoch(C.byref(mm), C.byref(dd), L.mfo_seteta(eta0, ep, gam), np.float32(lam), GB)
        n = C.c_int64()
        s = L.mfo_sse(C.byref(mm), C.byref(tt), GB, C.byref(n))
        assert abs(np.sqrt(s * 1.0 / n.value) - rmse[ep - 1]) < 2e-6
    if os.path.exists(ol.REF_BIN):  # the reference binary travels to the GPU box prebuilt
        ref = run(ol.REF_BIN, *args)
        ref_rmse = [float(x) for x in re.findall(r"tRMSE=([0-9.]+)", ref.stdout)]
        assert np.allclose(ref_rmse, rmse, atol=2e-6, rtol=0)
    fnv, fnu, fdim, flam, bv, phi, bu, theta = read_model_file(str(tmp_path / "out_3"))
    assert (fnv, fnu, fdim) == (nv, nu, dim) and np.float32(flam) == np.float32(lam)
    np.testing.assert_array_equal(theta, m.theta[:, :dim])
    np.testing.assert_array_equal(phi, m.phi[:, :dim])
    np.testing.assert_array_equal(bu, m.bu)
    np.testing.assert_array_equal(bv, m.bv)


@pytest.mark.gpu
def test_cli_all_algorithms_run_and_learn(tmp_path):
    nu, nv, dim = 3000, 700, 32
    tr, te, va = mb.generate(mb.gen_params(nu, nv, 200_000, test_frac=0.1, valid_frac=0.03, users_per_block=100))
    tp, sp, vp = tr.write(str(tmp_path / "train")), te.write(str(tmp_path / "test")), va.write(str(tmp_path / "valid"))
    base = ["--train", tp, "--test", sp, "--nu", nu, "--nv", nv, "--dim", dim, "--iter", 6, "--bias", GB]
    env = dict(os.environ, MF_SEED="7", MF_PRINT_LAMBDA="1")
    out = run(MF, "--alg", "mf", *base, env=env)
    assert out.returncode == 0, out.stderr
    r = [float(x) for x in re.findall(r"tRMSE=([0-9.]+)", out.stdout)]
    assert len(r) == 6 and r[-1] < r[0]
    out = run(MF, "--alg", "admf", "--valid", vp, "--eta_reg", 2e-2, *base, env=env)
    assert out.returncode == 0, out.stderr
    r = [float(x) for x in re.findall(r"tRMSE=([0-9.]+)", out.stdout)]
    lam = re.findall(r"lambda#6\t(\S+)\t(\S+)\t(\S+)\t(\S+)", out.stdout)
    assert len(r) == 6 and r[-1] < r[0] and len(lam) == 1 and all(float(x) >= 0 for x in lam[0])
    ntrain = tr.nratings
    out = run(MF, "--alg", "dpmf", "--eta", 2e-2 / ntrain, "--temp", 0.01, "--gam", 0.5, "--result",
              str(tmp_path / "dp"), *base, env=env)
    assert out.returncode == 0, out.stderr
    rows = re.findall(r"round #(\d+)\tRMSE=([0-9.]+)\ttRMSE=([0-9.]+)\t([0-9.]+)", out.stdout)  # model.cc:304-308
    assert [int(x[0]) for x in rows] == [1, 2, 3, 4, 5, 6]
    # with the reference's initial precisions (lambda_u = lambda_v = 1e2, model.cc:226) six SGLD
    # rounds mostly shrink the factors; the check here is the output contract, not convergence
    assert all(0.3 < float(x[1]) < 2.0 and 0.3 < float(x[2]) < 2.0 for x in rows)
    assert float(rows[-1][3]) >= float(rows[0][3])  # cumulative seconds
