"""checkpoints of the product's host mirror (experimental-mf_b200/csrc/model.cc MF / DPMF
save_model, read_model, read_hyper) against the REFERENCE's own save_model / read_model / read_hyper
(model.cc:75-195, run through oracle/_ref/libmf_ref.so) - byte for byte, both directions - and the resume
state (round, step size) the reference's format does not carry."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import mfb200 as mb
import oraclelib as ol
from oraclelib import _p, f32p

pytestmark = pytest.mark.gpu
TOOL = os.path.join(mb.HERE, "ckpt_tool")
MF = os.path.join(mb.HERE, "mf")
NU, NV, DIM, GB = 150, 70, 20, 2.76


@pytest.fixture(scope="module")
def files(tmp_path_factory):
    d = tmp_path_factory.mktemp("ckpt")
    train, test, _ = ol.make_ratings(NU, NV, 4000, seed=21)
    return str(d), train.write(str(d / "train")), test.write(str(d / "test"))


def need_ref():
    if not ol.have_ref():
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")


@pytest.mark.parametrize("alg", ["mf", "dpmf"])
def test_checkpoint_bytes_round_trip_with_the_reference(files, alg):
    need_ref()
    assert os.path.exists(TOOL), "run __graft_entry__.build()"
    d, tp, sp = files
    R = ol.ref()
    m = ol.Model(NU, NV, DIM, seed=33, scale=0.2)
    th, ph = m.dense()
    # 1. the reference writes a checkpoint of a seeded model
    R.ref_set_io_paths(("%s/ref_%s" % (d, alg)).encode(), None)
    if alg == "mf":
        h = R.ref_create_mf(tp.encode(), sp.encode(), DIM, 2e-2, 1.0, 7.5e-3, GB, NU, NV)
    else:
        h = R.ref_create_dpmf(tp.encode(), sp.encode(), DIM, 2e-10, 1.0, 5e-3, GB, NU, NV, 1.0, 100.0, 0.0, 0,
                              NV * (DIM + 1) + 20000, 1.0, 1e-13)
        hyp = np.r_[2.5, 31.0, 47.0, np.linspace(1, 2, DIM), np.linspace(3, 4, DIM)].astype(np.float32)
        R.ref_dpmf_set_hyper(h, _p(hyp, f32p))
    r = ol.Ref(h, NU, NV, DIM)
    r.set_factors(th, ph, m.bu, m.bv)
    R.ref_save_model(h, 7)
    ref_file = "%s/ref_%s_7" % (d, alg)
    # 2. the product reads it and writes it back: same bytes; no sidecar -> a fresh run (round 0)
    out = subprocess.run([TOOL, alg, tp, str(NU), str(NV), str(DIM), ref_file, "%s/ours_%s" % (d, alg), "9"],
                         capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    assert "start_round 0" in out.stdout
    ours = "%s/ours_%s_9" % (d, alg)
    assert open(ours, "rb").read() == open(ref_file, "rb").read()
    # 3. the sidecar carries the round; a second load of OUR checkpoint resumes there
    state = open(ours + ".state").read()
    assert "round 9" in state
    out = subprocess.run([TOOL, alg, tp, str(NU), str(NV), str(DIM), ours, "%s/again_%s" % (d, alg), "10"],
                         capture_output=True, text=True)
    assert out.returncode == 0 and "start_round 9" in out.stdout
    # 4. the reference reads the product's file: same factors (and hyper-parameters)
    R.ref_set_io_paths(None, ours.encode())
    if alg == "mf":
        h2 = R.ref_create_mf(tp.encode(), sp.encode(), DIM, 2e-2, 1.0, 5e-3, GB, NU, NV)
    else:
        h2 = R.ref_create_dpmf(tp.encode(), sp.encode(), DIM, 2e-10, 1.0, 5e-3, GB, NU, NV, 1.0, 100.0, 0.0, 0,
                               NV * (DIM + 1) + 20000, 1.0, 1e-13)
    R.ref_read_model(h2)
    r2 = ol.Ref(h2, NU, NV, DIM)
    t2, p2, bu2, bv2 = r2.get_factors()
    np.testing.assert_array_equal(t2, th)
    np.testing.assert_array_equal(p2, ph)
    np.testing.assert_array_equal(bu2, m.bu)
    np.testing.assert_array_equal(bv2, m.bv)
    if alg == "mf":
        assert np.float32(R.ref_get_lambda(h2)) == np.float32(7.5e-3)  # read_model adopts the file's lambda (model.cc:81)
    else:
        got = np.zeros(3 + 2 * DIM, np.float32)
        R.ref_dpmf_get_hyper(h2, _p(got, f32p))
        np.testing.assert_array_equal(got, hyp)
        # read_hyper (main.cc:57): the hyper-parameters alone
        h3 = R.ref_create_dpmf(tp.encode(), sp.encode(), DIM, 2e-10, 1.0, 5e-3, GB, NU, NV, 1.0, 100.0, 0.0, 0,
                               NV * (DIM + 1) + 20000, 1.0, 1e-13)
        R.ref_dpmf_read_hyper(h3)
        got3 = np.zeros(3 + 2 * DIM, np.float32)
        R.ref_dpmf_get_hyper(h3, _p(got3, f32p))
        np.testing.assert_array_equal(got3, hyp)
    R.ref_set_io_paths(None, None)


def test_resumed_run_continues_with_the_next_rounds_step_size(files):
    """mf --iter 4 in one go == mf --iter 2 (checkpoint) + mf --iter 4 --model <checkpoint> (ordered schedule:
    bit-exact); without the sidecar the reload starts over at eta0, as the reference would."""
    d, tp, sp = files
    common = ["--alg", "mf", "--train", tp, "--test", sp, "--nu", NU, "--nv", NV, "--dim", DIM, "--fly", 1,
              "--eta", 3e-2, "--lambda", 5e-3, "--gam", 1.0, "--bias", GB]

    def mf(*extra, env=None):
        out = subprocess.run([MF] + [str(a) for a in common + list(extra)], capture_output=True, text=True,
                             env=dict(os.environ, MF_SEED="5", **(env or {})))
        assert out.returncode == 0, out.stderr
        return [float(x.split("tRMSE=")[1]) for x in out.stdout.splitlines() if "tRMSE=" in x]

    whole = mf("--iter", 4)
    first = mf("--iter", 2, "--result", d + "/half", env={"MF_SAVE_EVERY": "2"})
    assert first == whole[:2]
    resumed = mf("--iter", 4, "--model", d + "/half_2")
    assert resumed == whole[2:]
    os.unlink(d + "/half_2.state")
    restarted = mf("--iter", 2, "--model", d + "/half_2")
    assert restarted != whole[2:] and len(restarted) == 2


def test_out_of_core_driver_prints_the_resident_drivers_lines(files):
    """MF_TILE_RATINGS: the `mf` driver trains straight from the file (two small device tile buffers), as the reference
    does; in the ordered schedule the printed test RMSE is the resident run's, digit for digit."""
    d, tp, sp = files
    common = [MF, "--alg", "mf", "--train", tp, "--test", sp, "--nu", NU, "--nv", NV, "--dim", DIM, "--fly", 1, "--iter", 3,
              "--eta", 3e-2, "--lambda", 5e-3, "--gam", 1.0, "--bias", GB]
    outs = []
    for env in ({}, {"MF_TILE_RATINGS": "1024"}):
        out = subprocess.run([str(a) for a in common], capture_output=True, text=True, env=dict(os.environ, MF_SEED="5", **env))
        assert out.returncode == 0, out.stderr
        outs.append([x.split("tRMSE=")[1] for x in out.stdout.splitlines() if "tRMSE=" in x])
    assert len(outs[0]) == 3 and outs[0] == outs[1]
