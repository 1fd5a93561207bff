"""The getdata converter (reference: data/getdata.cc) - host only, no GPU."""
import os
import subprocess

import numpy as np

import mfb200 as mb

HERE = os.path.dirname(os.path.abspath(__file__))
GETDATA = os.path.join(os.path.dirname(HERE), "experimental-mf_b200", "getdata")


def run(*args):
    return subprocess.run([GETDATA, *args], capture_output=True, text=True)


def test_protobuf_method_writes_the_reference_frame(tmp_path):
    """SURVEY 8c known-answer frame: one Block, one User uid=300, records (11, 5.0), (21, 3.0)."""
    txt, out = tmp_path / "u.txt", tmp_path / "b.bin"
    txt.write_text("300:\n11,5.000000\n21,3.000000\n")
    assert run("-r", str(txt), "-w", str(out), "--method", "protobuf").returncode == 0
    want = bytes.fromhex("17000000" "0a15" "08ac02" "1207" "080b" "15" "0000a040" "1207" "0815" "15" "00004040")
    assert out.read_bytes() == want


def test_userwise_then_protobuf_keeps_every_rating_and_block_size(tmp_path):
    rng = np.random.default_rng(0)
    n, nu, nv = 5000, 200, 90
    u, v = rng.integers(0, nu, n), rng.integers(0, nv, n)
    r = rng.integers(1, 6, n).astype(np.float32)
    raw, txt, out = tmp_path / "raw.txt", tmp_path / "user.txt", tmp_path / "train.bin"
    raw.write_text("%d\n" % n + "".join("%d,%d,%g,%d\n" % (a, b, c, 0) for a, b, c in zip(u, v, r)))
    assert run("-r", str(raw), "-w", str(txt), "--method", "userwise", "--split", "4").returncode == 0
    assert run("-r", str(txt), "-w", str(out), "--method", "protobuf", "--size", "37").returncode == 0
    b = mb.Blocks.read(str(out))
    assert b.nratings == n
    # Blocks of exactly --size users except the last (getdata.cc:98-110)
    users_per_block = np.diff(b.block_off)
    assert (users_per_block[:-1] == 37).all() and 0 < users_per_block[-1] <= 37
    # each of the 4 chunks holds every user at most once => a user appears at most 4 times
    assert np.bincount(b.run_uid).max() <= 4
    got = np.stack([np.repeat(b.run_uid, np.diff(b.run_off)), b.vid, b.rating.astype(np.int64)], 1)
    want = np.stack([u, v, r.astype(np.int64)], 1)
    assert sorted(map(tuple, got)) == sorted(map(tuple, want))


def test_synth_method_and_cli_errors(tmp_path):
    out = tmp_path / "ml"
    p = run("-w", str(out), "--method", "synth", "--nu", "300", "--nv", "80", "--nnz", "6000", "--valid", "0.05")
    assert p.returncode == 0 and p.stdout.startswith("train ")
    tr, te = mb.Blocks.read(str(out) + ".train"), mb.Blocks.read(str(out) + ".test")
    assert tr.nratings > 3000 and te.nratings > 0 and os.path.exists(str(out) + ".valid")
    assert run("--bogus").returncode == 1 and "unknown parameters." in run("--bogus").stdout  # getdata.cc:146-150
    assert run("-w", "x").returncode == 1  # getdata.cc:152-156
