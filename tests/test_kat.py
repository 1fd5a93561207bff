"""Known-answer tests that pin the pieces the reference has no tests for."""
import ctypes as C
import struct

import numpy as np
import pytest

import oraclelib as ol
from oraclelib import Dataset, _p, f32p

# one Block, one User uid=300, records (vid 11, 5.0), (vid 21, 3.0), produced with
# Python google.protobuf from blocks.proto:1-18
KAT_FRAME = bytes.fromhex("17000000" "0a15" "08ac02" "1207" "080b" "15" "0000a040" "1207" "0815" "15"
                          "00004040")


def test_wire_kat_decode(oracle_lib, tmp_path):
    p = tmp_path / "kat.bin"
    p.write_bytes(KAT_FRAME)
    ds = Dataset.read(str(p))
    assert ds.nblocks == 1 and ds.nruns == 1
    assert ds.run_uid.tolist() == [300] and ds.vid.tolist() == [11, 21]
    assert ds.rating.tolist() == [5.0, 3.0]


def test_wire_kat_encode(oracle_lib, tmp_path):
    ds = Dataset([0, 1], [300], [0, 2], [11, 21], [5.0, 3.0])
    p = tmp_path / "kat.bin"
    ds.write(str(p))
    assert p.read_bytes() == KAT_FRAME


def test_wire_against_python_protobuf(oracle_lib, tmp_path):
    """Encode with the real protobuf runtime (dynamic descriptor of blocks.proto:1-18)."""
    pb = pytest.importorskip("google.protobuf")
    from google.protobuf import descriptor_pb2, descriptor_pool, message_factory
    fdp = descriptor_pb2.FileDescriptorProto()
    fdp.name, fdp.package, fdp.syntax = "blocks.proto", "mf", "proto2"
    user = fdp.message_type.add()
    user.name = "User"
    f = user.field.add()
    f.name, f.number, f.label, f.type = "uid", 1, f.LABEL_REQUIRED, f.TYPE_INT32
    rec = user.nested_type.add()
    rec.name = "Record"
    f = rec.field.add()
    f.name, f.number, f.label, f.type = "vid", 1, f.LABEL_REQUIRED, f.TYPE_INT32
    f = rec.field.add()
    f.name, f.number, f.label, f.type = "rating", 2, f.LABEL_REQUIRED, f.TYPE_FLOAT
    f = user.field.add()
    f.name, f.number, f.label, f.type, f.type_name = "record", 2, f.LABEL_REPEATED, f.TYPE_MESSAGE, ".mf.User.Record"
    blk = fdp.message_type.add()
    blk.name = "Block"
    f = blk.field.add()
    f.name, f.number, f.label, f.type, f.type_name = "user", 1, f.LABEL_REPEATED, f.TYPE_MESSAGE, ".mf.User"
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fdp)
    Block = message_factory.GetMessageClass(pool.FindMessageTypeByName("mf.Block"))

    rng = np.random.default_rng(0)
    frames, want = b"", []
    for _ in range(3):
        b = Block()
        for _ in range(int(rng.integers(0, 5))):
            u = b.user.add()
            u.uid = int(rng.integers(-3, 1 << 20))  # negative ids are 10-byte varints
            recs = []
            for _ in range(int(rng.integers(0, 6))):
                r = u.record.add()
                r.vid = int(rng.integers(0, 1 << 17))
                r.rating = float(rng.integers(1, 6))
                recs.append((r.vid, r.rating))
            want.append((u.uid, recs))
        s = b.SerializeToString()
        frames += struct.pack("<I", len(s)) + s
    p = tmp_path / "pb.bin"
    p.write_bytes(frames)
    ds = Dataset.read(str(p))
    assert ds.nblocks == 3 and ds.nruns == len(want)
    for i, (uid, recs) in enumerate(want):
        assert ds.run_uid[i] == uid
        lo, hi = ds.run_off[i], ds.run_off[i + 1]
        assert list(zip(ds.vid[lo:hi].tolist(), ds.rating[lo:hi].tolist())) == recs
    q = tmp_path / "again.bin"
    ds.write(str(q))
    assert q.read_bytes() == frames  # our writer is byte-identical to protobuf's serializer


def test_philox4x32_10_kat(oracle_lib):
    # Random123 kat_vectors (philox4x32 10 rounds)
    def ph(c, k):
        cc, kk, o = (C.c_uint32 * 4)(*c), (C.c_uint32 * 2)(*k), (C.c_uint32 * 4)()
        oracle_lib.mfo_philox4x32_10(cc, kk, o)
        return [int(x) for x in o]
    assert ph([0] * 4, [0] * 2) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert ph([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert ph([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_philox_normals_moments(oracle_lib):
    z = np.zeros(4, np.float32)
    out = []
    for t in range(20000):
        oracle_lib.mfo_philox_normal4(0x4D46B200, 1, t & 1, t % 97, t, t % 32, _p(z, f32p))
        out.append(z.copy())
    x = np.concatenate(out).astype(np.float64)
    assert abs(x.mean()) < 0.02 and abs(x.var() - 1) < 0.03
    assert abs((x ** 3).mean()) < 0.06 and abs((x ** 4).mean() - 3) < 0.15


def test_philox_bias_normal_moments_and_independence(oracle_lib):
    """The bias normal of (kind,row,t) is built from the low bytes of chunks 0 and 1 - bits the
    factor normals (top 24 bits of each word) do not use: N(0,1) moments, and no correlation with
    the factor normals of the same two chunks."""
    z0, z1 = np.zeros(4, np.float32), np.zeros(4, np.float32)
    b, f = [], []
    for t in range(20000):
        args = (0x4D46B200, 2, t & 1, t % 89, t)
        b.append(oracle_lib.mfo_philox_bias_normal(*args))
        oracle_lib.mfo_philox_normal4(*args, 0, _p(z0, f32p))
        oracle_lib.mfo_philox_normal4(*args, 1, _p(z1, f32p))
        f.append(np.concatenate([z0, z1]))
    b, f = np.array(b, np.float64), np.array(f, np.float64)
    assert abs(b.mean()) < 0.03 and abs(b.var() - 1) < 0.04
    assert abs((b ** 3).mean()) < 0.1 and abs((b ** 4).mean() - 3) < 0.25
    corr = [abs(np.corrcoef(b, f[:, i])[0, 1]) for i in range(8)]
    assert max(corr) < 0.03, corr
    corr2 = [abs(np.corrcoef(b ** 2, f[:, i] ** 2)[0, 1]) for i in range(8)]
    assert max(corr2) < 0.03, corr2


def test_seteta_formula(oracle_lib):
    # model.cc:36-38 eta = eta0 / round^gam in double, narrowed
    for eta0, rnd, gam in ((2e-2, 1, 1.0), (2e-2, 7, 1.0), (4e-2, 3, 0.6)):
        want = np.float32(np.float64(np.float32(eta0)) / np.float64(rnd) ** np.float64(np.float32(gam)))
        assert np.float32(oracle_lib.mfo_seteta(eta0, rnd, gam)) == want
    assert oracle_lib.mfo_seteta_cutoff(2e-2, 1000000, 3.0, 1e-13) == pytest.approx(1e-13)


def test_one_hand_computed_sgd_step(oracle_lib):
    """mf.h:94-109 on one rating, against the closed form theta'=lameta*theta+e*phi, phi'=lameta*phi+e*theta."""
    m = ol.Model(1, 1, 4, seed=1, scale=0.5)
    th, ph, bu, bv = m.theta[0, :4].copy(), m.phi[0, :4].copy(), m.bu[0], m.bv[0]
    ds = Dataset([0, 1], [0], [0, 1], [0], [4.0])
    eta, lam, gb = np.float32(0.05), np.float32(0.1), np.float32(2.76)
    mm, dd = m.as_mfo(), ds.as_mfo()
    oracle_lib.mfo_sgd_epoch(C.byref(mm), C.byref(dd), eta, lam, gb)
    lameta = np.float32(1.0 - np.float64(eta * lam))
    e = eta * (np.float32(4.0) - np.dot(th.astype(np.float64), ph) - bu - bv - gb)
    np.testing.assert_allclose(m.theta[0, :4], lameta * th + e * ph, rtol=2e-6)
    np.testing.assert_allclose(m.phi[0, :4], lameta * ph + e * th, rtol=2e-6)
    np.testing.assert_allclose(m.bu[0], lameta * bu + e, rtol=2e-6)
    np.testing.assert_allclose(m.bv[0], lameta * bv + e, rtol=2e-6)
