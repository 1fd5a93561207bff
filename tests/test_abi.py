"""CPU-side checks of the C-ABI library: it loads, exports every symbol include/mf_b200.h
declares, and its host-only entry points (wire codec, generator, eta schedule) agree with the
oracle.  No compute call is made here (no GPU in the build container)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import mfb200 as mb
import oraclelib as ol

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "mf_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mfb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    L = mb.lib()
    syms = header_symbols()
    assert len(syms) >= 40
    for s in syms:
        assert hasattr(L, s), "libmf_b200.so does not export %s" % s
        assert s in mb.SIGNATURES, "python binding lacks %s" % s
    assert sorted(mb.SIGNATURES) == syms  # and the binding names nothing the header lacks
    assert b"sm_100a" in L.mfb_version()


def test_library_contains_sm100a_code_only():
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", mb.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    archs = set(re.findall(r"sm_\d+a?", out.stdout))
    assert archs == {"sm_100a"}, archs


def test_padding_and_eta_match_oracle(oracle_lib):
    L = mb.lib()
    for d in (1, 15, 16, 17, 20, 32, 64, 100, 128, 2048):
        assert L.mfb_padding(d) == oracle_lib.mfo_padding(d)
    for eta0, rnd, gam in ((2e-2, 1, 1.0), (2e-2, 9, 1.0), (4e-2, 5, 0.6), (2e-10, 3, 0.3)):
        assert np.float32(L.mfb_seteta(eta0, rnd, gam)) == np.float32(oracle_lib.mfo_seteta(eta0, rnd, gam))
        assert np.float32(L.mfb_seteta_cutoff(eta0, rnd, gam, 1e-9)) == \
            np.float32(oracle_lib.mfo_seteta_cutoff(eta0, rnd, gam, 1e-9))


def test_no_silent_cpu_fallback():
    """Without a CUDA device the product refuses to run (MFB_E_CUDA + message)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    h = C.c_void_p()
    rc = mb.lib().mfb_create(C.byref(h), 0, 10, 10, 8)
    assert rc == -2 and not h.value
    assert b"cuda" in mb.lib().mfb_last_error().lower()
    with pytest.raises(mb.MfbError):
        mb.Context(10, 10, 8)


def test_wire_decoder_on_reference_parsed_bytes(golden, tmp_path):
    p = tmp_path / "train.bin"
    golden["train_bytes"].tofile(p)
    b = mb.Blocks.read(str(p))
    for name in ("block_off", "run_uid", "run_off", "vid", "rating"):
        np.testing.assert_array_equal(getattr(b, name), golden["train_" + name])
    q = tmp_path / "again.bin"
    b.write(str(q))
    assert q.read_bytes() == p.read_bytes()


def test_wire_kat_and_errors(tmp_path):
    frame = bytes.fromhex("170000000a1508ac021207080b150000a04012070815150000" "4040")
    p = tmp_path / "kat.bin"
    p.write_bytes(frame)
    b = mb.Blocks.read(str(p))
    assert (b.nblocks, b.run_uid.tolist(), b.vid.tolist(), b.rating.tolist()) == (1, [300], [11, 21], [5.0, 3.0])
    # empty file, empty block, truncated frame, missing file
    (tmp_path / "empty.bin").write_bytes(b"")
    assert mb.Blocks.read(str(tmp_path / "empty.bin")).nblocks == 0
    (tmp_path / "emptyblock.bin").write_bytes(b"\x00\x00\x00\x00")
    e = mb.Blocks.read(str(tmp_path / "emptyblock.bin"))
    assert (e.nblocks, e.nruns, e.nratings) == (1, 0, 0)
    (tmp_path / "trunc.bin").write_bytes(frame[:-3])
    with pytest.raises(mb.MfbError):
        mb.Blocks.read(str(tmp_path / "trunc.bin"))
    with pytest.raises(mb.MfbError):
        mb.Blocks.read(str(tmp_path / "nope.bin"))


def test_host_half_of_the_device_decoder_indexes_frames_and_users(tmp_path):
    """mfb_wire_index_file = the walk the out-of-core epoch does on the host before the bytes go to the GPU: frames by
    their [u32] headers, one jump per serialized mf.User inside a Block, unknown top-level fields skipped; the counts
    must equal what the full host decoder finds, and damage to the framing must be reported"""
    import struct
    tr, _, _ = mb.generate(mb.gen_params(500, 200, 20000, test_frac=0.0, users_per_block=37))
    path = tr.write(str(tmp_path / "train.bin"))
    frames, users, ubytes = mb.wire_index_file(path)
    assert (frames, users) == (tr.nblocks, tr.nruns)
    # every user: tag 08 + uid varint, every record >= 9 bytes; the Block adds tag + length per user, the file 4 per frame
    raw = open(path, "rb").read()
    assert 9 * tr.nratings + 2 * tr.nruns <= ubytes < len(raw) - 4 * frames - 2 * users + 1
    # unknown top-level fields (varint, fixed64, bytes, fixed32) and an empty Block between real ones
    u1 = b"\x08\x07" + b"\x12\x07\x08\x03\x15" + struct.pack("<f", 4.0)
    u2 = b"\x08\x2a"
    blk = b"\x10\x05" + b"\x0a" + bytes([len(u1)]) + u1 + b"\x19" + bytes(8) + b"\x22\x02hi" + b"\x0a" + bytes([len(u2)]) + u2 + b"\x2d" + bytes(4)
    p2 = tmp_path / "odd.bin"
    p2.write_bytes(struct.pack("<I", len(blk)) + blk + struct.pack("<I", 0) + struct.pack("<I", len(blk)) + blk)
    assert mb.wire_index_file(str(p2)) == (3, 4, 2 * (len(u1) + len(u2)))
    b = mb.Blocks.read(str(p2))
    assert (b.nblocks, b.nruns, b.nratings) == (3, 4, 2)
    # a user that runs past its Block, a frame that runs past the file, a group wire type, a missing file
    for name, data in (("usr.bin", struct.pack("<I", 4) + b"\x0a\x20\x08\x01"), ("frame.bin", struct.pack("<I", 99) + blk),
                       ("group.bin", struct.pack("<I", 2) + b"\x0b\x00")):
        q = tmp_path / name
        q.write_bytes(data)
        with pytest.raises(mb.MfbError):
            mb.wire_index_file(str(q))
    with pytest.raises(mb.MfbError):
        mb.wire_index_file(str(tmp_path / "nope.bin"))
    (tmp_path / "empty.bin").write_bytes(b"")
    assert mb.wire_index_file(str(tmp_path / "empty.bin")) == (0, 0, 0)


def _varint(v):
    out = bytearray()
    v &= (1 << 64) - 1
    while v >= 0x80:
        out.append((v & 0x7F) | 0x80)
        v >>= 7
    out.append(v)
    return bytes(out)


def test_wire_decoder_fast_path_and_general_path_agree(tmp_path):
    """The decoder has a fast path for the canonical record (12 LL 08 <vid> 15 <f32>) and a general
    one; any legal encoding of the same messages (field order swapped, unknown fields, over-long
    varints, item ids of every varint width) must decode to the same arrays."""
    import struct
    vids = [0, 1, 127, 128, 300, 16383, 16384, 65535, 2097151, 2097152, 268435455, 2**31 - 1]
    ratings = [float(i % 5 + 1) for i in range(len(vids))]

    def user(uid, recs):
        body = b"\x08" + _varint(uid) + b"".join(recs)
        return b"\x0a" + _varint(len(body)) + body

    canon = [b"\x12" + _varint(len(_varint(v)) + 6) + b"\x08" + _varint(v) + b"\x15" + struct.pack("<f", r)
             for v, r in zip(vids, ratings)]
    odd = []
    for i, (v, r) in enumerate(zip(vids, ratings)):
        f1, f2 = b"\x08" + _varint(v), b"\x15" + struct.pack("<f", r)
        if i % 3 == 0:
            body = f2 + f1                                  # rating before vid
        elif i % 3 == 1:
            body = f1 + b"\x18\x07" + f2                    # unknown varint field 3
        else:
            body = b"\x08" + _varint(v)[:-1] + bytes([_varint(v)[-1] | 0x80, 0x00]) + f2  # over-long varint
        odd.append(b"\x12" + _varint(len(body)) + body)
    for name, recs in (("canon", canon), ("odd", odd)):
        blk = user(7, recs) + user(9, []) + user(300, recs[:3])
        path = tmp_path / (name + ".bin")
        path.write_bytes(struct.pack("<I", len(blk)) + blk)
        b = mb.Blocks.read(str(path))
        assert b.run_uid.tolist() == [7, 9, 300]
        assert b.run_off.tolist() == [0, len(vids), len(vids), len(vids) + 3]
        assert b.vid.tolist() == vids + vids[:3]
        assert b.rating.tolist() == ratings + ratings[:3]
    # a record cut short inside the fast path's window is an error, not a read past the frame
    blk = user(7, canon)[:-2]
    (tmp_path / "short.bin").write_bytes(struct.pack("<I", len(blk)) + blk)
    with pytest.raises(mb.MfbError):
        mb.Blocks.read(str(tmp_path / "short.bin"))


def test_multi_frame_file_is_decoded_in_file_order_and_a_bad_frame_is_reported(tmp_path):
    """The loader decodes the frames of a file on worker threads and stitches them in file order:
    a file of many frames must come back exactly (block boundaries, run offsets, records), twice
    in a row into the same arrays, and a malformed frame in the middle is an error that names it."""
    import struct
    tr, _, _ = mb.generate(mb.gen_params(3000, 400, 60000, users_per_block=20))
    assert tr.nblocks > 100
    path = str(tmp_path / "many.bin")
    tr.write(path)
    back = mb.Blocks.read(path)
    for name in ("block_off", "run_uid", "run_off", "vid", "rating"):
        np.testing.assert_array_equal(getattr(back, name), getattr(tr, name))
    raw = bytearray(open(path, "rb").read())
    # walk to frame 57 and break its first user message: length byte larger than the frame
    off = 0
    for _ in range(57):
        off += 4 + struct.unpack_from("<I", raw, off)[0]
    assert raw[off + 4] == 0x0A
    raw[off + 5] = 0xFF
    raw[off + 6] = 0x7F
    bad = tmp_path / "bad.bin"
    bad.write_bytes(bytes(raw))
    with pytest.raises(mb.MfbError, match="malformed mf.Block at offset %d" % (off + 4)):
        mb.Blocks.read(str(bad))


def test_wire_roundtrip_agrees_with_oracle_codec(oracle_lib, tmp_path):
    train, test, _ = ol.make_ratings(200, 90, 5000, seed=12)
    p = train.write(str(tmp_path / "o.bin"))  # written by the oracle's encoder
    b = mb.Blocks.read(p)
    np.testing.assert_array_equal(b.vid, train.vid)
    np.testing.assert_array_equal(b.run_off, train.run_off)
    np.testing.assert_array_equal(b.block_off, train.block_off)
    q = mb.Blocks.from_arrays(train.block_off, train.run_uid, train.run_off, train.vid, train.rating).write(
        str(tmp_path / "p.bin"))
    assert open(p, "rb").read() == open(q, "rb").read()
    d = ol.Dataset.read(q)  # and the oracle reads what the product wrote
    np.testing.assert_array_equal(d.rating, train.rating)


def test_generator_shape_determinism_and_sharding():
    nu, nv, nnz = 3000, 700, 120000
    p = mb.gen_params(nu, nv, nnz, test_frac=0.1, valid_frac=0.05, users_per_block=100)
    tr, te, va = mb.generate(p)
    total = tr.nratings + te.nratings + va.nratings
    assert abs(total - nnz) < 0.03 * nnz
    assert abs(te.nratings / total - 0.1) < 0.01 and abs(va.nratings / total - 0.05) < 0.01
    assert set(np.unique(tr.rating)) <= {1.0, 2.0, 3.0, 4.0, 5.0}
    assert tr.vid.min() >= 0 and tr.vid.max() < nv and tr.run_uid.max() < nu
    # no duplicate (u,i) across train+test+valid
    keys = []
    for b in (tr, te, va):
        u = np.repeat(b.run_uid, np.diff(b.run_off)).astype(np.int64)
        keys.append(u * nv + b.vid)
    keys = np.concatenate(keys)
    assert len(np.unique(keys)) == len(keys)
    # getdata-style layout: 4 chunks, each user at most once per chunk, blocks of <= 100 users
    assert np.diff(tr.block_off).max() <= 100
    counts = np.bincount(tr.run_uid, minlength=nu)
    assert counts.max() <= 4
    # popularity is skewed, degrees are skewed
    pop = np.sort(np.bincount(tr.vid, minlength=nv))[::-1]
    assert pop[:nv // 10].sum() > 0.3 * tr.nratings
    # same data for any thread count; a user range reproduces exactly that slice
    p1 = mb.gen_params(nu, nv, nnz, test_frac=0.1, valid_frac=0.05, users_per_block=100, threads=1)
    tr1, _, _ = mb.generate(p1)
    np.testing.assert_array_equal(tr1.vid, tr.vid)
    np.testing.assert_array_equal(tr1.rating, tr.rating)
    ps = mb.gen_params(nu, nv, nnz, test_frac=0.1, valid_frac=0.05, users_per_block=100,
                       user_begin=500, user_end=1700)
    trs, _, _ = mb.generate(ps)
    u_all = np.repeat(tr.run_uid, np.diff(tr.run_off))
    m = (u_all >= 500) & (u_all < 1700)
    np.testing.assert_array_equal(np.repeat(trs.run_uid, np.diff(trs.run_off)), u_all[m])
    np.testing.assert_array_equal(trs.vid, tr.vid[m])
    np.testing.assert_array_equal(trs.rating, tr.rating[m])
