"""the out-of-core epoch (mfb_sgd_epoch_from_file: raw bytes of a chunk -> pinned -> H2D -> records decoded
on the device (file_decode = 1, mfb_wire_decode.cu) or by the host cores (file_decode = 0) -> kernel; a fixed number of
device tile buffers whatever the file size) against the resident epoch on the same file."""
import struct

import numpy as np
import pytest

import mfb200 as mb
import oraclelib as ol
from gpu_common import ctx_from_model, model_equal, model_rel_err, oracle_sgd

pytestmark = pytest.mark.gpu
GB = 2.76


DECODERS = pytest.mark.parametrize("decode", [1, 0], ids=["device-decode", "host-decode"])


@DECODERS
def test_file_epoch_in_ordered_schedule_is_the_resident_epoch_bit_for_bit(tmp_path, decode):
    """file order is kept across chunks: many chunks (a tile far smaller than the file), ordered schedule ->
    identical to the epoch on the loaded file and to the CPU oracle"""
    nu, nv, dim = 400, 150, 32
    tr, _, _ = mb.generate(mb.gen_params(nu, nv, 30000, test_frac=0.0, users_per_block=20))
    path = tr.write(str(tmp_path / "train.bin"))
    train = ol.Dataset(tr.block_off, tr.run_uid, tr.run_off, tr.vid, tr.rating)
    m = ol.Model(nu, nv, dim, seed=4)
    c1, c2 = ctx_from_model(m), ctx_from_model(m)
    c1.set_option("file_decode", decode)
    d = c2.dataset_from_file(path)
    for ep in (1, 2):
        eta = mb.seteta(2e-2, ep, 1.0)
        n = c1.sgd_epoch_from_file(path, eta, 5e-3, GB, mb.MODE_ORDERED, tile_ratings=2000)
        assert n == tr.nratings
        c2.sgd_epoch(d, eta, 5e-3, GB, mb.MODE_ORDERED)
        oracle_sgd(m, train, eta, 5e-3, GB)
    for a, b in zip(c1.get_factors(), c2.get_factors()):
        np.testing.assert_array_equal(a, b)
    assert model_equal(c1, m)
    assert c1.launch_count() >= 2 * 10          # the file really went through many chunks
    assert c1.h2d_bytes() >= 2 * tr.nratings * 8
    c1.close()
    c2.close()


@DECODERS
@pytest.mark.parametrize("mode", [mb.MODE_ATOMIC, mb.MODE_HOGWILD])
def test_file_epoch_larger_than_the_tile_buffers_equals_resident_epoch_on_conflict_free_data(tmp_path, mode, decode):
    """a file 25x the configured tile buffer, production schedule, conflict-free data (the order cannot matter):
    the streamed epoch == the resident epoch == the serial oracle"""
    n, dim = 200_000, 64
    rng = np.random.default_rng(3)
    ds = ol.Dataset(np.r_[np.arange(0, n, 500), n], rng.permutation(n), np.arange(n + 1), rng.permutation(n),
                    rng.integers(1, 6, n))
    path = ds.write(str(tmp_path / "train.bin"))
    m = ol.Model(n, n, dim, seed=6, scale=0.3)
    c1, c2 = ctx_from_model(m), ctx_from_model(m)
    c1.set_option("file_decode", decode)
    d = c2.dataset_from_file(path)
    got = c1.sgd_epoch_from_file(path, 0.05, 0.02, GB, mode, tile_ratings=8000)
    assert got == n
    c2.sgd_epoch(d, 0.05, 0.02, GB, mode)
    oracle_sgd(m, ds, 0.05, 0.02, GB)
    assert model_rel_err(c1, m) <= 1e-5 and model_rel_err(c2, m) <= 1e-5
    c1.close()
    c2.close()


@pytest.mark.parametrize("with_epoch", [True, False], ids=["ingest fused with epoch 1", "ingest only"])
def test_streaming_ingest_leaves_the_dataset_the_loader_builds(tmp_path, with_epoch):
    """mfb_dataset_ingest_file: one pass over the file, records decoded on the GPU, chunks appended to the
    resident tiles behind their update kernels.  Ordered schedule: the fused first epoch + a resident second epoch ==
    load + finalize + two resident epochs == the oracle, bit for bit; same run / record / Block counts, same Block
    boundaries (an epoch over a Block range), same SSE."""
    nu, nv, dim = 400, 150, 32
    tr, te, _ = mb.generate(mb.gen_params(nu, nv, 30000, test_frac=0.1, users_per_block=20))
    path = tr.write(str(tmp_path / "train.bin"))
    train = ol.Dataset(tr.block_off, tr.run_uid, tr.run_off, tr.vid, tr.rating)
    m = ol.Model(nu, nv, dim, seed=4)
    c1, c2 = ctx_from_model(m), ctx_from_model(m)
    e1, e2 = mb.seteta(2e-2, 1, 1.0), mb.seteta(2e-2, 2, 1.0)
    d1, n = c1.dataset_ingest_file(path, with_epoch, e1, 5e-3, GB, mb.MODE_ORDERED, tile_ratings=2000)
    assert n == tr.nratings
    if not with_epoch:
        c1.sgd_epoch(d1, e1, 5e-3, GB, mb.MODE_ORDERED)
    d2 = c2.dataset_from_file(path)
    c2.sgd_epoch(d2, e1, 5e-3, GB, mb.MODE_ORDERED)
    assert (c1.num_ratings(d1), c1.num_runs(d1), c1.num_blocks(d1)) == (c2.num_ratings(d2), c2.num_runs(d2), c2.num_blocks(d2))
    assert (c1.num_ratings(d1), c1.num_runs(d1), c1.num_blocks(d1)) == (tr.nratings, tr.nruns, tr.nblocks)
    for c, d in ((c1, d1), (c2, d2)):
        c.sgd_epoch(d, e2, 5e-3, GB, mb.MODE_ORDERED)
        c.sgd_epoch_blocks(d, 3, 11, e2, 5e-3, GB, mb.MODE_ORDERED)   # Block boundaries
    oracle_sgd(m, train, e1, 5e-3, GB)
    oracle_sgd(m, train, e2, 5e-3, GB)
    oracle_sgd(m, train.block_range(3, 11), e2, 5e-3, GB)
    for a, b in zip(c1.get_factors(), c2.get_factors()):
        np.testing.assert_array_equal(a, b)
    assert model_equal(c1, m)
    s1, s2 = c1.sse(d1, GB), c2.sse(d2, GB)
    assert s1[1] == s2[1] and abs(s1[0] - s2[0]) <= 1e-12 * s2[0]   # (fp64 sums in the order the blocks finish)
    c1.close()
    c2.close()


def test_streaming_ingest_in_the_production_schedule_and_its_refusals(tmp_path):
    """conflict-free data (the order cannot matter): ingest fused with the first epoch == resident epoch == oracle
    to 1e-5; a dpmf context, a non-empty dataset and a damaged file are refused, and the context stays usable"""
    n, dim = 100_000, 64
    rng = np.random.default_rng(3)
    ds = ol.Dataset(np.r_[np.arange(0, n, 500), n], rng.permutation(n), np.arange(n + 1), rng.permutation(n),
                    rng.integers(1, 6, n))
    path = ds.write(str(tmp_path / "train.bin"))
    m = ol.Model(n, n, dim, seed=6, scale=0.3)
    c1, c2 = ctx_from_model(m), ctx_from_model(m)
    d1, got = c1.dataset_ingest_file(path, True, 0.05, 0.02, GB, mb.MODE_ATOMIC, tile_ratings=8000)
    assert got == n
    d2 = c2.dataset_from_file(path)
    for c, d, first in ((c1, d1, False), (c2, d2, True)):
        if first:
            c.sgd_epoch(d, 0.05, 0.02, GB, mb.MODE_ATOMIC)
        c.sgd_epoch(d, 0.03, 0.02, GB, mb.MODE_ATOMIC)
    oracle_sgd(m, ds, 0.05, 0.02, GB)
    oracle_sgd(m, ds, 0.03, 0.02, GB)
    assert model_rel_err(c1, m) <= 1e-5 and model_rel_err(c2, m) <= 1e-5
    # refusals
    bad = tmp_path / "bad.bin"
    bad.write_bytes(open(path, "rb").read()[:-7])
    with pytest.raises(mb.MfbError):
        c1.dataset_ingest_file(str(bad))
    d3, got3 = c1.dataset_ingest_file(path)          # ... and the context still ingests
    assert got3 == n and c1.num_runs(d3) == n
    c3 = mb.Context(50, 50, 16)
    c3.enable(2)
    with pytest.raises(mb.MfbError):
        c3.dataset_ingest_file(path)
    for c in (c1, c2, c3):
        c.close()


def _varint(v):
    out = bytearray()
    while v >= 0x80:
        out.append((v & 0x7F) | 0x80)
        v >>= 7
    out.append(v)
    return bytes(out)


def _padded_varint(v, nbytes):
    """the same value in a longer (still valid) encoding: continuation bytes with zero payload"""
    b = bytearray(_varint(v))
    while len(b) < nbytes:
        b[-1] |= 0x80
        b.append(0)
    return bytes(b)


def test_device_decoder_reads_every_valid_encoding_like_the_host_decoder(tmp_path):
    """protobuf accepts more than its own encoder writes: fields in any order, unknown fields of every wire type,
    over-long varints, a repeated scalar (the last value wins), missing fields (zero), users without records, empty
    Blocks.  Device decoder == host decoder == the intended content, checked through an ordered epoch bit for bit."""
    rng = np.random.default_rng(12)
    nu, nv, dim = 300, 70000, 16   # item ids up to 3-byte varints (and longer, padded)
    frames, want_uid, want_off, want_vid, want_rating, blocks = [], [], [0], [], [], [0]
    for blk in range(12):
        body = bytearray()
        if blk == 5:
            frames.append(b"")            # an empty Block
            blocks.append(len(want_uid))
            continue
        for _ in range(int(rng.integers(1, 30))):
            uid = int(rng.integers(0, nu))
            user = bytearray()
            nrec = int(rng.integers(0, 40)) if rng.random() > 0.1 else 0
            uid_first = rng.random() < 0.7
            if uid_first:
                user += b"\x08" + _varint(uid)
            else:
                user += b"\x08" + _varint((uid + 1) % nu)   # overwritten by the later occurrence
            for _ in range(nrec):
                vid = int(rng.choice([rng.integers(0, 128), rng.integers(128, 16384), rng.integers(16384, nv)]))
                rating = float(rng.integers(1, 11)) / 2
                style = rng.integers(0, 6)
                rec = bytearray()
                if style == 0:    # canonical
                    rec += b"\x08" + _varint(vid) + b"\x15" + struct.pack("<f", rating)
                elif style == 1:  # rating first
                    rec += b"\x15" + struct.pack("<f", rating) + b"\x08" + _varint(vid)
                elif style == 2:  # over-long varint
                    rec += b"\x08" + _padded_varint(vid, int(rng.integers(4, 8))) + b"\x15" + struct.pack("<f", rating)
                elif style == 3:  # unknown fields: varint (field 3), fixed64 (field 4), bytes (field 5), fixed32 (field 6)
                    rec += b"\x18" + _varint(int(rng.integers(0, 1 << 40))) + b"\x08" + _varint(vid)
                    rec += b"\x21" + bytes(8) + b"\x2a\x03abc" + b"\x15" + struct.pack("<f", rating) + b"\x35" + bytes(4)
                elif style == 4:  # repeated scalars: the last wins
                    rec += b"\x08" + _varint((vid + 7) % nv) + b"\x15" + struct.pack("<f", 9.0)
                    rec += b"\x08" + _varint(vid) + b"\x15" + struct.pack("<f", rating)
                else:             # vid missing -> 0
                    vid = 0
                    rec += b"\x15" + struct.pack("<f", rating)
                lenbytes = _varint(len(rec)) if rng.random() < 0.8 else _padded_varint(len(rec), 3)
                user += b"\x12" + lenbytes + rec
                want_vid.append(vid)
                want_rating.append(rating)
                if rng.random() < 0.05:
                    user += b"\x3a\x02xy"    # an unknown field between records
            if not uid_first:
                user += b"\x08" + _varint(uid)
            body += b"\x0a" + _varint(len(user)) + user
            if rng.random() < 0.1:
                body += b"\x10\x05"          # an unknown top-level field
            want_uid.append(uid)
            want_off.append(len(want_vid))
        frames.append(bytes(body))
        blocks.append(len(want_uid))
    path = tmp_path / "odd.bin"
    path.write_bytes(b"".join(struct.pack("<I", len(f)) + f for f in frames))
    ds = ol.Dataset(blocks, want_uid, want_off, want_vid, np.array(want_rating, np.float32))
    m = ol.Model(nu, nv, dim, seed=9)
    ctxs = [ctx_from_model(m) for _ in range(3)]
    ctxs[0].set_option("file_decode", 1)
    ctxs[1].set_option("file_decode", 0)
    d2 = ctxs[2].dataset_from_file(str(path))   # (the loader's decoder)
    assert ctxs[2].num_ratings(d2) == len(want_vid) and ctxs[2].num_runs(d2) == len(want_uid)
    for tile in (1024, 1 << 20):
        for c in ctxs[:2]:
            assert c.sgd_epoch_from_file(str(path), 0.02, 5e-3, GB, mb.MODE_ORDERED, tile_ratings=tile) == len(want_vid)
        ctxs[2].sgd_epoch(d2, 0.02, 5e-3, GB, mb.MODE_ORDERED)
        oracle_sgd(m, ds, 0.02, 5e-3, GB)
        for c in ctxs:
            assert model_equal(c, m)
    for c in ctxs:
        c.close()


def test_device_and_host_decoder_agree_on_damaged_files(tmp_path):
    """fuzz: a valid file with bytes flipped, inserted, removed or cut off.  Whatever the damage, the two decoders
    must do the same thing - both refuse the file, or both decode it to the same records (checked through an ordered
    epoch, bit for bit) - and the device decoder must never read or write out of bounds (the run would die)."""
    rng = np.random.default_rng(77)
    nu, nv, dim = 200, 300, 16
    tr, _, _ = mb.generate(mb.gen_params(nu, nv, 6000, test_frac=0.0, users_per_block=15))
    good = open(tr.write(str(tmp_path / "good.bin")), "rb").read()
    m = ol.Model(nu, nv, dim, seed=3)
    outcomes = {"both refuse": 0, "both decode": 0}
    for trial in range(40):
        b = bytearray(good)
        for _ in range(int(rng.integers(1, 4))):
            kind, at = rng.integers(0, 4), int(rng.integers(0, len(b)))
            if kind == 0:
                b[at] ^= 1 << int(rng.integers(0, 8))
            elif kind == 1:
                b[at:at] = bytes(rng.integers(0, 256, int(rng.integers(1, 6))).astype(np.uint8))
            elif kind == 2:
                del b[at:at + int(rng.integers(1, 6))]
            else:
                del b[max(at, len(b) - 200):]
        path = tmp_path / ("fuzz%d.bin" % trial)
        path.write_bytes(bytes(b))
        res = []
        for decode in (1, 0):
            c = ctx_from_model(m)
            c.set_option("file_decode", decode)
            try:
                n = c.sgd_epoch_from_file(str(path), 0.02, 5e-3, GB, mb.MODE_ORDERED, tile_ratings=1 << 20)
                res.append((n, [x.copy() for x in c.get_factors()]))
            except mb.MfbError:
                res.append(None)
            c.close()
        assert (res[0] is None) == (res[1] is None), "trial %d: device %s, host %s" % (
            trial, "refused" if res[0] is None else "decoded", "refused" if res[1] is None else "decoded")
        if res[0] is None:
            outcomes["both refuse"] += 1
        else:
            outcomes["both decode"] += 1
            assert res[0][0] == res[1][0]
            for a, bb in zip(res[0][1], res[1][1]):
                np.testing.assert_array_equal(a, bb)
    print(outcomes)
    assert outcomes["both refuse"] > 0 and outcomes["both decode"] > 0


@DECODERS
def test_file_epoch_reports_io_and_format_errors(tmp_path, decode):
    c = mb.Context(50, 50, 16)
    c.set_option("file_decode", decode)
    with pytest.raises(mb.MfbError):
        c.sgd_epoch_from_file(str(tmp_path / "missing.bin"), 0.01, 0.01, GB)
    bad = tmp_path / "bad.bin"
    bad.write_bytes(b"\x10\x00\x00\x00" + b"\xff" * 16)
    with pytest.raises(mb.MfbError):
        c.sgd_epoch_from_file(str(bad), 0.01, 0.01, GB)
    trunc = tmp_path / "trunc.bin"
    trunc.write_bytes(b"\x40\x00\x00\x00" + b"\x0a\x02\x08\x01")
    with pytest.raises(mb.MfbError):
        c.sgd_epoch_from_file(str(trunc), 0.01, 0.01, GB)
    # a user outside [0, nu), an item outside [0, nv)
    ds = ol.Dataset([0, 1], [77], [0, 1], [3], [4.0])
    p = ds.write(str(tmp_path / "range.bin"))
    with pytest.raises(mb.MfbError):
        c.sgd_epoch_from_file(p, 0.01, 0.01, GB)
    ds = ol.Dataset([0, 1], [7], [0, 2], [3, 50], [4.0, 1.0])
    p = ds.write(str(tmp_path / "range2.bin"))
    with pytest.raises(mb.MfbError):
        c.sgd_epoch_from_file(p, 0.01, 0.01, GB)
    # a record that runs past the end of its user, a user that runs past the end of its Block
    for name, body in (("rec.bin", b"\x0a\x08\x08\x01\x12\x09\x08\x02\x15\x00"), ("usr.bin", b"\x0a\x20\x08\x01")):
        q = tmp_path / name
        q.write_bytes(struct.pack("<I", len(body)) + body)
        with pytest.raises(mb.MfbError):
            c.sgd_epoch_from_file(str(q), 0.01, 0.01, GB)
    # the context is still usable after every one of these
    ds = ol.Dataset([0, 1], [7], [0, 2], [3, 4], [4.0, 1.0])
    assert c.sgd_epoch_from_file(ds.write(str(tmp_path / "good.bin")), 0.01, 0.01, GB) == 2
    empty = tmp_path / "empty.bin"
    empty.write_bytes(b"")
    assert c.sgd_epoch_from_file(str(empty), 0.01, 0.01, GB) == 0
    c.close()
