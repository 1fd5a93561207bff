"""the out-of-core epoch (mfb_sgd_epoch_from_file: decode -> pinned chunk -> H2D -> kernel, two device
tile buffers whatever the file size) against the resident epoch on the same file."""
import numpy as np
import pytest

import mfb200 as mb
import oraclelib as ol
from gpu_common import ctx_from_model, model_equal, model_rel_err, oracle_sgd

pytestmark = pytest.mark.gpu
GB = 2.76


def test_file_epoch_in_ordered_schedule_is_the_resident_epoch_bit_for_bit(tmp_path):
    """file order is kept across chunks: many chunks (a tile far smaller than the file), ordered schedule ->
    identical to the epoch on the loaded file and to the CPU oracle"""
    nu, nv, dim = 400, 150, 32
    tr, _, _ = mb.generate(mb.gen_params(nu, nv, 30000, test_frac=0.0, users_per_block=20))
    path = tr.write(str(tmp_path / "train.bin"))
    train = ol.Dataset(tr.block_off, tr.run_uid, tr.run_off, tr.vid, tr.rating)
    m = ol.Model(nu, nv, dim, seed=4)
    c1, c2 = ctx_from_model(m), ctx_from_model(m)
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


@pytest.mark.parametrize("mode", [mb.MODE_ATOMIC, mb.MODE_HOGWILD])
def test_file_epoch_larger_than_the_tile_buffers_equals_resident_epoch_on_conflict_free_data(tmp_path, mode):
    """a file 25x the configured tile buffer, production schedule, conflict-free data (the order cannot matter):
    the streamed epoch == the resident epoch == the serial oracle"""
    n, dim = 200_000, 64
    rng = np.random.default_rng(3)
    ds = ol.Dataset(np.r_[np.arange(0, n, 500), n], rng.permutation(n), np.arange(n + 1), rng.permutation(n),
                    rng.integers(1, 6, n))
    path = ds.write(str(tmp_path / "train.bin"))
    m = ol.Model(n, n, dim, seed=6, scale=0.3)
    c1, c2 = ctx_from_model(m), ctx_from_model(m)
    d = c2.dataset_from_file(path)
    got = c1.sgd_epoch_from_file(path, 0.05, 0.02, GB, mode, tile_ratings=8000)
    assert got == n
    c2.sgd_epoch(d, 0.05, 0.02, GB, mode)
    oracle_sgd(m, ds, 0.05, 0.02, GB)
    assert model_rel_err(c1, m) <= 1e-5 and model_rel_err(c2, m) <= 1e-5
    c1.close()
    c2.close()


def test_file_epoch_reports_io_and_format_errors(tmp_path):
    c = mb.Context(50, 50, 16)
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
    # a user outside [0, nu)
    ds = ol.Dataset([0, 1], [77], [0, 1], [3], [4.0])
    p = ds.write(str(tmp_path / "range.bin"))
    with pytest.raises(mb.MfbError):
        c.sgd_epoch_from_file(p, 0.01, 0.01, GB)
    empty = tmp_path / "empty.bin"
    empty.write_bytes(b"")
    assert c.sgd_epoch_from_file(str(empty), 0.01, 0.01, GB) == 0
    c.close()
