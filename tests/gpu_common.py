"""Helpers shared by the -m gpu parity tests (CUDA path through the C ABI vs the CPU oracle)."""
import ctypes as C

import numpy as np

import mfb200 as mb
import oraclelib as ol


def ctx_from_model(m, device=0):
    """Upload an oraclelib.Model into a fresh device context."""
    c = mb.Context(m.nu, m.nv, m.dim, device)
    th, ph = m.dense()
    c.set_factors(th, ph, m.bu, m.bv)
    return c


def upload_ds(c, ds):
    return c.dataset_from_arrays(ds.block_off, ds.run_uid, ds.run_off, ds.vid, ds.rating)


def row_rel_err(got, want):
    """max over rows of |got-want|_inf / max(|want|_inf, tiny): the north star's per-row metric."""
    got = np.asarray(got, np.float64).reshape(len(want), -1)
    want = np.asarray(want, np.float64).reshape(len(want), -1)
    num = np.abs(got - want).max(axis=1)
    den = np.maximum(np.abs(want).max(axis=1), 1e-30)
    return float((num / den).max())


def vec_rel_err(got, want):
    """bias vectors: a single entry may legitimately sit next to zero (lameta*b + e cancels), so
    the error is measured against the scale of the whole vector."""
    want = np.asarray(want, np.float64)
    return float(np.abs(np.asarray(got, np.float64) - want).max() / max(np.abs(want).max(), 1e-30))


def model_rel_err(c, m):
    """factor rows: per-row relative error (the north star's 1e-5 metric); biases: vec_rel_err"""
    th, ph, bu, bv = c.get_factors()
    d = m.dim
    return max(row_rel_err(th, m.theta[:, :d]), row_rel_err(ph, m.phi[:, :d]),
               vec_rel_err(bu, m.bu), vec_rel_err(bv, m.bv))


def model_equal(c, m):
    th, ph, bu, bv = c.get_factors()
    d = m.dim
    return (np.array_equal(th, m.theta[:, :d]) and np.array_equal(ph, m.phi[:, :d]) and
            np.array_equal(bu, m.bu) and np.array_equal(bv, m.bv))


def oracle_sgd(m, ds, eta, lam, gb):
    mm, dd = m.as_mfo(), ds.as_mfo()
    ol.oracle().mfo_sgd_epoch(C.byref(mm), C.byref(dd), eta, lam, gb)


def oracle_sse(m, ds, gb):
    mm, dd = m.as_mfo(), ds.as_mfo()
    n = C.c_int64()
    s = ol.oracle().mfo_sse(C.byref(mm), C.byref(dd), gb, C.byref(n))
    return float(s), n.value
