"""Parity metrics used by every test (stated once, SURVEY.md section 7 'value parity metric').

* integer / index arrays: bit-exact;
* vectors: |a - b| <= tol * max|b| (b the reference);
* matrix entries: |a - b| <= tol * max(|a|, |b|, ||row||_inf). The reference stores entries that
  are analytically zero (2 (n-1)^2 hypotenuse couplings of the P1 square mesh, the whole
  pressure-pressure block of Stokes); a per-entry relative test is meaningless there, so each
  entry is measured against the largest entry of its row.
"""
import numpy as np


def assert_int_equal(got, ref, name=""):
    got = np.asarray(got); ref = np.asarray(ref)
    assert got.shape == ref.shape, (name, got.shape, ref.shape)
    assert np.array_equal(got.astype(np.int64), ref.astype(np.int64)), \
        (name, int((got.astype(np.int64) != ref.astype(np.int64)).sum()))


def assert_values_close(got, ref, tol, name=""):
    got = np.asarray(got, dtype=np.float64); ref = np.asarray(ref, dtype=np.float64)
    assert got.shape == ref.shape, (name, got.shape, ref.shape)
    if ref.size == 0:
        return
    scale = max(np.abs(ref).max(), 1e-300)
    err = np.abs(got - ref).max()
    assert err <= tol * scale, (name, err, scale)


def matrix_entries_close(rowptr, got, ref, tol):
    got = np.asarray(got, dtype=np.float64); ref = np.asarray(ref, dtype=np.float64)
    assert got.shape == ref.shape
    rowptr = np.asarray(rowptr, dtype=np.int64)
    n = rowptr.shape[0] - 1
    lens = np.diff(rowptr)
    rows = np.repeat(np.arange(n), lens)
    rowmax = np.zeros(n)
    np.maximum.at(rowmax, rows, np.maximum(np.abs(got), np.abs(ref)))
    bound = tol * np.maximum(np.maximum(np.abs(got), np.abs(ref)), rowmax[rows])
    bad = np.abs(got - ref) > bound
    assert not bad.any(), (int(bad.sum()), float(np.abs(got - ref).max()))
