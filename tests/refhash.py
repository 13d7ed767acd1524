"""Fingerprints used by oracle/ref_driver.cpp --hash-only, recomputed with numpy (tests only)."""
import numpy as np

GOLD = np.uint64(0x9E3779B97F4A7C15)


def hash_int(a):
    """H = sum_i (a_i + 1) (2 i + 1) 0x9E3779B97F4A7C15  (mod 2^64); a_i taken as uint32"""
    a = np.ascontiguousarray(a).ravel()
    h = np.uint64(0)
    step = 1 << 24
    with np.errstate(over="ignore"):
        for s in range(0, a.shape[0], step):
            v = a[s:s + step].astype(np.int64).astype(np.uint32).astype(np.uint64) + np.uint64(1)
            i = np.arange(s, s + v.shape[0], dtype=np.uint64) * np.uint64(2) + np.uint64(1)
            h = h + np.sum(v * i * GOLD, dtype=np.uint64)
    return int(h)


def check(got, ref, prefix_keys, tol=1e-12):
    """got: arrays by name; ref: the reference's JSON line"""
    checked = 0
    for key, val in ref.items():
        if key.startswith("hash_"):
            name = key[5:]
            if name not in got:
                continue
            assert got[name].size == ref["len_" + name], (name, got[name].size, ref["len_" + name])
            assert hash_int(got[name]) == int(val), name
            checked += 1
        elif key.startswith("sumabs_"):
            name = key[7:]
            if name not in got:
                continue
            a = np.asarray(got[name], dtype=np.float64).ravel()
            assert a.size == ref["len_" + name], name
            scale = max(float(val), 1e-300)
            # the reference adds sequentially in double (ref_driver.cpp), numpy pairwise: allow n * eps of summation noise
            t = tol * 10 + 2.3e-16 * a.size
            assert abs(np.abs(a).sum() - float(val)) <= t * scale, (name, np.abs(a).sum(), val)
            assert abs(a.sum() - float(ref["sum_" + name])) <= t * scale, (name, a.sum(), ref["sum_" + name])
            assert abs((a * a).sum() - float(ref["sumsq_" + name])) <= t * max(float(ref["sumsq_" + name]), 1e-300), name
            checked += 1
    return checked
