"""TEST INFRASTRUCTURE — ctypes wrapper of oracle/gamma_oracle.c (the plain-C restatement of the
reference's CPU algorithm).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg
may import this; the product never does.  Build: `make -C oracle port`."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgamma_oracle.so")
FLT_MAX = float(np.finfo(np.float32).max)
K_DEL_IDX_MASK = -(1 << 63)
_lib = None


class RangeFilter(C.Structure):
    _fields_ = [("min_doc", C.c_int), ("max_doc", C.c_int), ("min_aligned", C.c_int), ("not_in", C.c_int),
                ("bitmap", C.c_void_p)]


class Ivfpq(C.Structure):
    _fields_ = [("d", C.c_int), ("raw_d", C.c_int), ("nlist", C.c_int), ("M", C.c_int), ("dsub", C.c_int),
                ("cent", C.c_void_p), ("pq", C.c_void_p), ("list_off", C.c_void_p), ("ids", C.c_void_p),
                ("codes", C.c_void_p), ("raw", C.c_void_p), ("nraw", C.c_long)]


def lib():
    global _lib
    if _lib is None:
        src = os.path.join(_HERE, "gamma_oracle.c")
        if not os.path.exists(LIB_PATH) or os.path.getmtime(src) > os.path.getmtime(LIB_PATH):
            subprocess.check_call(["make", "-C", _HERE, "port"], stdout=subprocess.DEVNULL)
        _lib = C.CDLL(LIB_PATH)
    return _lib


def _filters(filters):
    arr = (RangeFilter * max(len(filters), 1))()
    keep = []
    for i, (mn, mx, not_in, flags) in enumerate(filters):
        mn, mx = int(mn), int(mx)
        min_al, max_al = (mn // 8) * 8, (mx // 8 + 1) * 8 - 1
        bits = np.zeros(max_al - min_al + 1, np.uint8)
        bits[mn - min_al: mx - min_al + 1] = (np.asarray(flags) != 0)
        by = np.packbits(bits, bitorder="little")
        keep.append(by)
        arr[i] = RangeFilter(mn, mx, min_al, 1 if not_in else 0, by.ctypes.data)
    return arr, keep


def _deleted(deleted_docs, nbits):
    if deleted_docs is None or len(deleted_docs) == 0:
        return None, 0
    bits = np.zeros(nbits, np.uint8)
    bits[np.asarray(deleted_docs, dtype=np.int64)] = 1
    return np.packbits(bits, bitorder="little"), nbits


def coarse(xq, cent, nprobe):
    xq = np.ascontiguousarray(xq, np.float32)
    cent = np.ascontiguousarray(cent, np.float32)
    n, d = xq.shape
    cd = np.empty((n, nprobe), np.float32)
    keys = np.empty((n, nprobe), np.int64)
    lib().go_coarse(n, d, cent.shape[0], C.c_void_p(xq.ctypes.data), C.c_void_p(cent.ctypes.data), nprobe,
                    C.c_void_p(cd.ctypes.data), C.c_void_p(keys.ctypes.data))
    return cd, keys


def ivfpq_search(xq, cent, pq, lists, raw, k, nprobe, recall_num=100, metric="L2", has_rank=True,
                 min_score=-FLT_MAX, max_score=FLT_MAX, filters=(), deleted_docs=None, keys=None, coarse_dis=None):
    """lists: per-list (ids int64 with kDelIdxMask, codes u8 [len,M]) in the reference layout."""
    xq = np.ascontiguousarray(xq, np.float32)
    cent = np.ascontiguousarray(cent, np.float32)
    pq = np.ascontiguousarray(pq, np.float32)
    raw = np.ascontiguousarray(raw, np.float32)
    n, d = xq.shape
    M, _, dsub = pq.shape
    lens = np.array([len(i) for i, _ in lists], np.int64)
    off = np.zeros(len(lists) + 1, np.int64)
    off[1:] = np.cumsum(lens)
    ids = np.ascontiguousarray(np.concatenate([i for i, _ in lists]) if len(lists) else np.zeros(0), np.int64)
    codes = np.ascontiguousarray(np.concatenate([c.reshape(-1, M) for _, c in lists]), np.uint8)
    ix = Ivfpq(d, raw.shape[1], cent.shape[0], M, dsub, cent.ctypes.data, pq.ctypes.data, off.ctypes.data,
               ids.ctypes.data, codes.ctypes.data, raw.ctypes.data, raw.shape[0])
    D = np.zeros((n, k), np.float32)
    I = np.full((n, k), -1, np.int64)
    farr, keep = _filters(filters)
    dbm, dbits = _deleted(deleted_docs, max(raw.shape[0], int(ids.max() & 0x7fffffff) + 1 if ids.size else 1))
    kp = cp = None
    if keys is not None:
        keys = np.ascontiguousarray(keys, np.int64)
        coarse_dis = np.ascontiguousarray(coarse_dis, np.float32)
        kp, cp, nprobe = C.c_void_p(keys.ctypes.data), C.c_void_p(coarse_dis.ctypes.data), keys.shape[1]
    rc = lib().go_ivfpq_search(C.byref(ix), n, C.c_void_p(xq.ctypes.data), k, nprobe, recall_num,
                               0 if metric == "L2" else 1, 1 if has_rank else 0, C.c_float(min_score),
                               C.c_float(max_score), C.c_void_p(dbm.ctypes.data) if dbm is not None else None,
                               C.c_long(dbits), farr, len(filters), kp, cp, C.c_void_p(D.ctypes.data),
                               C.c_void_p(I.ctypes.data))
    assert rc == 0
    return D, I


def ivfpq_search_from_arrays(xq, cent, pq, list_no, vids, codes, raw, **kw):
    """postings given as flat arrays in arrival order (list order = arrival order per list)."""
    nlist = cent.shape[0]
    order = np.argsort(list_no, kind="stable")
    ln, v, c = np.asarray(list_no)[order], np.asarray(vids, np.int64)[order], np.asarray(codes)[order]
    bounds = np.searchsorted(ln, np.arange(nlist + 1))
    lists = [(v[bounds[l]:bounds[l + 1]], c[bounds[l]:bounds[l + 1]]) for l in range(nlist)]
    return ivfpq_search(xq, cent, pq, lists, raw, **kw)


def flat_search(xq, raw, k, metric="L2", min_score=-FLT_MAX, max_score=FLT_MAX, filters=(), deleted_docs=None):
    xq = np.ascontiguousarray(xq, np.float32)
    raw = np.ascontiguousarray(raw, np.float32)
    n, d = xq.shape
    D = np.zeros((n, k), np.float32)
    I = np.full((n, k), -1, np.int64)
    farr, keep = _filters(filters)
    dbm, dbits = _deleted(deleted_docs, raw.shape[0])
    rc = lib().go_flat_search(n, d, C.c_long(raw.shape[0]), C.c_void_p(raw.ctypes.data), C.c_void_p(xq.ctypes.data), k,
                              0 if metric == "L2" else 1, C.c_float(min_score), C.c_float(max_score),
                              C.c_void_p(dbm.ctypes.data) if dbm is not None else None, C.c_long(dbits), farr,
                              len(filters), C.c_void_p(D.ctypes.data), C.c_void_p(I.ctypes.data))
    assert rc == 0
    return D, I
