"""TEST INFRASTRUCTURE — ctypes wrapper over oracle/_ref/liboracle_ref.so.

That library is the reference's own CPU engine (GammaIVFPQIndex / GammaFLATIndex
over the vendored faiss 1.7.1) compiled unmodified by oracle/Makefile, with the
small C driver oracle/ref_driver.cc on top.  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference leg may import this module; the
product package (gamma_b200) never does.
"""
import ctypes as C
import os
import shutil
import tempfile

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "liboracle_ref.so")
_lib = None


def available():
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        # BLAS threads inside OpenMP regions oversubscribe (10x slowdown on 8 cores,
        # a stall on the 128-core GPU host); idle OpenMP workers must sleep, not spin.
        os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
        os.environ.setdefault("OMP_WAIT_POLICY", "PASSIVE")
        os.environ.setdefault("GOMP_SPINCOUNT", "0")
        _lib = C.CDLL(LIB_PATH)
        try:
            _lib.openblas_set_num_threads(1)
        except AttributeError:
            pass
        _lib.oref_open.restype = C.c_void_p
        _lib.oref_open.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_char_p, C.c_int, C.c_long]
        _lib.oref_close.argtypes = [C.c_void_p]
        _lib.oref_add_raw.argtypes = [C.c_void_p, C.c_long, C.c_void_p]
        _lib.oref_indexing.argtypes = [C.c_void_p]
        _lib.oref_add_to_index.argtypes = [C.c_void_p, C.c_long, C.c_int]
        _lib.oref_delete.argtypes = [C.c_void_p, C.c_long]
        _lib.oref_update.argtypes = [C.c_void_p, C.c_long, C.c_void_p]
        _lib.oref_info.restype = C.c_long
        _lib.oref_info.argtypes = [C.c_void_p, C.c_char_p]
        _lib.oref_get_centroids.argtypes = [C.c_void_p, C.c_void_p]
        _lib.oref_get_pq.argtypes = [C.c_void_p, C.c_void_p]
        _lib.oref_list_size.restype = C.c_long
        _lib.oref_list_size.argtypes = [C.c_void_p, C.c_long]
        _lib.oref_get_list.argtypes = [C.c_void_p, C.c_long, C.c_void_p, C.c_void_p]
        _lib.oref_opq_dims.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.oref_get_opq.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.oref_coarse.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        _lib.oref_set_trained.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.oref_inject_postings.argtypes = [C.c_void_p, C.c_long, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.oref_search.argtypes = [
            C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_char_p, C.c_int, C.c_int, C.c_float, C.c_float,
            C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
            C.c_void_p, C.c_void_p]
    return _lib


FLT_MAX = float(np.finfo(np.float32).max)


class RefIndex:
    """One reference RetrievalModel ("IVFPQ" or "FLAT") + its MemoryRawVector + deleted bitmap."""

    def __init__(self, d, retrieval_type="IVFPQ", model_json="", indexing_size=100000,
                 bitmap_bits=10_000_000, work_dir=None):
        self._own_dir = work_dir is None
        self.work_dir = work_dir or tempfile.mkdtemp(prefix="oref_")
        self.d = d
        self.h = lib().oref_open(self.work_dir.encode(), d, retrieval_type.encode(), model_json.encode(),
                                 indexing_size, bitmap_bits)
        if not self.h:
            raise RuntimeError("oref_open failed")

    def close(self):
        if self.h:
            lib().oref_close(self.h)
            self.h = None
        if self._own_dir:
            shutil.rmtree(self.work_dir, ignore_errors=True)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def info(self, key):
        return lib().oref_info(self.h, key.encode())

    def add_raw(self, x):
        x = np.ascontiguousarray(x, dtype=np.float32)
        assert x.shape[1] == self.d
        rc = lib().oref_add_raw(self.h, x.shape[0], x.ctypes.data)
        assert rc == 0, rc

    def indexing(self, threads=16):
        """train; k-means on small data degrades badly with >100 OpenMP threads, so cap them here"""
        prev = max_threads()
        set_threads(min(prev, threads))
        try:
            rc = lib().oref_indexing(self.h)
        finally:
            set_threads(prev)
        assert rc == 0, rc

    def set_trained(self, coarse, pq):
        coarse = np.ascontiguousarray(coarse, dtype=np.float32)
        pq = np.ascontiguousarray(pq, dtype=np.float32)
        rc = lib().oref_set_trained(self.h, coarse.ctypes.data, pq.ctypes.data)
        assert rc == 0, rc

    def inject_postings(self, list_no, vids, codes):
        ln = np.ascontiguousarray(list_no, dtype=np.int32)
        v = np.ascontiguousarray(vids, dtype=np.int64)
        c = np.ascontiguousarray(codes, dtype=np.uint8)
        rc = lib().oref_inject_postings(self.h, ln.size, ln.ctypes.data, v.ctypes.data, c.ctypes.data)
        assert rc == 0, rc

    def add_to_index(self, upto=-1, chunk=10000):
        rc = lib().oref_add_to_index(self.h, upto, chunk)
        assert rc == 0, rc

    def delete(self, docid):
        return lib().oref_delete(self.h, int(docid))

    def update(self, vid, x):
        x = np.ascontiguousarray(x, dtype=np.float32)
        return lib().oref_update(self.h, int(vid), x.ctypes.data)

    def centroids(self):
        out = np.empty((self.info("nlist"), self.d), np.float32)
        assert lib().oref_get_centroids(self.h, out.ctypes.data) == 0
        return out

    def pq_centroids(self):
        out = np.empty((self.info("M"), self.info("ksub"), self.info("dsub")), np.float32)
        assert lib().oref_get_pq(self.h, out.ctypes.data) == 0
        return out

    def opq(self):
        """(A [d_out, d_in], b [d_out] or None) of the model's OPQ pre-transform, or None without one"""
        di, do, hb = C.c_int(0), C.c_int(0), C.c_int(0)
        if not lib().oref_opq_dims(self.h, C.byref(di), C.byref(do), C.byref(hb)):
            return None
        A = np.empty((do.value, di.value), np.float32)
        b = np.empty(do.value, np.float32) if hb.value else None
        assert lib().oref_get_opq(self.h, A.ctypes.data, b.ctypes.data if b is not None else None) == 0
        return A, b

    def get_list(self, list_no):
        n = lib().oref_list_size(self.h, list_no)
        ids = np.empty(n, np.int64)
        codes = np.empty((n, self.info("code_size")), np.uint8)
        if n:
            assert lib().oref_get_list(self.h, list_no, ids.ctypes.data, codes.ctypes.data) == 0
        return ids, codes

    def lists(self):
        return [self.get_list(i) for i in range(self.info("nlist"))]

    def coarse(self, xq, nprobe):
        xq = np.ascontiguousarray(xq, dtype=np.float32)
        n = xq.shape[0]
        cd = np.empty((n, nprobe), np.float32)
        keys = np.empty((n, nprobe), np.int64)
        lib().oref_coarse(self.h, n, xq.ctypes.data, nprobe, cd.ctypes.data, keys.ctypes.data)
        return cd, keys

    def search(self, xq, k, retrieval_json="", has_rank=True, brute_force=False,
               min_score=-FLT_MAX, max_score=FLT_MAX, filters=(), keys=None, coarse_dis=None):
        """filters: sequence of (min_doc, max_doc, not_in, pass_flags[u8 over min..max])."""
        xq = np.ascontiguousarray(xq, dtype=np.float32)
        n = xq.shape[0]
        D = np.empty((n, k), np.float32)
        I = np.empty((n, k), np.int64)
        nf = len(filters)
        fmin = np.array([f[0] for f in filters], np.int32)
        fmax = np.array([f[1] for f in filters], np.int32)
        fnot = np.array([1 if f[2] else 0 for f in filters], np.int32)
        flags = [np.ascontiguousarray(f[3], dtype=np.uint8) for f in filters]
        ptrs = (C.c_void_p * max(nf, 1))(*[fl.ctypes.data for fl in flags])
        kp = cp = None
        npre = 0
        if keys is not None:
            keys = np.ascontiguousarray(keys, dtype=np.int64)
            coarse_dis = np.ascontiguousarray(coarse_dis, dtype=np.float32)
            kp, cp, npre = keys.ctypes.data, coarse_dis.ctypes.data, keys.shape[1]
        rc = lib().oref_search(self.h, n, xq.ctypes.data, k, retrieval_json.encode(), int(has_rank),
                               int(brute_force), min_score, max_score, nf, fmin.ctypes.data,
                               fmax.ctypes.data, fnot.ctypes.data, C.cast(ptrs, C.c_void_p), kp, cp, npre,
                               D.ctypes.data, I.ctypes.data)
        if rc != 0:
            raise RuntimeError("reference Search returned %d" % rc)
        return D, I


def set_threads(n):
    lib().oref_set_threads(int(n))


def max_threads():
    return lib().oref_max_threads()
