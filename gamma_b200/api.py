"""ctypes binding of include/gamma_b200.h — the host side used by tests/ and bench.py.

The classes mirror the reference's RetrievalModel surface for the search path
(index/retrieval_model.h:218-310; GammaIVFPQIndex / GammaFLATIndex): same method names
(Init / Add / Update / Delete / Search / GetTotalMemBytes), same argument meaning, same
error convention (0 = success, negative = error).  There is no Python or CPU fallback: if
libgamma_b200.so is missing or no sm_100 GPU is present, construction raises.
"""
import ctypes as C
import json
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GB200_LIB") or os.path.join(_HERE, "lib", "libgamma_b200.so")  # GB200_LIB: dev A/B builds

METRIC_IP, METRIC_L2 = 0, 1
FLT_MAX = float(np.finfo(np.float32).max)
K_DEL_IDX_MASK = -(1 << 63)  # realtime::kDelIdxMask as a signed int64


class IvfpqParams(C.Structure):
    _fields_ = [("device", C.c_int), ("d", C.c_int), ("raw_d", C.c_int), ("nlist", C.c_int),
                ("nsubvector", C.c_int), ("nbits", C.c_int), ("metric", C.c_int), ("nprobe", C.c_int),
                ("store_raw", C.c_int)]


class SearchParams(C.Structure):
    _fields_ = [("metric", C.c_int), ("nprobe", C.c_int), ("recall_num", C.c_int), ("has_rank", C.c_int),
                ("min_score", C.c_float), ("max_score", C.c_float)]


class RangeFilter(C.Structure):
    _fields_ = [("min_doc", C.c_int), ("max_doc", C.c_int), ("min_aligned", C.c_int), ("not_in", C.c_int),
                ("bitmap", C.c_void_p)]


_lib = None

EXPORTS = [
    "gb200_last_error", "gb200_device_count", "gb200_ivfpq_create", "gb200_flat_create", "gb200_destroy",
    "gb200_ivfpq_set_quantizers", "gb200_ivfpq_append", "gb200_ivfpq_update", "gb200_ivfpq_list_sizes",
    "gb200_ivfpq_get_list", "gb200_upload_raw", "gb200_upload_raw_dev", "gb200_raw_count", "gb200_set_deleted",
    "gb200_upload_deleted_bitmap", "gb200_ivfpq_search", "gb200_ivfpq_search_preassigned", "gb200_ivfpq_coarse",
    "gb200_flat_search", "gb200_ivfpq_search_dev", "gb200_flat_search_dev", "gb200_set_filters", "gb200_mem_bytes",
    "gb200_last_scanned_postings", "gb200_launch_count", "gb200_last_stage_ms", "gb200_set_profiling", "gb200_last_scan_kernel_ms", "gb200_sync",
    "gb200_debug_select", "gb200_debug_plan", "gb200_reload_tuning", "gb200_ivfpq_search_sharded_deferred", "gb200_comm_flush", "gb200_ivfpq_compact", "gb200_ivfpq_replace_list", "gb200_ivfpq_encode", "gb200_ivfpq_add_raw", "gb200_ivfpq_add_stored", "gb200_ivfpq_set_opq", "gb200_ivfflat_create", "gb200_ivfflat_set_quantizer", "gb200_ivfflat_append",
    "gb200_ivfflat_add_raw", "gb200_ivfflat_search", "gb200_comm_create", "gb200_comm_connect", "gb200_comm_destroy", "gb200_comm_slot_bytes", "gb200_comm_status", "gb200_comm_read",
    "gb200_comm_buffers", "gb200_comm_exchange", "gb200_ivfpq_search_sharded",
]


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("libgamma_b200.so is not built (%s); run `python -m gamma_b200.build` — "
                               "there is no fallback path" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        L.gb200_last_error.restype = C.c_char_p
        L.gb200_ivfpq_create.argtypes = [C.POINTER(IvfpqParams), C.POINTER(C.c_void_p)]
        L.gb200_flat_create.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
        L.gb200_destroy.argtypes = [C.c_void_p]
        L.gb200_ivfpq_set_quantizers.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.gb200_ivfpq_append.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
        L.gb200_ivfpq_update.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_void_p]
        L.gb200_ivfpq_list_sizes.argtypes = [C.c_void_p, C.c_void_p]
        L.gb200_ivfpq_get_list.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
        L.gb200_upload_raw.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p]
        L.gb200_upload_raw_dev.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p]
        L.gb200_raw_count.restype = C.c_int64
        L.gb200_raw_count.argtypes = [C.c_void_p]
        L.gb200_set_deleted.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int]
        L.gb200_upload_deleted_bitmap.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
        L.gb200_ivfpq_search.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.POINTER(SearchParams),
                                         C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.gb200_ivfpq_search_preassigned.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                                                     C.POINTER(SearchParams), C.c_void_p, C.c_int, C.c_void_p,
                                                     C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.gb200_ivfpq_coarse.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.gb200_flat_search.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.POINTER(SearchParams),
                                        C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.gb200_ivfpq_search_dev.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.POINTER(SearchParams),
                                             C.c_void_p, C.c_void_p, C.c_void_p]
        L.gb200_flat_search_dev.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.POINTER(SearchParams),
                                            C.c_void_p, C.c_void_p, C.c_void_p]
        L.gb200_set_filters.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.gb200_mem_bytes.restype = C.c_int64
        L.gb200_mem_bytes.argtypes = [C.c_void_p]
        L.gb200_last_scanned_postings.restype = C.c_int64
        L.gb200_last_scanned_postings.argtypes = [C.c_void_p]
        L.gb200_launch_count.restype = C.c_int64
        L.gb200_launch_count.argtypes = [C.c_void_p]
        L.gb200_last_stage_ms.argtypes = [C.c_void_p, C.c_void_p]
        L.gb200_set_profiling.argtypes = [C.c_void_p, C.c_int]
        L.gb200_reload_tuning.argtypes = [C.c_void_p]
        L.gb200_ivfpq_compact.argtypes = [C.c_void_p, C.c_int32, C.c_void_p]
        L.gb200_ivfpq_replace_list.argtypes = [C.c_void_p, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p]
        L.gb200_ivfpq_encode.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.gb200_ivfpq_add_raw.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
        L.gb200_comm_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int64, C.c_void_p, C.c_void_p]
        L.gb200_comm_connect.argtypes = [C.c_void_p, C.c_void_p]
        L.gb200_comm_destroy.argtypes = [C.c_void_p]
        L.gb200_comm_slot_bytes.argtypes = [C.c_void_p]
        L.gb200_comm_slot_bytes.restype = C.c_int64
        L.gb200_comm_status.argtypes = [C.c_void_p]
        L.gb200_comm_read.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int]
        L.gb200_comm_buffers.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.gb200_comm_exchange.argtypes = [C.c_void_p, C.c_int64, C.c_void_p]
        L.gb200_ivfpq_search_sharded.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p,
                                                 C.c_void_p, C.c_void_p, C.c_void_p]
        L.gb200_ivfpq_search_sharded_deferred.argtypes = L.gb200_ivfpq_search_sharded.argtypes
        L.gb200_comm_flush.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.gb200_ivfflat_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.gb200_ivfflat_set_quantizer.argtypes = [C.c_void_p, C.c_void_p]
        L.gb200_ivfflat_append.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
        L.gb200_ivfflat_add_raw.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]
        L.gb200_ivfflat_search.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                           C.c_void_p, C.c_void_p]
        L.gb200_ivfpq_set_opq.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.gb200_ivfpq_add_stored.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]
        L.gb200_last_scan_kernel_ms.argtypes = [C.c_void_p]
        L.gb200_last_scan_kernel_ms.restype = C.c_float
        L.gb200_sync.argtypes = [C.c_void_p]
        L.gb200_debug_select.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                         C.c_void_p]
        _lib = L
    return _lib


def debug_select(keys, R, cap, batch, threads=256, device=0):
    """test hook: the R smallest of `keys` (u64) through the kernels' streaming selection primitive"""
    keys = np.ascontiguousarray(keys, dtype=np.uint64)
    out = np.empty(R, np.uint64)
    n_out = C.c_int(0)
    rc = lib().gb200_debug_select(device, keys.ctypes.data, keys.size, R, cap, batch, threads, out.ctypes.data,
                                  C.byref(n_out))
    if rc != 0:
        raise RuntimeError("gb200_debug_select rc=%d %s" % (rc, lib().gb200_last_error().decode()))
    return out[:n_out.value]


class GammaB200Error(RuntimeError):
    pass


def _check(rc, what):
    if rc != 0:
        raise GammaB200Error("%s failed: rc=%d (%s)" % (what, rc, lib().gb200_last_error().decode()))


def _metric_of(name_or_int):
    if isinstance(name_or_int, str):
        return METRIC_L2 if name_or_int.lower() == "l2" else METRIC_IP
    return int(name_or_int)


def make_filters(filters):
    """filters: sequence of (min_doc, max_doc, not_in, pass_flags[u8 over min..max]) -> C array + keepalive.

    Builds the RangeQueryResult byte bitmaps exactly as table/range_query_result.h lays them out:
    bit (doc - min_aligned), min_aligned = (min/8)*8, bytes up to max_aligned.
    """
    arr = (RangeFilter * max(len(filters), 1))()
    keep = []
    for i, (mn, mx, not_in, flags) in enumerate(filters):
        mn, mx = int(mn), int(mx)
        min_al = (mn // 8) * 8
        max_al = (mx // 8 + 1) * 8 - 1
        bits = np.zeros(max_al - min_al + 1, np.uint8)
        bits[mn - min_al: mx - min_al + 1] = (np.asarray(flags) != 0)
        by = np.packbits(bits, bitorder="little")
        keep.append(by)
        arr[i] = RangeFilter(mn, mx, min_al, 1 if not_in else 0, by.ctypes.data)
    return arr, keep


class _Base:
    h = None

    def close(self):
        if self.h:
            lib().gb200_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def GetTotalMemBytes(self):
        return lib().gb200_mem_bytes(self.h)

    def upload_raw(self, x, first_vid=None):
        x = np.ascontiguousarray(x, dtype=np.float32)
        if first_vid is None:
            first_vid = lib().gb200_raw_count(self.h)
        _check(lib().gb200_upload_raw(self.h, int(first_vid), x.shape[0], x.ctypes.data), "upload_raw")

    def upload_raw_dev(self, ptr, n, first_vid=None):
        if first_vid is None:
            first_vid = lib().gb200_raw_count(self.h)
        _check(lib().gb200_upload_raw_dev(self.h, int(first_vid), int(n), C.c_void_p(ptr)), "upload_raw_dev")

    def set_deleted(self, docids, deleted=True):
        ids = np.ascontiguousarray(docids, dtype=np.int64).reshape(-1)
        _check(lib().gb200_set_deleted(self.h, ids.ctypes.data, ids.size, 1 if deleted else 0), "set_deleted")

    def Delete(self, ids):
        """RetrievalModel::Delete — liveness is the deleted bitmap (SURVEY §7 'Updates')."""
        self.set_deleted(ids, True)
        return 0

    def set_filters(self, filters):
        arr, keep = make_filters(filters)
        _check(lib().gb200_set_filters(self.h, C.cast(arr, C.c_void_p), len(filters)), "set_filters")

    def sync(self):
        _check(lib().gb200_sync(self.h), "sync")

    def last_scanned_postings(self):
        return lib().gb200_last_scanned_postings(self.h)

    def launch_count(self):
        return lib().gb200_launch_count(self.h)

    def set_profiling(self, on=True):
        lib().gb200_set_profiling(self.h, 1 if on else 0)

    def reload_tuning(self):
        """Re-read the GB200_* tuning environment variables (they are read once at index creation)."""
        lib().gb200_reload_tuning(self.h)

    def last_stage_ms(self):
        out = (C.c_float * 4)()
        lib().gb200_last_stage_ms(self.h, out)
        return dict(coarse=out[0], scan=out[1], rerank=out[2], total=out[3])

    def last_scan_kernel_ms(self):
        return float(lib().gb200_last_scan_kernel_ms(self.h))

    @staticmethod
    def _sp(metric, nprobe, recall_num, has_rank, min_score, max_score):
        return SearchParams(_metric_of(metric), int(nprobe), int(recall_num), 1 if has_rank else 0,
                            float(min_score), float(max_score))


class B200IVFPQ(_Base):
    """Device mirror + search of the reference's "IVFPQ" model (GammaIVFPQIndex)."""

    def __init__(self, device=0):
        self.device = device
        self.h = None

    def Init(self, model_parameters, d, raw_d=None):
        """model_parameters: the reference's JSON (ncentroids, nsubvector, nbits_per_idx, nprobe, metric_type)."""
        mp = json.loads(model_parameters) if isinstance(model_parameters, str) and model_parameters else \
            (model_parameters or {})
        self.nlist = int(mp.get("ncentroids", 2048))
        self.M = int(mp.get("nsubvector", 64))
        self.nbits = int(mp.get("nbits_per_idx", 8))
        self.nprobe = int(mp.get("nprobe", 80))
        mt = mp.get("metric_type", "InnerProduct")
        self.metric = METRIC_L2 if str(mt).lower() == "l2" else METRIC_IP
        self.d = int(d)
        self.raw_d = int(raw_d or d)
        p = IvfpqParams(self.device, self.d, self.raw_d, self.nlist, self.M, self.nbits, self.metric, self.nprobe, 1)
        h = C.c_void_p()
        rc = lib().gb200_ivfpq_create(C.byref(p), C.byref(h))
        if rc != 0:
            return rc
        self.h = h
        return 0

    def set_quantizers(self, coarse_centroids, pq_centroids):
        cc = np.ascontiguousarray(coarse_centroids, dtype=np.float32)
        pq = np.ascontiguousarray(pq_centroids, dtype=np.float32)
        assert cc.shape == (self.nlist, self.d), cc.shape
        assert pq.size == self.M * 256 * (self.d // self.M)
        _check(lib().gb200_ivfpq_set_quantizers(self.h, cc.ctypes.data, pq.ctypes.data), "set_quantizers")

    def append(self, list_no, vids, codes):
        ln = np.ascontiguousarray(list_no, dtype=np.int32)
        v = np.ascontiguousarray(vids, dtype=np.int64)
        c = np.ascontiguousarray(codes, dtype=np.uint8)
        assert c.shape == (ln.size, self.M)
        return lib().gb200_ivfpq_append(self.h, ln.size, ln.ctypes.data, v.ctypes.data, c.ctypes.data)

    def update(self, vid, new_list, code):
        c = np.ascontiguousarray(code, dtype=np.uint8)
        return lib().gb200_ivfpq_update(self.h, int(vid), int(new_list), c.ctypes.data)

    def set_opq(self, A, b=None):
        """OPQ pre-transform y = A x + b of the model (faiss::OPQMatrix), A [d, d]"""
        A = np.ascontiguousarray(A, dtype=np.float32)
        bb = None if b is None else np.ascontiguousarray(b, dtype=np.float32)
        _check(lib().gb200_ivfpq_set_opq(self.h, A.shape[1], A.shape[0], A.ctypes.data,
                                         None if bb is None else bb.ctypes.data), "set_opq")

    def encode(self, x):
        """stage 1 of GammaIVFPQIndex::Add on the device: (list_no [n] i32, codes [n, M] u8)"""
        x = np.ascontiguousarray(x, dtype=np.float32)
        n = x.shape[0]
        ln = np.empty(n, np.int32)
        cd = np.empty((n, self.M), np.uint8)
        _check(lib().gb200_ivfpq_encode(self.h, n, x.ctypes.data, x.shape[1], ln.ctypes.data, cd.ctypes.data), "encode")
        return ln, cd

    def add_raw(self, x, first_vid):
        """the whole Add: raw upload + device encode + append; returns (list_no, codes) as appended"""
        x = np.ascontiguousarray(x, dtype=np.float32)
        n = x.shape[0]
        ln = np.empty(n, np.int32)
        cd = np.empty((n, self.M), np.uint8)
        _check(lib().gb200_ivfpq_add_raw(self.h, int(first_vid), n, x.ctypes.data, ln.ctypes.data, cd.ctypes.data), "add_raw")
        return ln, cd

    def add_stored(self, first_vid, n, want_codes=False):
        """encode + append rows [first_vid, first_vid + n) of the device raw store"""
        if not want_codes:
            _check(lib().gb200_ivfpq_add_stored(self.h, int(first_vid), int(n), None, None), "add_stored")
            return None
        ln = np.empty(n, np.int32)
        cd = np.empty((n, self.M), np.uint8)
        _check(lib().gb200_ivfpq_add_stored(self.h, int(first_vid), int(n), ln.ctypes.data, cd.ctypes.data), "add_stored")
        return ln, cd

    def replace_list(self, list_no, ids, codes):
        """device copy of one list := the given reference-layout content (ids with kDelIdxMask in bit 63, AoS codes)"""
        i = np.ascontiguousarray(ids, dtype=np.int64)
        c = np.ascontiguousarray(codes, dtype=np.uint8)
        return lib().gb200_ivfpq_replace_list(self.h, int(list_no), int(i.size), i.ctypes.data, c.ctypes.data)

    def compact(self, list_no=-1):
        """RealTimeMemData::CompactBucket on the device; returns the number of postings dropped."""
        dropped = C.c_int64(0)
        _check(lib().gb200_ivfpq_compact(self.h, int(list_no), C.byref(dropped)), "compact")
        return int(dropped.value)

    def list_sizes(self):
        out = np.empty(self.nlist, np.int64)
        _check(lib().gb200_ivfpq_list_sizes(self.h, out.ctypes.data), "list_sizes")
        return out

    def get_list(self, list_no):
        n = int(self.list_sizes()[list_no])
        ids = np.empty(n, np.int64)
        codes = np.empty((n, self.M), np.uint8)
        _check(lib().gb200_ivfpq_get_list(self.h, list_no, ids.ctypes.data, codes.ctypes.data), "get_list")
        return ids, codes

    def coarse(self, xq, nprobe):
        xq = np.ascontiguousarray(xq, dtype=np.float32)
        n = xq.shape[0]
        cd = np.empty((n, nprobe), np.float32)
        keys = np.empty((n, nprobe), np.int64)
        _check(lib().gb200_ivfpq_coarse(self.h, n, xq.ctypes.data, nprobe, cd.ctypes.data, keys.ctypes.data), "coarse")
        return cd, keys

    def Search(self, xq, k, nprobe=-1, recall_num=100, metric=None, has_rank=True, min_score=-FLT_MAX,
               max_score=FLT_MAX, filters=(), keys=None, coarse_dis=None):
        """RetrievalModel::Search; returns (rc, distances[n,k] f32, labels[n,k] i64)."""
        xq = np.ascontiguousarray(xq, dtype=np.float32)
        n = xq.shape[0]
        D = np.empty((n, k), np.float32)
        I = np.empty((n, k), np.int64)
        sp = self._sp(self.metric if metric is None else metric, nprobe, recall_num, has_rank, min_score, max_score)
        arr, keep = make_filters(filters)
        if keys is None:
            rc = lib().gb200_ivfpq_search(self.h, n, xq.ctypes.data, k, C.byref(sp), C.cast(arr, C.c_void_p),
                                          len(filters), D.ctypes.data, I.ctypes.data)
        else:
            keys = np.ascontiguousarray(keys, dtype=np.int64)
            coarse_dis = np.ascontiguousarray(coarse_dis, dtype=np.float32)
            rc = lib().gb200_ivfpq_search_preassigned(self.h, n, xq.ctypes.data, k, C.byref(sp),
                                                      C.cast(arr, C.c_void_p), len(filters), keys.ctypes.data,
                                                      coarse_dis.ctypes.data, keys.shape[1], D.ctypes.data,
                                                      I.ctypes.data)
        return rc, D, I

    def search_dev(self, xq_ptr, n, k, D_ptr, I_ptr, stream_ptr, nprobe=-1, recall_num=100, metric=None,
                   has_rank=True, min_score=-FLT_MAX, max_score=FLT_MAX):
        sp = self._sp(self.metric if metric is None else metric, nprobe, recall_num, has_rank, min_score, max_score)
        return lib().gb200_ivfpq_search_dev(self.h, n, xq_ptr, k, C.byref(sp), D_ptr, I_ptr, stream_ptr)

    def flat_search(self, xq, k, metric=None, min_score=-FLT_MAX, max_score=FLT_MAX, filters=()):
        """brute_force_search / untrained fallback of the IVFPQ model (gamma_index_ivfpq.cc:529-537)."""
        return _flat_search(self, xq, k, self.metric if metric is None else metric, min_score, max_score, filters)


def _flat_search(obj, xq, k, metric, min_score, max_score, filters):
    xq = np.ascontiguousarray(xq, dtype=np.float32)
    n = xq.shape[0]
    D = np.empty((n, k), np.float32)
    I = np.empty((n, k), np.int64)
    sp = _Base._sp(metric, -1, 0, 0, min_score, max_score)
    arr, keep = make_filters(filters)
    rc = lib().gb200_flat_search(obj.h, n, xq.ctypes.data, k, C.byref(sp), C.cast(arr, C.c_void_p), len(filters),
                                 D.ctypes.data, I.ctypes.data)
    return rc, D, I


class B200IVFFLAT(B200IVFPQ):
    """Device mirror + search of the reference's "IVFFLAT" model (GammaIndexIVFFlat): lists of vids over the raw store."""

    def Init(self, model_parameters, d, raw_d=None):
        mp = json.loads(model_parameters) if isinstance(model_parameters, str) and model_parameters else \
            (model_parameters or {})
        self.d = self.raw_d = int(d)
        self.nlist = int(mp.get("ncentroids", 2048))
        self.M, self.nbits = 4, 8
        self.metric = METRIC_L2 if str(mp.get("metric_type", "L2")).lower() == "l2" else METRIC_IP
        self.nprobe = int(mp.get("nprobe", 80))
        h = C.c_void_p()
        rc = lib().gb200_ivfflat_create(self.device, self.d, self.nlist, self.metric, self.nprobe, C.byref(h))
        if rc != 0:
            return rc
        self.h = h
        return 0

    def set_quantizer(self, coarse):
        c = np.ascontiguousarray(coarse, dtype=np.float32)
        assert c.shape == (self.nlist, self.d)
        _check(lib().gb200_ivfflat_set_quantizer(self.h, c.ctypes.data), "ivfflat_set_quantizer")

    def append_vids(self, list_no, vids):
        ln = np.ascontiguousarray(list_no, dtype=np.int32)
        v = np.ascontiguousarray(vids, dtype=np.int64)
        return lib().gb200_ivfflat_append(self.h, ln.size, ln.ctypes.data, v.ctypes.data)

    def add_raw(self, x, first_vid):
        x = np.ascontiguousarray(x, dtype=np.float32)
        ln = np.empty(x.shape[0], np.int32)
        _check(lib().gb200_ivfflat_add_raw(self.h, int(first_vid), x.shape[0], x.ctypes.data, ln.ctypes.data), "ivfflat_add_raw")
        return ln

    def Search(self, xq, k, nprobe=-1, metric=None, min_score=-FLT_MAX, max_score=FLT_MAX, filters=()):
        xq = np.ascontiguousarray(xq, dtype=np.float32)
        n = xq.shape[0]
        D = np.empty((n, k), np.float32)
        I = np.empty((n, k), np.int64)
        sp = self._sp(self.metric if metric is None else metric, nprobe, 0, False, min_score, max_score)
        arr, keep = make_filters(filters)
        rc = lib().gb200_ivfflat_search(self.h, n, xq.ctypes.data, k, C.byref(sp), C.cast(arr, C.c_void_p), len(filters),
                                        D.ctypes.data, I.ctypes.data)
        return rc, D, I


class B200FLAT(_Base):
    """Device mirror + search of the reference's "FLAT" model (GammaFLATIndex)."""

    def __init__(self, device=0):
        self.device = device
        self.h = None

    def Init(self, model_parameters, d):
        mp = json.loads(model_parameters) if isinstance(model_parameters, str) and model_parameters else \
            (model_parameters or {})
        mt = mp.get("metric_type", "InnerProduct")  # FLATModelParams default (gamma_index_flat.cc:28-56)
        self.metric = METRIC_L2 if str(mt).lower() == "l2" else METRIC_IP
        self.d = int(d)
        h = C.c_void_p()
        rc = lib().gb200_flat_create(self.device, self.d, self.metric, C.byref(h))
        if rc != 0:
            return rc
        self.h = h
        return 0

    def Add(self, x):
        self.upload_raw(x)
        return True

    def Search(self, xq, k, metric=None, min_score=-FLT_MAX, max_score=FLT_MAX, filters=()):
        return _flat_search(self, xq, k, self.metric if metric is None else metric, min_score, max_score, filters)

    def search_dev(self, xq_ptr, n, k, D_ptr, I_ptr, stream_ptr, metric=None, min_score=-FLT_MAX, max_score=FLT_MAX):
        sp = self._sp(self.metric if metric is None else metric, -1, 0, 0, min_score, max_score)
        return lib().gb200_flat_search_dev(self.h, n, xq_ptr, k, C.byref(sp), D_ptr, I_ptr, stream_ptr)


COMM_HANDLE_BYTES = 128


class Comm:
    """One rank's endpoint of the multi-GPU result exchange (gb200_comm): peer stores over NVLink, no collective."""

    def __init__(self, device, rank, world, slot_bytes):
        self.rank, self.world = rank, world
        self.h = C.c_void_p()
        self.handle = (C.c_uint8 * COMM_HANDLE_BYTES)()
        _check(lib().gb200_comm_create(device, rank, world, int(slot_bytes), C.byref(self.h), self.handle), "comm_create")
        self.slot_bytes = int(lib().gb200_comm_slot_bytes(self.h))

    def handle_bytes(self):
        return bytes(self.handle)

    def connect(self, all_handles):
        """all_handles: world entries of COMM_HANDLE_BYTES bytes, rank-major"""
        blob = b"".join(all_handles)
        assert len(blob) == self.world * COMM_HANDLE_BYTES
        buf = (C.c_uint8 * len(blob)).from_buffer_copy(blob)
        _check(lib().gb200_comm_connect(self.h, buf), "comm_connect")

    def buffers(self):
        mine, allp = C.c_void_p(), C.c_void_p()
        _check(lib().gb200_comm_buffers(self.h, C.byref(mine), C.byref(allp)), "comm_buffers")
        return mine.value, allp.value

    def exchange(self, nbytes, stream_ptr):
        _check(lib().gb200_comm_exchange(self.h, int(nbytes), stream_ptr), "comm_exchange")

    def search_sharded(self, ix, xq_ptr, n, k, stream_ptr, nprobe=-1, recall_num=100, metric=None, has_rank=True,
                       deferred=False):
        """returns the device address of the gathered window: rank r's block ([n*k] f32, [n*k] i64) at + r * slot_bytes.
        deferred=True: the window of the PREVIOUS call (None on the first one); flush() gives the last one."""
        sp = ix._sp(ix.metric if metric is None else metric, nprobe, recall_num, has_rank, -FLT_MAX, FLT_MAX)
        D_all = C.c_void_p()
        fn = lib().gb200_ivfpq_search_sharded_deferred if deferred else lib().gb200_ivfpq_search_sharded
        _check(fn(ix.h, self.h, n, xq_ptr, k, C.byref(sp), C.byref(D_all), None, stream_ptr), "search_sharded")
        return D_all.value

    def flush(self, stream_ptr):
        """wait on the stream for the peers' results of the last exchange; returns that window's device address"""
        allp = C.c_void_p()
        _check(lib().gb200_comm_flush(self.h, C.byref(allp), stream_ptr), "comm_flush")
        return allp.value

    def status(self):
        return int(lib().gb200_comm_status(self.h))

    def read(self, dst_host_ptr, src_dev_ptr, nbytes, stream_ptr, sync=False):
        _check(lib().gb200_comm_read(self.h, dst_host_ptr, src_dev_ptr, int(nbytes), stream_ptr, 1 if sync else 0), "comm_read")

    def close(self):
        if self.h:
            lib().gb200_comm_destroy(self.h)
            self.h = None
