"""Python host-side mirror of the reference interface for the descriptor-matching path, over the C ABI.

The product is the C-ABI library (include/b200match.h) plus the C++ adaptors in alicevision_b200/adaptor/.
This module binds the same C entry points with ctypes so that the parity tests and the bench read like the
reference's own tests (same method names and argument meaning):

* ``ArrayMatcherB200``            <->  matching::ArrayMatcher<Scalar,Metric>      (matching/ArrayMatcher.hpp:19-66)
* ``ImageCollectionMatcherB200``  <->  matchingImageCollection::IImageCollectionMatcher (IImageCollectionMatcher.hpp:28-42)
* ``EMatcherType``                <->  matching::EMatcherType                      (matching/matcherType.hpp:15-22)

Nothing here computes distances: if the CUDA library is missing or no GPU is present, construction fails loudly.
"""
from __future__ import annotations

import atexit
import ctypes as C
import os
import weakref
from enum import IntEnum

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb200match.so")

F32, U8, BIN = 0, 1, 2
L2_SIMPLE, L2_VECTORIZED, HAMMING = 0, 1, 2
STAGE_DEVICE, STAGE_RAW, STAGE_FULL = 0, 1, 2

MATCH_DTYPE = np.dtype([("i", np.uint32), ("j", np.uint32), ("ratio", np.float32), ("dist", np.float32)])

# every symbol include/b200match.h declares (checked by tests/test_abi.py)
ABI_SYMBOLS = [
    "b200m_last_error", "b200m_version", "b200m_device_count", "b200m_ctx_create", "b200m_ctx_destroy", "b200m_ctx_set_host_threads",
    "b200m_ctx_set_force_exact", "b200m_ctx_set_tc_variant", "b200m_debug_trace", "b200m_debug_convert_f32_u8", "b200m_db_create", "b200m_db_destroy", "b200m_knn", "b200m_upload_view", "b200m_upload_views", "b200m_upload_views_async", "b200m_wait_uploads", "b200m_clear_views", "b200m_remove_view",
    "b200m_match_pairs", "b200m_result_num_pairs", "b200m_result_get", "b200m_result_free", "b200m_last_gpu_ms",
    "b200m_last_search_kernel_ms", "b200m_last_launches", "b200m_last_tc_pairs", "b200m_exactness_errors", "b200m_last_records", "b200m_last_real_tc_pairs", "b200m_last_fallback_rows",
    "b200m_shard_pairs", "b200m_shard_pairs_2d", "b200m_multi_create", "b200m_multi_destroy", "b200m_multi_num_devices", "b200m_multi_ctx", "b200m_multi_match",
    "b200m_multi_last_gpu_ms", "b200m_guided_match", "b200m_guided_match_model",
]


class EMatcherType(IntEnum):
    """matching/matcherType.hpp:15-22 plus the two values an integration adds for this engine."""
    BRUTE_FORCE_L2 = 0
    ANN_L2 = 1
    CASCADE_HASHING_L2 = 2
    FAST_CASCADE_HASHING_L2 = 3
    BRUTE_FORCE_HAMMING = 4
    BRUTE_FORCE_L2_B200 = 5
    BRUTE_FORCE_HAMMING_B200 = 6


class B200MatchError(RuntimeError):
    pass


_lib = None


def load_library() -> C.CDLL:
    """Load libb200match.so (no fallback: a missing library is an error)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise B200MatchError(f"{LIB_PATH} not found: run `python -c 'import __graft_entry__ as g; g.build()'` (there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    lib.b200m_last_error.restype = C.c_char_p
    lib.b200m_last_gpu_ms.restype = C.c_double
    lib.b200m_last_search_kernel_ms.restype = C.c_double
    lib.b200m_last_records.restype = C.c_int64
    lib.b200m_last_fallback_rows.restype = C.c_int64
    lib.b200m_exactness_errors.restype = C.c_uint
    lib.b200m_multi_last_gpu_ms.restype = C.c_double
    lib.b200m_multi_ctx.restype = C.c_void_p
    for name in ("b200m_ctx_destroy", "b200m_db_destroy", "b200m_result_free", "b200m_multi_destroy"):
        getattr(lib, name).restype = None
    _lib = lib
    return lib


def _check(rc: int, what: str) -> None:
    if rc != 0:
        raise B200MatchError(f"{what} failed (status {rc}): {load_library().b200m_last_error().decode()}")


def _dtype_code(a: np.ndarray, binary: bool) -> int:
    if a.dtype == np.float32:
        return F32
    if a.dtype == np.uint8:
        return BIN if binary else U8
    raise TypeError(f"unsupported descriptor dtype {a.dtype} (float32 / uint8 only)")


_live_contexts: "weakref.WeakSet[Context]" = weakref.WeakSet()


@atexit.register
def _drain_uploads_at_exit() -> None:
    """Asynchronous uploads read numpy memory from an engine thread: let them finish before the interpreter frees it."""
    for ctx in list(_live_contexts):
        try:
            if ctx._h:
                ctx.lib.b200m_wait_uploads(ctx._h)
        except Exception:
            pass


class Context:
    """One engine instance bound to one GPU (``b200m_ctx``)."""

    def __init__(self, device: int = 0, stream: int | None = None):
        self.lib = load_library()
        self._h = C.c_void_p()
        _check(self.lib.b200m_ctx_create(C.c_int(device), C.c_void_p(stream or 0), C.byref(self._h)), "b200m_ctx_create")
        _live_contexts.add(self)

    def close(self) -> None:
        if self._h:
            self.lib.b200m_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_host_threads(self, n: int) -> None:
        _check(self.lib.b200m_ctx_set_host_threads(self._h, C.c_int(n)), "b200m_ctx_set_host_threads")

    def set_force_exact(self, on: bool) -> None:
        _check(self.lib.b200m_ctx_set_force_exact(self._h, C.c_int(int(on))), "b200m_ctx_set_force_exact")

    def set_tc_variant(self, variant: int) -> None:
        _check(self.lib.b200m_ctx_set_tc_variant(self._h, C.c_int(variant)), "b200m_ctx_set_tc_variant")

    def debug_trace(self, enable: bool = True, read: bool = False):
        """Enable / read the per-tile pipeline trace (roles x 512 tiles x 4 stamps) of CTA 0 of the tensor-core kernel."""
        out = np.zeros((4, 512, 4), np.int64)
        _check(self.lib.b200m_debug_trace(self._h, C.c_int(int(enable)), out.ctypes.data_as(C.c_void_p) if read else None, C.c_int(out.size)), "b200m_debug_trace")
        return out

    # instrumentation
    def last_gpu_ms(self) -> float: return self.lib.b200m_last_gpu_ms(self._h)
    def last_search_kernel_ms(self) -> float: return self.lib.b200m_last_search_kernel_ms(self._h)
    def last_launches(self) -> int: return self.lib.b200m_last_launches(self._h)
    def last_tc_pairs(self) -> int: return self.lib.b200m_last_tc_pairs(self._h)
    def exactness_errors(self) -> int: return self.lib.b200m_exactness_errors(self._h)
    def last_records(self) -> int: return self.lib.b200m_last_records(self._h)
    def last_real_tc_pairs(self) -> int: return self.lib.b200m_last_real_tc_pairs(self._h)
    def last_fallback_rows(self) -> int: return self.lib.b200m_last_fallback_rows(self._h)


_default_ctx: dict[int, Context] = {}


def default_context(device: int = 0) -> Context:
    if device not in _default_ctx:
        _default_ctx[device] = Context(device)
    return _default_ctx[device]


class _ResultOwner:
    """Owns one b200m_result; freed when the last numpy view onto it is garbage collected."""

    def __init__(self, lib, handle):
        self.lib, self.handle = lib, handle

    def __del__(self):
        try:
            if self.handle:
                self.lib.b200m_result_free(self.handle)
                self.handle = None
        except Exception:
            pass


def _alias(addr, nbytes: int, dtype, owner) -> np.ndarray:
    """Zero-copy numpy view of engine-owned memory (keeps `owner` alive through the buffer object)."""
    if not addr or nbytes <= 0:
        return np.zeros(0, dtype)
    buf = (C.c_uint8 * nbytes).from_address(addr)
    buf._owner = owner
    return np.frombuffer(buf, dtype=dtype)


class ArrayMatcherB200:
    """Mirror of matching::ArrayMatcher (Build / SearchNeighbour / SearchNeighbours), bool returns, no exceptions
    on the matcher surface (ArrayMatcher_bruteForce.hpp:42-51,63-85,98-142)."""

    def __init__(self, metric: int = L2_SIMPLE, ctx: Context | None = None, binary: bool = False):
        self.ctx = ctx or default_context()
        self.metric = HAMMING if binary else metric
        self.binary = binary
        self._db = C.c_void_p()
        self._dim = 0
        self._dtype = None

    def Build(self, dataset: np.ndarray, nbRows: int | None = None, dimension: int | None = None) -> bool:
        self._release()
        a = np.ascontiguousarray(dataset)
        rows = a.shape[0] if nbRows is None else nbRows
        dim = (a.shape[1] if a.ndim == 2 else 0) if dimension is None else dimension
        if rows < 1:
            return False
        rc = self.ctx.lib.b200m_db_create(self.ctx._h, a.ctypes.data_as(C.c_void_p), C.c_int(rows), C.c_int(dim),
                                          C.c_int(_dtype_code(a, self.binary)), C.c_int(self.metric), C.byref(self._db))
        self._dim, self._dtype = dim, a.dtype
        return rc == 0

    def SearchNeighbours(self, query: np.ndarray, nbQuery: int | None = None, NN: int = 2):
        """Returns (ok, indices[nq,NN] int32 database rows, distances[nq,NN] float32|uint32)."""
        q = np.ascontiguousarray(query, dtype=self._dtype) if self._dtype is not None else np.ascontiguousarray(query)
        nq = (q.shape[0] if q.ndim == 2 else (1 if q.size else 0)) if nbQuery is None else nbQuery
        idx = np.zeros((max(nq, 1), NN), np.int32)
        dist = np.zeros((max(nq, 1), NN), np.uint32 if self.metric == HAMMING else np.float32)
        if not self._db:
            return False, idx[:0], dist[:0]
        rc = self.ctx.lib.b200m_knn(self.ctx._h, self._db, q.ctypes.data_as(C.c_void_p), C.c_int(nq), C.c_int(NN),
                                    idx.ctypes.data_as(C.c_void_p), dist.ctypes.data_as(C.c_void_p))
        if rc != 0:
            return False, idx[:0], dist[:0]
        return True, idx[:nq], dist[:nq]

    def SearchNeighbour(self, query: np.ndarray):
        """Single nearest neighbour: returns (ok, index, distance)."""
        ok, idx, dist = self.SearchNeighbours(np.ascontiguousarray(query).reshape(1, -1), 1, 1)
        if not ok:
            return False, -1, -1.0
        return True, int(idx[0, 0]), dist[0, 0]

    def _release(self):
        if self._db:
            self.ctx.lib.b200m_db_destroy(self._db)
            self._db = C.c_void_p()

    def __del__(self):
        try:
            self._release()
        except Exception:
            pass


def shard_pairs(pairs, n_shards: int) -> np.ndarray:
    """b200m_shard_pairs: shard index of every pair (database images dealt round-robin, alternating direction). Host only."""
    p = np.ascontiguousarray(pairs, np.uint32).reshape(-1, 2)
    out = np.zeros(max(p.shape[0], 1), np.int32)
    _check(load_library().b200m_shard_pairs(p.ctypes.data_as(C.c_void_p), C.c_int(p.shape[0]), C.c_int(n_shards), out.ctypes.data_as(C.c_void_p)),
           "b200m_shard_pairs")
    return out[: p.shape[0]]


def shard_pairs_2d(pairs, n_shards: int) -> np.ndarray:
    """b200m_shard_pairs_2d: shard index of every pair with 2-D block sharding (a shard needs only part of the views). Host only."""
    p = np.ascontiguousarray(pairs, np.uint32).reshape(-1, 2)
    out = np.zeros(max(p.shape[0], 1), np.int32)
    _check(load_library().b200m_shard_pairs_2d(p.ctypes.data_as(C.c_void_p), C.c_int(p.shape[0]), C.c_int(n_shards), out.ctypes.data_as(C.c_void_p)),
           "b200m_shard_pairs_2d")
    return out[: p.shape[0]]


class MultiContext:
    """``b200m_multi``: one engine context per device, driven by one host thread each (single-process multi-GPU)."""

    def __init__(self, devices):
        self.lib = load_library()
        self.devices = [int(d) for d in devices]
        self._h = C.c_void_p()
        arr = (C.c_int * len(self.devices))(*self.devices)
        _check(self.lib.b200m_multi_create(arr, C.c_int(len(self.devices)), C.byref(self._h)), "b200m_multi_create")

    def close(self) -> None:
        if self._h:
            self.lib.b200m_multi_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def last_gpu_ms(self) -> float: return self.lib.b200m_multi_last_gpu_ms(self._h)

    def exactness_errors(self) -> int:
        return sum(self.lib.b200m_exactness_errors(C.c_void_p(self.lib.b200m_multi_ctx(self._h, C.c_int(k)))) for k in range(len(self.devices)))


class ImageCollectionMatcherB200:
    """Mirror of IImageCollectionMatcher / ImageCollectionMatcher_generic (distRatio, crossMatching, matcherType).
    ``devices``: several CUDA ordinals -> the pair list is sharded over one context per device inside this process."""

    def __init__(self, distRatio: float = 0.8, crossMatching: bool = False, matcherType: EMatcherType = EMatcherType.BRUTE_FORCE_L2_B200,
                 ctx: Context | None = None, devices=None):
        if matcherType not in (EMatcherType.BRUTE_FORCE_L2_B200, EMatcherType.BRUTE_FORCE_HAMMING_B200, EMatcherType.BRUTE_FORCE_L2,
                               EMatcherType.BRUTE_FORCE_HAMMING):
            raise IndexError("Invalid matcherType enum")    # matchingCommon.cpp:41-42 throws std::out_of_range
        self.multi = MultiContext(devices) if devices is not None and len(devices) > 1 else None
        self.ctx = ctx or (default_context(devices[0] if devices else 0) if self.multi is None else None)
        self.distRatio, self.crossMatching, self.matcherType = float(distRatio), bool(crossMatching), matcherType
        self.hamming = matcherType in (EMatcherType.BRUTE_FORCE_HAMMING_B200, EMatcherType.BRUTE_FORCE_HAMMING)
        self._keepalive = []      # descriptor arrays of asynchronous uploads still in flight

    def upload(self, regionsPerView: dict) -> None:
        """regionsPerView: {viewId: (descriptors[n,dim], positions[n,2] or None)} for ONE descriptor type.
        Asynchronous (b200m_upload_views_async): the copies overlap the first batches of the next match call; the
        descriptor arrays are kept alive here until that call (or clear / wait_uploads) returns."""
        lib = self.ctx.lib
        groups = {}                                   # one bulk call per (dim, element type)
        for vid, (desc, xy) in regionsPerView.items():
            d = np.ascontiguousarray(desc)
            dim = d.shape[1] if d.ndim == 2 else 0
            binary = d.dtype == np.uint8 and self.hamming
            xya = None if xy is None else np.ascontiguousarray(xy, np.float32)
            groups.setdefault((max(dim, 1), _dtype_code(d, binary)), []).append((vid, d, xya))
        for (dim, code), items in groups.items():
            n = len(items)
            ids = np.array([v for v, _, _ in items], np.uint32)
            counts = np.array([d.shape[0] for _, d, _ in items], np.int32)
            dptr = (C.c_void_p * n)(*[d.ctypes.data if d.shape[0] else None for _, d, _ in items])
            have_xy = any(x is not None for _, _, x in items)
            xptr = (C.c_void_p * n)(*[(x.ctypes.data if x is not None and x.size else None) for _, _, x in items]) if have_xy else None
            self._keepalive.append(items)
            _check(lib.b200m_upload_views_async(self.ctx._h, C.c_int(n), ids.ctypes.data_as(C.c_void_p), dptr, counts.ctypes.data_as(C.c_void_p),
                                                C.c_int(dim), C.c_int(code), xptr), "b200m_upload_views_async")

    def wait_uploads(self) -> None:
        try:
            _check(self.ctx.lib.b200m_wait_uploads(self.ctx._h), "b200m_wait_uploads")
        finally:
            self._keepalive.clear()

    def clear(self) -> None:
        try:
            _check(self.ctx.lib.b200m_clear_views(self.ctx._h), "b200m_clear_views")
        finally:
            self._keepalive.clear()

    def match_uploaded(self, pairs, stage: int = STAGE_FULL):
        """Runs b200m_match_pairs on already-uploaded views. Returns (pair_ids[n,2], offsets[n+1], matches[MATCH_DTYPE])."""
        lib = self.ctx.lib
        p = np.ascontiguousarray(pairs, np.uint32).reshape(-1, 2)
        res = C.c_void_p()
        try:
            _check(lib.b200m_match_pairs(self.ctx._h, p.ctypes.data_as(C.c_void_p), C.c_int(p.shape[0]), C.c_float(self.distRatio),
                                         C.c_int(int(self.crossMatching)), C.c_int(stage), C.byref(res)), "b200m_match_pairs")
        finally:
            self._keepalive.clear()      # b200m_match_pairs returns after every pending upload has left caller memory
        return self._wrap_result(lib, res)

    def _wrap_result(self, lib, res):
        owner = _ResultOwner(lib, res)      # the numpy arrays below alias the result's memory; it is freed when they die
        n = lib.b200m_result_num_pairs(res)
        pid, off, mat = C.c_void_p(), C.c_void_p(), C.c_void_p()
        _check(lib.b200m_result_get(res, C.byref(pid), C.byref(off), C.byref(mat)), "b200m_result_get")
        offsets = _alias(off.value, (n + 1) * 8, np.int64, owner)
        pair_ids = _alias(pid.value, n * 8, np.uint32, owner).reshape(-1, 2)
        matches = _alias(mat.value, int(offsets[-1]) * MATCH_DTYPE.itemsize, MATCH_DTYPE, owner)
        return pair_ids, offsets, matches

    def _match_multi(self, regionsPerView: dict, pairs):
        """b200m_multi_match: all views of ONE element type, pair list sharded over the devices."""
        lib = self.multi.lib
        items = []
        for vid, (desc, xy) in regionsPerView.items():
            d = np.ascontiguousarray(desc)
            items.append((vid, d, None if xy is None else np.ascontiguousarray(xy, np.float32)))
        codes = {(_dtype_code(d, d.dtype == np.uint8 and self.hamming), d.shape[1] if d.ndim == 2 else 0) for _, d, _ in items}
        if len(codes) == 0:      # nothing referenced (empty pair list): the empty result
            return np.zeros((0, 2), np.uint32), np.zeros(1, np.int64), np.zeros(0, MATCH_DTYPE)
        if len(codes) != 1:
            raise B200MatchError("the multi-device path takes views of one descriptor type per call")
        (code, dim), = codes
        n = len(items)
        ids = np.array([v for v, _, _ in items], np.uint32)
        counts = np.array([d.shape[0] for _, d, _ in items], np.int32)
        dptr = (C.c_void_p * n)(*[d.ctypes.data if d.shape[0] else None for _, d, _ in items])
        xptr = (C.c_void_p * n)(*[(x.ctypes.data if x is not None and x.size else None) for _, _, x in items])
        p = np.ascontiguousarray(pairs, np.uint32).reshape(-1, 2)
        res = C.c_void_p()
        _check(lib.b200m_multi_match(self.multi._h, C.c_int(n), ids.ctypes.data_as(C.c_void_p), dptr, counts.ctypes.data_as(C.c_void_p), C.c_int(max(dim, 1)),
                                     C.c_int(code), xptr, p.ctypes.data_as(C.c_void_p), C.c_int(p.shape[0]), C.c_float(self.distRatio),
                                     C.c_int(int(self.crossMatching)), C.byref(res)), "b200m_multi_match")
        return self._wrap_result(lib, res)

    def Match(self, regionsPerView: dict, pairs, map_PutativesMatches: dict | None = None) -> dict:
        """IImageCollectionMatcher::Match: appends {(I, J): matches} for every pair with a non-empty result
        (ImageCollectionMatcher_generic.cpp:116-119: empty lists are not inserted; the output map is appended to).
        Only the regions the pair list references are touched (the reference calls getRegions for those only, :55,:72).
        Without an output map the result is a ``PairwiseMatches`` mapping (a dict view onto the engine's result arrays)."""
        p = np.ascontiguousarray(pairs, np.uint32).reshape(-1, 2)
        used = set(np.unique(p).tolist())
        needed = {v: r for v, r in regionsPerView.items() if v in used} if len(used) < len(regionsPerView) else regionsPerView
        if self.multi is not None:
            pair_ids, offsets, matches = self._match_multi(needed, p)
        else:
            self.upload(needed)
            pair_ids, offsets, matches = self.match_uploaded(p, STAGE_FULL)
        res = PairwiseMatches(pair_ids, offsets, matches)
        if map_PutativesMatches is None:
            return res
        map_PutativesMatches.update(res)
        return map_PutativesMatches


class PairwiseMatches(dict):
    """{(I, J): matches[MATCH_DTYPE]} of one Match call for the pairs with a non-empty result, in PairSet order
    (matching::PairwiseMatches, matching/IndMatch.hpp:148).  A dict whose entries are views into the engine's result arena;
    the entries are materialised on first access, so a caller that only forwards the arrays (``pair_ids``, ``offsets``,
    ``matches``: all pairs, CSR layout) pays nothing for the map."""

    def __init__(self, pair_ids, offsets, matches):
        super().__init__()
        self.pair_ids, self.offsets, self.matches = pair_ids, offsets, matches
        self._filled = False

    def _fill(self):
        if not self._filled:
            self._filled = True
            offs = self.offsets.tolist()
            m = self.matches
            for (i, j), a, b in zip(self.pair_ids.tolist(), offs[:-1], offs[1:]):
                if b > a:
                    dict.__setitem__(self, (i, j), m[a:b])
        return self

    def num_matches(self) -> int:
        return int(self.offsets[-1]) if len(self.offsets) else 0


def _lazy(name):
    def f(self, *a, **k):
        return getattr(dict, name)(self._fill(), *a, **k)
    f.__name__ = name
    return f


for _n in ("__getitem__", "__contains__", "__iter__", "__len__", "__eq__", "__ne__", "__repr__", "keys", "values", "items", "get", "copy", "__bool__",
           "pop", "__setitem__", "__delitem__", "setdefault", "update", "__reversed__", "__or__", "__ror__"):
    if hasattr(dict, _n):
        setattr(PairwiseMatches, _n, _lazy(_n))
PairwiseMatches.__bool__ = lambda self: len(self) > 0
PairwiseMatches.__hash__ = None


MODEL_FUNDAMENTAL, MODEL_HOMOGRAPHY = 0, 1


def guidedMatching(F, regions_left: "Regions", regions_right: "Regions", errorTh: float, distRatio: float, ctx: Context | None = None,
                   model: int = MODEL_FUNDAMENTAL) -> np.ndarray:
    """matching::guidedMatching<Mat3Model, FundamentalEpipolarDistanceError> (matching/guidedMatching.hpp:206-268) for cameras
    without distortion: F = 3x3 fundamental matrix (x_right^T F x_left = 0), errorTh / distRatio already squared as at the
    call site (GeometricFilterMatrix_F_AC.hpp:387-388).  model = MODEL_HOMOGRAPHY: F is a homography and the error is
    HomographyAsymmetricError (GeometricFilterMatrix_H_AC.hpp:217-225).  Returns matches[MATCH_DTYPE] with i = left, j = right feature."""
    ctx = ctx or default_context()
    lib = ctx.lib
    ids = (_new_view_id(), _new_view_id())
    Fm = np.ascontiguousarray(F, np.float64).reshape(9)
    up = []
    try:
        for vid, r in zip(ids, (regions_left, regions_right)):
            n = r.RegionCount()
            code = _dtype_code(r.descriptors, r.binary)
            _check(lib.b200m_upload_view(ctx._h, C.c_uint32(vid), r.descriptors.ctypes.data_as(C.c_void_p) if n else None, C.c_int(n),
                                         C.c_int(max(r.DescriptorLength(), 1)), C.c_int(code), r.positions.ctypes.data_as(C.c_void_p) if n else None), "b200m_upload_view")
            up.append(vid)
        res = C.c_void_p()
        _check(lib.b200m_guided_match_model(ctx._h, C.c_uint32(ids[0]), C.c_uint32(ids[1]), C.c_int(model), Fm.ctypes.data_as(C.c_void_p), C.c_double(errorTh),
                                            C.c_double(distRatio), C.byref(res)), "b200m_guided_match_model")
        owner = _ResultOwner(lib, res)
        off, mat = C.c_void_p(), C.c_void_p()
        _check(lib.b200m_result_get(res, None, C.byref(off), C.byref(mat)), "b200m_result_get")
        n = int(_alias(off.value, 16, np.int64, owner)[1])
        return _alias(mat.value, n * MATCH_DTYPE.itemsize, MATCH_DTYPE, owner).copy()
    finally:
        for vid in up:
            lib.b200m_remove_view(ctx._h, C.c_uint32(vid))


def createImageCollectionMatcher(matcherType: EMatcherType, distRatio: float, crossMatching: bool, ctx: Context | None = None):
    """matchingImageCollection/matchingCommon.cpp:19-47 restricted to the matcher types this engine implements."""
    return ImageCollectionMatcherB200(distRatio, crossMatching, matcherType, ctx)


# ---- Surface 1b: IRegionsMatcher / RegionsDatabaseMatcher / DistanceRatioMatch ----------------------------------------
_next_view_id = [0x40000000]


def _new_view_id() -> int:
    _next_view_id[0] += 1
    return _next_view_id[0]


class Regions:
    """The slice of feature::Regions the matchers read (feature/Regions.hpp:46-118): descriptors[n, L], positions[n, 2],
    IsBinary().  ``binary`` distinguishes AKAZE_BinaryRegions (uchar[64] bit strings) from SIFT_Regions (uchar[128])."""

    def __init__(self, descriptors: np.ndarray, positions: np.ndarray | None = None, binary: bool = False):
        self.descriptors = np.ascontiguousarray(descriptors)
        if self.descriptors.ndim != 2:
            raise ValueError("descriptors must be [n, L]")
        n = self.descriptors.shape[0]
        self.positions = np.zeros((n, 2), np.float32) if positions is None else np.ascontiguousarray(positions, np.float32).reshape(n, 2)
        self.binary = bool(binary)

    def RegionCount(self) -> int: return int(self.descriptors.shape[0])
    def DescriptorLength(self) -> int: return int(self.descriptors.shape[1])
    def IsBinary(self) -> bool: return self.binary
    def IsScalar(self) -> bool: return not self.binary
    def Type_id(self) -> str: return "f" if self.descriptors.dtype == np.float32 else "h"


class RegionsMatcherB200:
    """IRegionsMatcher (matching/RegionsMatcher.hpp:49-78) on the engine: the database Regions are uploaded once by the
    constructor, Match(f_dist_ratio, query_regions) runs RegionsMatcher::Match (:126-176) for that pair on the GPU and
    returns (ok, matches[MATCH_DTYPE]); ok == "the list is not empty" as in the reference (:175)."""

    def __init__(self, regions: Regions, ctx: Context | None = None):
        self.ctx = ctx or default_context()
        self.regions_ = regions
        self._code = _dtype_code(regions.descriptors, regions.binary)
        self._id = None
        if regions.RegionCount() > 0:
            self._id = _new_view_id()
            self._upload(self._id, regions)

    def _upload(self, vid: int, r: Regions) -> None:
        n = r.RegionCount()
        _check(self.ctx.lib.b200m_upload_view(self.ctx._h, C.c_uint32(vid), r.descriptors.ctypes.data_as(C.c_void_p) if n else None, C.c_int(n),
                                              C.c_int(max(r.DescriptorLength(), 1)), C.c_int(self._code),
                                              r.positions.ctypes.data_as(C.c_void_p) if n else None), "b200m_upload_view")

    def getDatabaseRegions(self) -> Regions:
        return self.regions_

    def Match(self, f_dist_ratio: float, query_regions: Regions):
        empty = np.zeros(0, MATCH_DTYPE)
        if query_regions.RegionCount() == 0 or self._id is None:
            return False, empty
        if (query_regions.descriptors.dtype != self.regions_.descriptors.dtype or query_regions.binary != self.regions_.binary
                or query_regions.DescriptorLength() != self.regions_.DescriptorLength()):
            return False, empty
        lib = self.ctx.lib
        qid = _new_view_id()
        self._upload(qid, query_regions)
        try:
            pair = np.array([[self._id, qid]], np.uint32)
            res = C.c_void_p()
            _check(lib.b200m_match_pairs(self.ctx._h, pair.ctypes.data_as(C.c_void_p), C.c_int(1), C.c_float(f_dist_ratio), C.c_int(0),
                                         C.c_int(STAGE_FULL), C.byref(res)), "b200m_match_pairs")
            owner = _ResultOwner(lib, res)
            off, mat = C.c_void_p(), C.c_void_p()
            _check(lib.b200m_result_get(res, None, C.byref(off), C.byref(mat)), "b200m_result_get")
            n = int(_alias(off.value, 16, np.int64, owner)[1])
            matches = _alias(mat.value, n * MATCH_DTYPE.itemsize, MATCH_DTYPE, owner).copy()
        finally:
            lib.b200m_remove_view(self.ctx._h, C.c_uint32(qid))
        return len(matches) > 0, matches

    def close(self) -> None:
        if self._id is not None and self.ctx._h:
            self.ctx.lib.b200m_remove_view(self.ctx._h, C.c_uint32(self._id))
        self._id = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def createRegionsMatcher(regions: Regions, matcherType: EMatcherType, ctx: Context | None = None):
    """matching/RegionsMatcher.cpp:54-176 for the matcher types this engine implements.  Invalid requests return None
    like the reference's null unique_ptr (:61-64): scalar regions + Hamming matcher, binary regions + L2 matcher."""
    hamming = matcherType in (EMatcherType.BRUTE_FORCE_HAMMING, EMatcherType.BRUTE_FORCE_HAMMING_B200)
    if matcherType not in (EMatcherType.BRUTE_FORCE_L2, EMatcherType.BRUTE_FORCE_L2_B200) and not hamming:
        return None
    if regions.IsScalar() and hamming:
        return None
    if regions.IsBinary() and not hamming:
        return None
    if regions.descriptors.dtype not in (np.float32, np.uint8):
        return None
    return RegionsMatcherB200(regions, ctx)


class RegionsDatabaseMatcherB200:
    """matching::RegionsDatabaseMatcher (RegionsMatcher.hpp:183-220, RegionsMatcher.cpp:30-52)."""

    def __init__(self, matcherType: EMatcherType = EMatcherType.BRUTE_FORCE_L2_B200, database_regions: Regions | None = None, ctx: Context | None = None):
        self._matcherType = matcherType
        self._regionsMatcher = None if database_regions is None else createRegionsMatcher(database_regions, matcherType, ctx)

    def Match(self, distRatio: float, queryRegions: Regions):
        if queryRegions.RegionCount() == 0 or self._regionsMatcher is None:      # RegionsMatcher.cpp:32-38
            return False, np.zeros(0, MATCH_DTYPE)
        return self._regionsMatcher.Match(distRatio, queryRegions)

    def getDatabaseRegions(self) -> Regions:
        return self._regionsMatcher.getDatabaseRegions()


def DistanceRatioMatch(f_dist_ratio: float, eMatcherType: EMatcherType, regions_I: Regions, regions_J: Regions, ctx: Context | None = None) -> np.ndarray:
    """matching::DistanceRatioMatch (RegionsMatcher.cpp:19-28): regions_I = database, regions_J = query."""
    matcher = RegionsDatabaseMatcherB200(eMatcherType, regions_I, ctx)
    _, matches = matcher.Match(f_dist_ratio, regions_J)
    return matches
