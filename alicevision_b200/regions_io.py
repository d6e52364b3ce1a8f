"""File formats on either side of the matching path, bound from include/b200io.h (host C++ in csrc/io.cpp).

Same function names and argument meaning as the reference:

* ``loadFeatsFromFile`` / ``saveFeatsToFile``          feature/PointFeature.hpp:88-122  (.feat, text)
* ``loadDescsFromBinFile`` / ``saveDescsToBinFile``    feature/Descriptor.hpp:244-307   (.desc, binary)
* ``Save`` / ``LoadMatchFile``                         matching/io.cpp:27-78,281-372    (matches.txt)

Errors the reference reports by throwing std::runtime_error are raised as ``IOError`` here.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, Tuple

import numpy as np

from .matching import F32, U8, MATCH_DTYPE, load_library

IO_SYMBOLS = ["b200io_last_error", "b200io_desc_count", "b200io_load_desc", "b200io_save_desc", "b200io_load_feat", "b200io_save_feat",
              "b200io_save_matches_txt", "b200io_load_matches_txt", "b200io_matches_num_blocks", "b200io_matches_block",
              "b200io_matches_data", "b200io_matches_free"]


def _lib():
    lib = load_library()
    lib.b200io_last_error.restype = C.c_char_p
    lib.b200io_matches_num_blocks.restype = C.c_int64
    lib.b200io_matches_data.restype = C.c_void_p
    lib.b200io_matches_free.restype = None
    return lib


def _ck(rc: int) -> None:
    if rc != 0:
        raise IOError(_lib().b200io_last_error().decode())


def _code(dtype) -> int:
    dtype = np.dtype(dtype)
    if dtype == np.float32:
        return F32
    if dtype == np.uint8:
        return U8
    raise TypeError(f"descriptor files hold float32 or uint8 elements, not {dtype}")


# ---- .feat ---------------------------------------------------------------------------------------------------------
def loadFeatsFromFile(sfileNameFeats: str) -> np.ndarray:
    """Returns features[n, 4] float32 = (x, y, scale, orientation)."""
    lib = _lib()
    n = C.c_int64()
    _ck(lib.b200io_load_feat(sfileNameFeats.encode(), None, C.c_int64(0), C.byref(n)))
    out = np.zeros((n.value, 4), np.float32)
    _ck(lib.b200io_load_feat(sfileNameFeats.encode(), out.ctypes.data_as(C.c_void_p), C.c_int64(n.value), C.byref(n)))
    return out


def saveFeatsToFile(sfileNameFeats: str, vec_feat: np.ndarray) -> None:
    f = np.ascontiguousarray(vec_feat, np.float32).reshape(-1, 4)
    _ck(_lib().b200io_save_feat(sfileNameFeats.encode(), f.ctypes.data_as(C.c_void_p), C.c_int64(f.shape[0])))


# ---- .desc ---------------------------------------------------------------------------------------------------------
def loadDescsFromBinFile(sfileNameDescs: str, dim: int, dtype=np.uint8, file_dtype=None, Nmax: int = 0) -> np.ndarray:
    """loadDescsFromBinFile<DescriptorT, FileDescriptorT>: ``dtype`` is the in-memory element type, ``file_dtype`` the one
    stored in the file (default: the same); Nmax != 0 limits the number of descriptors loaded (Descriptor.hpp:263-266)."""
    lib = _lib()
    n = C.c_int64()
    _ck(lib.b200io_desc_count(sfileNameDescs.encode(), C.byref(n)))
    rows = min(n.value, Nmax) if Nmax else n.value
    out = np.zeros((rows, dim), dtype)
    got = C.c_int64()
    _ck(lib.b200io_load_desc(sfileNameDescs.encode(), C.c_int(dim), C.c_int(_code(file_dtype if file_dtype is not None else dtype)), C.c_int(_code(dtype)),
                             out.ctypes.data_as(C.c_void_p), C.c_int64(rows), C.byref(got)))
    return out[: got.value]


def saveDescsToBinFile(sfileNameDescs: str, vec_desc: np.ndarray) -> None:
    d = np.ascontiguousarray(vec_desc)
    if d.ndim != 2:
        raise ValueError("descriptors must be [n, L]")
    _ck(_lib().b200io_save_desc(sfileNameDescs.encode(), d.ctypes.data_as(C.c_void_p), C.c_int64(d.shape[0]), C.c_int(d.shape[1]), C.c_int(_code(d.dtype))))


# ---- matches.txt ---------------------------------------------------------------------------------------------------
PairwiseMatches = Dict[Tuple[int, int], Dict[str, np.ndarray]]      # {(I, J): {descTypeName: matches[MATCH_DTYPE]}}


def _save_txt(path: str, matches: PairwiseMatches, keys) -> None:
    lib = _lib()
    names = sorted({d for k in keys for d in matches[k]}, key=_desc_order)
    pair_ids = np.array([[k[0], k[1]] for k in keys], np.uint32).reshape(-1, 2)
    offs, datas = [], []
    for d in names:
        lens = [len(matches[k].get(d, ())) for k in keys]
        offs.append(np.concatenate([[0], np.cumsum(lens)]).astype(np.int64))
        parts = [np.ascontiguousarray(matches[k][d], MATCH_DTYPE) for k in keys if d in matches[k] and len(matches[k][d])]
        datas.append(np.concatenate(parts) if parts else np.zeros(0, MATCH_DTYPE))
    nd = len(names)
    c_names = (C.c_char_p * max(nd, 1))(*[n.encode() for n in names])
    c_offs = (C.c_void_p * max(nd, 1))(*[o.ctypes.data for o in offs])
    c_data = (C.c_void_p * max(nd, 1))(*[(m.ctypes.data if len(m) else None) for m in datas])
    _ck(lib.b200io_save_matches_txt(path.encode(), C.c_int64(len(keys)), pair_ids.ctypes.data_as(C.c_void_p), C.c_int(nd), c_names, c_offs, c_data))


# EImageDescriberType values (feature/imageDescriberCommon.hpp:19-52) and names (imageDescriberCommon.cpp:44-90):
# MatchesPerDescType is a std::map keyed by the enum, so descriptor types are written in this order
_DESC_ENUM = {"unknown": 0, "sift": 10, "sift_float": 11, "sift_upright": 12, "dspsift": 13, "akaze": 20, "akaze_liop": 21, "akaze_mldb": 22,
              "cctag3": 30, "cctag4": 31, "sift_ocv": 40, "akaze_ocv": 41, "tag16h5": 50}


def _desc_order(name: str):
    return (_DESC_ENUM.get(name, 1000), name)


def Save(matches: PairwiseMatches, folder: str, extension: str = "txt", matchFilePerImage: bool = False, prefix: str = "") -> bool:
    """matching::Save (io.cpp:361-372): ``<prefix>matches.<extension>`` in ``folder``, or one ``<I>.<prefix>matches.<extension>``
    per first view id when matchFilePerImage (saveOneFilePerImage, :325-349)."""
    if extension != "txt":
        raise RuntimeError("Unknown matching file format: ." + extension)
    filename = prefix + "matches." + extension
    keys = sorted(matches)
    if not matchFilePerImage:
        _save_txt(os.path.join(folder, filename), matches, keys)
        return True
    for first in sorted({k[0] for k in keys}):
        _save_txt(os.path.join(folder, f"{first}.{filename}"), matches, [k for k in keys if k[0] == first])
    return True


def LoadMatchFile(matches: PairwiseMatches, filepath: str) -> bool:
    """matching::LoadMatchFile (io.cpp:27-78): merges the file into ``matches``; False when the file does not exist or is
    not a .txt.  Loaded matches carry i and j only (ratio / distance are zero, IndMatch.hpp:69)."""
    if not os.path.exists(filepath) or os.path.splitext(filepath)[1] != ".txt":
        return False
    lib = _lib()
    h = C.c_void_p()
    _ck(lib.b200io_load_matches_txt(filepath.encode(), C.byref(h)))
    try:
        nb = lib.b200io_matches_num_blocks(h)
        base = lib.b200io_matches_data(h)
        total = 0
        blocks = []
        for b in range(nb):
            I, J, name, a, e = C.c_uint32(), C.c_uint32(), C.c_char_p(), C.c_int64(), C.c_int64()
            _ck(lib.b200io_matches_block(h, C.c_int64(b), C.byref(I), C.byref(J), C.byref(name), C.byref(a), C.byref(e)))
            blocks.append((I.value, J.value, name.value.decode(), a.value, e.value))
            total = max(total, e.value)
        data = np.frombuffer((C.c_uint8 * (total * MATCH_DTYPE.itemsize)).from_address(base), dtype=MATCH_DTYPE).copy() if total else np.zeros(0, MATCH_DTYPE)
        for I, J, name, a, e in blocks:
            matches.setdefault((I, J), {})[name] = data[a:e]
    finally:
        lib.b200io_matches_free(h)
    return True


# ---- the small filters main_featureMatching applies between matching and export (matching/io.cpp:82-130) --------------------
def filterMatchesByViews(matches: PairwiseMatches, viewsKeys) -> None:
    """Keep the pairs whose two views are both in viewsKeys (io.cpp:82-93); in place."""
    keys = set(int(v) for v in viewsKeys)
    for k in [k for k in matches if k[0] not in keys or k[1] not in keys]:
        del matches[k]


def filterTopMatches(allMatches: PairwiseMatches, maxNum: int, minNum: int) -> None:
    """io.cpp:95-114: lists shorter than minNum are emptied, lists longer than maxNum truncated (they are already ordered);
    the (then empty) entries stay in the map, as in the reference."""
    if maxNum <= 0 and minNum <= 0:
        return
    if maxNum > 0 and minNum > maxNum:
        raise RuntimeError("The minimum number of matches is higher than the maximum.")
    for perDesc in allMatches.values():
        for d in list(perDesc):
            m = perDesc[d]
            if minNum > 0 and len(m) < minNum:
                perDesc[d] = m[:0]
            elif maxNum > 0 and len(m) > maxNum:
                perDesc[d] = m[:maxNum]


def filterMatchesByDesc(allMatches: PairwiseMatches, descTypesFilter) -> None:
    """io.cpp:116-130: keep only the listed descriptor types; pairs left without any type disappear."""
    keep = set(descTypesFilter)
    for k in list(allMatches):
        kept = {d: m for d, m in allMatches[k].items() if d in keep}
        if kept:
            allMatches[k] = kept
        else:
            del allMatches[k]


# ---- regions of a set of views (sfm/pipeline/regionsIO.cpp:25-78,197-251) ------------------------------------------------
# imageDescriberType name (feature/imageDescriberCommon.cpp:46-90) -> (element type, descriptor length, binary), i.e. the Regions class its
# ImageDescriber allocates (feature/regionsFactory.hpp:15-33)
DESCRIBER_REGIONS = {
    "sift": (np.uint8, 128, False), "sift_upright": (np.uint8, 128, False), "dspsift": (np.uint8, 128, False), "sift_ocv": (np.uint8, 128, False),
    "sift_float": (np.float32, 128, False),
    "akaze": (np.float32, 64, False), "akaze_ocv": (np.float32, 64, False), "akaze_liop": (np.uint8, 144, False), "akaze_mldb": (np.uint8, 64, True),
    "cctag3": (np.uint8, 128, False), "cctag4": (np.uint8, 128, False), "tag16h5": (np.uint8, 150, False),
}


def loadRegions(folders, viewId: int, imageDescriberType: str):
    """sfm::loadRegions (regionsIO.cpp:25-78): ``<viewId>.<describerType>.feat`` / ``.desc`` from the LAST folder that holds both; raises like the
    reference when no folder does or a file is invalid.  Returns ``matching.Regions`` (descriptors, positions, binary flag) with the full
    features[n, 4] kept as ``.features``."""
    from .matching import Regions
    if imageDescriberType not in DESCRIBER_REGIONS:
        raise ValueError(f"unknown image describer type '{imageDescriberType}'")
    dtype, dim, binary = DESCRIBER_REGIONS[imageDescriberType]
    base = f"{int(viewId)}.{imageDescriberType}"
    feat = desc = None
    for folder in folders:
        f, d = os.path.join(folder, base + ".feat"), os.path.join(folder, base + ".desc")
        if os.path.exists(f) and os.path.exists(d):
            feat, desc = f, d
    if feat is None:
        raise IOError(f"Can't find view {int(viewId)} region files in folders {', '.join(folders)}")
    feats = loadFeatsFromFile(feat)
    descs = loadDescsFromBinFile(desc, dim, dtype)
    if len(feats) != len(descs):
        raise IOError(f"Invalid {imageDescriberType} regions files for the view {int(viewId)}: {len(feats)} features, {len(descs)} descriptors")
    r = Regions(descs, feats[:, :2], binary=binary)
    r.features = feats
    return r


def loadRegionsPerView(viewIds, folders, imageDescriberTypes, viewIdFilter=None, threads: int = 8):
    """sfm::loadRegionsPerView (regionsIO.cpp:197-251) without the SfMData container: ``viewIds`` are the views of the scene, ``folders`` the
    features folders (sfmData's first, then the user's, as the caller concatenates them).  Returns ``(ok, {viewId: {describerType: Regions}})``;
    ok is False as soon as one view cannot be loaded (the reference logs the error and returns false).  Files are read in parallel
    (the reference uses 3 OpenMP threads; the parsing is the C library's and releases the GIL)."""
    from concurrent.futures import ThreadPoolExecutor
    folders = list(dict.fromkeys(folders))                       # std::unique on the concatenated list
    wanted = [int(v) for v in viewIds if viewIdFilter is None or len(viewIdFilter) == 0 or int(v) in viewIdFilter]
    jobs = [(v, t) for v in wanted for t in imageDescriberTypes]
    out: Dict[int, Dict[str, object]] = {}
    ok = True

    def one(job):
        try:
            return job, loadRegions(folders, job[0], job[1]), None
        except Exception as e:      # noqa: BLE001
            return job, None, e

    with ThreadPoolExecutor(max_workers=max(1, threads)) as ex:
        for (v, t), r, err in ex.map(one, jobs):
            if err is not None:
                ok = False
                continue
            out.setdefault(v, {})[t] = r
    return ok, out
