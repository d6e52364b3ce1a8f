"""Vocabulary-tree pair-list producer (what aliceVision_imageMatching does before featureMatching), bound from
include/b200voc.h.  Names follow the reference:

* ``VocabularyTree``  <->  voctree::VocabularyTree<Descriptor<float,128>>   (voctree/VocabularyTree.hpp:96-296)
* ``Database``        <->  voctree::Database                                 (voctree/Database.hpp:52-160)
* ``conditionVocTree``<->  imageMatching::conditionVocTree in mode a/a       (imageMatching/ImageMatching.cpp:239-357)

Quantisation and all-against-all scoring run on the GPU; there is no CPU path for them.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Sequence, Tuple

import numpy as np

from .matching import F32, U8, B200MatchError, _check, load_library

VOC_SYMBOLS = ["b200v_tree_create", "b200v_tree_load", "b200v_tree_save", "b200v_tree_destroy", "b200v_tree_levels", "b200v_tree_splits",
               "b200v_tree_words", "b200v_quantize", "b200v_db_create", "b200v_db_destroy", "b200v_db_insert_descriptors", "b200v_db_insert_words",
               "b200v_db_size", "b200v_db_document", "b200v_db_compute_tfidf", "b200v_db_query_all", "b200v_db_last_scores", "b200v_db_last_gpu_ms",
               "b200v_convert_matches_to_pairs"]


def _lib():
    lib = load_library()
    lib.b200v_tree_destroy.restype = None
    lib.b200v_db_destroy.restype = None
    lib.b200v_db_size.restype = C.c_int64
    lib.b200v_db_last_gpu_ms.restype = C.c_double
    for n in ("b200v_tree_levels", "b200v_tree_splits", "b200v_tree_words"):
        getattr(lib, n).restype = C.c_uint32
    return lib


def _code(a: np.ndarray) -> int:
    if a.dtype == np.float32:
        return F32
    if a.dtype == np.uint8:
        return U8
    raise TypeError(f"descriptors must be float32 or uint8, not {a.dtype}")


class VocabularyTree:
    """Tree of float centers in the reference's node order; ``load`` / ``save`` use the reference's file layout."""

    def __init__(self, k: int = 0, levels: int = 0, centers: np.ndarray | None = None, valid: np.ndarray | None = None, file: str | None = None, dim: int = 128):
        self.lib = _lib()
        self._h = C.c_void_p()
        self.dim = dim
        if file is not None:
            _check(self.lib.b200v_tree_load(file.encode(), C.c_int(dim), C.byref(self._h)), "b200v_tree_load")
        elif centers is not None:
            c = np.ascontiguousarray(centers, np.float32)
            v = np.ones(c.shape[0], np.uint8) if valid is None else np.ascontiguousarray(valid, np.uint8)
            self.dim = c.shape[1]
            _check(self.lib.b200v_tree_create(C.c_uint32(k), C.c_uint32(levels), C.c_int(c.shape[1]), c.ctypes.data_as(C.c_void_p), v.ctypes.data_as(C.c_void_p),
                                              C.c_uint32(c.shape[0]), C.byref(self._h)), "b200v_tree_create")

    def save(self, file: str) -> None:
        _check(self.lib.b200v_tree_save(self._h, file.encode()), "b200v_tree_save")

    def levels(self) -> int: return self.lib.b200v_tree_levels(self._h)
    def splits(self) -> int: return self.lib.b200v_tree_splits(self._h)
    def words(self) -> int: return self.lib.b200v_tree_words(self._h)

    def quantize(self, features: np.ndarray, device: int = 0) -> np.ndarray:
        """VocabularyTree::quantize(std::vector<DescriptorT>): one visual word per descriptor."""
        f = np.ascontiguousarray(features)
        words = np.zeros(f.shape[0], np.int32)
        _check(self.lib.b200v_quantize(C.c_int(device), self._h, f.ctypes.data_as(C.c_void_p), C.c_int64(f.shape[0]), C.c_int(_code(f)),
                                       words.ctypes.data_as(C.c_void_p)), "b200v_quantize")
        return words

    def __del__(self):
        try:
            if self._h:
                self.lib.b200v_tree_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass


class Database:
    """voctree::Database: insert documents, computeTfIdfWeights, find for every document at once."""

    def __init__(self, tree: VocabularyTree, device: int = 0):
        self.lib = _lib()
        self.tree = tree
        self._h = C.c_void_p()
        _check(self.lib.b200v_db_create(tree._h, C.c_int(device), C.byref(self._h)), "b200v_db_create")

    def insert(self, doc_id: int, descriptors: np.ndarray, Nmax: int = 0) -> None:
        """populateDatabase step for one view: quantizeToSparse + insert."""
        d = np.ascontiguousarray(descriptors)
        _check(self.lib.b200v_db_insert_descriptors(self._h, C.c_uint32(doc_id), d.ctypes.data_as(C.c_void_p), C.c_int64(d.shape[0]), C.c_int(_code(d)),
                                                    C.c_int64(Nmax)), "b200v_db_insert_descriptors")

    def insert_words(self, doc_id: int, words: np.ndarray) -> None:
        w = np.ascontiguousarray(words, np.int32)
        _check(self.lib.b200v_db_insert_words(self._h, C.c_uint32(doc_id), w.ctypes.data_as(C.c_void_p), C.c_int64(w.size)), "b200v_db_insert_words")

    def size(self) -> int: return self.lib.b200v_db_size(self._h)

    def document(self, doc_id: int) -> np.ndarray:
        p, n = C.c_void_p(), C.c_int64()
        _check(self.lib.b200v_db_document(self._h, C.c_uint32(doc_id), C.byref(p), C.byref(n)), "b200v_db_document")
        return np.frombuffer((C.c_int32 * n.value).from_address(p.value), np.int32).copy() if n.value else np.zeros(0, np.int32)

    def computeTfIdfWeights(self, default_weight: float = 1.0) -> np.ndarray:
        w = np.zeros(self.tree.words(), np.float32)
        _check(self.lib.b200v_db_compute_tfidf(self._h, C.c_float(default_weight), w.ctypes.data_as(C.c_void_p)), "b200v_db_compute_tfidf")
        return w

    def find_all(self, N: int = 0, distanceMethod: str = "strongCommonPoints"):
        """Database::find for every inserted document. Returns (query_ids[n], match_ids[n, keep], scores[n, keep])."""
        n = self.size()
        keep = n if N == 0 else min(N, n)
        q = np.zeros(n, np.uint32); ids = np.zeros((n, keep), np.uint32); sc = np.zeros((n, keep), np.float32)
        nk = C.c_size_t()
        _check(self.lib.b200v_db_query_all(self._h, C.c_size_t(N), distanceMethod.encode(), q.ctypes.data_as(C.c_void_p), ids.ctypes.data_as(C.c_void_p),
                                           sc.ctypes.data_as(C.c_void_p), C.byref(nk)), "b200v_db_query_all")
        assert nk.value == keep
        return q, ids, sc

    def last_scores(self) -> np.ndarray:
        p, n = C.c_void_p(), C.c_int64()
        _check(self.lib.b200v_db_last_scores(self._h, C.byref(p), C.byref(n)), "b200v_db_last_scores")
        return np.frombuffer((C.c_int32 * (n.value * n.value)).from_address(p.value), np.int32).reshape(n.value, n.value).copy() if n.value else np.zeros((0, 0), np.int32)

    def last_gpu_ms(self) -> float: return self.lib.b200v_db_last_gpu_ms(self._h)

    def __del__(self):
        try:
            if self._h:
                self.lib.b200v_db_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass


def convertAllMatchesToPairList(query_ids: np.ndarray, match_ids: np.ndarray, numMatches: int) -> np.ndarray:
    """imageMatching::convertAllMatchesToPairList, flattened to (I, J) rows in OrderedPairList order."""
    lib = _lib()
    q = np.ascontiguousarray(query_ids, np.uint32)
    m = np.ascontiguousarray(match_ids, np.uint32)
    m = m.reshape(len(q), m.size // len(q)) if len(q) else m.reshape(0, 0)
    n = C.c_int64()
    _check(lib.b200v_convert_matches_to_pairs(q.ctypes.data_as(C.c_void_p), m.ctypes.data_as(C.c_void_p), C.c_size_t(len(q)), C.c_size_t(m.shape[1]),
                                              C.c_size_t(numMatches), None, C.c_int64(0), C.byref(n)), "b200v_convert_matches_to_pairs")
    out = np.zeros((n.value, 2), np.uint32)
    _check(lib.b200v_convert_matches_to_pairs(q.ctypes.data_as(C.c_void_p), m.ctypes.data_as(C.c_void_p), C.c_size_t(len(q)), C.c_size_t(m.shape[1]),
                                              C.c_size_t(numMatches), out.ctypes.data_as(C.c_void_p), C.c_int64(n.value), C.byref(n)), "b200v_convert_matches_to_pairs")
    return out


def conditionVocTree(tree: VocabularyTree, descriptorsPerView: Dict[int, np.ndarray], nbMaxDescriptors: int = 0, numImageQuery: int = 0,
                     distanceMethod: str = "strongCommonPoints", device: int = 0):
    """imageMatching::conditionVocTree (mode a/a, no weights file): populate the database, TF-IDF weights, query every
    document, convert to the pair list.  Returns (pairs[n, 2], database)."""
    db = Database(tree, device)
    for vid in sorted(descriptorsPerView):
        db.insert(vid, descriptorsPerView[vid], nbMaxDescriptors)
    db.computeTfIdfWeights()
    q, ids, _ = db.find_all(numImageQuery, distanceMethod)
    return convertAllMatchesToPairList(q, ids, numImageQuery), db
