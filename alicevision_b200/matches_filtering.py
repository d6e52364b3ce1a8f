"""The post-match filters `aliceVision_featureMatching` applies between matching and export (matching/matchesFiltering.cpp), on the match arrays
this package produces (MATCH_DTYPE: i, j, ratio, dist) and features[n, 4] = (x, y, scale, orientation) as loaded by regions_io.

Host-side index arithmetic, restated from the reference (its translation unit pulls in sfmData and cannot be compiled here, so these are
checked by hand-computed known answers and properties, not against compiled reference code):

* ``filterMatchesByMin2DMotion``   matchesFiltering.cpp:196-241  (main_featureMatching.cpp:352, --minRequired2DMotion)
* ``sortMatches_byDistanceRatio``  :45-53                        (:371, before homography growing)
* ``sortMatches_byFeaturesScale``  :12-43   ``thresholdMatches`` :60-66
* ``matchesGridFiltering``         :68-143  ``matchesGridFilteringForAllPairs`` :145-194 (:515, --useGridSort / --maxMatches)

Where the reference uses std::sort (unstable) the order of EQUAL keys is unspecified there; a stable sort is used here.
"""
from __future__ import annotations

import numpy as np


def filterMatchesByMin2DMotion(mapPutativesMatches: dict, featuresPerView: dict, minRequired2DMotion: float) -> None:
    """Drops the matches whose two features are closer than minRequired2DMotion * 2**max(scale_i, scale_j) pixels (float arithmetic as in the
    reference: Vec2f difference norm, float pow).  ``mapPutativesMatches``: {(I, J): {descType: matches}}; ``featuresPerView``: {view: {descType:
    features[n, 4]}}.  In place; negative threshold = disabled (:198-199)."""
    if minRequired2DMotion < 0.0:
        return
    for (vi, vj), per_desc in mapPutativesMatches.items():
        for desc, m in list(per_desc.items()):
            fi = np.asarray(featuresPerView[vi][desc], np.float32); fj = np.asarray(featuresPerView[vj][desc], np.float32)
            pi, pj = fi[m["i"], :2], fj[m["j"], :2]
            scale = np.maximum(fi[m["i"], 2], fj[m["j"], 2])
            coeff = np.power(np.float32(2.0), scale).astype(np.float32)                  # float coeff = pow(2, scale)
            d = pi - pj
            norm = np.sqrt((d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]).astype(np.float32)).astype(np.float32)
            keep = ~(norm.astype(np.float64) < np.float64(minRequired2DMotion) * coeff.astype(np.float64))   # float norm vs double product
            per_desc[desc] = m[keep]


def sortMatches_byDistanceRatio(matches: np.ndarray) -> np.ndarray:
    """Increasing Lowe ratio (:45-53)."""
    return matches[np.argsort(matches["ratio"], kind="stable")]


def sortMatches_byFeaturesScale(inputMatches: np.ndarray, featuresI: np.ndarray, featuresJ: np.ndarray) -> np.ndarray:
    """Decreasing mean scale of the two features, (scale1 + scale2) / 2.0 evaluated in double and stored as float (:28-30, matchCompare :55-58)."""
    s = ((np.asarray(featuresI, np.float32)[inputMatches["i"], 2].astype(np.float64) + np.asarray(featuresJ, np.float32)[inputMatches["j"], 2].astype(np.float64)) / 2.0).astype(np.float32)
    return inputMatches[np.argsort(-s, kind="stable")]


def thresholdMatches(matches: np.ndarray, uNumMatchesToKeep: int) -> np.ndarray:
    return matches[:uNumMatchesToKeep] if len(matches) > uNumMatchesToKeep else matches


def _divide_round_up(x: int, y: int) -> int:
    return (x + y - 1) // y


def matchesGridFiltering(lFeatures: np.ndarray, lImgSize, rFeatures: np.ndarray, rImgSize, matches: np.ndarray, gridSize: int = 3) -> np.ndarray:
    """Re-orders the matches so that they are spread over a gridSize x gridSize grid of both images (:68-143): every match goes to its left-image
    cell or its right-image cell, whichever is shorter at that moment (left on ties); the cells are then interleaved round-robin (left cells
    first, row-major).  The input order matters, as in the reference."""
    lw, lh = _divide_round_up(int(lImgSize[0]), gridSize), _divide_round_up(int(lImgSize[1]), gridSize)
    rw, rh = _divide_round_up(int(rImgSize[0]), gridSize), _divide_round_up(int(rImgSize[1]), gridSize)
    lf = np.asarray(lFeatures, np.float32); rf = np.asarray(rFeatures, np.float32)
    f32 = np.float32
    li = np.floor(lf[matches["i"], 0] / f32(lw)) + np.floor(lf[matches["i"], 1] / f32(lh)) * f32(gridSize)
    ri = np.floor(rf[matches["j"], 0] / f32(rw)) + np.floor(rf[matches["j"], 1] / f32(rh)) * f32(gridSize)
    # clamp(index, 0, gridSize - 1) - the reference clamps the COMBINED index to gridSize - 1, not to gridSize^2 - 1 (:98-99); restated as written
    li = np.clip(li, 0, gridSize - 1).astype(np.int64); ri = np.clip(ri, 0, gridSize - 1).astype(np.int64)
    n_cells = gridSize * gridSize
    cells = [[] for _ in range(2 * n_cells)]
    for k in range(len(matches)):
        cl, cr = cells[li[k]], cells[ri[k] + n_cells]
        (cl if len(cl) <= len(cr) else cr).append(k)
    order = []
    for c in range(max((len(x) for x in cells), default=0)):
        for cell in cells:
            if c < len(cell):
                order.append(cell[c])
    return matches[np.array(order, np.int64)] if order else matches[:0]


def matchesGridFilteringForAllPairs(geometricMatches: dict, imgSizePerView: dict, featuresPerView: dict, useGridSort: bool, numMatchesToKeep: int) -> dict:
    """:145-194: per pair and descriptor type: sort by feature scale, optionally grid-order, keep the first numMatchesToKeep (0 = all).
    ``imgSizePerView``: {view: (width, height)}.  Returns a new {(I, J): {descType: matches}}."""
    out = {}
    for (vi, vj), per_desc in geometricMatches.items():
        for desc, m in per_desc.items():
            fl, fr = featuresPerView[vi][desc], featuresPerView[vj][desc]
            o = sortMatches_byFeaturesScale(m, fl, fr)
            if useGridSort:
                o = matchesGridFiltering(fl, imgSizePerView[vi], fr, imgSizePerView[vj], o)
            if numMatchesToKeep > 0:
                o = o[: min(numMatchesToKeep, len(o))]
            out.setdefault((vi, vj), {})[desc] = o
    return out
