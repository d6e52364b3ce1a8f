"""Pair-list construction and text IO for the matching step, host-side and tiny.

Restates, for plain integer view ids,
* ``exhaustivePairs``            matchingImageCollection/pairBuilder.cpp:22-46  (upper triangle in view-id order, optional
                                 --rangeStart/--rangeSize chunk of FIRST images as Meshroom uses it),
* ``loadPairs`` / ``savePairs``  matchingImageCollection/ImagePairListIO.cpp:17-69,71-95 (one line per first image
                                 "I J1 J2 ...", pairs normalised to I<J on load, self pairs rejected).
A pair set is a sorted list of (I, J) tuples, i.e. the iteration order of the reference's ``PairSet`` (std::set).
"""
from __future__ import annotations

from typing import Iterable, List, Optional, Sequence, Tuple

Pair = Tuple[int, int]


def exhaustivePairs(view_ids: Sequence[int], rangeStart: int = -1, rangeSize: int = 0) -> List[Pair]:
    ids = sorted(int(v) for v in view_ids)                       # sfmData::Views is an ordered map keyed by view id
    first = range(len(ids))
    if rangeStart != -1 and rangeSize != 0:
        if rangeStart >= len(ids):
            return []
        first = range(rangeStart, min(rangeStart + rangeSize, len(ids)))
    return sorted({(ids[a], ids[b]) for a in first for b in range(a + 1, len(ids))})


def loadPairs(text: str, rangeStart: int = -1, rangeSize: int = 0) -> Optional[List[Pair]]:
    """Returns the sorted pair set, or None where the reference returns false (a line with fewer than two ids, or a
    self pair)."""
    pairs = set()
    for nb_line, line in enumerate(text.splitlines()):           # std::getline: a trailing newline adds no empty line
        if rangeStart != -1 and rangeSize != 0:
            if nb_line < rangeStart:
                continue
            if nb_line >= rangeStart + rangeSize:
                break
        tok = line.strip().replace("\t", " ").split()            # boost::trim + split on tab/space with token_compress
        if len(tok) < 2:
            return None
        i = int(tok[0])
        for t in tok[1:]:
            j = int(t)
            if i == j:
                return None
            pairs.add((i, j) if i < j else (j, i))
    return sorted(pairs)


def savePairs(pairs: Iterable[Pair]) -> str:
    ps = sorted({(int(a), int(b)) for a, b in pairs})
    if not ps:
        return ""
    out, prev = [f"{ps[0][0]} {ps[0][1]}"], ps[0][0]
    for a, b in ps[1:]:
        if a == prev:
            out.append(f" {b}")
        else:
            out.append(f"\n{a} {b}")
            prev = a
    return "".join(out) + "\n"


def loadPairsFromFile(path: str, rangeStart: int = -1, rangeSize: int = 0) -> Optional[List[Pair]]:
    with open(path) as f:
        return loadPairs(f.read(), rangeStart, rangeSize)


def savePairsToFile(path: str, pairs: Iterable[Pair]) -> bool:
    with open(path, "w") as f:
        f.write(savePairs(pairs))
    return True


# ---- the other pair generators of aliceVision_imageMatching (imageMatching/ImageMatching.cpp:145-189), for plain ids ----------
def generateSequentialMatches(imagePathPerView: dict, nbMatches: int) -> List[Pair]:
    """ImageMatching.cpp:145-164: views sorted by image path, each matched with its next nbMatches neighbours; pairs as
    (min id, max id).  imagePathPerView: {viewId: imagePath}."""
    order = [v for _, v in sorted((p, int(v)) for v, p in imagePathPerView.items())]
    out = set()
    for i in range(len(order)):
        for n in range(i + 1, min(i + nbMatches + 1, len(order))):
            a, b = order[i], order[n]
            out.add((min(a, b), max(a, b)))
    return sorted(out)


def generateAllMatchesInOneMap(viewIds: Iterable[int]) -> List[Pair]:
    """ImageMatching.cpp:166-186: every pair (a, b) with b > a."""
    ids = sorted(set(int(v) for v in viewIds))
    return [(a, b) for k, a in enumerate(ids) for b in ids[k + 1:]]


def generateAllMatchesBetweenTwoMap(viewIdsA: Iterable[int], viewIdsB: Iterable[int]) -> List[Pair]:
    """ImageMatching.cpp:188-206: every (a, b) with a in A and b in B, as given (a may be larger than b, or equal)."""
    A = sorted(set(int(v) for v in viewIdsA)); B = sorted(set(int(v) for v in viewIdsB))
    return [(a, b) for a in A for b in B]
