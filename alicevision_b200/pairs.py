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
