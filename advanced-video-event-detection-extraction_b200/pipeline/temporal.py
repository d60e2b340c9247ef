"""Temporal post-processing of hit lists.

`apply_temporal_consistency` mirrors Phase3Advanced._apply_temporal_consistency
(/root/reference/src/pipeline/phase3_advanced.py:37-81): hits are visited in timestamp order; a hit whose segment
overlaps an already kept one by more than half of either segment is dropped unless it has strictly higher
confidence, in which case it replaces the kept one.  Segments are [start_time, end_time] when present, else
timestamp +- 2.5 s.  It runs on the handful of hits K4 returns (<= top_k dicts), so it is host code; the kernels
upstream of it are what this package accelerates.  The result keeps the reference's order (insertion order of the
survivors, i.e. ascending timestamp) -- phase3_advanced.py:31 re-sorts by confidence afterwards, as does
`merge_hits` below."""
from __future__ import annotations

from typing import Dict, List


def _segment(hit: Dict):
    return (hit.get("start_time", hit["timestamp"] - 2.5), hit.get("end_time", hit["timestamp"] + 2.5))


def apply_temporal_consistency(results: List[Dict]) -> List[Dict]:
    if len(results) <= 1:                                   # phase3_advanced.py:39-40
        return results
    kept: List[Dict] = []
    for cur in sorted(results, key=lambda x: x["timestamp"]):       # stable: equal timestamps keep list order
        add = True
        cs, ce = _segment(cur)
        # the reference iterates over the list it removes from (phase3_advanced.py:49-73): after a removal the next
        # element is skipped.  Reproduced with an explicit index so the quirk is visible.
        i = 0
        while i < len(kept):
            ex = kept[i]
            es, ee = _segment(ex)
            overlap = max(0, min(ce, ee) - max(cs, es))
            if overlap > 0.5 * (ce - cs) or overlap > 0.5 * (ee - es):
                if cur["confidence"] <= ex["confidence"]:
                    add = False
                    break
                kept.pop(i)          # list.remove(existing) inside `for existing in list`: the iterator index stays,
                i += 1               # so the element that slid into slot i is never examined
                continue
            i += 1
        if add:
            kept.append(cur)
    return kept


def merge_hits(results: List[Dict]) -> List[Dict]:
    """Temporal consistency followed by the confidence sort of phase3_advanced.py:30-31: the 'segment merge' that
    turns frame hits into non-overlapping clips."""
    out = apply_temporal_consistency(list(results))
    out.sort(key=lambda x: x["confidence"], reverse=True)
    return out
