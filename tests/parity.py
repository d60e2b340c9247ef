"""Shared parity helpers (tolerances from BASELINE.json north_star)."""
import numpy as np

COS_MIN = 0.999       # embedding cosine, bf16 tensor-core path vs fp32 oracle
SCORE_TOL = 1e-2      # absolute similarity-score tolerance


def cosine_rows(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    a = a / np.linalg.norm(a, axis=-1, keepdims=True)
    b = b / np.linalg.norm(b, axis=-1, keepdims=True)
    return (a * b).sum(-1)


def assert_topk_equivalent(ref_scores: np.ndarray, got_idx, got_scores, k: int, tol: float = SCORE_TOL):
    """north_star: 'top-k frame indices identical whenever the score gap exceeds the tolerance'.
    ref_scores: the oracle's full score vector.  For every rank r < k: the returned score must be within tol of
    the oracle's r-th best score, the returned index's oracle score must be within tol of it too, and if the
    oracle's r-th score is separated from BOTH neighbours by more than 2*tol the index must be identical."""
    order = np.argsort(ref_scores)[::-1]
    k = min(k, len(ref_scores))
    srt = ref_scores[order]
    for r in range(k):
        gi = int(got_idx[r])
        assert 0 <= gi < len(ref_scores), f"rank {r}: index {gi} out of range"
        assert abs(float(got_scores[r]) - float(srt[r])) <= tol, f"rank {r}: score {got_scores[r]} vs {srt[r]}"
        assert abs(float(ref_scores[gi]) - float(srt[r])) <= tol, f"rank {r}: picked a frame {gi} outside tolerance"
        gap_up = srt[r - 1] - srt[r] if r > 0 else np.inf
        gap_dn = srt[r] - srt[r + 1] if r + 1 < len(srt) else np.inf
        if gap_up > 2 * tol and gap_dn > 2 * tol:
            assert gi == int(order[r]), f"rank {r}: index {gi} != oracle {order[r]} despite a clear score gap"
    assert len(set(int(i) for i in got_idx[:k])) == k, "duplicate indices in top-k"
