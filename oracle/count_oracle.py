"""ORACLE — test infrastructure, not product code.

Pure-Python restatement of ``pred_to_count`` (workoutdetector/utils/inference_count.py:114-165) and of the two count
metrics (utils/eval.py:11-24 ``obo_mae``; datasets/repcount_dataset.py:212-251 ``RepcountHelper.eval_count``).
Pinned by the reference's own known-answer vectors (tests/test_inference_count.py:8-48, the doctest at
inference_count.py:140-143) in tests/test_oracle_golden.py, and by outputs of the reference function itself on
random state sequences (tests/golden/count_vectors.json, written by oracle/gen_golden.py).  ``vote_states`` restates
the 7-frame vote of count_by_image_model (inference_count.py:213-224); pinned by tests/golden/image_vote.json, which
oracle/gen_golden.py writes by running the reference's own count_by_image_model loop on seeded per-frame scores.
"""
from typing import List, Sequence, Tuple


def pred_to_count(preds: Sequence[int], step: int) -> Tuple[int, List[int]]:
    """A repetition is an even -> odd transition inside one action (state 2k -> 2k+1); -1 is skipped.
    Returns (count, [start_1, end_1, ...]) in frames (window index * step)."""
    count = 0
    reps: List[int] = []
    have_last = False
    last = 0
    start_idx = 0  # inference_count.py:149 prev_state_start_idx
    for idx, pred in enumerate(preds):
        if pred == -1:                                   # :151-152
            continue
        if have_last and last != pred:                   # :154
            if pred % 2 == 1 and last == pred - 1:       # :155
                count += 1
                reps.append(start_idx * step)            # :157
                reps.append(idx * step)                  # :158
        last, have_last = pred, True                     # :159
        if pred != preds[start_idx]:                     # :160-162 (preds[start_idx] may be -1 at the start)
            start_idx = idx
    return count, reps


def vote_states(labels: Sequence[int], window: int = 7, votes: int = 4) -> List[int]:
    """The majority vote of count_by_image_model (utils/inference_count.py:213-224): a deque(maxlen=window) of the
    per-frame arg-max labels; state = sum(deque) >= votes (a bool there; 0 / 1 here — pred_to_count treats them alike)."""
    out: List[int] = []
    for f in range(len(labels)):
        out.append(1 if sum(labels[max(0, f - window + 1):f + 1]) >= votes else 0)   # :222-224
    return out


def count_by_image_labels(labels: Sequence[int]) -> Tuple[int, List[int], List[int]]:
    """Labels -> vote (window 7, votes 4) -> pred_to_count(step=7) (utils/inference_count.py:213-235)."""
    st = vote_states(labels, 7, 4)
    c, r = pred_to_count(st, 7)
    return c, r, st


def obo_mae(preds: Sequence[int], targets: Sequence[int]) -> Tuple[float, float]:
    """utils/eval.py:11-24 with ratio=True: un-normalised MAE; 'off by one' counts |diff| == 1 exactly."""
    mae = sum(abs(p - t) for p, t in zip(preds, targets))
    obo = sum(1 for p, t in zip(preds, targets) if abs(p - t) == 1)
    return mae / len(preds), obo / len(preds)


def helper_eval_count(pred: dict, gt: dict, n_items: int) -> Tuple[float, float]:
    """datasets/repcount_dataset.py:232-251: MAE normalised by the GT count (0 when GT is 0), OBO = diff <= 1,
    both divided by the number of items in the split (not the number of predictions)."""
    total_mae = 0.0
    total_obo = 0.0
    for name, count in pred.items():
        g = gt[name]
        diff = abs(count - g)
        total_mae += diff / g if g > 0 else 0
        total_obo += diff <= 1
    return total_mae / n_items, total_obo / n_items
