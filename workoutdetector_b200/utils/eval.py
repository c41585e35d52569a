"""Drop-in for ``workoutdetector.utils.eval`` (reference: workoutdetector/utils/eval.py) — the ``eval_count`` entry
point: score JSONs -> softmax -> arg-max / threshold -> pred_to_count -> MAE / off-by-one.
``main`` runs the softmax / threshold of every window of every video in one launch (wd_scores_to_states) and the
counting of every video in a second one (wd_count_reps)."""
import json
import os
from typing import Dict, List, Optional, Tuple, Union

import numpy as np
import pandas as pd
import torch
import torch.nn.functional as F

from ..engine import scores_to_states
from .inference_count import pred_to_count_batch


def to_softmax(d: Dict[str, float]) -> Dict[str, float]:
    """Raw scores of one window -> softmax scores, same keys (reference utils/visualize.py:140-150)."""
    p = F.softmax(torch.tensor(list(d.values()), dtype=torch.float32), dim=0)
    return dict(zip(d.keys(), p.numpy()))


def obo_mae(preds: List[int], targets: List[int], ratio: bool = True
            ) -> Union[Tuple[float, int], Tuple[float, float]]:
    """MAE (un-normalised) and off-by-one, which here counts |diff| == 1 only (reference eval.py:11-24)."""
    diffs = [abs(p - t) for p, t in zip(preds, targets)]
    mae = float(sum(diffs)) / len(preds)
    obo = float(sum(1 for d in diffs if d == 1))
    return (mae, obo / len(preds)) if ratio else (mae, obo)


def states_from_scores(scores: Dict[str, Dict[str, float]], threshold: float, softmax: bool) -> List[int]:
    """First-max arg-max per window, -1 when the top score is below threshold (reference eval.py:153-164)."""
    pred = []
    for v in scores.values():
        if softmax:
            v = to_softmax(v)
        class_id, score = max(v.items(), key=lambda x: x[1])
        pred.append(int(class_id) if score >= threshold else -1)
    return pred


def main(json_dir: str, anno_path: str, out_csv: Optional[str], softmax: bool = False) -> Tuple[float, float]:
    """Evaluate every ``*.json`` score file in json_dir (reference eval.py:117-180); also returns (mae, obo)."""
    threshold, step = 0.5, 8
    files = sorted(f for f in os.listdir(json_dir) if f.endswith('.json'))
    anno = pd.read_csv(anno_path, index_col='name')
    per_video = []
    for f in files:
        with open(os.path.join(json_dir, f)) as fp:
            data = json.load(fp)
        rows = [list(v.values()) for v in data['scores'].values()]
        keys = [list(v.keys()) for v in data['scores'].values()]
        per_video.append((f.split('.')[0] + '.mp4', data, rows, keys))
    counts = [0] * len(per_video)
    reps_list: List[List[int]] = [[] for _ in per_video]
    lens = [len(r) for _, _, r, _ in per_video]
    if sum(lens) > 0:
        if not torch.cuda.is_available():
            raise RuntimeError("utils.eval.main scores and counts on the GPU; no CUDA device is visible")
        flat = torch.tensor([r for _, _, rows, _ in per_video for r in rows], dtype=torch.float32, device='cuda')
        _, st_flat = scores_to_states(flat, threshold, softmax)            # launch 1: all windows of all videos
        # state = position of the winning entry; map it to the class id stored as the JSON key
        st_flat = st_flat.cpu().tolist()
        st = torch.full((len(per_video), max(lens)), -1, dtype=torch.int32)
        o = 0
        for i, (_, _, rows, keys) in enumerate(per_video):
            for w in range(len(rows)):
                s_ = st_flat[o + w]
                st[i, w] = int(keys[w][s_]) if s_ >= 0 else -1
            o += len(rows)
        c, r, rl = pred_to_count_batch(st, torch.tensor(lens, dtype=torch.int32).cuda(), step)   # launch 2
        c, r, rl = c.cpu(), r.cpu(), rl.cpu()
        counts = c.tolist()
        reps_list = [r[i, :int(rl[i])].tolist() for i in range(len(per_video))]
    out, preds, gts = [], [], []
    for (name, data, _, _), cnt, rep in zip(per_video, counts, reps_list):
        row = anno.loc[name]
        gt_count = int(row['count'])
        preds.append(cnt)
        gts.append(gt_count)
        out.append([name, gt_count, cnt, row['reps'], rep, row['split'], data['action']])
    mae, obo = obo_mae(preds, gts) if preds else (0.0, 0.0)
    df = pd.DataFrame(out, columns=['name', 'gt_count', 'pred_count', 'gt_rep', 'pred_rep', 'split', 'action'])
    if out_csv:
        df.to_csv(out_csv)
        print(f'Done. csv file saved to {out_csv}')
    print(f'=====Mean absolute error: {mae:.4f}, OBO acc: {obo:.4f}=====')
    return mae, obo


def analyze_count(csv: str, out_csv: Optional[str]) -> pd.DataFrame:
    """Per-(split, action) MAE / OBO table from the csv ``main`` writes (reference eval.py:58-114). The reference
    uses DataFrame.append, removed in pandas 2; this builds the same rows with pd.concat."""
    df = pd.read_csv(csv, index_col='name')
    rows = []
    totals: Dict[str, dict] = {}
    for split in df.split.unique():
        agg = totals.setdefault(split, dict(mae=0, obo=0, total=0, avg_count=0.0))
        for action in df.action.unique():
            sub = df.loc[(df.action == action) & (df.split == split)]
            if len(sub) == 0:
                continue
            gt, pred = sub.gt_count.values, sub.pred_count.values
            mae, obo = obo_mae(list(pred), list(gt), ratio=False)
            rows.append([action, split, mae, obo, len(sub), float(np.mean(gt))])
            agg['mae'] += int(mae * len(sub))
            agg['obo'] += int(obo)
            agg['total'] += len(sub)
            agg['avg_count'] += gt.sum()
    df_out = pd.DataFrame(rows, columns=['action', 'split', 'mae', 'obo_acc', 'total', 'avg_count'])
    for split, agg in totals.items():
        print(f'{split}: {agg}')
        total = agg['total']
        row = pd.DataFrame({'action': 'all', 'split': split, 'mae': agg['mae'] / total, 'obo_acc': agg['obo'],
                            'total': total, 'avg_count': agg['avg_count'] / total}, index=[0])
        df_out = pd.concat([df_out, row], ignore_index=True)
    if out_csv:
        df_out.to_csv(out_csv)
    print(df_out)
    return df_out
