"""Drop-in for the annotation helper of ``workoutdetector.datasets.repcount_dataset`` (reference:
workoutdetector/datasets/repcount_dataset.py:104-251). The torch Dataset classes there feed training and are out of
the hot path; RepcountHelper and the item records are the data model of eval_dataset / inference_dataset."""
import os
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import pandas as pd

ACTIONS = ['situp', 'push_up', 'pull_up', 'jump_jack', 'squat', 'front_raise']


def eval_count(preds: List[int], targets: List[int]) -> Tuple[float, float]:
    """Mean absolute error and the fraction with |error| == 1 (repcount_dataset.py:104-112)."""
    n = len(preds)
    diffs = [abs(p - t) for p, t in zip(preds, targets)]
    return sum(diffs) / n, sum(1.0 for d in diffs if d == 1) / n


@dataclass
class RepcountItem:
    """One RepCount video (repcount_dataset.py:115-138)."""
    video_path: str
    frames_path: str
    total_frames: int
    class_: str
    count: int
    reps: List[int]  # start_1, end_1, start_2, end_2, ...
    split: str
    video_name: str
    ytb_id: Optional[str] = None
    ytb_start_sec: Optional[int] = None
    ytb_end_sec: Optional[int] = None

    def __str__(self):
        return f'{self.video_name}\n{self.class_}\n{self.count}\n{self.reps}'

    def __getitem__(self, key):
        return self.__dict__[key]

    def __iter__(self):
        return iter(self.__dict__.items())


@dataclass
class RepcountItemWithPred(RepcountItem):
    """RepcountItem plus a prediction (repcount_dataset.py:141-149)."""
    pred_count: int = 0
    pred_reps: Optional[List[int]] = None
    mae: float = 0
    obo_acc: bool = False
    model_type: Optional[str] = None


class RepcountHelper:
    """Annotation lookup and count evaluation for RepCount (repcount_dataset.py:152-251).

    Args:
        data_root: e.g. 'data/RepCount' (videos/<split>/<name>, rawframes/<split>/<name without ext>)
        anno_file: annotation.csv with columns class_, split, name, vid, start, end, count, reps
    """

    def __init__(self, data_root: str, anno_file: str):
        self.anno_file = anno_file
        self.data_root = data_root
        self.classes = list(ACTIONS)

    def get_rep_data(self, split: List[str] = ['test'], action: List[str] = ['situp']) -> Dict[str, RepcountItem]:
        assert len(split) > 0, 'split must be specified, e.g. ["train", "val"]'
        assert len(action) > 0, 'action must be specified, e.g. ["pull_up", "squat"]'
        split = [s.lower() for s in split]
        action = [a.lower() for a in action]
        if 'all' in action:
            action = self.classes
        df = pd.read_csv(self.anno_file, index_col=0)
        df = df[df['split'].isin(split) & df['class_'].isin(action)].reset_index(drop=True)
        items: Dict[str, RepcountItem] = {}
        for row in df.itertuples(index=False):
            name = row.name
            frame_dir = os.path.join(self.data_root, 'rawframes', row.split, name.split('.')[0])
            total = len(os.listdir(frame_dir)) if os.path.isdir(frame_dir) else -1
            count = int(row.count)
            reps = [int(x) for x in row.reps.split()] if count > 0 else []
            items[name] = RepcountItem(os.path.join(self.data_root, 'videos', row.split, name), frame_dir, total,
                                       row.class_, count, reps, row.split, name, row.vid, row.start, row.end)
        return items

    def eval_count(self, pred_reps: Dict[str, int], split: List[str] = ['test'], action: List[str] = []
                   ) -> Tuple[float, float, Dict[str, RepcountItemWithPred]]:
        """MAE normalised by the ground-truth count (0 when GT is 0) and off-by-one accuracy (|diff| <= 1), both
        divided by the number of items of the split — as the reference does (repcount_dataset.py:232-251)."""
        items = self.get_rep_data(split=split, action=action)
        total_mae, total_obo = 0.0, 0.0
        out: Dict[str, RepcountItemWithPred] = {}
        for name, count in pred_reps.items():
            gt = items[name].count
            diff = abs(count - gt)
            mae = diff / gt if gt > 0 else 0
            obo = diff <= 1
            total_mae += mae
            total_obo += obo
            out[name] = RepcountItemWithPred(**items[name].__dict__, pred_count=count, pred_reps=[], mae=mae,
                                             obo_acc=obo)
        return total_mae / len(items), total_obo / len(items), out
