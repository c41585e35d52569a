"""Drop-in for ``workoutdetector.utils.inference_count`` (reference: workoutdetector/utils/inference_count.py).

Same entry points and return types; the per-window Python loop of the reference (one batch-1 model call per
8-frame window) becomes: upload the video once, gather + preprocess every window on the GPU, run the windows through
the engine in batches, threshold on the GPU, count repetitions with the batched counter kernel.

``model`` may be
  * a ``workoutdetector_b200.models.TSM`` (the B200 engine) — the fused path, or
  * any object with the onnxruntime.InferenceSession call surface (``get_inputs()`` / ``run()``), which is fed
    exactly as the reference feeds it (inference_count.py:265-276).
Where the reference at HEAD has dead code (the mmaction branch of inference_video uses an undefined name, :277-280;
count_by_video_model passes no transform, :326, and reads ``pred[0][0]`` of an unsorted list, :327) the drop-in does
what the surrounding code evidently intends (top-1 class) and says so in the docstring.
"""
import argparse
import json
import os
import os.path as osp
import time
from bisect import bisect_left
from typing import Callable, Dict, List, Optional, Sequence, Tuple, Union

import numpy as np
import torch
from torch import Tensor

from ..engine import count_reps, scores_to_states, vote_states
from ..settings import PROJ_ROOT, REPCOUNT_ANNO_PATH

COLORS = {
    'red': (0, 0, 255),
    'green': (0, 255, 0),
    'blue': (255, 0, 0),
    'yellow': (0, 255, 255),
    'white': (255, 255, 255),
    'black': (0, 0, 0),
    'orange': (12, 136, 237),
}

CLASSES = ['situp', 'push_up', 'pull_up', 'jump_jack', 'squat', 'front_raise']


# --------------------------------------------------------------------------------------------------
# counting
# --------------------------------------------------------------------------------------------------
def pred_to_count_batch(states: Tensor, lens: Optional[Tensor], step: int) -> Tuple[Tensor, Tensor, Tensor]:
    """Batched pred_to_count on the GPU: states int32 [V, W] (+ lens [V]) -> (counts [V], reps [V, W+1],
    reps_len [V]). Bit-exact with the reference state machine (inference_count.py:146-165)."""
    if not torch.cuda.is_available():
        raise RuntimeError("pred_to_count runs on the GPU counter kernel; no CUDA device is visible")
    states = states.to("cuda", torch.int32) if not states.is_cuda else states
    return count_reps(states, lens, step)


def pred_to_count(preds: Sequence[int], step: int) -> Tuple[int, List[int]]:
    """Convert a list of per-window predictions to a repetition count (reference inference_count.py:114-165).

    preds: one state per window, -1 = no action; states 2k / 2k+1 are the two halves of action k, and an
    even -> odd change inside one action counts one repetition.
    Returns (count, [start_1, end_1, start_2, end_2, ...]) with indices in frames (window index * step).

        >>> preds = [-1, -1, 6, 6, 6, 7, 6, 6, 6, 7, 6, 6, 7, 7, 6, 6, 7, 7, 6, 6, 7, 7, 6, 6, 7, 7, -1]
        >>> pred_to_count(preds, step=8)
        (6, [16, 40, 48, 72, 80, 96, 112, 128, 144, 160, 176, 192])
    """
    vals = [int(p) for p in preds]
    if len(vals) == 0:
        return 0, []
    states = torch.tensor([vals], dtype=torch.int32)
    counts, reps, reps_len = pred_to_count_batch(states, None, step)
    n = int(reps_len[0])
    count = int(counts[0])
    out = reps[0, :n].tolist()
    assert count * 2 == len(out)
    return count, out


# --------------------------------------------------------------------------------------------------
# window scoring on the engine
# --------------------------------------------------------------------------------------------------
def window_index_table(total_frames: int, stride: int = 8, span: int = 16, step: int = 2) -> Tensor:
    """int32 [W, span/step] frame indices of the sliding windows of inference_dataset (inference_count.py:411-414):
    window w covers frames 8w, 8w+2, ..., 8w+14; frames past the end are -1 (the reference pads with zero frames)."""
    starts = torch.arange(0, total_frames, stride, dtype=torch.int32).unsqueeze(1)
    offs = torch.arange(0, span, step, dtype=torch.int32).unsqueeze(0)
    idx = starts + offs
    return torch.where(idx < total_frames, idx, torch.full_like(idx, -1))


def queue_index_table(total_frames: int) -> Tensor:
    """int32 [W, 8]: the non-overlapping 8-frame queues of count_by_video_model (inference_count.py:313-329); a
    trailing partial queue is dropped."""
    w = total_frames // 8
    return torch.arange(0, w * 8, dtype=torch.int32).view(w, 8)


def _is_engine_model(model) -> bool:
    return hasattr(model, "engine") and hasattr(model, "num_segments")


def _is_ort_like(model) -> bool:
    return hasattr(model, "get_inputs") and hasattr(model, "run")


def score_windows(model, frames_u8: Tensor, index_table: Tensor, in_scale: float = 1.0 / 255.0,
                  threshold: float = 0.5, softmax: bool = True, batch: int = 64
                  ) -> Tuple[Tensor, Tensor, Tensor]:
    """Every window of one video through the engine.

    frames_u8: uint8 [F,H,W,3] (host or device); index_table int32 [W,8].
    Returns (logits [W,C] f32, probs [W,C] f32, states [W] i32) on the GPU.
    """
    if not _is_engine_model(model):
        raise TypeError("score_windows needs a workoutdetector_b200 TSM model")
    W = index_table.shape[0]
    eng = model.engine(min(batch, max(W, 1)))
    dev = eng.device
    frames = frames_u8.to(dev, non_blocking=True)
    table = index_table.to(dev)
    logits, probs, states = [], [], []
    for w0 in range(0, W, eng.max_clips):
        idx = table[w0:w0 + eng.max_clips].reshape(-1)
        x = eng.preprocess_u8(frames, idx, in_scale=in_scale)
        lg, pb, st = eng.forward(x, threshold=threshold, softmax=softmax)
        logits.append(lg)
        probs.append(pb)
        states.append(st)
    if not logits:
        c = model.num_class
        return (torch.empty(0, c, device=dev), torch.empty(0, c, device=dev),
                torch.empty(0, dtype=torch.int32, device=dev))
    return torch.cat(logits), torch.cat(probs), torch.cat(states)


def _as_u8_hwc(inputs: Union[Tensor, np.ndarray]) -> Tuple[Tensor, float]:
    """Clip as uint8 HWC + the scale the reference transform would effectively apply.
    uint8 -> ConvertImageDtype divides by 255; float input passes through unscaled (datasets/build.py:132), which is
    what happens to every clip in inference_dataset because of the float32 zero padding (inference_count.py:413)."""
    x = torch.from_numpy(inputs) if isinstance(inputs, np.ndarray) else inputs
    if x.dim() != 4 or x.shape[-1] != 3:
        raise ValueError(f"expected a clip [T,H,W,3], got {tuple(x.shape)}")
    if x.dtype == torch.uint8:
        return x.contiguous(), 1.0 / 255.0
    xf = x.to(torch.float32)
    xr = xf.round()
    if not bool(((xf == xr) & (xf >= 0) & (xf <= 255)).all()):
        raise NotImplementedError("float clips must hold integral values in 0..255 (frames promoted from uint8)")
    return xr.to(torch.uint8).contiguous(), 1.0


def inference_video(model, inputs: Union[Tensor, np.ndarray], threshold: float = 0.5,
                    transform: Callable = None) -> List[Tuple[int, float]]:
    """Score one 8-frame clip (reference inference_count.py:246-282).

    inputs: clip [8,H,W,3] (Tensor, as inference_dataset passes) or np.ndarray.
    Returns [(class_id, score)] in class order — raw consensus logits, like ``list(enumerate(score))`` at :276.
    With a B200 TSM model the resize/crop/normalize of ``build_test_transform`` runs fused on the GPU and
    ``transform`` is not needed; with an onnxruntime-style session the reference's call sequence is used verbatim.
    """
    if _is_engine_model(model):
        u8, scale = _as_u8_hwc(inputs)
        table = torch.arange(u8.shape[0], dtype=torch.int32).view(1, -1)
        logits, _, _ = score_windows(model, u8, table, in_scale=scale, threshold=threshold)
        return list(enumerate(logits[0].tolist()))
    if _is_ort_like(model):
        if type(inputs) is not Tensor:
            x = torch.from_numpy(inputs).float()
        else:
            x = inputs.permute(0, 3, 1, 2)
        assert transform is not None
        x = torch.unsqueeze(transform(x), 0)
        input_name = model.get_inputs()[0].name
        ort_outs = model.run(None, {input_name: x.cpu().numpy()})
        return list(enumerate(ort_outs[0][0].tolist()))
    raise TypeError(f"unsupported model type {type(model)}")


def read_video_frames(video_path: str, rgb: bool = True) -> Tensor:
    """Decode a whole video with OpenCV to uint8 [F,H,W,3] (RGB). torchvision.io.read_video, which the reference
    uses (inference_count.py:400), no longer exists in current torchvision."""
    import cv2
    cap = cv2.VideoCapture(video_path)
    if not cap.isOpened():
        raise IOError(f'Failed to open {video_path}')
    frames = []
    while True:
        ret, frame = cap.read()
        if not ret:
            break
        frames.append(cv2.cvtColor(frame, cv2.COLOR_BGR2RGB) if rgb else frame)
    cap.release()
    if not frames:
        return torch.empty((0, 0, 0, 3), dtype=torch.uint8)
    return torch.from_numpy(np.stack(frames))


def count_by_video_model(model, video_path: str, ground_truth: Optional[list] = None,
                         video_out_path: Optional[str] = None, threshold: float = 0.5) -> Tuple[int, List[int]]:
    """Count repetitions in a video with a video model (reference inference_count.py:285-339).

    Decodes with OpenCV, BGR->RGB (:322), scores non-overlapping 8-frame windows (:313-329), takes the top-1 state
    per window (softmax >= threshold, else -1 — the evident intent of ``pred[0][0]``, see module docstring) and runs
    ``pred_to_count(step=8)``. Returns (count, reps).
    """
    print(f'{video_path}')
    frames = read_video_frames(video_path)
    table = queue_index_table(frames.shape[0])
    if _is_engine_model(model):
        _, _, st = score_windows(model, frames, table, threshold=threshold)
        states = st.tolist()
    else:
        from ..datasets.build import build_test_transform
        transform = build_test_transform(person_crop=False)
        states = []
        for w in range(table.shape[0]):
            pred = inference_video(model, frames[table[w].long()], transform=transform)
            states.append(max(pred, key=lambda p: p[1])[0])
    count, reps = pred_to_count(preds=states, step=8)
    gt_count = len(ground_truth) // 2 if ground_truth else -1
    correct = (abs(gt_count - count) <= 1)
    print(f'count={count}, gt_count={gt_count}, correct={correct}')
    if video_out_path is not None:
        write_to_video(video_path, video_out_path, reps, states=states, step=8)
    return count, reps


# --------------------------------------------------------------------------------------------------
# image-model counting (per-frame classifier -> 7-frame majority vote -> counter)
# --------------------------------------------------------------------------------------------------
def _image_transform():
    """The reference's module-level ``data_transform`` (inference_count.py:27-34), built lazily."""
    import torchvision.transforms as T
    return T.Compose([T.ToPILImage(), T.Resize(256), T.CenterCrop(224), T.ToTensor(),
                      T.Normalize([0.485, 0.456, 0.406], [0.229, 0.224, 0.225])])


def inference_image(model, frame: np.ndarray) -> np.ndarray:
    """Scores of one frame from an image classifier (reference inference_count.py:168-189).

    model: an onnxruntime-style session (``get_inputs()`` / ``run()``), a torch module, or any callable that maps
    the transformed ``[1,3,224,224]`` tensor to scores ``[1,C]``.  frame: cv2 image (H, W, 3), fed as the reference
    feeds it (no BGR->RGB conversion there).  Returns float32 [C].
    """
    x = _image_transform()(frame).unsqueeze(0)
    if _is_ort_like(model):
        input_name = model.get_inputs()[0].name
        score = model.run(None, {input_name: x.numpy()})[0][0]
    else:
        if isinstance(model, torch.nn.Module):
            p = next(model.parameters(), None)
            if p is not None:
                x = x.to(p.device)
        with torch.no_grad():
            score = model(x)
        score = score.detach().cpu().numpy()[0] if isinstance(score, Tensor) else np.asarray(score)[0]
    return np.asarray(score).astype(np.float32)


def vote_and_count(frame_scores: Union[Tensor, np.ndarray], lens: Optional[Tensor] = None, window: int = 7,
                   votes: int = 4, step: int = 7) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """The tail of count_by_image_model (inference_count.py:211-235), batched over videos on the GPU.

    frame_scores: per-frame scores [V, F, C] (float) or per-frame arg-max labels [V, F] (integer).
    Returns (states [V,F] i32 — the 0/1 vote, -1 past lens —, counts [V], reps [V,F+1], reps_len [V]).
    Three kernels: first-max arg-max (scores_to_states_kernel), vote_states_kernel, count_reps_kernel.
    """
    if not torch.cuda.is_available():
        raise RuntimeError("vote_and_count runs on the GPU kernels; no CUDA device is visible")
    x = torch.from_numpy(frame_scores) if isinstance(frame_scores, np.ndarray) else frame_scores
    x = x.to("cuda")
    if x.dim() == 3:
        V, F, C = x.shape
        # numpy's argmax = first maximum; no softmax, no threshold (the reference's `threshold` argument is unused)
        _, labels = scores_to_states(x.reshape(V * F, C).float(), threshold=float("-inf"), softmax=False)
        labels = labels.view(V, F)
    else:
        labels = x.to(torch.int32)
    states = vote_states(labels, lens, window, votes)
    counts, reps, reps_len = count_reps(states, lens, step)
    return states, counts, reps, reps_len


def count_by_image_model(model, video_path: str, ground_truth: Optional[List[int]] = None,
                         video_out_path: Optional[str] = None, pred_out_path: Optional[str] = None,
                         threshold: float = 0.1) -> Tuple[int, List[int]]:
    """Count repetitions with a per-frame image classifier (reference inference_count.py:192-243).

    Every frame is scored by ``model`` (see inference_image; a model with a ``score_frames(frames_bgr_u8) -> [F,C]``
    method is called once for the whole video instead); the arg-max labels go through the 7-frame majority vote
    (state = sum of the last 7 labels >= 4) and ``pred_to_count(step=7)`` on the GPU.  ``threshold`` is accepted and,
    as in the reference, not used.
    """
    print(f'{video_path}')
    frames = read_video_frames(video_path, rgb=False)      # the reference feeds cv2's BGR frames as they are
    if hasattr(model, "score_frames"):
        sc = model.score_frames(frames)
        scores = np.asarray(sc.detach().cpu() if isinstance(sc, Tensor) else sc, dtype=np.float32)
    else:
        scores = np.stack([inference_image(model, f) for f in frames.numpy()]) if frames.shape[0] else \
            np.zeros((0, 1), np.float32)
    if scores.shape[0] == 0:
        states, count, reps = [], 0, []
    else:
        st, counts, reps_t, reps_len = vote_and_count(torch.from_numpy(scores).unsqueeze(0))
        states = st[0].tolist()
        count = int(counts[0])
        reps = reps_t[0, :int(reps_len[0])].tolist()
    gt_count = len(ground_truth) // 2 if ground_truth else -1
    correct = (abs(count - gt_count) <= 1)
    print(f'count={count} gt_count={gt_count} correct={correct}')
    if pred_out_path:
        save_scores_to_json(list(scores), pred_out_path, video_path, step=1)
    if video_out_path:
        write_to_video(video_path, video_out_path, reps, states, step=7)
    return count, reps


# --------------------------------------------------------------------------------------------------
# score files
# --------------------------------------------------------------------------------------------------
def save_scores_to_json(scores: List[np.ndarray], output_path: str, video_path: str, step: int) -> None:
    """Same file format as the reference (inference_count.py:47-67)."""
    if not output_path.endswith('.json'):
        output_path += '.json'
    assert not osp.exists(output_path), f'{output_path} already exists.'
    d = {'video_path': video_path, 'step': step}
    d['scores'] = {i: np.asarray(score).tolist() for i, score in enumerate(scores)}
    json.dump(d, open(output_path, 'w'))


def write_to_video(video_path: str, output_path: str, reps: List[int], states: List[int], step: int = 8) -> None:
    """Overlay state and running count on the video (reference inference_count.py:70-111)."""
    import cv2
    cap = cv2.VideoCapture(video_path)
    if not cap.isOpened():
        raise IOError(f'Failed to open {video_path}')
    fps = cap.get(cv2.CAP_PROP_FPS)
    width = int(cap.get(cv2.CAP_PROP_FRAME_WIDTH))
    height = int(cap.get(cv2.CAP_PROP_FRAME_HEIGHT))
    if output_path.endswith('.webm'):
        out = cv2.VideoWriter(output_path, cv2.VideoWriter_fourcc(*'vp80'), fps, (width, height))
    else:
        if not output_path.endswith('.mp4'):
            output_path += '.mp4'
        out = cv2.VideoWriter(output_path, cv2.VideoWriter_fourcc(*'mp4v'), fps, (width, height))
    for idx, res in enumerate(np.repeat(states, step)):
        ret, frame = cap.read()
        if not ret:
            break
        count_idx = bisect_left(reps[::2], idx)
        cv2.putText(frame, f'class {res}', (int(width * 0.2), int(height * 0.25)), cv2.FONT_HERSHEY_SIMPLEX, 1,
                    COLORS['red'], 2)
        cv2.putText(frame, f'count {count_idx}', (int(width * 0.25), int(height * 0.5)), cv2.FONT_HERSHEY_SIMPLEX,
                    1, COLORS['orange'], 2)
        out.write(frame)
    cap.release()
    out.release()


def score_video_to_dict(model, frames_u8: Tensor, item_meta: dict, checkpoint: str, quirk_float_promotion: bool = True
                        ) -> dict:
    """The per-video record inference_dataset writes (inference_count.py:401-420): raw logits keyed by the window's
    start frame. quirk_float_promotion=True reproduces the reference at HEAD, where the zero padding promotes every
    clip to float32 and the 1/255 scaling is skipped (:413-414); False applies the intended uint8 semantics."""
    table = window_index_table(frames_u8.shape[0])
    logits, _, _ = score_windows(model, frames_u8, table, in_scale=1.0 if quirk_float_promotion else 1.0 / 255.0)
    lg = logits.cpu()
    res = dict(item_meta)
    res.update(model='video_model', input_shape=[1, 8, 3, 224, 224], checkpoint=checkpoint,
               total_frames=int(frames_u8.shape[0]))
    res['scores'] = {int(8 * w): {c: float(v) for c, v in enumerate(lg[w].tolist())} for w in range(lg.shape[0])}
    return res


def inference_dataset(model, splits: List[str], out_dir: str, checkpoint: str, person_crop: bool = False,
                      data_root: Optional[str] = None, quirk_float_promotion: bool = True) -> None:
    """Score every RepCount video of ``splits`` and save ``{video_name}.score.json`` (reference
    inference_count.py:342-421; same JSON keys). Only person_crop=False runs on the engine."""
    from ..datasets import RepcountHelper
    if person_crop:
        raise NotImplementedError("person_crop=True needs the Faster-RCNN detector, which is outside the B200 path")
    os.makedirs(out_dir, exist_ok=True)
    data_root = data_root or osp.expanduser('~/data/RepCount/')
    helper = RepcountHelper(data_root, osp.join(data_root, 'annotation.csv'))
    data = helper.get_rep_data(splits, action=['all'])
    for item in data.values():
        vid = read_video_frames(item.video_path)
        meta = dict(video_name=item.video_name, ground_truth=item.reps, action=item.class_)
        res = score_video_to_dict(model, vid, meta, checkpoint, quirk_float_promotion)
        out_path = os.path.join(out_dir, f'{item.video_name}.score.json')
        json.dump(res, open(out_path, 'w'))
        print(f'{item.video_name} result saved to {out_path}')


def eval_dataset(model, action: List[str], split: str, model_type: str = 'video', output_dir: Optional[str] = None,
                 csv_name: str = None, save_video: bool = False, threshold: float = 0.7,
                 data_root: Optional[str] = None, anno_path: Optional[str] = None) -> None:
    """Count every video of a split and report MAE / OBO (reference inference_count.py:424-512)."""
    import pandas as pd
    from ..datasets import RepcountHelper
    data_root = data_root or os.path.join(PROJ_ROOT, 'data/RepCount/')
    helper = RepcountHelper(data_root, anno_path or REPCOUNT_ANNO_PATH)
    repcount_items = helper.get_rep_data(split=[split], action=action)
    pred_dict = dict()
    for name, item in repcount_items.items():
        assert os.path.exists(item['video_path']), f'{item["video_path"]} not exists'
        if save_video and output_dir is not None:
            assert os.path.isdir(output_dir)
            assert name.endswith('.mp4')
            output_path = os.path.join(output_dir, name)
        else:
            output_path = None
        if model_type == 'video':
            count, reps = count_by_video_model(model, item.video_path, ground_truth=item.reps,
                                               video_out_path=output_path)
        elif model_type == 'image':
            count, reps = count_by_image_model(model, item.video_path, ground_truth=item.reps,
                                               video_out_path=output_path, pred_out_path=None, threshold=threshold)
        else:
            raise ValueError(f'Invalid model type: {model_type}')
        pred_dict[name] = count
    mae, obo_acc, eval_res = helper.eval_count(pred_dict, action=action, split=[split])
    print(f'MAE={mae}, OBO_ACC={obo_acc}, SPLIT={split}, ACTION={action}')
    if output_dir is not None:
        res = []
        for item in eval_res.values():
            dict_ = dict(item.__dict__)
            dict_.pop('video_path')
            dict_.pop('frames_path')
            res.append(dict_)
        df = pd.DataFrame.from_dict(res)
        if csv_name is None:
            csv_name = f'eval_count_{model_type}_model.csv'
        if os.path.isfile(csv_name):
            csv_name = csv_name.split('.')[0] + '_' + str(time.time()) + '.csv'
        df.to_csv(os.path.join(output_dir, csv_name))
        print(f'Saved to {os.path.join(output_dir, csv_name)}')


def parse_args(argv=None) -> argparse.Namespace:
    """Same flags as the reference CLI (inference_count.py:560-595); -ckpt takes a TSM state_dict checkpoint."""
    parser = argparse.ArgumentParser(description='Evaluate RepCount')
    parser.add_argument('-ckpt', '--checkpoint', help='TSM checkpoint (.pth / .ckpt with state_dict)', required=True)
    parser.add_argument('--num-class', type=int, default=12)
    parser.add_argument('-i', '--video', help='video path', required=False)
    parser.add_argument('--eval', help='evaluate dataset', action='store_true')
    parser.add_argument('-t', '--threshold', help='threshold', type=float, default=0.5)
    parser.add_argument('-o', '--output', help='video output path. If evaluate dataset, it is output_dir',
                        required=False)
    parser.add_argument('-m', '--model-type', default='video', choices=['image', 'video'])
    parser.add_argument('-a', '--action', default='situp', choices=CLASSES + ['all'])
    parser.add_argument('-s', '--split', default='test', choices=['test', 'train', 'val'])
    return parser.parse_args(argv)


def main(args) -> None:
    from ..models import create_model
    model = create_model(num_class=args.num_class, num_segments=8, base_model='resnet50',
                         checkpoint=args.checkpoint, device='cuda')
    if not args.eval and args.video is not None:
        if args.model_type == 'image':
            raise SystemExit("the CLI builds a TSM video model; count_by_image_model takes a per-frame classifier "
                             "(call it from Python)")
        count_by_video_model(model, args.video, ground_truth=[], video_out_path=args.output,
                             threshold=args.threshold)
    elif args.eval:
        action = CLASSES if args.action == 'all' else [args.action]
        csv_name = args.checkpoint.split('.')[0].split('/')[-1] + '.csv'
        eval_dataset(model, action=action, split=args.split, model_type=args.model_type, output_dir=args.output,
                     csv_name=csv_name)


if __name__ == '__main__':
    main(parse_args())
