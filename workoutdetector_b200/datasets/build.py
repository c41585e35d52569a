"""Drop-in for the test-time transform of ``workoutdetector.datasets.build`` (reference:
workoutdetector/datasets/build.py:66-68, 115-136). Training transforms and dataset builders are out of the path."""
import torch

from ..engine import Engine

MEAN_STD = dict(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225])
INPUT_SIZE = (224, 224)


class B200TestTransform:
    """ConvertImageDtype(float32) -> Resize(256) -> CenterCrop(224) -> Normalize as ONE GPU kernel
    (csrc/wd_aux_kernels.cuh: preprocess_u8_kernel). Callable like the torchvision Compose it replaces:
    ``[T,3,H,W]`` uint8 (or float holding integral 0..255 values, which — as in the reference — are NOT rescaled)
    -> ``[T,3,224,224]`` float32 on the GPU."""

    def __init__(self, device: int = 0):
        self.device = device
        self._eng = None

    def _engine(self) -> Engine:
        if self._eng is None:
            self._eng = Engine(num_class=1, max_clips=1, mode="fp32", device=self.device)
        return self._eng

    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        if x.dim() != 4 or x.shape[1] != 3:
            raise ValueError(f"expected [T,3,H,W], got {tuple(x.shape)}")
        eng = self._engine()
        if x.dtype == torch.uint8:
            scale = 1.0 / 255.0
        else:
            xr = x.round()
            if not bool(((x == xr) & (x >= 0) & (x <= 255)).all()):
                raise NotImplementedError("float frames must hold integral values in 0..255")
            x, scale = xr.to(torch.uint8), 1.0
        hwc = x.to(eng.device).permute(0, 2, 3, 1).contiguous()
        out = eng.preprocess_u8(hwc, None, in_scale=scale)          # [T,224,224,4] fp32
        return eng.image_view(out).permute(0, 3, 1, 2).contiguous()

    def __repr__(self):
        return "B200TestTransform(ConvertImageDtype(float32), Resize(256), CenterCrop(224), Normalize(ImageNet))"


def build_test_transform(person_crop: bool = False) -> B200TestTransform:
    """build.py:115-136. person_crop=True would need the Faster-RCNN person detector (datasets/transform.py:128-262),
    which is outside the hot path."""
    if person_crop:
        raise NotImplementedError("person_crop=True is outside the B200 hot path")
    return B200TestTransform()
