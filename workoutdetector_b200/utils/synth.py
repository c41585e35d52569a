"""Seeded synthetic inputs of the shapes BASELINE.json names (there is no dataset or network in this environment).
Pure torch on the CPU, deterministic in ``seed``; used by bench.py, the tests and oracle/gen_golden.py.
Low-frequency structure + motion is deliberate: i.i.d. noise frames self-average in the global pool and make every
clip score the same, which would turn parity checks vacuous (SURVEY.md §7 "vacuous parity")."""
import math

import torch
import torch.nn.functional as F


def _field(g: torch.Generator, n: int, grid: int = 7) -> torch.Tensor:
    """n smooth RGB fields [n,3,224,224] in 0..1."""
    z = torch.rand(n, 3, grid, grid, generator=g)
    f = F.interpolate(z, size=(224, 224), mode="bicubic", align_corners=False)
    return f.clamp(0, 1)


def synth_clips_u8(n_clips: int, seed: int) -> torch.Tensor:
    """[n_clips*8, 224, 224, 3] uint8: per clip a smooth background, a moving bright blob and a per-clip tint."""
    g = torch.Generator().manual_seed(seed)
    ys, xs = torch.meshgrid(torch.arange(224.0), torch.arange(224.0), indexing="ij")
    out = []
    for _ in range(n_clips):
        bg = _field(g, 1)[0]
        tint = torch.rand(3, 1, 1, generator=g) * 0.8 + 0.2
        contrast = float(torch.rand(1, generator=g)) * 0.8 + 0.2
        cx, cy = (torch.rand(2, generator=g) * 160 + 32).tolist()
        vx, vy = ((torch.rand(2, generator=g) - 0.5) * 24).tolist()
        rad = float(torch.rand(1, generator=g)) * 30 + 15
        for t in range(8):
            blob = torch.exp(-(((xs - cx - vx * t) ** 2 + (ys - cy - vy * t) ** 2) / (2 * rad * rad)))
            img = (bg * contrast + (1 - contrast) * 0.5) * tint + blob.unsqueeze(0) * (1 - tint)
            out.append((img.clamp(0, 1) * 255).round().to(torch.uint8).permute(1, 2, 0))
    return torch.stack(out).contiguous()


def synth_video_u8(n_frames: int, seed: int, period: float = 48.0, H: int = 224, W: int = 224) -> torch.Tensor:
    """[n_frames, H, W, 3] uint8 RepCount-shaped video: two smooth "poses" blended with a raised cosine of the given
    period (one repetition per period) plus a little per-frame noise."""
    g = torch.Generator().manual_seed(seed)
    a, b = _field(g, 2)
    if (H, W) != (224, 224):
        a, b = F.interpolate(torch.stack([a, b]), size=(H, W), mode="bilinear", align_corners=False)
    out = []
    for f in range(n_frames):
        w = 0.5 * (1 - math.cos(2 * math.pi * f / period))
        img = a * (1 - w) + b * w + (torch.rand(3, H, W, generator=g) - 0.5) * 0.04
        out.append((img.clamp(0, 1) * 255).round().to(torch.uint8).permute(1, 2, 0))
    return torch.stack(out).contiguous()


def synth_frames_u8(n: int, H: int, W: int, seed: int) -> torch.Tensor:
    """[n, H, W, 3] uint8 frames of any geometry: a sinusoidal pattern (integer-exact, platform independent) plus seeded
    integer noise — the input of the preprocessing goldens for non-224 geometries."""
    g = torch.Generator().manual_seed(seed)
    yy = torch.arange(H, dtype=torch.int64).view(H, 1, 1)
    xx = torch.arange(W, dtype=torch.int64).view(1, W, 1)
    ch = torch.arange(3, dtype=torch.int64).view(1, 1, 3)
    base = ((yy * 7 + xx * 5 + ch * 40) % 200) + ((yy // 9 + xx // 7) % 2) * 30       # integer arithmetic only
    noise = torch.randint(-12, 13, (n, H, W, 3), generator=g)
    return (base.unsqueeze(0) + noise).clamp(0, 255).to(torch.uint8)
