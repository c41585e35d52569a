import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import tsm_oracle as O
from workoutdetector_b200.models import create_model
model = create_model(num_class=12, device="cuda")
eng = model.engine(8)
for (H, W) in [(224, 224), (240, 200)]:
    g = torch.Generator().manual_seed(H * 7 + W)
    fr = torch.randint(0, 256, (3, H, W, 3), generator=g, dtype=torch.uint8)
    ref = O.preprocess_u8(fr)                      # [3,3,224,224] fp32
    refb = ref.bfloat16().float()
    for env in ("1", "0"):
        os.environ["WD_PRE_PAIR"] = env
        out = eng.preprocess_u8(fr.cuda()).float().cpu()
        got = eng.image_view(out).permute(0, 3, 1, 2)
        d = (got - refb).abs()
        ulp = refb.abs().clamp_min(1e-9).log2().floor().exp2() * 2.0 ** -7
        bad = d > ulp * 1.001
        print(H, W, "pair" if env == "1" else "rows", "mismatch frac", float((d > 0).float().mean()), "max", float(d.max()), ">1ulp:", int(bad.sum()))
        if bad.any():
            idx = bad.nonzero()[:8]
            for i in idx:
                f, c, y, x = i.tolist()
                print("   at", i.tolist(), "got", float(got[f, c, y, x]), "ref", float(ref[f, c, y, x]))
