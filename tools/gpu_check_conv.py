"""GPU bring-up check (run on the B200 box): every conv shape class of TSM-R50 through wd_debug_conv versus a
torch fp32 convolution on bf16-rounded operands. Each case runs in its own process with a timeout so a hang or a
fault in one case does not take the others down. Usage: python tools/gpu_check_conv.py [case-index]"""
import subprocess
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

CASES = [
    # name, clips, H, Cin, Cout, k, stride, fold, relu, residual, mode, tile_n
    ("1x1 k64 single kblock", 2, 8, 64, 64, 1, 1, 0, 0, 0, "gather", 64),
    ("1x1 k256", 2, 8, 256, 64, 1, 1, 0, 1, 0, "gather", 64),
    ("1x1 k256 tma", 2, 8, 256, 64, 1, 1, 0, 1, 0, "tma", 64),
    ("1x1 k64 tma", 2, 8, 64, 64, 1, 1, 0, 0, 0, "tma", 64),
    ("1x1 fold8 gather", 2, 8, 64, 64, 1, 1, 8, 1, 0, "gather", 64),
    ("1x1 fold32 gather", 2, 8, 256, 128, 1, 1, 32, 1, 0, "gather", 128),
    ("1x1 fold64 gather", 1, 8, 512, 128, 1, 1, 64, 1, 0, "gather", 128),
    ("1x1 fold64 tma", 1, 8, 512, 128, 1, 1, 64, 1, 0, "tma", 128),
    ("1x1 fold128 tma n256", 1, 14, 1024, 256, 1, 1, 128, 1, 0, "tma", 256),
    ("3x3 s1 c64", 2, 8, 64, 64, 3, 1, 0, 1, 0, "gather", 64),
    ("3x3 s2 c128", 2, 8, 128, 128, 3, 2, 0, 1, 0, "gather", 128),
    ("3x3 s1 c256 n256", 1, 14, 256, 256, 3, 1, 0, 1, 0, "gather", 256),
    ("1x1 s2 downsample", 2, 8, 256, 512, 1, 2, 0, 0, 0, "gather", 256),
    ("1x1 residual relu tma 2 n-tiles", 1, 7, 512, 2048, 1, 1, 0, 1, 1, "tma", 256),
    ("1x1 residual relu gather tail", 1, 7, 128, 512, 1, 1, 0, 1, 1, "gather", 128),
    ("3x3 7x7 tail", 1, 7, 512, 512, 3, 1, 0, 1, 0, "gather", 256),
    ("1x1 many tiles tma res", 40, 14, 256, 1024, 1, 1, 0, 1, 1, "tma", 256),
    ("3x3 many tiles gather", 24, 14, 128, 128, 3, 1, 0, 1, 0, "gather", 128),
    ("1x1 k64 n256 many tiles res", 16, 28, 64, 256, 1, 1, 0, 1, 1, "tma", 256),
]


def run_case(i):
    import torch
    import torch.nn.functional as F
    from workoutdetector_b200.engine import debug_conv
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    name, clips, H, Cin, Cout, k, stride, fold, relu, res, mode, tile_n = CASES[i]
    g = torch.Generator(device="cpu").manual_seed(100 + i)
    x = torch.randn(clips, H, H, 8, Cin, generator=g).to(torch.bfloat16).cuda()
    w = (torch.randn(Cout, Cin, k, k, generator=g) / (Cin * k * k) ** 0.5).to(torch.bfloat16).float()
    b = torch.randn(Cout, generator=g)
    Ho = (H + 2 * (k // 2) - k) // stride + 1
    r = torch.randn(clips, Ho, Ho, 8, Cout, generator=g).to(torch.bfloat16).cuda() if res else None
    persistent = os.environ.get('WD_PERSISTENT', '1') == '1'
    y = debug_conv(x, w, b, r, stride, fold, bool(relu), mode, tile_n, persistent)
    # reference: frames NCHW fp32
    xf = x.float().permute(0, 3, 4, 1, 2).reshape(clips, 8, Cin, H, H)
    if fold:
        out = torch.zeros_like(xf)
        out[:, :-1, :fold] = xf[:, 1:, :fold]
        out[:, 1:, fold:2 * fold] = xf[:, :-1, fold:2 * fold]
        out[:, :, 2 * fold:] = xf[:, :, 2 * fold:]
        xf = out
    ref = F.conv2d(xf.reshape(clips * 8, Cin, H, H), w.cuda(), b.cuda(), stride=stride, padding=k // 2)
    ref = ref.reshape(clips, 8, Cout, Ho, Ho).permute(0, 3, 4, 1, 2)
    if res:
        ref = ref + r.float()
    if relu:
        ref = ref.relu()
    got = y.float()
    err = (got - ref).abs()
    tol = 0.02 + 0.01 * ref.abs()
    bad = err > tol
    nbad = int(bad.sum())
    print(f"case {i:2d} {name:34s} max_err {float(err.max()):.4g} ref_absmax {float(ref.abs().max()):.3g} "
          f"bad {nbad}/{bad.numel()} {'OK' if nbad == 0 else 'FAIL'}", flush=True)
    if nbad:
        idx = bad.nonzero()
        rows = (bad.reshape(-1, Cout).any(dim=1)).nonzero().flatten()
        cols = (bad.reshape(-1, Cout).any(dim=0)).nonzero().flatten()
        print("   bad rows (first 24):", rows[:24].tolist(), "n_bad_rows", rows.numel(), "of", bad.numel() // Cout)
        print("   bad cols (first 24):", cols[:24].tolist(), "n_bad_cols", cols.numel(), "of", Cout)
        f0 = idx[0].tolist()
        print("   first bad", f0, "got", float(got[tuple(f0)]), "ref", float(ref[tuple(f0)]))
    return nbad == 0


if __name__ == "__main__":
    if len(sys.argv) > 1:
        ok = run_case(int(sys.argv[1]))
        sys.exit(0 if ok else 1)
    fails = 0
    for i in range(len(CASES)):
        try:
            p = subprocess.run([sys.executable, __file__, str(i)], timeout=120, capture_output=True, text=True)
            sys.stdout.write(p.stdout)
            if p.returncode != 0:
                fails += 1
                sys.stdout.write(f"case {i} rc={p.returncode}\n" + p.stderr[-1500:] + "\n")
        except subprocess.TimeoutExpired as e:
            fails += 1
            print(f"case {i} {CASES[i][0]} TIMEOUT (hang)", flush=True)
            if e.stdout:
                print(e.stdout[-500:] if isinstance(e.stdout, str) else e.stdout.decode()[-500:])
    print(f"conv check: {len(CASES) - fails}/{len(CASES)} passed")
    sys.exit(1 if fails else 0)
