"""In-kernel timeline of conv_tc_kernel's CTA 0 (kt_debug_set_trace): per role / tile clock64() stamps.
roles: 0/1 producer groups (ev0 stage free, ev1 image staged), 2 MMA issuer (ev0 start, ev1 accumulator free,
ev2 first image ready, ev3 all MMAs issued + commit), 3 epilogue (ev0 start, ev1 accumulator full, ev2 TMEM drained,
ev3 tile stored)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from kantts_b200 import ops, _lib
from kantts_b200._lib import KT_ACT_LRELU

LAYERS = {
    "gen_32_32_k7": (dict(c_in=32, c_out=32, kernel=7, pad_left=6), 16, 8192, 0, True),
    "gen_64_64_k7": (dict(c_in=64, c_out=64, kernel=7, pad_left=6), 16, 4096, 0, True),
    "gen_128_128_k11": (dict(c_in=128, c_out=128, kernel=11, pad_left=10), 16, 2048, 0, True),
    "gen_256_256_k3": (dict(c_in=256, c_out=256, kernel=3, pad_left=2), 16, 256, 0, True),
    "mpd_1024_1024_k5_p3": (dict(c_in=1024, c_out=1024, kernel=5, pad_left=2, pad_right=2), 32, 34, 3, False),
    "msd_128_128_k41_g4": (dict(c_in=128, c_out=128, kernel=41, stride=4, pad_left=20, pad_right=20, groups=4), 16, 8192, 0, False),
}
lib = _lib.load()
for name, (kw, B, T, period, resid) in LAYERS.items():
    spec = ops.ConvSpec(**kw)
    if resid:
        spec.act_in, spec.act_in_slope = KT_ACT_LRELU, 0.1
    else:
        spec.act_out, spec.act_out_slope = KT_ACT_LRELU, 0.1
    g = torch.Generator().manual_seed(1)
    v = torch.nn.Parameter((torch.randn((spec.c_out, spec.c_in // spec.groups, spec.kernel), generator=g) * 0.05).cuda(), requires_grad=False)
    bias = torch.zeros(spec.c_out, device="cuda")
    xs = (B, T, period, spec.c_in) if period else (B, T, spec.c_in)
    x = torch.randn(xs, generator=g).cuda()
    cache = ops.PreparedWeight()
    with torch.no_grad():
        y = ops.conv(x, spec, cache, v, None, bias)
        r = torch.randn_like(y) if resid else None
        for _ in range(3):
            y = ops.conv(x, spec, cache, v, None, bias, r)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            y = ops.conv(x, spec, cache, v, None, bias, r)
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 100
        tr = torch.zeros(8 * 16 * 4, dtype=torch.int64, device="cuda")
        lib.kt_debug_set_trace(tr.data_ptr())
        y = ops.conv(x, spec, cache, v, None, bias, r)
        torch.cuda.synchronize()
        lib.kt_debug_set_trace(None)
    t = tr.cpu().view(8, 16, 4)
    t0 = int(t[:6][t[:6] > 0].min())
    rel = lambda a: "   -  " if a == 0 else f"{(a - t0) / 1.9e3:6.1f}"
    print(f"=== {name}: {us:.1f} us/launch (warm, back to back); CTA0 timeline in us (clock64 / 1.9 GHz)")
    for ti in range(10):
        if int(t[:6, ti].max()) == 0:
            break
        row = []
        for role, nm in ((0, "P0"), (1, "P1"), (2, "MMA"), (3, "EPI")):
            row.append(nm + ":" + " ".join(rel(int(t[role, ti, e])) for e in range(4 if role >= 2 else 2)))
        print(f"  tile {ti}: " + " | ".join(row) + f" | issuer waited: images {int(t[6, ti, 0]) / 1.9e3:5.1f} us, weights {int(t[6, ti, 1]) / 1.9e3:5.1f} us (steps so far {int(t[6, ti, 2])}, stages a/b {int(t[6, ti, 3])})")
    for role, ttl in ((4, "tile 0"), (5, "tile 1")):
        chunks = [n for n in range(16) if int(t[role, n].max()) > 0]
        if chunks:
            print(f"  epilogue warp 0, {ttl}, per 32-column chunk (us): start / TMEM loaded (+fuse2 add) / transposed / stored")
            for n in chunks:
                print(f"    chunk {n:2d}: " + " ".join(rel(int(t[role, n, e])) for e in range(4)))


# ---- epilogue ablation (kt_debug_set_flags): bit 0 no residual loads, 1 no stores, 2 no transposition, 3 no TMEM loads
print("=== epilogue ablation: us/launch and CTA0 epilogue window of tile 0 (accumulator full -> tile stored)")
for name in ("gen_128_128_k11", "gen_32_32_k7", "gen_64_64_k7"):
    kw, B, T, period, resid = LAYERS[name]
    spec = ops.ConvSpec(**kw)
    spec.act_in, spec.act_in_slope = KT_ACT_LRELU, 0.1
    g = torch.Generator().manual_seed(1)
    v = torch.nn.Parameter((torch.randn((spec.c_out, spec.c_in, spec.kernel), generator=g) * 0.05).cuda(), requires_grad=False)
    bias = torch.zeros(spec.c_out, device="cuda")
    x = torch.randn((B, T, spec.c_in), generator=g).cuda()
    cache = ops.PreparedWeight()
    with torch.no_grad():
        y = ops.conv(x, spec, cache, v, None, bias)
        r = torch.randn_like(y)
        for flags in (0, 1, 2, 3, 4, 8, 12, 15):
            lib.kt_debug_set_flags(flags)
            for _ in range(2):
                y = ops.conv(x, spec, cache, v, None, bias, r)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                y = ops.conv(x, spec, cache, v, None, bias, r)
            e1.record(); torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 100
            tr = torch.zeros(8 * 16 * 4, dtype=torch.int64, device="cuda")
            lib.kt_debug_set_trace(tr.data_ptr())
            y = ops.conv(x, spec, cache, v, None, bias, r)
            torch.cuda.synchronize()
            lib.kt_debug_set_trace(None)
            t = tr.cpu().view(8, 16, 4)
            ep = [(int(t[3, ti, 3]) - int(t[3, ti, 1])) / 1.9e3 for ti in range(2)]
            mm = [(int(t[2, ti, 3]) - int(t[2, ti, 2])) / 1.9e3 for ti in range(2)]
            print(f"  {name:18s} flags {flags:2d}: {us:6.1f} us/launch | epilogue tile0 {ep[0]:5.1f} us tile1 {ep[1]:5.1f} us | MMA phase tile0 {mm[0]:5.1f} tile1 {mm[1]:5.1f}")
        lib.kt_debug_set_flags(0)
