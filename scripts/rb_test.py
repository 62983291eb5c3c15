"""Fused ResidualBlock unit (kt_resblock_fwd) vs the unfused product path (two ops.conv calls) -- development check + timing."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from kantts_b200 import ops, _lib
from kantts_b200._lib import KT_ACT_LRELU, KtResblockDesc, ptr, stream_ptr, check

lib = _lib.load()
torch.manual_seed(0)


def run(C, k, d, causal, B, T, save_h=True, iters=10):
    p1 = (k - 1) * d if causal else (k - 1) * d // 2
    p2 = (k - 1) if causal else (k - 1) // 2
    s1 = ops.ConvSpec(c_in=C, c_out=C, kernel=k, dilation=d, pad_left=p1, pad_right=(k - 1) * d - p1, act_in=KT_ACT_LRELU, act_in_slope=0.1)
    s2 = ops.ConvSpec(c_in=C, c_out=C, kernel=k, dilation=1, pad_left=p2, pad_right=(k - 1) - p2, act_in=KT_ACT_LRELU, act_in_slope=0.1)
    mk = lambda: torch.nn.Parameter((torch.randn(C, C, k) * (1.0 / (C * k) ** 0.5)).cuda(), requires_grad=False)
    v1, v2 = mk(), mk()
    b1, b2 = torch.randn(C).cuda() * 0.1, torch.randn(C).cuda() * 0.1
    x = torch.randn(B, T, C).cuda()
    c1, c2 = ops.PreparedWeight(), ops.PreparedWeight()
    with torch.no_grad():
        h_ref = ops.conv(x, s1, c1, v1, None, b1)
        y_ref = ops.conv(h_ref, s2, c2, v2, None, b2, x)
    d_ = KtResblockDesc(batch=B, t=T, channels=C, kernel=k, dilation=d, pad_left1=p1, pad_left2=p2, slope=0.1, path=0)
    assert lib.kt_resblock_plan(ctypes.byref(d_)) == 1
    nbytes = int(lib.kt_resblock_image_bytes(ctypes.byref(d_)))
    img1 = torch.empty(nbytes // 2, device="cuda", dtype=torch.bfloat16)
    img2 = torch.empty(nbytes // 2, device="cuda", dtype=torch.bfloat16)
    check(lib.kt_resblock_pack(ctypes.byref(d_), ptr(c1.w_fwd), ptr(img1, True), stream_ptr()), "pack1")
    check(lib.kt_resblock_pack(ctypes.byref(d_), ptr(c2.w_fwd), ptr(img2, True), stream_ptr()), "pack2")
    h = torch.full_like(x, float("nan")) if save_h else None
    y = torch.full_like(x, float("nan"))
    check(lib.kt_resblock_fwd(ctypes.byref(d_), ptr(x), ptr(img1, True), ptr(b1), ptr(img2, True), ptr(b2), ptr(h), ptr(y), stream_ptr()), "fwd")
    torch.cuda.synchronize()
    rel = lambda a, b: float((a - b).norm() / b.norm())
    eh = rel(h, h_ref) if save_h else -1.0
    ey = rel(y, y_ref)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        check(lib.kt_resblock_fwd(ctypes.byref(d_), ptr(x), ptr(img1, True), ptr(b1), ptr(img2, True), ptr(b2), ptr(h), ptr(y), stream_ptr()), "fwd")
    e1.record(); torch.cuda.synchronize()
    t_f = e0.elapsed_time(e1) / iters * 1e3
    e0.record()
    with torch.no_grad():
        for _ in range(iters):
            hh = ops.conv(x, s1, c1, v1, None, b1)
            yy = ops.conv(hh, s2, c2, v2, None, b2, x)
    e1.record(); torch.cuda.synchronize()
    t_u = e0.elapsed_time(e1) / iters * 1e3
    gb = 2 * x.numel() * 4 / 1e9 + (x.numel() * 4 / 1e9 if save_h else 0)
    print(f"C={C} k={k} d={d} causal={causal} B={B} T={T} save_h={save_h}: rel err h {eh:.2e} y {ey:.2e} | fused {t_f:.1f} us "
          f"({gb / (t_f * 1e-6):.0f} GB/s algorithmic) vs unfused pair {t_u:.1f} us", flush=True)
    return ey


ok = True
SMALL = [] if os.environ.get("RB_ONLY_BIG") == "1" else None
for cfg in SMALL if SMALL is not None else [(64, 3, 1, True, 2, 300), (64, 7, 3, True, 2, 515), (32, 3, 1, True, 2, 300), (32, 7, 5, True, 3, 1000), (32, 11, 5, False, 2, 777),
            (64, 11, 5, True, 2, 999), (64, 11, 1, False, 1, 64)]:
    ok &= run(*cfg) < 1e-4
for cfg in [(32, 3, 1, True, 16, 8192), (32, 7, 3, True, 16, 8192), (32, 11, 5, True, 16, 8192), (64, 3, 1, True, 16, 4096),
            (64, 7, 3, True, 16, 4096), (64, 11, 5, True, 16, 4096)]:
    ok &= run(*cfg) < 1e-4
    run(*cfg, save_h=False)
print("RB_TEST_OK" if ok else "RB_TEST_FAILED")
