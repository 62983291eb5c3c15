"""autograd wrappers of the SAM-BERT entry points of libkantts_b200.so (LayerNorm, multi-head attention,
FSMN memory block, LengthRegulator gather).  Same contract as ops.py: CUDA fp32 tensors only, explicit
stream, RuntimeError on any failure -- no PyTorch / CPU fallback."""
import ctypes

import torch

from . import _lib
from ._lib import KtAttnDesc, check, ptr, stream_ptr
from .ops import _count


def _u8(mask):
    """bool mask -> uint8 view for the C ABI (zero-copy)."""
    if mask is None:
        return None
    m = mask.contiguous()
    return m.view(torch.uint8) if m.dtype == torch.bool else m


class LayerNormFn(torch.autograd.Function):
    """nn.LayerNorm over the last dim (kt_layernorm_fwd / kt_layernorm_bwd)."""

    @staticmethod
    def forward(ctx, x, gamma, beta, eps):
        lib = _lib.load()
        x = x.contiguous()
        c = x.shape[-1]
        rows = x.numel() // c
        y = torch.empty_like(x)
        mean = torch.empty(rows, device=x.device, dtype=torch.float32)
        rstd = torch.empty(rows, device=x.device, dtype=torch.float32)
        check(lib.kt_layernorm_fwd(ptr(x), ptr(gamma.detach()), ptr(beta.detach()), ptr(y), ptr(mean), ptr(rstd),
                                   rows, c, float(eps), stream_ptr()), "kt_layernorm_fwd")
        _count()
        ctx.save_for_backward(x, gamma, mean, rstd)
        return y

    @staticmethod
    def backward(ctx, dy):
        lib = _lib.load()
        x, gamma, mean, rstd = ctx.saved_tensors
        c = x.shape[-1]
        rows = x.numel() // c
        dy = dy.contiguous()
        dx = torch.empty_like(x)
        dgamma = torch.empty_like(gamma)
        dbeta = torch.empty_like(gamma)
        n = int(lib.kt_layernorm_bwd_workspace(rows, c))
        ws = torch.empty(n, device=x.device, dtype=torch.float32)
        check(lib.kt_layernorm_bwd(ptr(dy), ptr(x), ptr(gamma.detach()), ptr(mean), ptr(rstd), ptr(dx), ptr(dgamma),
                                   ptr(dbeta), ptr(ws), n, rows, c, stream_ptr()), "kt_layernorm_bwd")
        _count(2)
        return dx, dgamma, dbeta, None


def layer_norm(x, gamma, beta, eps):
    return LayerNormFn.apply(x, gamma, beta, eps)


def _attn_desc(B, H, D, lq, lk, qs, ks, vs, os_, mask, p_drop=0.0):
    d = KtAttnDesc(batch=B, heads=H, d_head=D, lq=lq, lk=lk, q_stride=qs, k_stride=ks, v_stride=vs, o_stride=os_,
                   mask_q_stride=0, mask_b_stride=0, scale=float(D) ** -0.5, keep_scale=1.0 / (1.0 - p_drop))
    if mask is not None:
        if mask.dim() == 2:                      # (B, Lk) key-padding mask, broadcast over the queries
            assert mask.shape == (B, lk), (mask.shape, B, lk)
            d.mask_q_stride, d.mask_b_stride = 0, lk
        else:                                    # (B or 1, Lq, Lk)
            assert mask.shape[1:] == (lq, lk) and mask.shape[0] in (1, B), (mask.shape, B, lq, lk)
            d.mask_q_stride, d.mask_b_stride = lk, (lq * lk if mask.shape[0] == B else 0)
    return d


def _keep_mask(p_drop, shape, device):
    """nn.Dropout(p) on the attention probabilities (sambert/__init__.py:26): Bernoulli keep mask, uint8."""
    if p_drop <= 0.0:
        return None
    return (torch.rand(shape, device=device) >= p_drop).view(torch.uint8)


def _off(t, col):
    """device address of column `col` of the first row of a contiguous row tensor."""
    return ptr(t) + 4 * col


class SelfAttnFn(torch.autograd.Function):
    """softmax(Q K^T / sqrt(d) + mask) V over all heads, straight from the fused QKV projection
    (sambert/__init__.py:80-100).  qkv: (B, L, 3*H*D) -> out (B, L, H*D), probs (H*B, L, L)."""

    @staticmethod
    def forward(ctx, qkv, mask, n_head, p_drop=0.0, keep=None):
        lib = _lib.load()
        qkv = qkv.contiguous()
        B, L, w = qkv.shape
        hd = w // 3
        D = hd // n_head
        m = _u8(mask)
        d = _attn_desc(B, n_head, D, L, L, w, w, w, hd, m, p_drop)
        out = torch.empty(B, L, hd, device=qkv.device, dtype=torch.float32)
        probs = torch.empty(n_head * B, L, L, device=qkv.device, dtype=torch.float32)
        keep = _u8(keep) if keep is not None else _keep_mask(p_drop, probs.shape, qkv.device)
        dropped = torch.empty_like(probs) if keep is not None else None
        check(lib.kt_attention_fwd(ctypes.byref(d), _off(qkv, 0), _off(qkv, hd), _off(qkv, 2 * hd), ptr(m, True), ptr(keep, True),
                                   ptr(out), ptr(probs), ptr(dropped), stream_ptr()), "kt_attention_fwd")
        _count()
        ctx.d, ctx.hd = d, hd
        ctx.save_for_backward(qkv, probs, keep)
        attn = probs if dropped is None else dropped
        ctx.mark_non_differentiable(attn)
        return out, attn

    @staticmethod
    def backward(ctx, dout, _dprobs):
        lib = _lib.load()
        qkv, probs, keep = ctx.saved_tensors
        d, hd = ctx.d, ctx.hd
        dout = dout.contiguous()
        dqkv = torch.empty_like(qkv)
        delta = torch.empty(d.heads * d.batch * d.lq, device=qkv.device, dtype=torch.float32)
        check(lib.kt_attention_bwd(ctypes.byref(d), _off(qkv, 0), _off(qkv, hd), _off(qkv, 2 * hd), ptr(probs),
                                   ptr(keep, True), ptr(dout), _off(dqkv, 0), _off(dqkv, hd), _off(dqkv, 2 * hd),
                                   ptr(delta), 0, stream_ptr()), "kt_attention_bwd")
        _count(2)
        return dqkv, None, None, None, None


class PncaAttnFn(torch.autograd.Function):
    """The two attentions of MultiHeadPNCAAttention (sambert/__init__.py:269-300) sharing their queries:
    x_qkv (B, L, 3HD) self part under the causal band mask, h_kv (B, Lh, 2HD) memory part under the
    look-ahead band mask.  -> out_x, out_h (B, L, HD), probs_x (H*B, L, L), probs_h (H*B, L, Lh)."""

    @staticmethod
    def forward(ctx, x_qkv, h_kv, mask_x, mask_h, n_head, p_drop=0.0, keep_x=None, keep_h=None):
        lib = _lib.load()
        x_qkv, h_kv = x_qkv.contiguous(), h_kv.contiguous()
        B, L, w = x_qkv.shape
        hd = w // 3
        D = hd // n_head
        Lh = h_kv.shape[1]
        mx, mh = _u8(mask_x), _u8(mask_h)
        dx = _attn_desc(B, n_head, D, L, L, w, w, w, hd, mx, p_drop)
        dh = _attn_desc(B, n_head, D, L, Lh, w, 2 * hd, 2 * hd, hd, mh, p_drop)
        out_x = torch.empty(B, L, hd, device=x_qkv.device, dtype=torch.float32)
        out_h = torch.empty_like(out_x)
        px = torch.empty(n_head * B, L, L, device=x_qkv.device, dtype=torch.float32)
        ph = torch.empty(n_head * B, L, Lh, device=x_qkv.device, dtype=torch.float32)
        kx = _u8(keep_x) if keep_x is not None else _keep_mask(p_drop, px.shape, x_qkv.device)
        kh = _u8(keep_h) if keep_h is not None else _keep_mask(p_drop, ph.shape, x_qkv.device)
        pxd = torch.empty_like(px) if kx is not None else None
        phd = torch.empty_like(ph) if kh is not None else None
        st = stream_ptr()
        check(lib.kt_attention_fwd(ctypes.byref(dx), _off(x_qkv, 0), _off(x_qkv, hd), _off(x_qkv, 2 * hd), ptr(mx, True),
                                   ptr(kx, True), ptr(out_x), ptr(px), ptr(pxd), st), "kt_attention_fwd")
        check(lib.kt_attention_fwd(ctypes.byref(dh), _off(x_qkv, 0), _off(h_kv, 0), _off(h_kv, hd), ptr(mh, True),
                                   ptr(kh, True), ptr(out_h), ptr(ph), ptr(phd), st), "kt_attention_fwd")
        _count(2)
        ctx.dx, ctx.dh, ctx.hd = dx, dh, hd
        ctx.save_for_backward(x_qkv, h_kv, px, ph, kx, kh)
        ax = px if pxd is None else pxd
        ah = ph if phd is None else phd
        ctx.mark_non_differentiable(ax, ah)
        return out_x, out_h, ax, ah

    @staticmethod
    def backward(ctx, dox, doh, _dpx, _dph):
        lib = _lib.load()
        x_qkv, h_kv, px, ph, kx, kh = ctx.saved_tensors
        dx, dh, hd = ctx.dx, ctx.dh, ctx.hd
        dox, doh = dox.contiguous(), doh.contiguous()
        dqkv = torch.empty_like(x_qkv)
        dhkv = torch.empty_like(h_kv)
        delta = torch.empty(dx.heads * dx.batch * dx.lq, device=x_qkv.device, dtype=torch.float32)
        st = stream_ptr()
        check(lib.kt_attention_bwd(ctypes.byref(dx), _off(x_qkv, 0), _off(x_qkv, hd), _off(x_qkv, 2 * hd), ptr(px),
                                   ptr(kx, True), ptr(dox), _off(dqkv, 0), _off(dqkv, hd), _off(dqkv, 2 * hd), ptr(delta), 0,
                                   st), "kt_attention_bwd")
        check(lib.kt_attention_bwd(ctypes.byref(dh), _off(x_qkv, 0), _off(h_kv, 0), _off(h_kv, hd), ptr(ph),
                                   ptr(kh, True), ptr(doh), _off(dqkv, 0), _off(dhkv, 0), _off(dhkv, hd), ptr(delta), 1, st),
              "kt_attention_bwd")
        _count(4)
        return dqkv, dhkv, None, None, None, None, None, None


def pnca_attn_step(q_row, x_cache, h_kv, mask_x, mask_h, n_head):
    """One free-running decoder step of MultiHeadPNCAAttention (sambert/__init__.py:212-300), inference only.
    q_row (B, 1, 3HD): this step's fused QKV projection (only its q block is read here; the caller has already written
    the row into ``x_cache``); x_cache (B, Lmax, 3HD): the preallocated K/V state of the self part (rows beyond the
    current step are masked by ``mask_x`` (B or 1, 1, Lmax), so the reference's ``torch.cat`` growth is never needed);
    h_kv (B, Lh, 2HD): the memory keys / values projected once; mask_h (B or 1, 1, Lh).
    -> out_x, out_h (B, 1, HD), probs_x (H*B, 1, Lmax), probs_h (H*B, 1, Lh)."""
    lib = _lib.load()
    assert q_row.is_contiguous() and x_cache.is_contiguous() and h_kv.is_contiguous()
    B, _, w = q_row.shape
    hd = w // 3
    D = hd // n_head
    lmax, lh = x_cache.shape[1], h_kv.shape[1]
    mx, mh = _u8(mask_x), _u8(mask_h)
    dx = _attn_desc(B, n_head, D, 1, lmax, w, w, w, hd, mx)
    dh = _attn_desc(B, n_head, D, 1, lh, w, 2 * hd, 2 * hd, hd, mh)
    out_x = torch.empty(B, 1, hd, device=q_row.device, dtype=torch.float32)
    out_h = torch.empty_like(out_x)
    px = torch.empty(n_head * B, 1, lmax, device=q_row.device, dtype=torch.float32)
    ph = torch.empty(n_head * B, 1, lh, device=q_row.device, dtype=torch.float32)
    st = stream_ptr()
    check(lib.kt_attention_fwd(ctypes.byref(dx), _off(q_row, 0), _off(x_cache, hd), _off(x_cache, 2 * hd), ptr(mx, True), None,
                               ptr(out_x), ptr(px), None, st), "kt_attention_fwd")
    check(lib.kt_attention_fwd(ctypes.byref(dh), _off(q_row, 0), _off(h_kv, 0), _off(h_kv, hd), ptr(mh, True), None,
                               ptr(out_h), ptr(ph), None, st), "kt_attention_fwd")
    _count(2)
    return out_x, out_h, px, ph


class FsmnMemoryFn(torch.autograd.Function):
    """MemoryBlockV2 (fsmn.py:46-77).  x (B, T, C), w (C, 1, K), mask (B, T) bool or None."""

    @staticmethod
    def forward(ctx, x, w, mask, pad_left):
        lib = _lib.load()
        x = x.contiguous()
        B, T, C = x.shape
        K = w.shape[-1]
        m = _u8(mask)
        y = torch.empty_like(x)
        check(lib.kt_fsmn_fwd(ptr(x), ptr(w.detach().contiguous()), ptr(m, True), ptr(y), B, T, C, K, pad_left,
                              stream_ptr()), "kt_fsmn_fwd")
        _count()
        ctx.pad_left = pad_left
        ctx.save_for_backward(x, w, m)
        return y

    @staticmethod
    def backward(ctx, dy):
        lib = _lib.load()
        x, w, m = ctx.saved_tensors
        B, T, C = x.shape
        K = w.shape[-1]
        dy = dy.contiguous()
        dx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        dw = ws = None
        n = 0
        if ctx.needs_input_grad[1]:
            dw = torch.empty_like(w)
            n = int(lib.kt_fsmn_bwd_workspace(B, T, C, K))
            ws = torch.empty(n, device=x.device, dtype=torch.float32)
        check(lib.kt_fsmn_bwd(ptr(x), ptr(dy), ptr(w.detach().contiguous()), ptr(m, True), ptr(dx), ptr(dw), ptr(ws), n,
                              B, T, C, K, ctx.pad_left, stream_ptr()), "kt_fsmn_bwd")
        _count(3)
        return dx, dw, None, None


class RowsGatherFn(torch.autograd.Function):
    """LengthRegulator expansion (adaptors.py:15-37) as a gather; idx (B, T_out) int32 (-1 = zero row),
    start / count (B, T_in) int32 = each input row's span of output rows."""

    @staticmethod
    def forward(ctx, x, idx, start, count):
        lib = _lib.load()
        x = x.contiguous()
        B, T_in, C = x.shape
        T_out = idx.shape[1]
        out = torch.empty(B, T_out, C, device=x.device, dtype=torch.float32)
        check(lib.kt_rows_gather_fwd(ptr(x), ptr(idx, True), ptr(out), B, T_out, T_in, C, stream_ptr()),
              "kt_rows_gather_fwd")
        _count()
        ctx.save_for_backward(idx, start, count)
        ctx.t_in = T_in
        return out

    @staticmethod
    def backward(ctx, dout):
        lib = _lib.load()
        idx, start, count = ctx.saved_tensors
        dout = dout.contiguous()
        B, T_out, C = dout.shape
        din = torch.empty(B, ctx.t_in, C, device=dout.device, dtype=torch.float32)
        check(lib.kt_rows_gather_bwd(ptr(dout), ptr(idx, True), ptr(start, True), ptr(count, True), ptr(din), B, T_out, ctx.t_in, C,
                                     stream_ptr()), "kt_rows_gather_bwd")
        _count()
        return din, None, None, None
