"""Layer-level conv oracle on channels-first CPU tensors (TEST INFRASTRUCTURE): restates the
KtConv1dDesc semantics of include/kantts_b200.h with torch CPU ops, i.e. exactly what the
reference's layers dispatch (layers.py:44-46,82-88,123,161; hifigan.py:85,254-267)."""
import torch
import torch.nn.functional as F


def conv_layer(x, w, bias=None, resid=None, *, stride=1, dilation=1, pad_left=0, pad_right=0, groups=1,
               transposed=False, upsample=1, crop=0, act_in=None, act_out=None, t_out=None):
    """x: (B, Cin, T) (or (B, Cin, H, P) for the period layout: conv along H).  w: reference layout."""
    four_d = x.dim() == 4
    if four_d:
        B, C, H, P = x.shape
        x = x.permute(0, 3, 1, 2).reshape(B * P, C, H)
        if resid is not None:
            resid = resid.permute(0, 3, 1, 2).reshape(B * P, resid.shape[1], resid.shape[2])
    if act_in is not None:
        x = F.leaky_relu(x, act_in)
    if upsample > 1:
        x = F.interpolate(x, scale_factor=upsample, mode="nearest")
    w3 = w.reshape(w.shape[0], w.shape[1], -1)
    if transposed:
        y = F.conv_transpose1d(x, w3, bias, stride=stride, padding=pad_left, dilation=dilation)
        if crop:
            y = y[..., : y.shape[-1] - crop]
    else:
        y = F.conv1d(F.pad(x, (pad_left, pad_right)), w3, bias, stride=stride, dilation=dilation, groups=groups)
    if act_out == "tanh":
        y = torch.tanh(y)
    elif act_out is not None:
        y = F.leaky_relu(y, act_out)
    if resid is not None:
        y = y + resid
    if four_d:
        y = y.reshape(B, P, y.shape[1], y.shape[2]).permute(0, 2, 3, 1)
    return y
