"""db3 single-level analysis DWT, zero-padding mode (TEST INFRASTRUCTURE).

Restates ``pytorch_wavelets.DWT1DForward(J=1, wave="db3", mode="zero")`` as it
is called by the reference at kantts/models/hifigan/hifigan.py:445-448,469-471.
pytorch_wavelets (unpinned git HEAD in environment.yaml) and PyWavelets 1.3.0
are NOT vendored in /root/reference and not installed here: PARITY UNPINNED for
this function; the algorithm below follows the package's published
``lowlevel.afb1d(mode='zero')``:

    N even:  p = 2*(outsize-1) - N + L,  outsize = floor((N + L - 1) / 2)
             y = conv1d(pad(x, p//2 both sides), flip(dec), stride=2)
    i.e. yl[n] = sum_j dec_lo[j] * x[2n + 1 - j]   (full convolution, odd taps)

The filter taps are PyWavelets' ``Wavelet('db3').dec_lo / dec_hi``.
"""
import torch
import torch.nn.functional as F

DEC_LO = [0.035226291882100656, -0.08544127388224149, -0.13501102001039084,
          0.4598775021193313, 0.8068915093133388, 0.3326705529509569]
DEC_HI = [-0.3326705529509569, 0.8068915093133388, -0.4598775021193313,
          -0.13501102001039084, 0.08544127388224149, 0.035226291882100656]
L = 6


def dwt_out_len(n: int) -> int:
    """pywt.dwt_coeff_len(n, 6, 'zero') = floor((n + 5) / 2)."""
    return (n + L - 1) // 2


def dwt_db3_zero(x: torch.Tensor):
    """x: (B, C, N) -> (yl, yh) each (B, C, floor((N+5)/2)).

    pytorch_wavelets pads odd-length signals by one trailing zero first (afb1d:
    ``if N % 2 == 1: x = cat(x, 0)``), which leaves the formula below unchanged.
    """
    B, C, N = x.shape
    out = dwt_out_len(N)
    p = 2 * (out - 1) - N + L
    # afb1d: symmetric zero padding p//2 on both sides (+1 trailing when p odd)
    xp = F.pad(x, (p // 2, p - p // 2))
    h0 = torch.tensor(DEC_LO[::-1], dtype=x.dtype, device=x.device).view(1, 1, L)
    h1 = torch.tensor(DEC_HI[::-1], dtype=x.dtype, device=x.device).view(1, 1, L)
    xr = xp.reshape(B * C, 1, -1)
    yl = F.conv1d(xr, h0, stride=2).reshape(B, C, -1)
    yh = F.conv1d(xr, h1, stride=2).reshape(B, C, -1)
    assert yl.shape[-1] == out, (yl.shape, out)
    return yl, yh


class DWT1DForward(torch.nn.Module):
    """Drop-in for pytorch_wavelets.DWT1DForward(J=1, wave='db3', mode='zero').

    Registers the same two buffers (``h0``, ``h1`` of shape (1,1,6), the flipped
    analysis filters) so the reference MSD state_dict keeps its
    ``meanpools.{i}.{h0,h1}`` keys.
    Returns ``(yl, [yh])`` like the package.
    """

    def __init__(self, J=1, wave="db3", mode="zero"):
        super().__init__()
        assert J == 1 and wave == "db3" and mode == "zero"
        self.register_buffer("h0", torch.tensor(DEC_LO[::-1], dtype=torch.float32).view(1, 1, L))
        self.register_buffer("h1", torch.tensor(DEC_HI[::-1], dtype=torch.float32).view(1, 1, L))
        self.J, self.mode = J, mode

    def forward(self, x):
        yl, yh = dwt_db3_zero(x)
        return yl, [yh]
