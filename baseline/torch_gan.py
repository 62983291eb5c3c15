"""Stock-PyTorch (cuDNN / cuBLAS / cuFFT, fp32, TF32 off) HiFi-GAN v1 train step on one GPU -- the "GPU library baseline"
BASELINE.md section 3 asks for ("the practical kernel to beat").  MEASUREMENT BASELINE ONLY: it is neither the product
(kantts_b200 never imports it) nor the parity oracle (tests never import it); `bench.py --impl torch_gpu` times it and
reports the number under the informational key `gpu_library_baseline`.

It is a from-scratch plain `torch.nn` statement of the same workload bench.py's CONFIG describes (class-default causal
Generator 512 ch / scales 8-8-2-2 with nearest-upsample + conv "repeat" branches, MultiPeriodDiscriminator (2,3,5,7,11),
MultiScaleDiscriminator x3 with db3-DWT pooling and spectral norm on scale 0, mel-L1 x45 + LSGAN + feature matching x2,
three Adam optimisers), executed in the reference trainer's order (kantts/train/trainer.py:469-589: generator phase with
D forwards on (y_, y), then discriminator phase on a re-generated y_).  Weights are random (timing only)."""
import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.nn.utils import spectral_norm, weight_norm

SLOPE = 0.1


class CConv(nn.Module):          # causal weight-normed conv (left pad (k-1)*d)
    def __init__(self, cin, cout, k, d=1):
        super().__init__()
        self.pad = (k - 1) * d
        self.c = weight_norm(nn.Conv1d(cin, cout, k, dilation=d))

    def forward(self, x):
        return self.c(F.pad(x, (self.pad, 0)))


class ResBlock(nn.Module):
    def __init__(self, ch, k, dil=(1, 3, 5)):
        super().__init__()
        self.c1 = nn.ModuleList([CConv(ch, ch, k, d) for d in dil])
        self.c2 = nn.ModuleList([CConv(ch, ch, k, 1) for _ in dil])

    def forward(self, x):
        for a, b in zip(self.c1, self.c2):
            x = b(F.leaky_relu(a(F.leaky_relu(x, SLOPE)), SLOPE)) + x
        return x


class Generator(nn.Module):
    def __init__(self, ch=512, scales=(8, 8, 2, 2), ks=(16, 16, 4, 4), rk=(3, 7, 11)):
        super().__init__()
        self.pre = CConv(80, ch, 7)
        self.ups, self.reps, self.blocks = nn.ModuleList(), nn.ModuleList(), nn.ModuleList()
        for i, (s, k) in enumerate(zip(scales, ks)):
            cin, cout = ch // 2 ** i, ch // 2 ** (i + 1)
            self.ups.append(weight_norm(nn.ConvTranspose1d(cin, cout, k, s)))
            self.reps.append(nn.Sequential(nn.Upsample(scale_factor=s, mode="nearest"), CConv(cin, cout, 7)))
            self.blocks.append(nn.ModuleList([ResBlock(cout, k_) for k_ in rk]))
        self.scales, self.ks = scales, ks
        self.post = CConv(ch // 2 ** len(scales), 1, 7)

    def forward(self, x):
        x = self.pre(x)
        for up, rep, blocks, s, k in zip(self.ups, self.reps, self.blocks, self.scales, self.ks):
            x = F.leaky_relu(x, SLOPE)
            x = up(x)[..., :-(k - s)] + rep(x)
            x = torch.sin(x) + x
            x = sum(b(x) for b in blocks) / len(blocks)
        return torch.tanh(self.post(F.leaky_relu(x)))


class PeriodD(nn.Module):
    def __init__(self, p):
        super().__init__()
        self.p = p
        chans = [1, 32, 128, 512, 1024, 1024]
        self.convs = nn.ModuleList([weight_norm(nn.Conv2d(chans[i], chans[i + 1], (5, 1), (3 if i < 4 else 1, 1), (2, 0)))
                                    for i in range(5)])
        self.post = nn.Conv2d(1024, 1, (2, 1), 1, (1, 0))

    def forward(self, x):
        b, c, t = x.shape
        if t % self.p:
            x = F.pad(x, (0, self.p - t % self.p), "reflect")
        x = x.view(b, c, -1, self.p)
        fm = []
        for l in self.convs:
            x = F.leaky_relu(l(x), SLOPE)
            fm.append(x)
        x = self.post(x)
        fm.append(x)
        return x.flatten(1), fm


class ScaleD(nn.Module):
    def __init__(self, cin, spectral):
        super().__init__()
        norm = spectral_norm if spectral else weight_norm
        spec = [(cin, 128, 15, 1, 1), (128, 128, 41, 4, 4), (128, 256, 41, 4, 16), (256, 512, 41, 4, 16),
                (512, 1024, 41, 4, 16), (1024, 1024, 41, 1, 16), (1024, 1024, 5, 1, 1)]
        self.convs = nn.ModuleList([norm(nn.Conv1d(a, b, k, s, (k - 1) // 2, groups=g)) for a, b, k, s, g in spec])
        self.post = norm(nn.Conv1d(1024, 1, 3, 1, 1))

    def forward(self, x):
        fm = []
        for l in self.convs:
            x = F.leaky_relu(l(x), SLOPE)
            fm.append(x)
        x = self.post(x)
        fm.append(x)
        return x.flatten(1), fm


class DWT(nn.Module):
    LO = [0.035226291882100656, -0.08544127388224149, -0.13501102001039084, 0.4598775021193313, 0.8068915093133388,
          0.3326705529509569]

    def __init__(self):
        super().__init__()
        lo = torch.tensor(self.LO[::-1])
        hi = torch.tensor([(-1) ** (k + 1) * self.LO[::-1][k] for k in range(6)][::-1])
        self.register_buffer("w", torch.stack([lo, hi]).view(2, 1, 6))

    def forward(self, x):            # (B, C, T) -> (B, 2C, (T+5)//2)
        b, c, t = x.shape
        y = F.conv1d(F.pad(x.reshape(b * c, 1, t), (4, 4 + t % 2)), self.w, stride=2)
        return y.view(b, c, 2, -1).transpose(1, 2).reshape(b, 2 * c, -1)


class MSD(nn.Module):
    def __init__(self):
        super().__init__()
        self.ds = nn.ModuleList([ScaleD(1, True), ScaleD(1, False), ScaleD(1, False)])
        self.pool = DWT()
        self.aux = nn.ModuleList([weight_norm(nn.Conv1d(2, 1, 15, 1, 7)), weight_norm(nn.Conv1d(4, 1, 15, 1, 7))])

    def forward(self, y):
        outs, fms = [], []
        o, f = self.ds[0](y)
        outs.append(o); fms.append(f)
        x = y
        for i in range(2):
            x = self.pool(x)
            o, f = self.ds[i + 1](self.aux[i](x))
            outs.append(o); fms.append(f)
        return outs, fms


class MPD(nn.Module):
    def __init__(self):
        super().__init__()
        self.ds = nn.ModuleList([PeriodD(p) for p in (2, 3, 5, 7, 11)])

    def forward(self, y):
        r = [d(y) for d in self.ds]
        return [o for o, _ in r], [f for _, f in r]


class Mel(nn.Module):
    def __init__(self, fs=22050, n_fft=1024, hop=256, n_mels=80):
        super().__init__()
        self.n_fft, self.hop = n_fft, hop
        self.register_buffer("win", torch.hann_window(n_fft))
        # (timing baseline: any fixed 513 x 80 projection costs the same as the Slaney filterbank)
        self.register_buffer("fb", torch.rand(n_fft // 2 + 1, n_mels) / 64)

    def forward(self, y):
        s = torch.stft(y.squeeze(1), self.n_fft, self.hop, self.n_fft, self.win, center=True, pad_mode="constant",
                       return_complex=True)
        amp = torch.sqrt(torch.clamp(s.real ** 2 + s.imag ** 2, min=1e-10))
        return torch.log10(torch.clamp(amp.transpose(1, 2) @ self.fb, min=1e-10))


class TorchGanStep:
    def __init__(self, device):
        torch.backends.cudnn.benchmark = True
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
        self.g, self.msd, self.mpd, self.mel = Generator().to(device), MSD().to(device), MPD().to(device), Mel().to(device)
        mk = lambda m: torch.optim.Adam(m.parameters(), lr=2e-4, betas=(0.5, 0.9))
        self.og, self.os, self.op = mk(self.g), mk(self.msd), mk(self.mpd)

    def step(self, batch):
        y, x = batch
        # ---- generator phase (trainer.py:473-512)
        y_ = self.g(x)
        loss = 45.0 * F.l1_loss(self.mel(y_), self.mel(y))
        for d in (self.msd, self.mpd):
            p_, fm_ = d(y_)
            with torch.no_grad():
                _, fm = d(y)
            loss = loss + sum(F.mse_loss(o, torch.ones_like(o)) for o in p_)
            loss = loss + 2.0 * sum(F.l1_loss(a, b.detach()) for fa, fb in zip(fm_, fm) for a, b in zip(fa, fb))
        self.og.zero_grad(set_to_none=True); self.os.zero_grad(set_to_none=True); self.op.zero_grad(set_to_none=True)
        loss.backward()
        self.og.step()
        # ---- discriminator phase (trainer.py:514-589)
        with torch.no_grad():
            y_ = self.g(x)
        dl = 0.0
        for d in (self.msd, self.mpd):
            p, _ = d(y)
            p_, _ = d(y_.detach())
            dl = dl + sum(F.mse_loss(o, torch.ones_like(o)) for o in p) + sum(F.mse_loss(o, torch.zeros_like(o)) for o in p_)
        self.os.zero_grad(set_to_none=True); self.op.zero_grad(set_to_none=True)
        dl.backward()
        self.os.step(); self.op.step()
        return loss.detach(), dl.detach()
