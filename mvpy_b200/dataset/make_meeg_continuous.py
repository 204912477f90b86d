"""Synthetic continuous M/EEG generator -- the benchmark / parity input of the TRF path.

Host-side plumbing (plain torch, no kernels): BASELINE.json quotes its metric on
``make_meeg_continuous``-style data, and ``/root/reference`` does not exist on the GPU box, so the
generator is restated here.  It draws from the global torch RNG in the same order as the reference
(``mvpy/dataset/make_meeg_continuous.py:16-197``: randint for stimulus frequencies, rand for
background phases, one randperm per source for the sensor dipoles, randn for the innovations), so a
given ``torch.manual_seed`` reproduces the reference's arrays on CPU (checked bit-for-bit by
``tests/test_host_logic.py`` against the reference's checksums).

Model: ``n_background`` 1/f oscillatory sources plus ``n_features`` AR(1) stimulus series (active
only in [t_baseline, t_baseline + t_duration)) are convolved with Hann-tapered sinusoidal response
functions, projected through a smooth random sensor pattern, summed, and normalised to unit
standard deviation.  Returns X (trials, features, time) and y (trials, channels, time).
"""
from __future__ import annotations

import math
from typing import Tuple

import torch


def make_meeg_layout(n_channels: int, device: str = 'cpu') -> torch.Tensor:
    """Sensor positions on the upper unit hemisphere: a centred square grid pushed onto the disc with
    the concentric (Shirley-Chiu) map, z lifted to the sphere.
    reference: ``mvpy/dataset/make_meeg_layout.py:11-76``."""
    n_cols = math.ceil(math.sqrt(n_channels))
    n_rows = math.ceil(n_channels / n_cols)
    gx = torch.linspace(0.5 / n_cols, 1 - 0.5 / n_cols, n_cols, device=device)
    gy = torch.linspace(0.5 / n_rows, 1 - 0.5 / n_rows, n_rows, device=device)
    mx, my = torch.meshgrid(gx, gy, indexing='xy')
    sq = 2.0 * torch.stack([mx.flatten(), my.flatten()], dim=-1) - 1.0
    a, b = sq[:, 0], sq[:, 1]
    nz = (a != 0.0) | (b != 0.0)
    an, bn = a[nz], b[nz]
    wide = an.abs() > bn.abs()
    ang_w = math.pi / 4 * (bn[wide] / (an[wide] + 1e-12))
    xw = torch.sign(an[wide]) * an[wide].abs() * torch.cos(ang_w)
    yw = torch.sign(an[wide]) * an[wide].abs() * torch.sin(ang_w)
    tall = ~wide
    ang_t = math.pi / 2 - math.pi / 4 * (an[tall] / (bn[tall] + 1e-12))
    xt = torch.sign(bn[tall]) * bn[tall].abs() * torch.cos(ang_t)
    yt = torch.sign(bn[tall]) * bn[tall].abs() * torch.sin(ang_t)
    dx, dy = torch.zeros_like(a), torch.zeros_like(b)
    dx[nz] = torch.cat([xw, xt])  # same (wide-first) placement as the reference
    dy[nz] = torch.cat([yw, yt])
    dz = torch.sqrt(torch.clamp(1 - (dx * dx + dy * dy), min=0.0))
    return torch.stack([dx, dy, dz], dim=-1)[:n_channels]


def _rbf(x: torch.Tensor, gamma: float) -> torch.Tensor:
    """exp(-gamma |xi - xj|^2), reference: ``mvpy/math/kernel_rbf.py``."""
    sq = (x * x).sum(1, keepdim=True)
    d2 = sq + sq.T - 2.0 * (x @ x.T)
    d2.clamp_(min=0.0)
    return d2.mul_(-gamma).exp_()


def _hann(n: int, device: str) -> torch.Tensor:
    """0.5 - 0.5 cos(2 pi linspace(0,1,n)), reference: ``mvpy/signal/raisedcosinewindow.py``."""
    return 0.5 - (1 - 0.5) * torch.cos(2 * torch.pi * torch.linspace(0, 1, n, device=device))


def make_meeg_continuous(n_trials: int = 120, n_channels: int = 64, t_padding: float = 1.0,
                         t_baseline: float = 0.25, t_duration: float = 1.0, t_length: float = 2.0,
                         fs: int = 200, n_background: int = 50, n_features: int = 3, n_cycles: int = 2,
                         snr: float = 0.1, gamma: float = 1.0, phi: float = 0.5, return_Xyß: bool = False,
                         device: str = 'cpu') -> Tuple[torch.Tensor, ...]:
    s_pad, s_base, s_dur = int(t_padding * fs), int(t_baseline * fs), int(t_duration * fs)
    T_all = int(t_length * fs) + s_pad
    n_src = n_background + n_features

    # source spectra: background 1..n_background Hz with random phase, stimuli 2-4 Hz at zero phase
    freq = torch.cat((torch.arange(1, n_background + 1, device=device).float(),
                      torch.randint(low=2, high=5, size=(n_features,), device=device).float()))
    phase = torch.cat((torch.rand(n_background, device=device), torch.zeros(n_features, device=device)))
    amp = 1 / freq

    # sensor patterns: smooth covariance x random +1/-1 dipole per source, smoothed again
    K = _rbf(make_meeg_layout(n_channels, device=device), gamma)
    dipole = torch.tensor([0] * (n_channels - 2) + [1, -1], device=device).float()
    sens = torch.stack([dipole[torch.randperm(n_channels)] for _ in range(n_src)], dim=1)  # (ch, src)
    mix = torch.bmm(K[None] * sens.T[:, None, :], K.expand(n_src, n_channels, n_channels))  # (src, ch, ch)

    # response functions: n_cycles periods of a Hann-tapered sinusoid per source, padded to equal length
    L = (n_cycles * (1 / freq * fs)).clip(max=T_all).long()
    Lmax = int(L.max())
    trf = torch.stack([
        torch.nn.functional.pad(
            _hann(int(L[i]), device) * amp[i]
            * torch.sin(2 * torch.pi * freq[i] * torch.linspace(0, 1, int(L[i]), device=device) + phase[None, i] * torch.pi),
            (0, Lmax - int(L[i])))
        for i in range(n_src)], dim=0)
    beta = trf.expand(n_channels, n_src, Lmax)
    beta = torch.bmm(beta.permute(1, 2, 0), mix).permute(2, 0, 1)  # (ch, src, lag)

    # scale the background responses so that stimulus-to-background response mass equals snr
    mass = beta.abs().sum(-1).sum(0)
    ratio = snr / (mass[n_background:].sum() / mass[:n_background].sum())
    beta[:, :n_background, :] = beta[:, :n_background, :] / ratio

    # AR(1) sources; stimulus innovations only inside the active window
    E = torch.randn((n_trials, n_src, T_all), device=device)
    E[:, n_background:, 0:s_pad + s_base] = 0.0
    E[:, n_background:, s_pad + s_base + s_dur:] = 0.0
    S = torch.zeros_like(E)
    for t in range(1, T_all):
        S[..., t] = phi * S[..., t - 1] + E[..., t]
    act = slice(s_pad + s_base, s_pad + s_base + s_dur)
    e_bg, e_fg = S[:, :n_background, act].abs().sum(), S[:, n_background:, act].abs().sum()
    S[:, :n_background, :] = S[:, :n_background, :] / (snr / (e_fg / e_bg))

    # FFT convolution of every source with its response, summed over sources
    Sp = torch.nn.functional.pad(S, (0, Lmax - int(L.min())))
    n_fft = 1 << (Sp.shape[-1] + 1).bit_length()
    FS = torch.fft.rfft(Sp, n=n_fft, dim=-1)
    FB = torch.fft.rfft(beta, n=n_fft, dim=-1)
    n_keep = S.shape[-1]
    if n_src * n_trials * n_channels * n_keep * 4 <= (2 << 30):
        y = torch.stack([torch.fft.irfft(FS[:, i, ...] * FB[:, i, None, ...], n=n_fft, dim=-1)[..., :n_keep]
                         for i in range(n_src)], dim=0).sum(0).swapaxes(0, 1)
    else:
        # same sum, accumulated source by source: the stacked form needs n_src x the output size in RAM
        # (~10 GB at 306 channels, ~91 GB at config 5 -- SURVEY.md section 8d)
        y = torch.zeros((n_channels, n_trials, n_keep), device=device)
        for i in range(n_src):
            y += torch.fft.irfft(FS[:, i, ...] * FB[:, i, None, ...], n=n_fft, dim=-1)[..., :n_keep]
        y = y.swapaxes(0, 1)
    y = y / y.std()

    X = S[:, n_background:, s_pad:]
    y = y[..., s_pad:]
    if return_Xyß:
        return X, y, beta[:, n_background:, :]
    return X, y
