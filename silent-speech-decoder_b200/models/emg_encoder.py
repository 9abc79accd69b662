"""EMG encoder with the reference's interface (``src/models/emg_encoder.py:13-113``), kept in PyTorch.

north_star keeps the encoder's dense layers in PyTorch; this module exists so that BASELINE.json's config 5
(the full evaluate pipeline: encoder forward + native beam-50) can run on a box that has no copy of the
reference.  Same constructor arguments, same parameter names and creation order -- a reference checkpoint's
``state_dict`` loads unchanged, and under the same ``torch.manual_seed`` the random initialisation is the
reference's (``tests/test_encoder_mirror.py`` pins both against outputs of the reference module).
The Conformer stack is ``torchaudio.models.Conformer``, the third-party module the reference itself uses.
"""

from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Tuple

import torch
import torch.nn as nn


@dataclass
class EncoderConfig:
    input_dim: int
    d_model: int = 256
    num_layers: int = 6
    num_heads: int = 4
    ffn_dim: int = 512
    depthwise_conv_kernel_size: int = 15
    dropout: float = 0.1
    subsample_factor: int = 4


class Conv1dSubsampler(nn.Module):
    """log2(factor) stride-2 convolutions (kernel 5, ReLU after each); factor 1 is a 1x1 projection."""

    def __init__(self, input_dim: int, output_dim: int, factor: int = 4, kernel_size: int = 5):
        super().__init__()
        if factor < 1:
            raise ValueError("factor must be >= 1")
        self._geometry: List[Tuple[int, int, int]] = []  # (kernel, stride, padding) per convolution
        stages: List[nn.Module] = []
        width, left = input_dim, factor
        while left > 1:
            stages += [nn.Conv1d(width, output_dim, kernel_size=kernel_size, stride=2, padding=kernel_size // 2), nn.ReLU()]
            self._geometry.append((kernel_size, 2, kernel_size // 2))
            width, left = output_dim, left // 2
        if not stages:
            stages.append(nn.Conv1d(width, output_dim, kernel_size=1))
            self._geometry.append((1, 1, 0))
        self.net = nn.Sequential(*stages)
        self.factor = factor

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.net(x.transpose(1, 2)).transpose(1, 2)  # (batch, time, feat) -> (batch, time', output_dim)

    def output_lengths(self, lengths: torch.Tensor) -> torch.Tensor:
        out = lengths.clone()
        for k, s, p in self._geometry:
            out = torch.div(out + 2 * p - k, s, rounding_mode="floor") + 1
        return out


class EMGConformerEncoder(nn.Module):
    def __init__(self, cfg: EncoderConfig):
        super().__init__()
        import torchaudio  # the reference's own dependency for the Conformer stack

        self.subsample = Conv1dSubsampler(cfg.input_dim, cfg.d_model, factor=cfg.subsample_factor)
        self.encoder = torchaudio.models.Conformer(
            input_dim=cfg.d_model, num_heads=cfg.num_heads, ffn_dim=cfg.ffn_dim, num_layers=cfg.num_layers,
            depthwise_conv_kernel_size=cfg.depthwise_conv_kernel_size, dropout=cfg.dropout)
        self.subsample_factor = cfg.subsample_factor

    def forward(self, x: torch.Tensor, lengths: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
        """x (batch, time, feat), lengths (batch,) -> (batch, time', d_model), adjusted lengths."""
        x = self.subsample(x)
        out_lengths = None
        if lengths is not None:
            out_lengths = torch.clamp(self.subsample.output_lengths(lengths), max=x.shape[1])
        return self.encoder(x, out_lengths)


MPS_FAST_BASIC = dict(input_dim=640, d_model=256, num_layers=6, num_heads=4, ffn_dim=512, depthwise_conv_kernel_size=15,
                      dropout=0.1, subsample_factor=2)    # configs/mps_fast_basic.yaml:24-34
MPS_FAST_PLUS = dict(input_dim=640, d_model=288, num_layers=6, num_heads=6, ffn_dim=1152, depthwise_conv_kernel_size=15,
                     dropout=0.12, subsample_factor=2)    # configs/mps_fast_plus.yaml:25-35 (ctc_dropout 0.12)
