"""CTC head with the reference's interface (``src/models/heads.py:23-37``): dropout, linear, log_softmax.

The head is the PRODUCER of the hot path's input and stays PyTorch (scope row a4).  ``logits`` is the
same module stopped before ``log_softmax`` -- what the logits entry of the loss consumes (scope row f-2:
``DistillationCTCLoss(..., from_logits=True)`` / ``ctc_loss(..., from_logits=True)``), and what the decoders
accept unchanged (argmax and the beam search's own fp32 log-softmax are invariant to a per-row constant).
"""

from __future__ import annotations

import torch
import torch.nn as nn


class CTCHead(nn.Module):
    def __init__(self, input_dim: int, vocab_size: int, dropout: float = 0.1):
        super().__init__()
        self.dropout = nn.Dropout(dropout)
        self.fc = nn.Linear(input_dim, vocab_size)

    def logits(self, x: torch.Tensor) -> torch.Tensor:
        """(batch, time, dim) -> (batch, time, vocab) pre-softmax scores."""
        return self.fc(self.dropout(x))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """(batch, time, dim) -> (batch, time, vocab) log-probabilities, as the reference."""
        return torch.log_softmax(self.logits(x), dim=-1)
