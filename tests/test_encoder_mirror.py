"""The PyTorch-side mirror of the reference's encoder and CTC head (used by the config-5 pipeline) against
outputs of the reference modules themselves (tests/golden/encoder_ref.npz, oracle/make_golden_r2.py)."""
import os

import numpy as np
import torch

from silent_speech_decoder_b200.models.emg_encoder import MPS_FAST_PLUS, EMGConformerEncoder, EncoderConfig
from silent_speech_decoder_b200.models.heads import CTCHead


def test_small_encoder_reproduces_reference(golden_dir):
    from oracle.make_golden_r2 import SMALL_ENCODER

    z = np.load(os.path.join(golden_dir, "encoder_ref.npz"))
    torch.manual_seed(3201)  # same seed, same parameter creation order -> the reference's weights
    enc = EMGConformerEncoder(EncoderConfig(**SMALL_ENCODER)).eval()
    head = CTCHead(SMALL_ENCODER["d_model"], 48, dropout=0.1).eval()
    with torch.no_grad():
        y, n = enc(torch.from_numpy(z["x"]), torch.from_numpy(z["lengths"]))
        lp = head(y)
        logits = head.logits(y)
    assert n.tolist() == z["out_lengths"].tolist()
    np.testing.assert_allclose(y.numpy(), z["y"], rtol=0, atol=1e-5)
    np.testing.assert_allclose(lp.numpy(), z["lp"], rtol=0, atol=1e-5)
    np.testing.assert_allclose(torch.log_softmax(logits, -1).numpy(), z["lp"], rtol=0, atol=1e-5)


def test_plus_model_has_the_reference_parameters(golden_dir):
    z = np.load(os.path.join(golden_dir, "encoder_ref.npz"))
    sd = EMGConformerEncoder(EncoderConfig(**MPS_FAST_PLUS)).state_dict()
    assert list(sd.keys()) == list(z["plus_keys"])  # a reference checkpoint loads unchanged
    assert [str(tuple(v.shape)) for v in sd.values()] == list(z["plus_shapes"])
