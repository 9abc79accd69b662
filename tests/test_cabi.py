"""CPU tier: the C-ABI library builds, loads and exports every symbol include/ssd_ctc.h declares
(no compute calls without a GPU), and the product fails loudly instead of falling back."""
import ctypes
import os
import re

import pytest
import torch

from silent_speech_decoder_b200 import _native, ops
from silent_speech_decoder_b200.build import build_native
from silent_speech_decoder_b200.data.vocab import Vocab
from silent_speech_decoder_b200.decoding import ctc
from silent_speech_decoder_b200.models.losses import CTCLoss

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    return _native.Library(build_native())


def test_header_symbols_are_exported(lib):
    header = open(os.path.join(REPO, "include", "ssd_ctc.h")).read()
    declared = set(re.findall(r"\b(ssd_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_native.SIGNATURES), declared ^ set(_native.SIGNATURES)
    raw = ctypes.CDLL(str(lib.path))
    for name in declared:
        assert hasattr(raw, name), name
    assert lib.ssd_abi_version() == _native.ABI_VERSION == int(re.search(r"SSD_ABI_VERSION (\d+)", header).group(1))


def test_argument_validation_needs_no_device(lib):
    # rejected before any CUDA call
    assert lib.ssd_ctc_greedy(None, 1, 4, 0, 0, 0, None, 0, 0.0, None, None, None) == _native.SSD_ERR_BAD_ARG
    assert b"dimension" in lib.ssd_last_error()
    assert lib.ssd_ctc_greedy(None, 1, 4, 4096, 0, 0, None, 0, 0.0, None, None, None) == _native.SSD_ERR_UNSUPPORTED
    assert lib.ssd_ctc_beam(None, 0, 4, 48, 192, 48, None, None, 99, 0, 3, 0.0, 50, -10.0, -5.0, None, None, None, None, 0, None) == _native.SSD_ERR_BAD_ARG
    with pytest.raises(ValueError):
        lib.check(_native.SSD_ERR_BAD_ARG)
    with pytest.raises(_native.NativeError):
        lib.check(_native.SSD_ERR_CUDA)
    assert lib.ssd_edit_counts(None, None, None, None, -1, 0, 0, None, None) == _native.SSD_ERR_BAD_ARG
    assert lib.ssd_edit_counts(None, None, None, None, 4, 10, 10, None, None) == _native.SSD_ERR_BAD_ARG  # NULL offsets
    assert lib.ssd_edit_counts(None, None, None, None, 0, 9000, 1, None, None) == _native.SSD_ERR_UNSUPPORTED
    assert b"SSD_MAX_EDIT_LEN" in lib.ssd_last_error()
    assert lib.ssd_edit_counts(None, None, None, None, 0, 0, 0, None, None) == _native.SSD_OK  # nothing to do
    assert lib.ssd_ctc_loss_workspace_bytes(256, 1000, 48, 200) > 256
    assert lib.ssd_ctc_beam_workspace_bytes(16, 500, 48, 50) > 256


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    v = Vocab.default()
    lp = torch.zeros(1, 4, 48)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ctc.build_decoder("greedy", v)(lp, torch.tensor([4]))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ctc.build_decoder("beam", v)(lp, torch.tensor([4]))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        CTCLoss(blank=1)(lp.transpose(0, 1), torch.tensor([[2, 3]]), torch.tensor([4]), torch.tensor([2]))
    with pytest.raises(FileNotFoundError):  # an LM path is honoured (row f-4): a missing file is an error, not ignored
        ctc.build_decoder("beam", v, lm_path="no_such_file.arpa")


def test_missing_library_is_an_error(monkeypatch, tmp_path):
    monkeypatch.setattr(_native, "_LIB", None)
    monkeypatch.setattr(_native, "LIB_PATH", tmp_path / "libssd_ctc.so")
    with pytest.raises(_native.NativeError, match="has not been built"):
        _native.load()


def test_install_as_src_aliases_reference_names():
    import importlib
    import sys

    import silent_speech_decoder_b200 as pkg

    saved = {k: sys.modules.get(k) for k in list(sys.modules) if k == "src" or k.startswith("src.")}
    try:
        pkg.install_as_src()
        assert importlib.import_module("src.decoding.ctc").build_decoder is ctc.build_decoder
        assert importlib.import_module("src.models.losses").DistillationCTCLoss.__module__.endswith("models.losses")
        assert importlib.import_module("src.data.vocab").Vocab is Vocab
    finally:
        for k in [k for k in sys.modules if k == "src" or k.startswith("src.")]:
            del sys.modules[k]
        sys.modules.update({k: m for k, m in saved.items() if m is not None})
