"""Host-side transcript assembly (decoding/ctc.py::_rows_to_strings) against the reference's token-by-token
rendering: Vocab.decode (src/data/vocab.py:44-53) for greedy, the pyctcdecode label strings + whitespace
normalisation (src/decoding/ctc.py:64-65,97 and the upstream ``" ".join(text.split())``) for beam."""
import numpy as np
import torch

from silent_speech_decoder_b200.data.vocab import Vocab
from silent_speech_decoder_b200.decoding import ctc


def _cases():
    rng = np.random.default_rng(3)
    for B, T, hi in ((1, 6, 6), (9, 40, 40), (300, 120, 90), (4, 10, 0)):
        n = rng.integers(0, hi + 1, B).astype(np.int32)
        tok = torch.from_numpy(rng.integers(0, 48, (B, T)).astype(np.int32))
        yield tok, torch.from_numpy(n)


def test_rows_to_strings_matches_token_by_token_rendering():
    v = Vocab.default()
    text, _ = ctc._beam_label_strings(v)
    greedy_text = ["" if i in (v.blank_id, v.pad_id) else v.tokens[i] for i in range(v.size)]
    for tok, cnt in _cases():
        rows = ctc._rows_to_lists(tok, cnt)
        want_beam = [" ".join("".join(text[i] for i in seq).split()) for seq in rows]
        want_greedy = [v.decode(seq) for seq in rows]
        assert ctc._rows_to_strings(tok, cnt, text, normalise_space=True) == want_beam
        assert ctc._rows_to_strings(tok, cnt, greedy_text, normalise_space=False) == want_greedy


def test_rows_to_strings_unk_and_spaces():
    v = Vocab.default()
    sp = v.token_to_id[" "]
    a = v.token_to_id["a"]
    tok = torch.tensor([[sp, a, sp, sp, v.unk_id, sp], [v.unk_id, v.unk_id, a, 0, 0, 0]], dtype=torch.int32)
    cnt = torch.tensor([6, 3], dtype=torch.int32)
    text, _ = ctc._beam_label_strings(v)
    assert ctc._rows_to_strings(tok, cnt, text, True) == ["a " + ctc.UNK_GLYPH, ctc.UNK_GLYPH * 2 + "a"]
    greedy_text = ["" if i in (v.blank_id, v.pad_id) else v.tokens[i] for i in range(v.size)]
    assert ctc._rows_to_strings(tok, cnt, greedy_text, False) == [" a  <unk> ", "<unk><unk>a"]
