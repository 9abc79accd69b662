"""CPU oracle for the CTC hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import this package.  The product
package (``silent-speech-decoder_b200/``) never does: it fails loudly when the
CUDA extension is missing instead of falling back to anything in here.

Pinning status:

* ``ctc_loss``  -- pinned against the reference's own ``DistillationCTCLoss``
  (``src/models/losses.py``) run in this container -> ``tests/golden/loss_*.npz``.
* ``greedy``    -- pinned against the reference's own ``_greedy_decode``
  (``src/decoding/ctc.py:16-36``) and its known-answer test
  (``tests/test_evaluation.py:6-14``) -> ``tests/golden/greedy_*.npz``.
* ``beam``      -- PARITY UNPINNED: the search lives in third-party
  ``pyctcdecode`` (un-pinned in ``environment.yml:59``; 0.5.0 is the latest
  release), which is neither vendored in the reference nor installable here
  (no network).  ``beam.py`` restates its published algorithm from memory and
  follows the reference's own call site ``src/decoding/ctc.py:47-99`` for the
  wrapper.  The reference holds no test or golden vector for beam search.
* ``lm``        -- PARITY UNPINNED: ``kenlm`` back-off scoring and pyctcdecode's ``LanguageModel`` /
  ``_get_lm_beams`` restated from memory (scope row f-4); checked only against hand-computed
  answers on a three-gram file.  ``make_beam_golden_upstream.py`` writes upstream LM fixtures
  where both libraries exist.
"""
