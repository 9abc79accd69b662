"""Evaluation loop around the native decoders (reference ``src/evaluation/evaluate.py``)."""
