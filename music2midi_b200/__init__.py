"""music2midi_b200 — B200-native (sm_100a) implementation of Music2MIDI's inference hot path:
log-mel frontend -> T5 encoder -> KV-cached greedy decode -> tokens -> notes/MIDI.

Same Python API as the reference package (``music2midi.input / transformer / model / tokenizer /
utils``; the top-level ``music2midi`` package of this repository re-exports these modules), with the
arithmetic in hand-written CUDA behind the C ABI of ``libm2m_b200.so`` (include/m2m_b200.h).
Importing the package does not load the shared library; using the hot path does, and fails loudly if
it is missing (build it with ``python -m music2midi_b200.build``).
"""
__version__ = "0.1.0"
