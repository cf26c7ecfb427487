"""Drop-in alias: music2midi.tokenizer of the reference, served by music2midi_b200.tokenizer."""
from music2midi_b200.tokenizer import *  # noqa: F401,F403
from music2midi_b200 import tokenizer as _impl

__all__ = [n for n in dir(_impl) if not n.startswith("_")]
