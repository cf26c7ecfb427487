"""Drop-in alias: music2midi.transformer of the reference, served by music2midi_b200.transformer."""
from music2midi_b200.transformer import *  # noqa: F401,F403
from music2midi_b200 import transformer as _impl

__all__ = [n for n in dir(_impl) if not n.startswith("_")]
