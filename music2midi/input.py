"""Drop-in alias: music2midi.input of the reference, served by music2midi_b200.input."""
from music2midi_b200.input import *  # noqa: F401,F403
from music2midi_b200 import input as _impl

__all__ = [n for n in dir(_impl) if not n.startswith("_")]
