"""Drop-in alias: music2midi.model of the reference, served by music2midi_b200.model."""
from music2midi_b200.model import *  # noqa: F401,F403
from music2midi_b200 import model as _impl

__all__ = [n for n in dir(_impl) if not n.startswith("_")]
