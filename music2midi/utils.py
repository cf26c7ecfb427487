"""Drop-in alias: music2midi.utils of the reference, served by music2midi_b200.utils."""
from music2midi_b200.utils import *  # noqa: F401,F403
from music2midi_b200 import utils as _impl

__all__ = [n for n in dir(_impl) if not n.startswith("_")]
