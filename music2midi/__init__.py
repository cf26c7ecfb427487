"""Import-compatible stand-in for the reference's ``music2midi`` package (hot path only):
``from music2midi.model import Music2MIDI`` etc. resolve to the B200-native implementation in
``music2midi_b200``.  Modules of the reference outside the inference path (dataset, evaluation,
plot_midi, webui_utils) are intentionally not provided."""
