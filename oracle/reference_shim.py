"""TEST INFRASTRUCTURE ONLY — imports the *real* reference classes from /root/reference.

The reference (ytinyui/music2midi) is pure Python; its hot-path arithmetic lives in
torchaudio (`MelSpectrogram`) and HF transformers (`T5ForConditionalGeneration`).  Two
small shims are needed to import it in this image:

* ``omegaconf`` is absent -> a stand-in whose ``OmegaConf.load`` returns an attribute
  dict over ``yaml.safe_load`` (the reference only ever reads the config).
* ``np.float_`` was removed in numpy 2 (used at music2midi/tokenizer.py:200).

``music2midi.model`` / ``music2midi.utils`` cannot be imported (pytorch_lightning, librosa,
pretty_midi, more_itertools, mir_eval are absent); ``sample_tokens`` is restated in
oracle/hf_path.py instead.

/root/reference exists only in the authoring container.  Nothing that runs on the GPU box
(`-m gpu` tests, smoke(), bench.py) may import this module; it is used by
tests/golden/make_golden.py and by the CPU tests that re-validate the oracle port against
the live reference when the reference tree is present.
"""
from __future__ import annotations

import os
import sys
import types

REFERENCE_ROOT = os.environ.get("M2M_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "music2midi", "transformer.py"))


class _AttrDict(dict):
    """dict with attribute access, recursive; enough of DictConfig for the reference."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:  # pragma: no cover
            raise AttributeError(k) from e


def _wrap(o):
    if isinstance(o, dict):
        return _AttrDict({k: _wrap(v) for k, v in o.items()})
    if isinstance(o, list):
        return [_wrap(v) for v in o]
    return o


def install_shims() -> None:
    import numpy as np
    import yaml

    if "omegaconf" not in sys.modules:
        m = types.ModuleType("omegaconf")

        class OmegaConf:  # noqa: D401 - stand-in
            @staticmethod
            def load(path):
                with open(path) as f:
                    return _wrap(yaml.safe_load(f))

        m.OmegaConf = OmegaConf
        m.DictConfig = _AttrDict
        sys.modules["omegaconf"] = m
    if not hasattr(np, "float_"):
        np.float_ = np.float64


def import_reference():
    """Returns (input_module, transformer_module, tokenizer_module) of the real reference."""
    if not available():
        raise RuntimeError(f"reference tree not present at {REFERENCE_ROOT}")
    install_shims()
    import importlib.util

    # Load the three hot-path modules under a private package name so that the
    # reference's `music2midi` never shadows (or is shadowed by) this repo's drop-in
    # `music2midi` alias package.
    pkg_name = "_m2m_reference"
    if pkg_name not in sys.modules:
        pkg = types.ModuleType(pkg_name)
        pkg.__path__ = [os.path.join(REFERENCE_ROOT, "music2midi")]
        sys.modules[pkg_name] = pkg
    mods = []
    for name in ("input", "tokenizer", "transformer"):
        full = f"{pkg_name}.{name}"
        if full not in sys.modules:
            spec = importlib.util.spec_from_file_location(
                full, os.path.join(REFERENCE_ROOT, "music2midi", f"{name}.py")
            )
            mod = importlib.util.module_from_spec(spec)
            sys.modules[full] = mod
            spec.loader.exec_module(mod)
        mods.append(sys.modules[full])
    return mods[0], mods[2], mods[1]


def reference_config_path() -> str:
    return os.path.join(REFERENCE_ROOT, "config.yaml")
