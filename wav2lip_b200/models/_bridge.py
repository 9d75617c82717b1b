"""Locates wav2lip_b200._lib whether this package is imported as `wav2lip_b200.models` or, in
drop-in mode, as the top-level `models` package the reference scripts import by bare name
(PYTHONPATH=<repo>/wav2lip_b200)."""
import importlib
import os
import sys


def _load():
    try:
        return importlib.import_module("wav2lip_b200._lib")
    except ModuleNotFoundError:
        root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        if root not in sys.path:
            sys.path.append(root)
        return importlib.import_module("wav2lip_b200._lib")


lib = _load()
