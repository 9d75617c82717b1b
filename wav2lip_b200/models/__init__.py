"""Drop-in for the reference's `models` package (/root/reference/models/__init__.py:1-2)."""
from .wav2lip import Wav2Lip, Wav2Lip_disc_qual
from .syncnet import SyncNet_color

__all__ = ["Wav2Lip", "Wav2Lip_disc_qual", "SyncNet_color"]
