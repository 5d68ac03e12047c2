"""smilei_b200 — B200-native PIC time-step hot path behind Smilei's operator surface.

Only what the hot path needs lives here: csrc/ (hand-written sm_100a CUDA kernels + the
C ABI of include/smilei_b200.h) and the host-side mirror of the reference's operator /
namelist interface.  Importing the package does not load the CUDA library; the first use
does, and fails loudly if it is missing (there is no CPU fallback).
"""
from .capi import Patch, SmileiB200Error, FIELDS, PUSHERS, device_count, LIB_PATH  # noqa: F401

__version__ = "0.1.0"
