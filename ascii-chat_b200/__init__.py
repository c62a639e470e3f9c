"""ascii-chat_b200 — B200-native RGB -> glyph/ANSI render path behind libasciichat's C entry points.

The product is the C-ABI shared library ``lib/libasciichat_b200.so`` (hand-written sm_100a CUDA,
sources in ``csrc/``, public header ``include/asciichat_b200.h``).  This Python package is only the
thin ctypes mirror of that ABI used by tests, bench.py and the multi-GPU harness; it contains no
rendering logic and no fallback: if the library is missing or no CUDA device is usable, calls raise.

The directory name carries a hyphen (the project name); import it through the top-level shim:
    import ascii_chat_b200 as acb
"""
from .binding import *  # noqa: F401,F403
from .binding import lib, build_library  # noqa: F401
