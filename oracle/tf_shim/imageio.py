"""Empty stand-in for `imageio` (imported by tfrt/distributions.py:17, unused on the hot path)."""
