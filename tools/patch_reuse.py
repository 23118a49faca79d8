#!/usr/bin/env python
"""CLI of tensorflowraytrace_b200/_reuse_patch.py (tuning experiments on tools/k1_bench binaries)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tensorflowraytrace_b200 import _reuse_patch  # noqa: E402

_reuse_patch.main()
