"""Regenerates tests/golden/*.npy from the CPU oracle (python tests/golden/make_golden.py).

The reference (C#/.NET 9) cannot run in this image, so these fixtures freeze the ORACLE's output — they guard the
oracle against drift; they are not outputs of the reference binary (see DESIGN.md "parity unpinned").
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests import test_oracle_kats as T  # noqa: E402

here = os.path.dirname(os.path.abspath(__file__))
for name, fn in (("c2_small", T._golden_c2), ("c3_small", T._golden_c3), ("f3_small", T._golden_f3)):
    y = fn()
    np.save(os.path.join(here, name + ".npy"), y)
    print(name, y.shape, float(np.abs(y).max()))
