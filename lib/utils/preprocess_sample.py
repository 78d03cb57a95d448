"""Overlay of lib/utils/preprocess_sample.py: the test-time image path on the device (see detectorch_b200/utils/preprocess_sample.py)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from detectorch_b200.utils.preprocess_sample import preprocess_sample  # noqa: E402,F401
