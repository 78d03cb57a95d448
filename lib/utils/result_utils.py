"""Overlay of the hot-path entries of lib/utils/result_utils.py (postprocess_output and segm_results on the device).  Everything
else the notebooks import from utils.result_utils (empty_results, extend_results) still comes from the reference tree: append the
reference's lib/ AFTER this overlay on sys.path and import those names from there."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from detectorch_b200.utils.result_utils import postprocess_output, segm_results  # noqa: E402,F401
