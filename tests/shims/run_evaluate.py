"""Launcher of the reference's UNMODIFIED projects/tools/evaluate.py (TEST INFRASTRUCTURE):  python run_evaluate.py <path/to/evaluate.py> [its flags]
Run with cwd = a directory whose `projects/` resolves to this repository's drop-in package (tests/test_evaluate_drive.py builds it from symlinks)
and with tests/shims on PYTHONPATH.  Without a CUDA device the GPU engine and pixel decoders are replaced by the shape-correct fakes of cpu_stubs."""
import json
import os
import runpy
import sys

import torch

script = sys.argv[1]
sys.argv = [script] + sys.argv[2:]
import cv2
try:                                    # headless OpenCV builds have no highgui: the visualiser's clean-up call (visulize.py:74) must not abort the run
    cv2.destroyAllWindows()
except cv2.error:
    cv2.destroyAllWindows = lambda: None
stubbed = not torch.cuda.is_available()
if stubbed:
    import cpu_stubs
    cpu_stubs.install()
runpy.run_path(script, run_name="__main__")
if stubbed:
    print("STUB_CALLS " + json.dumps(cpu_stubs.CALLS))
