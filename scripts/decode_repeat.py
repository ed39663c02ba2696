"""Decode timing studies in a fresh process (allocation layouts differ from process to process):
   python scripts/decode_repeat.py BEAMS        -> time_decode(32, BEAMS) without the general route
   python scripts/decode_repeat.py BEAMS skip   -> in-graph cost of each kernel kind (step timed with one kind left out)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scripts import gpu_check_decode as K

beams = int(sys.argv[1]) if len(sys.argv) > 1 else 5
if len(sys.argv) > 2 and sys.argv[2] == "skip":
    K.time_decode_skip(32, beams)
else:
    K.time_decode(32, beams, general=False)
