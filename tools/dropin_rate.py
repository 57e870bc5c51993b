"""Developer helper (GPU box): steps/s of the drop-in Kalman_Filter class stepped the way the reference driver steps it."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from optistate_b200.synth import make_streams  # noqa: E402

if __name__ == "__main__":
    print(bench.dropin_rate(make_streams(range(1), 400)))
