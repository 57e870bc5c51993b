"""Developer helper (GPU box): steps/s of the drop-in Kalman_Filter class stepped the way the reference driver steps it."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from optistate_b200.synth import make_streams  # noqa: E402

if __name__ == "__main__":
    st = make_streams(range(1), 400)
    for mode in ("copy", "store", "mapped"):
        print(bench.dropin_rate(st, transfer=mode))
    if "--profile" in sys.argv:
        import cProfile
        import pstats

        cProfile.run("bench.dropin_rate(st)", "/tmp/dropin.prof")
        pstats.Stats("/tmp/dropin.prof").sort_stats("tottime").print_stats(12)
