"""TEST INFRASTRUCTURE ONLY - stages the UNMODIFIED reference files the checker needs on the GPU box.

    python -m oracle.make_ref            (also run by __graft_entry__.build() when /root/reference is present)

/root/reference exists only in the build container.  The GPU box gets a snapshot of this repository, and the
north-star asks for "the reference CPU filter timed on the same box's host cores in the same run" (bench.py
`cpu_baseline.reference_class_*`) and tests/test_reference_driver_gpu.py runs the reference's own conversion driver
against the drop-in class.  Both need the reference's Python files at run time, so they are copied VERBATIM, keeping
their directory layout, into oracle/_ref/ - which is git-ignored (never part of the history, never edited) but not
gpurun-ignored, so it travels like the built .so files.  Nothing under oracle/_ref is imported by the product package.
"""
from __future__ import annotations

import filecmp
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.environ.get("OPTISTATE_REF_SRC", "/root/reference")
REF_DST = os.path.join(HERE, "_ref")
FILES = [
    "kalman_filter/kalman_filter.py",                           # the filter class (hot path)
    "misc/force_controller.py",                                 # next_state + StanceController set-up
    "settings.py",                                              # INITIAL_PARAMS
    "data_collection/data_conversion_Kalman_to_Training.py",    # the KF driver
    "data_collection/trajectories/Q_R.pkl",                     # the only data fixture the reference ships
    "gru/gru_model.py",                                         # consumer of config 5
    "LICENSE",
]


def stage(verbose: bool = False) -> str | None:
    """Copies the files when the reference tree is present; returns the staged root (or None if nothing is available)."""
    if os.path.isdir(REF_SRC):
        for rel in FILES:
            src, dst = os.path.join(REF_SRC, rel), os.path.join(REF_DST, rel)
            if not os.path.isfile(src):
                continue
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            if not (os.path.isfile(dst) and filecmp.cmp(src, dst, shallow=False)):
                shutil.copyfile(src, dst)
                if verbose:
                    print("staged", rel)
    return REF_DST if os.path.isfile(os.path.join(REF_DST, FILES[0])) else None


def root() -> str | None:
    """Where the unmodified reference can be imported from: the live tree in the build container, else the staged copy."""
    if os.path.isfile(os.path.join(REF_SRC, FILES[0])):
        return REF_SRC
    return REF_DST if os.path.isfile(os.path.join(REF_DST, FILES[0])) else None


if __name__ == "__main__":
    print(stage(verbose=True))
