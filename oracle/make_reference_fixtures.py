"""Writes tests/golden/reference_tf.npz: outputs of the UNMODIFIED reference under TensorFlow on the oracle's tiny synthetic
case (oracle.make_golden.tiny_case), for boxes that have TensorFlow.  Run from the repo root:

    VQB_REFERENCE_DIR=/path/to/reference python -m oracle.make_reference_fixtures

The build container has no TensorFlow, so the file does not exist yet; tests/test_reference_tf.py::test_oracle_matches_
reference_fixture compares the oracle (and, with -m gpu, the CUDA path) against it as soon as it does.  Commit the file
together with the TensorFlow version it prints."""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "tests", "golden", "reference_tf.npz")


def main():
    from oracle import reference_tf as RT
    from oracle.make_golden import tiny_case
    ok, why = RT.available()
    if not ok:
        print("cannot generate reference fixtures:", why)
        return 1
    spec, weights, vq, x = tiny_case()
    out = RT.run_case(spec, weights, vq, x)
    out["x"] = x
    out["tf_version"] = np.array(RT.load().tf.__version__)
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes; TensorFlow", RT.load().tf.__version__)
    return 0


if __name__ == "__main__":
    sys.exit(main())
