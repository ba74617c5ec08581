"""Generates tests/golden/*.npz by running the reference's own sources (oracle/_ref, built by
oracle/build_ref.sh from /root/reference) on the small parity cases of tests/util_parity.py.

    python tests/golden/make_golden.py            # regenerate every fixture
    python tests/golden/make_golden.py NAME ...   # only the named cases

Each fixture holds, per zone: final values[9] per node, node flags, border faces, the local
bases of the last step, and for step `tap_step` the owner tetrahedron of every characteristic
foot (last alpha try) and the number of alpha tries.  The fixtures are what the GPU tests fall
back to when oracle/_ref did not travel, and what pins the oracle restatements.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import util_parity as up  # noqa: E402


def main():
    for name in (sys.argv[1:] or up.CASE_NAMES):
        case = up.make_case(name)
        res = up.run_oracle(case, tap=True)
        d = dict(n_zones=len(res))
        for k, r in enumerate(res):
            for w in ("values", "flags", "faces", "bases", "owners", "tries"):
                d["%s%d" % (w, k)] = r[w]
            if name in up.DESTR_CASES:      # ElasticNode::destruction_criterias after the last step
                d["destr%d" % k] = r["destr"]
        np.savez_compressed(up.golden_path(name), **d)
        print(name, [int(up.local_mask(r["flags"]).sum()) for r in res], "local nodes;",
              os.path.getsize(up.golden_path(name)), "bytes")


if __name__ == "__main__":
    main()
