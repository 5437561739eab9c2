"""Golden trajectory with FOUR agents in a close train, from the UNMODIFIED reference (run in the build container):

    python tests/golden/make_golden_a4.py   ->  tests/golden/traj_a4_train.npz

base_classes.Simulator with 4 agents 0.6-1.2 m apart nose to tail on example_map: every ego sees up to three opponents
(ray_cast_agents over several opponents in index order, front and rear windows at once), collision_multiple with more
than one contact candidate, check_collision's agent ordering.  Same recorder and file layout as make_golden.run_traj.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402  (loads the reference through oracle/ref_import.py; generates nothing at import)


def train(rng):
    k = int(rng.integers(0, mg.WP.shape[0]))
    poses, back = [], 0
    for i in range(4):
        p = mg.wp_pose(k - back)
        if i:
            p[2] += rng.uniform(-0.25, 0.25)
        poses.append(p)
        back += int(rng.integers(3, 7))
    return np.stack(poses)


if __name__ == '__main__':
    mg.run_traj('traj_a4_train.npz', 4, 4, 80, train, 41, scan_every=5)
