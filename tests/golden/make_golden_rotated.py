"""Golden scans on a map whose yaml origin has a yaw (xy_2_rc rotation terms, laser_models.py:75-78): the reference
ScanSimulator2D on examples/example_map.png with origin (ox + 3, oy - 2, 0.35).  Container only (needs /root/reference).
   python tests/golden/make_golden_rotated.py"""
import os
import shutil
import sys
import tempfile

import numpy as np
import yaml

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, '..', '..'))
from oracle import ref_import  # noqa: E402

ns = ref_import.load()
ex = os.path.join(ref_import.REF_ROOT, 'examples')
d = tempfile.mkdtemp()
shutil.copy(os.path.join(ex, 'example_map.png'), d)
y = yaml.safe_load(open(os.path.join(ex, 'example_map.yaml')))
ox0, oy0 = y['origin'][0], y['origin'][1]
th = 0.35
y['origin'] = [ox0 + 3.0, oy0 - 2.0, th]
yaml.safe_dump(y, open(os.path.join(d, 'example_map.yaml'), 'w'))
s = ns.laser_models.ScanSimulator2D(1080, 4.7)
s.set_map(os.path.join(d, 'example_map.yaml'), '.png')
WP = np.loadtxt(ns.example_waypoints, delimiter=';', skiprows=3)
c, sn = np.cos(th), np.sin(th)
rng = np.random.default_rng(5)
poses = []
for k in (0, 97, 211, 333, 480, 555, 690, 760):            # on-track poses of the unrotated map, moved into the new frame
    mx, my = WP[k, 1] - ox0, WP[k, 2] - oy0
    poses.append([c * mx - sn * my + y['origin'][0], sn * mx + c * my + y['origin'][1],
                  WP[k, 3] + np.pi / 2 + th + rng.uniform(-0.5, 0.5)])
poses = np.array(poses + [[500.0, 0.0, 1.0]])                 # and one pose off the map
scans = np.stack([s.scan(p, None) for p in poses])
np.savez_compressed(os.path.join(HERE, 'scans_rotated_origin.npz'), poses=poses, scan_1080=scans,
                    origin=np.array(y['origin']), resolution=y['resolution'])
print('scans_rotated_origin.npz', scans.shape, float((scans < 29).mean()))
