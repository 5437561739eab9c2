"""Wide scan goldens from the UNMODIFIED reference (laser_models.ScanSimulator2D.scan, noise off): 208 poses per
bundled map -- on-track (example_map) and free-space poses, poses anywhere on the map (also inside obstacles), poses hugging the map border
from both sides, and poses far outside / at absurd coordinates (the dt[-1,-1] wrap of laser_models.py:79-81).

Run in the build container only (needs /root/reference):   python tests/golden/make_golden_scans_wide.py
Writes scans_wide_<map>.npz: poses [208,3]; example_map: scan_1080 [208,1080] (+ scan_270 / scan_2160 for the
first 48 poses); the 0.05 m maps: beam_idx [208,270] and scan_1080_sub [208,270] (every 4th beam, offset pose % 4).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, '..', '..'))
from oracle import ref_import  # noqa: E402

ns = ref_import.load()
lm = ns.laser_models


def poses_for(s, rng, waypoints=None):
    H, W = s.dt.shape
    res, ox, oy = s.map_resolution, s.orig_x, s.orig_y
    out = []
    free = np.argwhere(s.dt > 0.3)
    n_free = 96
    if waypoints is not None:                                     # on the track: raceline points with jitter
        n_free = 32
        for k in rng.choice(waypoints.shape[0], 64, replace=False):
            out.append([waypoints[k, 1] + rng.uniform(-0.5, 0.5), waypoints[k, 2] + rng.uniform(-0.5, 0.5),
                        rng.uniform(-7, 7)])
    sel = free[rng.choice(free.shape[0], n_free, replace=False)]
    for r, c in sel:                                             # free space (on / around the track)
        out.append([c * res + ox + rng.uniform(0, res), r * res + oy + rng.uniform(0, res), rng.uniform(-7, 7)])
    for _ in range(48):                                          # anywhere, also inside walls
        out.append([ox + rng.uniform(0, W * res), oy + rng.uniform(0, H * res), rng.uniform(0, 2 * np.pi)])
    for k in range(32):                                          # hugging the border, inside and outside
        t = rng.uniform(0, 1)
        d = rng.uniform(-3, 3) * res
        side = k % 4
        x = ox + (t * W * res if side < 2 else (d if side == 2 else W * res + d))
        y = oy + ((d if side == 0 else H * res + d) if side < 2 else t * H * res)
        out.append([x, y, rng.uniform(0, 2 * np.pi)])
    for k in range(30):                                          # far outside
        out.append([ox + rng.uniform(-200, 200 + W * res), oy + rng.uniform(-200, 200 + H * res), rng.uniform(0, 2 * np.pi)])
    out.append([3.0e8, -1.0, 0.5])                               # absurd coordinates
    out.append([-2.0, -4.0e9, 2.5])
    return np.array(out)


def main():
    rng = np.random.default_rng(20260924)
    for name in ('example_map', 'berlin', 'skirk', 'vegas', 'stata_basement', 'levine'):
        yaml = ns.example_map if name == 'example_map' else os.path.join(ns.maps_dir, name + '.yaml')
        ext = '.pgm' if name == 'levine' else '.png'         # levine ships as .pgm (f110_env.py map_ext)
        s = lm.ScanSimulator2D(1080, 4.7)
        s.set_map(yaml, ext)
        wps = np.loadtxt(ns.example_waypoints, delimiter=';', skiprows=3) if name == 'example_map' else None
        poses = poses_for(s, rng, wps)
        full = np.stack([s.scan(p, None) for p in poses])
        out = {'poses': poses}
        if name == 'example_map':
            out['scan_1080'] = full
            for B in (270, 2160):
                sb = lm.ScanSimulator2D(B, 4.7)
                sb.set_map(yaml, ext)
                out['scan_%d' % B] = np.stack([sb.scan(p, None) for p in poses[:48]])
        else:
            idx = np.stack([np.arange(270) * 4 + (k % 4) for k in range(poses.shape[0])])
            out['beam_idx'] = idx.astype(np.int32)
            out['scan_1080_sub'] = np.take_along_axis(full, idx, axis=1)
        path = os.path.join(HERE, 'scans_wide_%s.npz' % name)
        np.savez_compressed(path, **out)
        print('%-32s %8.1f KB  (max range hits %d, min %.4f)' % (os.path.basename(path), os.path.getsize(path) / 1024,
                                                                  int((full >= 30.0).sum()), full.min()))


if __name__ == '__main__':
    main()
