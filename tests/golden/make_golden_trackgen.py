"""Golden centerlines of the reference's random track generator (unittest/random_trackgen.py:56-165).

Run in the build container only (needs /root/reference):   python tests/golden/make_golden_trackgen.py
The reference module needs cv2 / shapely / matplotlib (absent here) only for the wall offsetting and the
rendering AFTER the centerline is finished; they are replaced by inert stand-ins so that the UNMODIFIED
`create_track()` runs up to and including its closed-loop search and gluing test.  `shp.Polygon(track_xy)` is
the hand-over point: the stand-in records the centerline it is given.  For every seed the module is executed
afresh (its module-level `np.random.seed(args.seed)` is the seeding the reference does) and `create_track()`
is called `CALLS` times in a row, exactly as the reference's main loop does; failures (`False`) are recorded
as empty centerlines so that the RNG consumption across retries is pinned as well.
"""
import io
import os
import sys
import tempfile
import types
import contextlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(os.environ.get('F110_REF', '/root/reference'), 'gym', 'f110_gym', 'unittest', 'random_trackgen.py')
SEEDS = (123, 1, 7, 2024)
CALLS = 4


class _Poly:
    def __init__(self, xy):
        self.xy = np.asarray(xy)
        self.exterior = self.xy

    def buffer(self, _w):
        return self


def _stubs():
    shp_pkg = types.ModuleType('shapely')
    shp = types.ModuleType('shapely.geometry')
    shp.Polygon = _Poly
    shp_pkg.geometry = shp
    mpl = types.ModuleType('matplotlib')
    plt = types.ModuleType('matplotlib.pyplot')
    patches = types.ModuleType('matplotlib.patches')
    patches.Polygon = object
    coll = types.ModuleType('matplotlib.collections')
    coll.PatchCollection = object
    mpl.pyplot, mpl.patches, mpl.collections = plt, patches, coll
    cv2 = types.ModuleType('cv2')
    return {'shapely': shp_pkg, 'shapely.geometry': shp, 'matplotlib': mpl, 'matplotlib.pyplot': plt,
            'matplotlib.patches': patches, 'matplotlib.collections': coll, 'cv2': cv2}


def run_reference(seed, calls):
    saved = {k: sys.modules.get(k) for k in _stubs()}
    sys.modules.update(_stubs())
    argv, cwd = sys.argv, os.getcwd()
    out = []
    with tempfile.TemporaryDirectory() as tmp:
        try:
            os.chdir(tmp)                                   # the module creates maps/ and centerline/ in the cwd
            sys.argv = ['random_trackgen.py', '--seed', str(seed)]
            glb = {'__name__': 'random_trackgen_ref'}
            with contextlib.redirect_stdout(io.StringIO()):
                exec(compile(open(REF).read(), REF, 'exec'), glb)        # module body: argparse + np.random.seed
                for _ in range(calls):
                    try:
                        r = glb['create_track']()
                    except AssertionError:
                        r = False
                    out.append(np.zeros((0, 2)) if r is False else np.asarray(r[0], dtype=np.float64))
        finally:
            os.chdir(cwd)
            sys.argv = argv
            for k, v in saved.items():
                if v is None:
                    sys.modules.pop(k, None)
                else:
                    sys.modules[k] = v
    return out


if __name__ == '__main__':
    arrays = {'seeds': np.array(SEEDS), 'calls': np.array(CALLS)}
    for s in SEEDS:
        for c, xy in enumerate(run_reference(s, CALLS)):
            arrays['seed%d_call%d' % (s, c)] = xy
            print('seed %5d call %d: %4d centerline points' % (s, c, xy.shape[0]))
    path = os.path.join(HERE, 'trackgen_centerlines.npz')
    np.savez_compressed(path, **arrays)
    print('%s %.1f KB' % (path, os.path.getsize(path) / 1024))
