"""Recipe for oracle/_ref/: the UNMODIFIED reference modules of the hot path, copied from where they lie under
/root/reference so that the GPU box (which has no /root/reference) can time the reference's own numba path.

    python oracle/make_ref.py            # in the build container; __graft_entry__.build() runs it when the tree exists

TEST INFRASTRUCTURE ONLY.  oracle/_ref/ is git-ignored (no reference source enters the history) but travels to the
GPU box with the gpurun snapshot, like the built .so files.  Only bench.py's reference arm / cpu_baseline leg
(oracle/ref_runner.py) and tests read it; the product never does.
Copied, byte for byte (sha256 recorded in oracle/_ref/MANIFEST.json):
    gym/f110_gym/envs/{dynamic_models,laser_models,collision_models,base_classes}.py
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get('F110_REF_SRC', '/root/reference')
DST = os.path.join(HERE, '_ref')
FILES = ['dynamic_models.py', 'laser_models.py', 'collision_models.py', 'base_classes.py']


def make(verbose=True):
    src_dir = os.path.join(SRC, 'gym', 'f110_gym', 'envs')
    if not os.path.isdir(src_dir):
        if verbose:
            print('make_ref: %s not present; keeping whatever oracle/_ref holds' % src_dir)
        return os.path.isdir(os.path.join(DST, 'gym', 'f110_gym', 'envs'))
    dst_dir = os.path.join(DST, 'gym', 'f110_gym', 'envs')
    os.makedirs(dst_dir, exist_ok=True)
    manifest = {}
    for f in FILES:
        shutil.copyfile(os.path.join(src_dir, f), os.path.join(dst_dir, f))
        with open(os.path.join(dst_dir, f), 'rb') as fh:
            manifest[f] = hashlib.sha256(fh.read()).hexdigest()
    with open(os.path.join(DST, 'MANIFEST.json'), 'w') as fh:
        json.dump({'source': src_dir, 'sha256': manifest}, fh, indent=1)
    with open(os.path.join(DST, 'README.txt'), 'w') as fh:
        fh.write('oracle/_ref/ holds UNMODIFIED modules of the reference (f1tenth/f1tenth_gym), copied byte for byte by\n'
                 'oracle/make_ref.py when __graft_entry__.build() runs in the build container (VERDICT r1, next-round item 6).\n'
                 'The directory is git-ignored: nothing in it is part of this repository or of the product.  It exists so that\n'
                 'bench.py --impl reference / cpu_baseline can time the reference\'s own numba path on the GPU box, which has no\n'
                 '/root/reference.  sha256 of every file: MANIFEST.json.\n')
    if verbose:
        print('make_ref: %d reference modules -> %s' % (len(FILES), dst_dir))
    return True


if __name__ == '__main__':
    sys.exit(0 if make() else 1)
