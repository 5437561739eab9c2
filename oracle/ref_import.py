"""Import the UNMODIFIED reference (f1tenth_gym) kernels from /root/reference by path.

TEST INFRASTRUCTURE ONLY.  Used in the build container to (a) validate the C
restatement in oracle/f110_oracle.c against the reference's own numba path and
(b) generate the committed golden fixtures under tests/golden/.  /root/reference
does not exist on the GPU box, so nothing executed there may import this module.

The package __init__ files of the reference import `gym`/`pyglet` (absent here);
we pre-register empty namespace modules so the kernel modules
(dynamic_models / laser_models / collision_models / base_classes) import unmodified.
"""
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))


def _default_root():
    """The reference tree when it is mounted (build container), else the copy oracle/make_ref.py left in
    oracle/_ref/ (git-ignored; travels to the GPU box with the snapshot)."""
    if "F110_REF" in os.environ:
        return os.environ["F110_REF"]
    if os.path.isdir("/root/reference/gym/f110_gym/envs"):
        return "/root/reference"
    return os.path.join(_HERE, "_ref")


REF_ROOT = _default_root()


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "gym", "f110_gym", "envs"))


def load():
    """Returns a namespace with Simulator, Integrator, RaceCar and the three kernel modules."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/numba_cache_f110")
    pkg = os.path.join(REF_ROOT, "gym", "f110_gym")
    for name, path in (("f110_gym", pkg), ("f110_gym.envs", os.path.join(pkg, "envs"))):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__path__ = [path]
            sys.modules[name] = m
    import warnings
    warnings.filterwarnings("ignore")
    from f110_gym.envs import base_classes, dynamic_models, laser_models, collision_models
    ns = types.SimpleNamespace(
        base_classes=base_classes, dynamic_models=dynamic_models,
        laser_models=laser_models, collision_models=collision_models,
        Simulator=base_classes.Simulator, Integrator=base_classes.Integrator,
        RaceCar=base_classes.RaceCar,
        example_map=os.path.join(REF_ROOT, "examples", "example_map.yaml"),
        example_waypoints=os.path.join(REF_ROOT, "examples", "example_waypoints.csv"),
        maps_dir=os.path.join(pkg, "envs", "maps"),
    )
    return ns


def new_simulator(ns, params, num_agents, map_yaml, map_ext=".png", seed=12345, timestep=0.01,
                  integrator=None, lidar_dist=0.0, num_beams=1080, fov=4.7):
    """Build a reference Simulator with a fresh scan-simulator singleton (RaceCar.scan_simulator is
    a process-wide class attribute, base_classes.py:64,118)."""
    ns.RaceCar.scan_simulator = None
    if integrator is None:
        integrator = ns.Integrator.RK4
    if (num_beams, fov) != (1080, 4.7):
        # num_beams/fov are not plumbed through Simulator (base_classes.py:493-496)
        raise ValueError("reference Simulator only supports 1080 beams / fov 4.7")
    sim = ns.Simulator(params, num_agents, seed, time_step=timestep, integrator=integrator,
                       lidar_dist=lidar_dist)
    sim.set_map(map_yaml, map_ext)
    return sim


def reset_noise_off(sim, poses):
    sim.reset(poses)
    for a in sim.agents:
        a.scan_rng = None   # ScanSimulator2D.scan accepts rng=None (laser_models.py:450)
