"""Times the UNMODIFIED reference (f1tenth_gym numba path: base_classes.Simulator.step) on the host cores.

TEST / MEASUREMENT INFRASTRUCTURE ONLY (bench.py `--impl reference` and the `cpu_baseline` leg).  The reference is
single-threaded (`@njit`, no parallel=True), so SURVEY.md 8(d)'s policy applies: ONE OS PROCESS PER USABLE HOST CORE,
each owning its own reference Simulators; benchmark policy = example_map, default vehicle parameters, RK4, dt 0.01,
scan noise off (`agent.scan_rng = None` after every reset), start pose = raceline waypoint k ~ U (second agent 23
waypoints behind), actions steer ~ U[-0.4189, 0.4189], speed ~ U[0, 8] i.i.d. per tick, env reset to a fresh pose
when the ego's collision flag fires.

The modules come from /root/reference when it is mounted, else from oracle/_ref/ (oracle/make_ref.py).
A "step" of the bench contract = `ticks_per_step` ticks of every process's `envs_per_proc` Simulators, i.e. a bounded
sample of the workload's env batch advanced in lockstep between two barriers.
"""
import multiprocessing as mp
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)

PARAMS = {'mu': 1.0489, 'C_Sf': 4.718, 'C_Sr': 5.4562, 'lf': 0.15875, 'lr': 0.17145, 'h': 0.074, 'm': 3.74,
          'I': 0.04712, 's_min': -0.4189, 's_max': 0.4189, 'sv_min': -3.2, 'sv_max': 3.2, 'v_switch': 7.319,
          'a_max': 9.51, 'v_min': -5.0, 'v_max': 20.0, 'width': 0.31, 'length': 0.58}
MAP_YAML = os.path.join(ROOT, 'f1tenth_gym_b200', 'maps', 'example_map.yaml')
WAYPOINTS = os.path.join(ROOT, 'f1tenth_gym_b200', 'maps', 'example_waypoints.csv')
POSE_GAP = 23


def available():
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    from oracle import ref_import
    return ref_import.available()


def usable_cpus():
    """Hardware threads this process may actually use: affinity mask capped by the cgroup CPU quota."""
    n = os.cpu_count() or 1
    if hasattr(os, 'sched_getaffinity'):
        n = min(n, len(os.sched_getaffinity(0)))
    try:
        with open('/sys/fs/cgroup/cpu.max') as f:
            quota, period = f.read().split()[:2]
        if quota != 'max':
            n = max(1, min(n, int(-(-int(quota) // int(period)))))
    except Exception:
        pass
    return n


def _load(cache_dir):
    os.environ['NUMBA_CACHE_DIR'] = cache_dir
    os.environ.setdefault('NUMBA_NUM_THREADS', '1')
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    from oracle import ref_import
    return ref_import, ref_import.load()


class _Envs(object):
    """`count` independent reference Simulators with the benchmark policy."""

    def __init__(self, ref_import, ns, count, num_agents, seed):
        wp = np.loadtxt(WAYPOINTS, delimiter=';', skiprows=3)
        self.wp = np.stack([wp[:, 1], wp[:, 2], wp[:, 3] + np.pi / 2], axis=1)
        self.rng = np.random.default_rng(seed)
        self.A = num_agents
        self.ref_import = ref_import
        self.sims = []
        for i in range(count):
            sim = ref_import.new_simulator(ns, PARAMS, num_agents, MAP_YAML) if i == 0 else \
                ns.Simulator(PARAMS, num_agents, 12345, time_step=0.01, integrator=ns.Integrator.RK4)
            if i > 0:
                sim.set_map(MAP_YAML, '.png')      # RaceCar.scan_simulator is a class-level singleton: the map is shared
            self.sims.append(sim)
            self._reset(sim)

    def _reset(self, sim):
        k = int(self.rng.integers(0, self.wp.shape[0]))
        poses = np.stack([self.wp[(k - POSE_GAP * i) % self.wp.shape[0]] for i in range(self.A)])
        self.ref_import.reset_noise_off(sim, poses)

    def run(self, ticks):
        A = self.A
        for sim in self.sims:
            for _ in range(ticks):
                act = np.stack([self.rng.uniform(-0.4189, 0.4189, A), self.rng.uniform(0.0, 8.0, A)], axis=1)
                obs = sim.step(act)
                if obs['collisions'][0]:
                    self._reset(sim)
        return len(self.sims) * ticks * A


def _worker(rank, num_agents, envs_per_proc, ticks_per_step, steps, cache_dir, barrier, seed):
    ref_import, ns = _load(cache_dir)
    envs = _Envs(ref_import, ns, envs_per_proc, num_agents, seed + rank)
    envs.run(5)                                 # JIT (from the warm cache) + first-call work, untimed
    barrier.wait()
    for _ in range(steps):
        barrier.wait()
        envs.run(ticks_per_step)
        barrier.wait()


def calibrate(num_agents, cache_dir, ticks=60):
    """Runs in the PARENT: fills the numba on-disk cache (one cold JIT, ~20-60 s) and returns env-ticks/s of one
    process, so that the step size can be chosen before the workers start."""
    ref_import, ns = _load(cache_dir)
    envs = _Envs(ref_import, ns, 2, num_agents, 999)
    envs.run(5)
    t0 = time.perf_counter()
    envs.run(ticks)
    return 2 * ticks / (time.perf_counter() - t0)


def run(num_agents, steps, warmup, procs=0, target_step_s=None, cache_dir=None, seed=12345):
    """-> dict(value agent-steps/s over the timed steps, ms_per_step, procs, envs_per_proc, ticks_per_step, ...)."""
    cache_dir = cache_dir or os.path.join('/tmp', 'numba_cache_f110_%d' % os.getuid())
    os.makedirs(cache_dir, exist_ok=True)
    procs = procs or usable_cpus()
    rate1 = calibrate(num_agents, cache_dir)                      # env-ticks/s, one process, cache now warm
    if target_step_s is None:                                      # whole run (K + W steps) within ~40-60 s
        target_step_s = min(1.0, max(0.1, 40.0 / max(steps + warmup, 1)))
    envs_per_proc = 4
    ticks_per_step = max(1, int(round(target_step_s * rate1 / envs_per_proc)))
    ctx = mp.get_context('spawn')
    barrier = ctx.Barrier(procs + 1)
    total = steps + warmup
    ws = [ctx.Process(target=_worker, args=(r, num_agents, envs_per_proc, ticks_per_step, total, cache_dir, barrier, seed),
                      daemon=True) for r in range(procs)]
    for w in ws:
        w.start()
    barrier.wait(timeout=900)                                      # all workers built + warmed
    times = []
    for s in range(total):
        barrier.wait(timeout=900)
        t0 = time.perf_counter()
        barrier.wait(timeout=900)
        times.append(time.perf_counter() - t0)
    for w in ws:
        w.join(timeout=60)
    timed = np.array(times[warmup:])
    agent_steps_per_step = procs * envs_per_proc * ticks_per_step * num_agents
    return {'value': agent_steps_per_step * len(timed) / float(timed.sum()), 'ms_per_step': 1e3 * float(timed.mean()),
            'seconds': float(timed.sum()), 'procs': procs, 'envs_per_proc': envs_per_proc,
            'ticks_per_step': ticks_per_step, 'agent_steps_per_step': agent_steps_per_step,
            'one_process_env_ticks_per_s': rate1, 'host_cpus': os.cpu_count()}


if __name__ == '__main__':
    import json
    A = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    print(json.dumps(run(A, steps=5, warmup=2)))
