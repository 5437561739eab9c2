"""Generate the committed golden fixtures by running the UNMODIFIED reference (numba path).

Run in the build container only (needs /root/reference):   python tests/golden/make_golden.py
Outputs small .npz files next to this script.  The GPU box has no reference tree; tests there
read only these fixtures.  Scan noise is disabled (agent.scan_rng = None after reset) so the
reference is bit-deterministic.

Fixtures
  kat_reference_tests.npz   the reference's own known-answer vectors (dynamic_models.py:257-266,
                            collision_models.py:274-324), copied as data.
  kat_kernels.npz           per-kernel input/output pairs: dynamics RHS, pid, get_vertices, GJK,
                            ray_cast, check_ttc, get_blocked_view_indices on random inputs.
  scans_<map>.npz           get_scan at fixed poses for B in {270,540,1080,2160} (example_map) and
                            1080 beams on berlin/skirk/vegas/stata_basement (res 0.05 / 0.0504 maps).
  scans_rotated_origin.npz  get_scan on example_map.png under a yaml origin with yaw 0.35 (made ad hoc, see
                            tests/golden/make_golden_rotated.py).
  traj_a1_random.npz        Simulator, 1 agent, random actions: state every tick, scans every 8th.
  traj_a2_random.npz        Simulator, 2 agents 4.6 m apart, random actions.
  traj_a2_close.npz         Simulator, 2 agents 0.6-1.2 m apart (GJK contact, opponent occlusion,
                            rear-cut window case), several episodes.
  traj_a3_euler.npz         Simulator, 3 agents, Euler integrator, lidar_dist 0.1.
  traj_a2_params.npz        Simulator, 2 close agents, update_params(p2, agent_idx=1) before the run.
  traj_{berlin,vegas}_a2.npz  Simulator, 2 agents 0.7-1.1 m apart on the 0.05 m maps (vegas = the reference default).
  env_race2.npz             real F110Env, 2 pure-pursuit cars (BASELINE configs[0]) until the ego rams the slower opponent.
  env_laps.npz              real F110Env (gym/pyglet stubbed) + PurePursuitPlanner, 1 agent 2 laps:
                            actions, states, lap_times/lap_counts/done/toggles every tick.
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, '..', '..'))
from oracle import ref_import  # noqa: E402

ns = ref_import.load()
lm, dm, cm, bc = ns.laser_models, ns.dynamic_models, ns.collision_models, ns.base_classes

PARAMS = {'mu': 1.0489, 'C_Sf': 4.718, 'C_Sr': 5.4562, 'lf': 0.15875, 'lr': 0.17145, 'h': 0.074,
          'm': 3.74, 'I': 0.04712, 's_min': -0.4189, 's_max': 0.4189, 'sv_min': -3.2, 'sv_max': 3.2,
          'v_switch': 7.319, 'a_max': 9.51, 'v_min': -5.0, 'v_max': 20.0, 'width': 0.31, 'length': 0.58}
PKEYS = ['mu', 'C_Sf', 'C_Sr', 'lf', 'lr', 'h', 'm', 'I', 's_min', 's_max', 'sv_min', 'sv_max',
         'v_switch', 'a_max', 'v_min', 'v_max']
WP = np.loadtxt(ns.example_waypoints, delimiter=';', skiprows=3)


def wp_pose(k):
    k = k % WP.shape[0]
    return np.array([WP[k, 1], WP[k, 2], WP[k, 3] + np.pi / 2])


def save(name, **arrays):
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **arrays)
    print('%-28s %8.1f KB' % (name, os.path.getsize(path) / 1024))


# ----------------------------------------------------------------------------- reference KATs
def kat_reference_tests():
    # dynamic_models.py:232-266 (DynamicsTest.setUp + test_derivatives) — data values only
    tp = dict(mu=1.0489, C_Sf=21.92 / 1.0489, C_Sr=21.92 / 1.0489, lf=0.3048 * 3.793293,
              lr=0.3048 * 4.667707, h=0.3048 * 2.01355, m=4.4482216152605 / 0.3048 * 74.91452,
              I=4.4482216152605 * 0.3048 * 1321.416, s_min=-1.066, s_max=1.066, sv_min=-0.4,
              sv_max=0.4, v_switch=7.319, a_max=11.5, v_min=-13.6, v_max=50.8)
    pvec = np.array([tp[k] for k in PKEYS] + [0.31, 0.58])
    f_ks_gt = np.array([16.3475935934250209, 0.4819314886013121, 0.1500000000000000,
                        5.1464424102339752, 0.2401426578627629])
    f_st_gt = np.array([15.7213512030862397, 0.0925527979719355, 0.1500000000000000,
                        5.3536773276413925, 0.0529001056654038, 0.6435589397748606,
                        0.0313297971641291])
    x_ks = np.array([3.9579422297936526, 0.0391650102771405, 0.0378491427211811,
                     16.3546957860883566, 0.0294717351052816])
    x_st = np.array([2.0233348142065677, 0.0041907137716636, 0.0197545248559617,
                     15.7216236334290116, 0.0025857914776859, 0.0529001056654038,
                     0.0033012170610298])
    u = np.array([0.15, 0.63 * 9.81])
    # collision_models.py:274-324 (seed 1234; test_multiple_collisions draws a..f in order)
    np.random.seed(1234)
    v1 = np.asarray([[4, 11.], [5, 5], [9, 9], [10, 10]])
    bodies = [v1 + np.random.normal(size=v1.shape) / 100. for _ in range(6)] + [v1 + 10.]
    allv = np.stack(bodies)
    col, idx = cm.collision_multiple(allv)
    assert np.all(col == np.array([1., 1., 1., 1., 1., 1., 0.]))
    assert np.all(idx == np.array([5., 5., 5., 5., 5., 4., -1.]))
    # test_random_collision: 1000 jittered pairs must all collide
    pairs = np.stack([np.stack([v1 + np.random.normal(size=v1.shape) / 100.,
                                v1 + np.random.normal(size=v1.shape) / 100.]) for _ in range(1000)])
    assert all(cm.collision(np.ascontiguousarray(p[0]), np.ascontiguousarray(p[1])) for p in pairs)
    save('kat_reference_tests.npz', pvec=pvec, f_ks_gt=f_ks_gt, f_st_gt=f_st_gt, x_ks=x_ks, x_st=x_st,
         u=u, multi_vertices=allv, multi_collisions=col, multi_collision_idx=idx, jitter_pairs=pairs)


# ----------------------------------------------------------------------------- kernel KATs
def kat_kernels():
    rng = np.random.default_rng(2024)
    p16 = [PARAMS[k] for k in PKEYS]
    pvec = np.array(p16 + [PARAMS['width'], PARAMS['length']])
    # dynamics RHS on random states (both branches), + pid
    n = 512
    X = np.zeros((n, 7))
    X[:, 0:2] = rng.uniform(-50, 50, (n, 2))
    X[:, 2] = rng.uniform(-0.45, 0.45, n)
    X[:, 3] = np.where(rng.random(n) < 0.3, rng.uniform(-0.6, 0.6, n), rng.uniform(-5.5, 21, n))
    X[:, 4] = rng.uniform(0, 2 * np.pi, n)
    X[:, 5] = rng.uniform(-3, 3, n)
    X[:, 6] = rng.uniform(-0.3, 0.3, n)
    U = np.stack([rng.uniform(-4, 4, n), rng.uniform(-12, 12, n)], axis=1)
    F = np.stack([dm.vehicle_dynamics_st(X[i], U[i], *p16) for i in range(n)])
    pid_in = np.stack([rng.uniform(-6, 21, n), rng.uniform(-0.5, 0.5, n), rng.uniform(-6, 21, n),
                       rng.uniform(-0.45, 0.45, n)], axis=1)
    pid_in[:16, 1] = pid_in[:16, 3] + rng.uniform(-1.5e-4, 1.5e-4, 16)   # around the 1e-4 dead band
    pid_out = np.array([dm.pid(r[0], r[1], r[2], r[3], PARAMS['sv_max'], PARAMS['a_max'],
                               PARAMS['v_max'], PARAMS['v_min']) for r in pid_in])   # (accl, sv)
    # vertices + GJK on random car pairs (centre distance 0..1 m)
    m = 2000
    pa = np.stack([rng.uniform(-5, 5, m), rng.uniform(-5, 5, m), rng.uniform(0, 2 * np.pi, m)], axis=1)
    off_r, off_a = rng.uniform(0, 1.0, m), rng.uniform(0, 2 * np.pi, m)
    pb = np.stack([pa[:, 0] + off_r * np.cos(off_a), pa[:, 1] + off_r * np.sin(off_a),
                   rng.uniform(0, 2 * np.pi, m)], axis=1)
    pb[:20] = pa[:20]   # identical poses: d == 0 branch (collision_models.py:133-134)
    va = np.stack([cm.get_vertices(p, 0.58, 0.31) for p in pa])
    vb = np.stack([cm.get_vertices(p, 0.58, 0.31) for p in pb])
    hit = np.array([cm.collision(np.ascontiguousarray(va[i]), np.ascontiguousarray(vb[i])) for i in range(m)])
    # opponent ray-cast + window indices on random close pairs
    tabs = _beam_tables(1080, 4.7)
    k = 160
    ego = np.stack([rng.uniform(-5, 5, k), rng.uniform(-5, 5, k), rng.uniform(0, 2 * np.pi, k)], axis=1)
    ego[:20, 2] = 0.0      # yaw zeroed by an iTTC hit (base_classes.py:246-249)
    r_, a_ = rng.uniform(0.2, 6.0, k), rng.uniform(0, 2 * np.pi, k)
    r_[:60] = rng.uniform(0.05, 0.5, 60)   # straddling the rear cut
    opp = np.stack([ego[:, 0] + r_ * np.cos(a_), ego[:, 1] + r_ * np.sin(a_), rng.uniform(0, 2 * np.pi, k)], axis=1)
    ov = np.stack([cm.get_vertices(p, 0.58, 0.31) for p in opp])
    base_scan = rng.uniform(0.3, 12.0, (k, 1080)).astype(np.float32).astype(np.float64)
    rc = np.stack([lm.ray_cast(ego[i], base_scan[i].copy(), tabs[0], ov[i]) for i in range(k)])
    win = np.array([lm.get_blocked_view_indices(ego[i], ov[i], tabs[0]) for i in range(k)])
    # iTTC
    q = 160
    ttc_scan = rng.uniform(0.1, 10.0, (q, 1080))
    ttc_scan[:80] = tabs[2][None, :] + rng.uniform(0.0, 0.1, (80, 1080))
    ttc_scan = ttc_scan.astype(np.float32).astype(np.float64)
    ttc_vel = rng.uniform(-5, 20, q)
    ttc_vel[:10] = 0.0
    ttc = np.array([lm.check_ttc_jit(ttc_scan[i], ttc_vel[i], tabs[0], tabs[1], tabs[2], 0.005) for i in range(q)])
    save('kat_kernels.npz', pvec=pvec, X=X, U=U, F=F, pid_in=pid_in, pid_out=pid_out, pose_a=pa, pose_b=pb,
         verts_a=va, verts_b=vb, gjk=hit, rc_ego=ego, rc_opp_verts=ov, rc_scan_in=base_scan.astype(np.float32),
         rc_scan_out=rc, rc_window=win, ttc_scan=ttc_scan.astype(np.float32), ttc_vel=ttc_vel, ttc=ttc,
         scan_angles=tabs[0], cosines=tabs[1], side_distances=tabs[2])


def _beam_tables(num_beams, fov):
    ns.RaceCar.scan_simulator = None
    bc.RaceCar(PARAMS, 12345, num_beams=num_beams, fov=fov)
    return ns.RaceCar.scan_angles.copy(), ns.RaceCar.cosines.copy(), ns.RaceCar.side_distances.copy()


# ----------------------------------------------------------------------------- scans
def scans():
    rng = np.random.default_rng(99)
    poses = np.stack([wp_pose(k) for k in (0, 97, 211, 333, 480, 555, 690, 760)])
    poses[4:, 2] = rng.uniform(0, 2 * np.pi, 4)            # random headings
    poses = np.concatenate([poses, [[-90.0, 0.0, 0.3], [30.0, 60.0, 4.0], [0.7, 0.0, -7.5], [0.7, 0.0, 13.0]]])
    out = {'poses': poses}
    for B in (270, 540, 1080, 2160):
        s = lm.ScanSimulator2D(B, 4.7)
        s.set_map(ns.example_map, '.png')
        out['scan_%d' % B] = np.stack([s.scan(p, None) for p in poses])
    save('scans_example_map.npz', **out)
    for name, ext in (('berlin', '.png'), ('skirk', '.png'), ('vegas', '.png'), ('stata_basement', '.png')):
        s = lm.ScanSimulator2D(1080, 4.7)
        s.set_map(os.path.join(ns.maps_dir, name + '.yaml'), ext)
        free = np.argwhere(s.dt > 0.4)
        sel = free[rng.choice(free.shape[0], 6, replace=False)]
        ps = np.stack([sel[:, 1] * s.map_resolution + s.orig_x + 0.013, sel[:, 0] * s.map_resolution + s.orig_y + 0.007,
                       rng.uniform(0, 2 * np.pi, 6)], axis=1)
        save('scans_%s.npz' % name, poses=ps, scan_1080=np.stack([s.scan(p, None) for p in ps]))


# ----------------------------------------------------------------------------- trajectories
def run_traj(name, num_agents, episodes, ticks, gap_fn, seed, integrator=None, lidar_dist=0.0,
             speed_hi=8.0, scan_every=8, map_yaml=None, params_update=None):
    rng = np.random.default_rng(seed)
    sim = ref_import.new_simulator(ns, PARAMS, num_agents, map_yaml or ns.example_map, integrator=integrator,
                                   lidar_dist=lidar_dist)
    if params_update is not None:                      # Simulator.update_params (base_classes.py:514-534)
        sim.update_params(params_update[1], agent_idx=params_update[0])
    A = num_agents
    poses0, actions, states, cols, cidx, scans_, scan_ticks = [], [], [], [], [], [], []
    for ep in range(episodes):
        poses = gap_fn(rng)
        ref_import.reset_noise_off(sim, poses)
        poses0.append(poses)
        ea, es, ec, ei, esc = [], [], [], [], []
        for t in range(ticks):
            act = np.stack([rng.uniform(-0.4189, 0.4189, A), rng.uniform(0, speed_hi, A)], axis=1)
            obs = sim.step(act)
            ea.append(act)
            es.append(np.array([a.state.copy() for a in sim.agents]))
            ec.append(obs['collisions'].copy())
            ei.append(sim.collision_idx.copy())
            if t % scan_every == 0 or obs['collisions'].any():
                esc.append(np.array(obs['scans']))
                if ep == 0:
                    pass
                scan_ticks.append((ep, t))
        actions.append(ea); states.append(es); cols.append(ec); cidx.append(ei); scans_.extend(esc)
    save(name, poses0=np.array(poses0), actions=np.array(actions), states=np.array(states),
         collisions=np.array(cols), collision_idx=np.array(cidx), scans=np.array(scans_),
         scan_ticks=np.array(scan_ticks), lidar_dist=lidar_dist,
         integrator=1 if integrator in (None, ns.Integrator.RK4) else 2)


def trajectories():
    def far(A):
        def f(rng):
            k = int(rng.integers(0, WP.shape[0]))
            return np.stack([wp_pose(k - 23 * i) for i in range(A)])
        return f

    def close(rng):
        k = int(rng.integers(0, WP.shape[0]))
        g = int(rng.integers(3, 7))
        p = np.stack([wp_pose(k), wp_pose(k - g)])
        p[1, 2] += rng.uniform(-0.3, 0.3)
        return p

    run_traj('traj_a1_random.npz', 1, 3, 260, far(1), 11)
    run_traj('traj_a2_random.npz', 2, 2, 260, far(2), 12)
    run_traj('traj_a2_close.npz', 2, 6, 90, close, 13, scan_every=6)
    run_traj('traj_a3_euler.npz', 3, 2, 150, far(3), 14, integrator=ns.Integrator.Euler, lidar_dist=0.1)


def trajectory_params():
    """update_params on one agent: a heavier, shorter, grippier car in slot 1 (dynamics AND its body for GJK/ray-cast)."""
    def close(rng):
        k = int(rng.integers(0, WP.shape[0]))
        p = np.stack([wp_pose(k), wp_pose(k - int(rng.integers(4, 8)))])
        return p
    p2 = dict(PARAMS, mu=0.8, m=4.5, lf=0.17, lr=0.16, C_Sf=5.1, I=0.05, width=0.28, length=0.50, a_max=7.0)
    run_traj('traj_a2_params.npz', 2, 3, 120, close, 15, scan_every=6, params_update=(1, p2))


def trajectories_other_maps():
    """Simulator trajectories on the 0.05 m maps (vegas is the reference's default map): the metre-unit march, the
    general xy_2_rc and the opponent ray-cast on a resolution that is not a power of two."""
    for name, seed in (('berlin', 31), ('vegas', 32)):
        yaml_path = os.path.join(ns.maps_dir, name + '.yaml')
        s = lm.ScanSimulator2D(1080, 4.7)
        s.set_map(yaml_path, '.png')
        free = np.argwhere(s.dt > 0.7)

        def pair(rng, s=s, free=free):
            r, c = free[rng.integers(0, free.shape[0])]
            x, y = c * s.map_resolution + s.orig_x + 0.011, r * s.map_resolution + s.orig_y + 0.017
            th = rng.uniform(0, 2 * np.pi)
            d = rng.uniform(0.7, 1.1)
            return np.array([[x, y, th], [x - d * np.cos(th), y - d * np.sin(th), th + rng.uniform(-0.3, 0.3)]])
        run_traj('traj_%s_a2.npz' % name, 2, 4, 70, pair, seed, speed_hi=6.0, scan_every=5, map_yaml=yaml_path)


# ----------------------------------------------------------------------------- F110Env laps
def _run_real_env(fname, num_agents, pose_fn, gains, max_ticks=20000):
    """Real F110Env (gym / pyglet stubbed, SURVEY app. B ii) driven by the example's PurePursuitPlanner, one planner
    call per agent with its own (lookahead, vgain); every tick is recorded until `done`."""
    gym = types.ModuleType('gym')
    gym.Env = object
    for sub in ('error', 'spaces', 'utils'):
        setattr(gym, sub, types.ModuleType('gym.' + sub)); sys.modules['gym.' + sub] = getattr(gym, sub)
    seeding = types.ModuleType('gym.utils.seeding'); gym.utils.seeding = seeding
    sys.modules['gym.utils.seeding'] = seeding
    envs = types.ModuleType('gym.envs'); reg = types.ModuleType('gym.envs.registration')
    reg.register = lambda **k: None; envs.registration = reg; gym.envs = envs
    sys.modules.update({'gym': gym, 'gym.envs': envs, 'gym.envs.registration': reg})
    pyglet = types.ModuleType('pyglet'); pyglet.options = {}
    gl = types.ModuleType('pyglet.gl'); gl.GL_POINTS = 0; pyglet.gl = gl
    sys.modules.update({'pyglet': pyglet, 'pyglet.gl': gl})
    from f110_gym.envs import f110_env
    if os.path.join(ref_import.REF_ROOT, 'examples') not in sys.path:
        sys.path.insert(0, os.path.join(ref_import.REF_ROOT, 'examples'))
    import waypoint_follow as wf
    from argparse import Namespace
    import yaml
    ex = os.path.join(ref_import.REF_ROOT, 'examples')
    cwd = os.getcwd()
    os.chdir(ex)
    try:
        with open('config_example_map.yaml') as f:
            conf = Namespace(**yaml.safe_load(f))
        ns.RaceCar.scan_simulator = None
        A = num_agents
        env = f110_env.F110Env(map=conf.map_path, map_ext=conf.map_ext, num_agents=A, timestep=0.01,
                               integrator=ns.Integrator.RK4)
        planner = wf.PurePursuitPlanner(conf, 0.17145 + 0.15875)
        pose0 = pose_fn(conf)
        # noise must be off for determinism; F110Env.reset -> Simulator.reset recreates the rng, so
        # patch RaceCar.reset's effect by wrapping sim.reset.
        orig_reset = env.sim.reset

        def reset_noise_off(p):
            orig_reset(p)
            for a in env.sim.agents:
                a.scan_rng = None
        env.sim.reset = reset_noise_off
        obs, rew, done, info = env.reset(pose0)
        rec = dict(actions=[], states=[], lap_times=[], lap_counts=[], done=[], toggles=[], collisions=[])

        def record(act, done, info):
            rec['actions'].append(act)
            rec['states'].append(np.array([a.state.copy() for a in env.sim.agents]) if A > 1
                                 else env.sim.agents[0].state.copy())
            rec['lap_times'].append(env.lap_times.copy()); rec['lap_counts'].append(env.lap_counts.copy())
            rec['done'].append(done); rec['toggles'].append(env.toggle_list.copy())
            rec['collisions'].append(obs['collisions'].copy())
        record(np.zeros((A, 2)), done, info)     # the tick executed inside reset
        while not done and len(rec['done']) < max_ticks:
            act = np.zeros((A, 2))
            for i in range(A):
                speed, steer = planner.plan(obs['poses_x'][i], obs['poses_y'][i], obs['poses_theta'][i], gains[i][0], gains[i][1])
                act[i] = [steer, speed]
            obs, rew, done, info = env.step(act)
            record(act, done, info)
    finally:
        os.chdir(cwd)
    save(fname, pose0=pose0, gains=np.array(gains), **{k: np.array(v) for k, v in rec.items()})
    print(fname, 'ticks', len(rec['done']), 'done', rec['done'][-1], 'lap_times', rec['lap_times'][-1], 'lap_counts',
          rec['lap_counts'][-1], 'collisions', rec['collisions'][-1])


def env_laps():
    _run_real_env('env_laps.npz', 1, lambda conf: np.array([[conf.sx, conf.sy, conf.stheta]]),
                  [(0.82461887897713965, 1.375)])


def env_race2():
    """BASELINE configs[0]: two pure-pursuit cars on example_map in the real F110Env.  The ego starts 200 waypoints
    (40 m) behind a slower opponent, laps once and then runs into it: lap toggles of both cars, opponent occlusion in
    the ego's scan while closing in, then done through the ego's collision."""
    def poses(conf):
        return np.stack([wp_pose(0), wp_pose(200)])
    _run_real_env('env_race2.npz', 2, poses, [(0.82461887897713965, 1.375), (0.82461887897713965, 1.1)])


# ----------------------------------------------------------------------------- planner
def kat_planner():
    """PurePursuitPlanner.plan of examples/waypoint_follow.py at on-track, off-track (re-acquire) and far poses."""
    import yaml
    from argparse import Namespace
    for name in ('gym', 'pyglet', 'pyglet.gl'):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules['pyglet.gl'].GL_POINTS = 0
    sys.modules['pyglet'].gl = sys.modules['pyglet.gl']
    ex = os.path.join(ref_import.REF_ROOT, 'examples')
    if ex not in sys.path:
        sys.path.insert(0, ex)
    import waypoint_follow as wf
    with open(os.path.join(ex, 'config_example_map.yaml')) as f:
        conf = Namespace(**yaml.safe_load(f))
    conf.wpt_path = os.path.join(ex, 'example_waypoints.csv')
    wb = 0.17145 + 0.15875
    pl = wf.PurePursuitPlanner(conf, wb)
    W = pl.waypoints
    rng = np.random.default_rng(77)
    M = 2000
    k = rng.integers(0, W.shape[0], M)
    off = np.where(rng.random(M) < 0.7, rng.normal(0, 0.3, M), rng.uniform(-30, 30, M))
    ang = rng.uniform(0, 2 * np.pi, M)
    poses = np.stack([W[k, 1] + off * np.cos(ang), W[k, 2] + off * np.sin(ang), rng.uniform(-np.pi, 2 * np.pi, M)], axis=1)
    poses[:40, :2] = W[k[:40], 1:3]                      # exactly on a waypoint
    poses[40:60, :2] = W[-3:, 1:3].mean(axis=0)          # near the end of the list: wrap-around search
    tlad, vgain = 0.82461887897713965, 1.375
    out = np.array([pl.plan(p[0], p[1], p[2], tlad, vgain) for p in poses])     # (speed, steer)
    out2 = np.array([pl.plan(p[0], p[1], p[2], 2.5, 0.9) for p in poses[:400]])
    save('kat_planner.npz', poses=poses, speed_steer=out, speed_steer_l25=out2, tlad=tlad, vgain=vgain, wheelbase=wb)


if __name__ == '__main__':
    groups = {'kat_planner': kat_planner, 'kat_reference_tests': kat_reference_tests, 'kat_kernels': kat_kernels,
              'scans': scans, 'trajectories': trajectories, 'trajectories_other_maps': trajectories_other_maps, 'trajectory_params': trajectory_params,
              'env_laps': env_laps, 'env_race2': env_race2}
    for name in (sys.argv[1:] or list(groups)):          # python make_golden.py [group ...]
        groups[name]()
