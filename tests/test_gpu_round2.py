"""GPU parity tests added in round 2 (VERDICT r1 "untested product configurations"), all through the C ABI:

  * 208 reference scans per map (on-track, free space, inside walls, map border, far outside, absurd coordinates):
    the stand-alone scan kernel bit-exact in fp64, and the PRODUCTION TICK PATH (k_dynamics -> k_march_lean ->
    k_tail) at the same poses equal to fp32(reference fp64 value) beam for beam, with the oracle's lookup count;
  * the tick path at 270 / 540 / 2160 beams with two agents per env (GJK + opponent ray-cast live), incl. the number
    of DT lookups;
  * every march kernel variant that can be selected (lean fp64 / lean rank-coded / 48-warp builds / the TMA-tile
    kernel with 128- and 160-cell tiles / the round-1 persistent and coded kernels / block-per-tile / literal) against
    the oracle;
  * 64-beam work items (march_item_beams=64);
  * CUDA replay of tests/golden/traj_a2_params.npz (update_params with another body size on one agent);
  * the reference's kinematic single-track known-answer vector on the GPU (f110_vehicle_dynamics_ks);
  * the latched `terminated` flag of the fused tick with auto-reset.
"""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

G = os.path.join(os.path.dirname(__file__), 'golden')
MAPS = os.path.join(os.path.dirname(__file__), '..', 'f1tenth_gym_b200', 'maps')
TOL_STATE = 1e-9
TOL_SCAN32 = 4e-6


def g(name):
    return np.load(os.path.join(G, name))


def cpu(t):
    return t.detach().cpu().numpy()


@pytest.fixture(scope='module')
def f110():
    import f1tenth_gym_b200 as f
    return f


@pytest.fixture(scope='module')
def dev():
    return torch.device('cuda:0')


@pytest.fixture(scope='module')
def example_map(f110, dev):
    return f110.DeviceMap.from_yaml(os.path.join(MAPS, 'example_map.yaml'), '.png', dev)


@pytest.fixture
def variant(f110):
    """Selects a march kernel variant for one test and restores the default afterwards."""
    L = f110._native.lib()

    def set_variant(v):
        L.f110_debug_set_variant(int(v))
    yield set_variant
    L.f110_debug_set_variant(0)


def _omap(dmap):
    import oracle
    h = dmap.host
    return oracle.OracleMap(h.dt, h.resolution, (h.orig_x, h.orig_y, 0.0))


def _start_poses(f110, rng, N, A, gap=23):
    wp = f110.maps.load_waypoints()
    k = rng.integers(0, wp.shape[0], N)
    return np.stack([np.stack([wp[(kk - gap * i) % wp.shape[0]] for i in range(A)]) for kk in k])


# ----------------------------------------------------------------------------- wide scan goldens
@pytest.mark.parametrize('name', ['example_map', 'berlin', 'skirk', 'vegas', 'stata_basement', 'levine'])
def test_scans_wide_standalone(f110, dev, name):
    k = g('scans_wide_%s.npz' % name)
    sim = f110.ScanSimulator2D(1080, 4.7, device=dev)
    sim.set_map(os.path.join(MAPS, name + '.yaml'), '.pgm' if name == 'levine' else '.png')
    s64 = cpu(sim.scan(k['poses'], out_f64=True))
    if name == 'example_map':
        assert np.array_equal(s64, k['scan_1080'])
        for B in (270, 2160):
            sb = f110.ScanSimulator2D(B, 4.7, device=dev)
            sb.set_map(os.path.join(MAPS, name + '.yaml'), '.png')
            assert np.array_equal(cpu(sb.scan(k['poses'][:48], out_f64=True)), k['scan_%d' % B])
    else:
        assert np.array_equal(np.take_along_axis(s64, k['beam_idx'].astype(np.int64), axis=1), k['scan_1080_sub'])


@pytest.mark.parametrize('name,v', [('example_map', 0), ('example_map', 20), ('example_map', 21), ('example_map', 22),
                                    ('example_map', 30), ('example_map', 31), ('example_map', 40), ('example_map', 41),
                                    ('example_map', 1), ('example_map', 6), ('example_map', 7), ('example_map', 13),
                                    ('berlin', 0), ('berlin', 1), ('skirk', 0), ('vegas', 0), ('stata_basement', 0),
                                    ('levine', 0)])
def test_tick_path_at_wide_poses(f110, dev, variant, name, v):
    """The production tick path with a car standing at each golden pose (zero action, zero speed: the pose does not
    move, only the single-shot yaw wrap of base_classes.py:400-404 applies, which the oracle reproduces): every beam
    must be the fp32 rounding of the oracle's fp64 value, and the DT lookup counts must agree."""
    import oracle
    k = g('scans_wide_%s.npz' % name)
    poses = k['poses']
    N = poses.shape[0]
    dmap = f110.DeviceMap.from_yaml(os.path.join(MAPS, name + '.yaml'), '.pgm' if name == 'levine' else '.png', dev)
    variant(v)
    sim = f110.Simulator(f110.maps.DEFAULT_PARAMS, 1, 1, num_envs=N, device=dev, count_lookups=True)
    sim.set_device_map(dmap)
    sim.reset(poses[:, None, :])
    omap = _omap(dmap)
    for t in range(2):                                    # tick 2 runs with the queue ordered by tick 1's costs
        obs = sim.tick(np.zeros((N, 1, 2)), env_level=False)
    sc = cpu(obs['scans'])[:, 0]
    nl = 0
    for j in range(N):
        o = oracle.OracleSim(omap, num_agents=1)
        o.reset(poses[j:j + 1])
        o.step(np.zeros((1, 2)))
        o.step(np.zeros((1, 2)))
        assert np.array_equal(sc[j], o.scans[0].astype(np.float32)), (name, v, j)
        nl += o.nlook
        # within the yaw range the wrap leaves untouched, the oracle value IS the reference golden
        if name == 'example_map' and 0.0 <= poses[j, 2] <= 2 * np.pi:
            assert np.array_equal(o.scans[0], k['scan_1080'][j]), j
    assert sim.lookups() == nl


# ----------------------------------------------------------------------------- tick path at other beam counts, variants
def _rollout_vs_oracle(f110, dev, dmap, N, A, B, T, gap, seed, all_tick=False, **simkw):
    import oracle
    rng = np.random.default_rng(seed)
    omap = _omap(dmap)
    osims = [oracle.OracleSim(omap, num_agents=A, num_beams=B) for _ in range(N)]
    sim = f110.Simulator(f110.maps.DEFAULT_PARAMS, A, 12345, num_envs=N, num_beams=B, device=dev, count_lookups=True, **simkw)
    sim.set_device_map(dmap)
    poses = _start_poses(f110, rng, N, A, gap)
    sim.reset(poses)
    for e in range(N):
        osims[e].reset(poses[e])
    worst_state, n_col, n_occ = 0.0, 0, 0
    for t in range(T):
        act = np.stack([rng.uniform(-0.4189, 0.4189, (N, A)), rng.uniform(0, 8, (N, A))], axis=2)
        obs = sim.tick(act, env_level=False) if (t % 2 or all_tick) else sim.step(act)
        st = cpu(sim.state).reshape(7, N, A).transpose(1, 2, 0)
        sc = cpu(obs['scans'])
        col = cpu(obs['collisions'])
        for e in range(N):
            osims[e].step(act[e])
            worst_state = max(worst_state, np.abs(st[e] - osims[e].state).max())
            d = np.abs(sc[e].astype(np.float64) - osims[e].scans)
            assert d.max() < TOL_SCAN32, (t, e, d.max())
            n_occ += int((sc[e] != osims[e].scans.astype(np.float32)).sum())     # only opponent-occluded beams may differ in the last bits
            assert np.array_equal(col[e], osims[e].collisions), (t, e)
            n_col += int(col[e].sum())
    assert worst_state < TOL_STATE, worst_state
    assert sim.lookups() == sum(o.nlook for o in osims)
    return n_col, n_occ


@pytest.mark.parametrize('B', [270, 540, 1080, 2160])
def test_tick_path_beam_counts_two_agents(f110, dev, example_map, B):
    """BASELINE configs[4] beam counts on the production tick path (k_march_lean; partial last 32-beam slice at 270 /
    1080, LUT-bin duplication at 2160) with GJK and the opponent ray-cast live."""
    _rollout_vs_oracle(f110, dev, example_map, N=16, A=2, B=B, T=60, gap=4, seed=700 + B)


@pytest.mark.parametrize('v', [0, 20, 21, 22, 30, 31, 40, 41, 1, 6, 7, 9, 13])
def test_march_variants_vs_oracle(f110, dev, example_map, variant, v):
    variant(v)
    _rollout_vs_oracle(f110, dev, example_map, N=12, A=2, B=1080, T=40, gap=23, seed=900 + v)


@pytest.mark.parametrize('name,v', [('berlin', 0), ('berlin', 40), ('berlin', 1), ('vegas', 0)])
def test_metre_march_variants_vs_oracle(f110, dev, variant, name, v):
    """0.05 m maps: the metre-unit flavour of the lean kernel (and of the round-1 kernel) on the tick path."""
    import oracle
    variant(v)
    dmap = f110.DeviceMap.from_yaml(os.path.join(MAPS, name + '.yaml'), '.png', dev)
    h = dmap.host
    rng = np.random.default_rng(77)
    free = np.argwhere(h.dt > 0.5)
    N, A, T = 16, 1, 40
    sel = free[rng.choice(free.shape[0], N, replace=False)]
    poses = np.stack([sel[:, 1] * h.resolution + h.orig_x, sel[:, 0] * h.resolution + h.orig_y,
                      rng.uniform(0, 2 * np.pi, N)], axis=1)[:, None, :]
    omap = _omap(dmap)
    osims = [oracle.OracleSim(omap, num_agents=A) for _ in range(N)]
    sim = f110.Simulator(f110.maps.DEFAULT_PARAMS, A, 1, num_envs=N, device=dev, count_lookups=True)
    sim.set_device_map(dmap)
    sim.reset(poses)
    for e in range(N):
        osims[e].reset(poses[e])
    for t in range(T):
        act = np.stack([rng.uniform(-0.4, 0.4, (N, A)), rng.uniform(0, 6, (N, A))], axis=2)
        obs = sim.step(act)
        sc = cpu(obs['scans'])
        for e in range(N):
            osims[e].step(act[e])
            assert np.array_equal(sc[e], osims[e].scans.astype(np.float32)), (t, e)
    assert sim.lookups() == sum(o.nlook for o in osims)


@pytest.mark.parametrize('v', [60, 61, 62, 63, 64, 65, 66, 43, 44])
def test_block_shape_and_cluster_variants_bit_identical(f110, dev, example_map, variant, v):
    """Block shapes (2 x 1024, 8 x 256 threads per SM) and the thread-block-cluster launches that share one ticket counter
    through distributed shared memory: same scans, state and collisions as the default launch, bit for bit."""
    N, A, T = 96, 2, 12
    rng = np.random.default_rng(4242)
    poses = _start_poses(f110, rng, N, A, 5)
    acts = np.stack([rng.uniform(-0.4189, 0.4189, (T, N, A)), rng.uniform(0, 8, (T, N, A))], axis=3)
    out = []
    for vv in (0, v):
        variant(vv)
        sim = f110.Simulator(f110.maps.DEFAULT_PARAMS, A, 1, num_envs=N, device=dev)
        sim.set_device_map(example_map)
        sim.env_reset(poses)
        for t in range(T):
            sim.tick(acts[t], env_level=True)
        torch.cuda.synchronize()
        out.append((cpu(sim.scans).copy(), cpu(sim.state).copy(), cpu(sim.collisions).copy()))
    for a, b in zip(out[0], out[1]):
        assert np.array_equal(a, b)


def test_dynamic_queue_large_batch(f110, dev, example_map, variant):
    """The default hands out half of the queue dynamically -- but only when every block has enough runs, which the small
    batches of the other tests never reach.  2048 envs (69632 items = 14 runs per block: 7 static, 7 dynamic): identical to
    the static launch (variant 66) bit for bit, and a sample of envs is checked against the oracle."""
    import oracle
    N, A, T = 2048, 1, 6
    rng = np.random.default_rng(99)
    poses = _start_poses(f110, rng, N, A)
    acts = np.stack([rng.uniform(-0.4189, 0.4189, (T, N, A)), rng.uniform(0, 8, (T, N, A))], axis=3)
    out = {}
    for vv in (66, 0, 42, 43, 44, 45, 57, 58, 81, 82, 83, 84, 85, 86):
        variant(vv)
        sim = f110.Simulator(f110.maps.DEFAULT_PARAMS, A, 1, num_envs=N, device=dev, count_lookups=(vv != 42))
        sim.set_device_map(example_map)
        sim.reset(poses)
        for t in range(T):
            sim.tick(acts[t], env_level=False)
        torch.cuda.synchronize()
        out[vv] = (cpu(sim.scans).copy(), cpu(sim.state).copy(), sim.lookups())
    assert np.array_equal(out[0][0], out[66][0]) and np.array_equal(out[0][1], out[66][1]) and out[0][2] == out[66][2]
    assert np.array_equal(out[42][0], out[66][0])
    for vv in (43, 44, 45, 57, 58, 81, 82, 83, 84, 85, 86):     # several entries per ticket, 48 warps per SM, the three ring hand-offs
        assert np.array_equal(out[vv][0], out[66][0]) and np.array_equal(out[vv][1], out[66][1]) and out[vv][2] == out[66][2]
    omap = _omap(example_map)
    for e in range(0, N, 97):
        o = oracle.OracleSim(omap, num_agents=A)
        o.reset(poses[e])
        for t in range(T):
            o.step(acts[t, e])
        assert np.array_equal(out[0][0][e], o.scans[0].astype(np.float32)), e


@pytest.mark.parametrize('chunk', [3, 2, 4])
def test_ticket_sizes_bit_identical(f110, dev, example_map, variant, chunk):
    """The default launch lets a warp draw 1, 2, 4 or 8 consecutive queue entries per ticket, by the class of the run (very heavy /
    heavy / light) and for the dynamic tail (forced here through f110_debug_set_ipt, with runs of 4, 8 and 16 entries): every
    combination gives the scans, states and lookup count of the static launch, bit for bit."""
    N, A, T = 2048, 1, 4
    L = f110._native.lib()
    rng = np.random.default_rng(7)
    poses = _start_poses(f110, rng, N, A)
    acts = np.stack([rng.uniform(-0.4189, 0.4189, (T, N, A)), rng.uniform(0, 8, (T, N, A))], axis=3)

    def go(v, ipt):
        variant(v)
        L.f110_debug_set_ipt(*ipt)
        sim = f110.Simulator(f110.maps.DEFAULT_PARAMS, A, 1, num_envs=N, device=dev, count_lookups=True)
        sim.set_device_map(example_map)
        sim.reset(poses)
        for t in range(T):
            sim.tick(acts[t], env_level=False)
        torch.cuda.synchronize()
        return cpu(sim.scans).copy(), cpu(sim.state).copy(), sim.lookups()
    try:
        L.f110_debug_set_chunk(chunk)
        ref = go(66, (-1, -1, -1, -1))
        for ipt in [(-1, -1, -1, -1), (0, 0, 0, 0), (0, 1, 2, 0), (0, 1, 2, 1), (2, 1, 0, 2), (1, 1, 1, 1), (2, 2, 2, 2), (0, 1, 3, 3),
                    (3, 3, 3, 3), (0, 2, 3, -1), (1, 0, 2, 0)]:
            got = go(0, ipt)
            assert np.array_equal(got[0], ref[0]) and np.array_equal(got[1], ref[1]) and got[2] == ref[2], ipt
    finally:
        L.f110_debug_set_ipt(-1, -1, -1, -1)
        L.f110_debug_set_chunk(3)


@pytest.mark.parametrize('A,gap', [(2, 3), (3, 4), (4, 3), (2, 23)])
def test_two_phase_tail_vs_oracle(f110, dev, example_map, A, gap):
    """k_tail2 (the fused tick's tail for 2-4 agents per env: thread per (ego, opponent) pair for the scalar work, warp per ego
    for the beams) on close trains of cars -- GJK contacts, wall hits, front and rear (all-beams) occlusion windows -- against the
    oracle, every tick through f110_tick; then the same ticks with the one-warp-per-agent k_tail, bit for bit."""
    L = f110._native.lib()
    L.f110_debug_set_tail(-2)                      # k_tail2 (the default for 2 <= A <= 4)
    try:
        n_col, n_occ = _rollout_vs_oracle(f110, dev, example_map, N=20, A=A, B=1080, T=50, gap=gap, seed=3100 + 10 * A + gap,
                                          all_tick=True)
    finally:
        L.f110_debug_set_tail(8)
    if gap <= 3:                                   # 0.6 m apart: the bodies touch
        assert n_col > 0
    rng = np.random.default_rng(77 + A)
    N, T = 40, 30
    poses = _start_poses(f110, rng, N, A, gap)
    acts = np.stack([rng.uniform(-0.4189, 0.4189, (T, N, A)), rng.uniform(0, 8, (T, N, A))], axis=3)
    wp = torch.from_numpy(f110.maps.load_waypoints()).to(dev)
    out = []
    try:
        for mode in (-2, -1):                      # -2: k_tail2, -1: k_tail
            L.f110_debug_set_tail(mode)
            sim = f110.Simulator(f110.maps.DEFAULT_PARAMS, A, 1, num_envs=N, device=dev)
            sim.set_device_map(example_map)
            sim.env_reset(poses)
            for t in range(T):
                sim.tick(acts[t], env_level=True, autoreset_poses=wp, pose_gap=gap)
            torch.cuda.synchronize()
            out.append([cpu(x).copy() for x in (sim.scans, sim.state, sim.collisions, sim.collision_idx, sim.done, sim.lap_times,
                                                sim.toggle_list, sim.current_time)])
    finally:
        L.f110_debug_set_tail(8)
    for x, y in zip(out[0], out[1]):
        assert np.array_equal(x, y)


def test_march_item_beams_64(f110, dev, example_map, variant):
    """64-beam work items (Simulator(march_item_beams=64)) run on the round-1 persistent kernel <SUB=2>."""
    for B in (1080, 270):
        _rollout_vs_oracle(f110, dev, example_map, N=10, A=2, B=B, T=30, gap=5, seed=1200 + B, march_item_beams=64)


def test_lean_kernel_is_the_default(f110, dev, example_map):
    """The tick path must run k_march_lean on example_map (per-agent record bound, padded tables present)."""
    sim = f110.Simulator(f110.maps.DEFAULT_PARAMS, 1, 1, num_envs=4, device=dev)
    assert sim.march_rec is not None and sim.c.march_rec
    assert example_map.c.dt_cells_pad and example_map.c.dt_codes_pad and example_map.c.sincos2
    assert example_map.c.dt_min_positive == example_map.host.resolution
    assert sim.beams.c.side_max > 0.2


# ----------------------------------------------------------------------------- update_params golden on CUDA
def test_trajectory_with_updated_params_cuda(f110, dev, example_map):
    """tests/golden/traj_a2_params.npz: reference Simulator.update_params(p2, agent_idx=1) (base_classes.py:514-534) with a
    different width / length: slot 1 integrates and ray-casts opponents with its own body (:206-227) while GJK keeps
    the Simulator-level size (:536-550)."""
    k = g('traj_a2_params.npz')
    E, T, A = k['actions'].shape[:3]
    p2 = dict(f110.maps.DEFAULT_PARAMS, mu=0.8, m=4.5, lf=0.17, lr=0.16, C_Sf=5.1, I=0.05, width=0.28, length=0.50, a_max=7.0)
    sim = f110.Simulator(f110.maps.DEFAULT_PARAMS, A, 12345, num_envs=E, device=dev)
    sim.set_device_map(example_map)
    sim.update_params(p2, agent_idx=1)
    sim.reset(k['poses0'])
    ticks = {(int(e), int(t)): i for i, (e, t) in enumerate(k['scan_ticks'])}
    worst_state = worst_scan = 0.0
    n_col = 0
    for t in range(T):
        obs = sim.step(k['actions'][:, t])
        st = cpu(sim.state).reshape(7, E, A).transpose(1, 2, 0)
        worst_state = max(worst_state, np.abs(st - k['states'][:, t]).max())
        assert np.array_equal(cpu(obs['collisions']), k['collisions'][:, t]), t
        assert np.array_equal(cpu(sim.collision_idx).reshape(E, A), k['collision_idx'][:, t]), t
        n_col += int(k['collisions'][:, t].sum())
        sc = None
        for e in range(E):
            if (e, t) in ticks:
                sc = cpu(obs['scans']) if sc is None else sc
                worst_scan = max(worst_scan, np.abs(sc[e].astype(np.float64) - k['scans'][ticks[(e, t)]]).max())
    assert worst_state < TOL_STATE, worst_state
    assert worst_scan < TOL_SCAN32, worst_scan
    assert n_col > 0


# ----------------------------------------------------------------------------- KS known-answer vector on the GPU
def test_reference_ks_kat(f110):
    """dynamic_models.py:257 f_ks_gt (inputs :262-266): the reference's own kinematic single-track vector."""
    k = g('kat_reference_tests.npz')
    f = cpu(f110.kernels.vehicle_dynamics_ks(k['x_ks'][None], k['u'][None], k['pvec']))[0]
    assert np.abs(f - k['f_ks_gt']).max() < 5e-8
    import oracle
    rng = np.random.default_rng(3)
    X = np.stack([rng.uniform(-5, 5, 256), rng.uniform(-5, 5, 256), rng.uniform(-0.45, 0.45, 256), rng.uniform(-3, 12, 256),
                  rng.uniform(0, 6.28, 256)], axis=1)
    U = np.stack([rng.uniform(-4, 4, 256), rng.uniform(-12, 12, 256)], axis=1)
    pv = f110.maps.params_vector(f110.maps.DEFAULT_PARAMS)
    F = cpu(f110.kernels.vehicle_dynamics_ks(X, U, pv))
    Fo = np.stack([oracle.vehicle_dynamics_ks(X[i], U[i], pv) for i in range(256)])
    assert np.abs(F - Fo).max() <= 1e-12 * max(1.0, np.abs(Fo).max())


def test_levine_env_pgm(f110, dev):
    """F110Env(map='levine', map_ext='.pgm') (f110_env.py:108-120 bundled name) loads and steps."""
    env = f110.F110Env(map='levine', map_ext='.pgm', num_agents=1, scan_noise_std=0.0, device=dev)
    obs, rew, done, info = env.reset(np.array([[0.0, 0.0, 0.3]]))
    assert len(obs['scans']) == 1 and obs['scans'][0].shape == (1080,) and obs['scans'][0].dtype == np.float64
    obs, rew, done, info = env.step(np.array([[0.1, 2.0]]))
    assert rew == 0.01 and isinstance(done, bool) and info['checkpoint_done'].shape == (1,)
    assert isinstance(obs['poses_x'][0], float) and obs['lap_counts'].shape == (1,)


def test_packed_u24_host_observation(f110, dev, example_map):
    """Opt-in narrow host observation (f110_pack_scans_u24 inside f110_step_host_async): the decoded ranges equal
    round(fp32 range * 2^19) * 2^-19, i.e. within 9.6e-7 m of the fp32 scan; everything else is the fp32 pipeline's."""
    N, A, T = 32, 2, 6
    rng = np.random.default_rng(11)
    poses = _start_poses(f110, rng, N, A, 5)
    sims = []
    for packed in (False, True):
        sim = f110.Simulator(f110.maps.DEFAULT_PARAMS, A, 1, num_envs=N, device=dev)
        sim.set_device_map(example_map)
        sim.env_reset(poses)
        sims.append((sim, sim.make_host_pipeline(depth=2, packed_scans=packed)))
    acts = np.stack([rng.uniform(-0.4189, 0.4189, (T, N * A)), rng.uniform(0, 8, (T, N * A))], axis=2)
    for t in range(T):
        for sim, sets in sims:
            io = sets[t % 2]
            sim.wait_host(io)
            io['actions'].numpy()[:] = acts[t]
            sim.step_host_async(io)
    for sim, sets in sims:
        for io in sets:
            sim.wait_host(io)
    last = (T - 1) % 2
    f32 = sims[0][1][last]['scans'].numpy()
    dec = f110.Simulator.unpack_scans_u24(sims[1][1][last]['scans_u24'])
    assert dec.shape == f32.shape and dec.dtype == np.float32
    expect = (np.rint(np.clip(f32, 0, None).astype(np.float64) * 2.0 ** 19) * 2.0 ** -19).astype(np.float32)
    assert np.array_equal(dec, expect)
    assert np.abs(dec.astype(np.float64) - f32).max() <= 2.0 ** -20 + 1e-12
    assert np.array_equal(sims[0][1][last]['state'].numpy(), sims[1][1][last]['state'].numpy())


# ----------------------------------------------------------------------------- auto-reset keeps the episode end visible
def test_tick_autoreset_latches_done(f110, dev, example_map):
    """with the fused tick + auto-reset a collision-terminated episode must still show done = 1 (the
    env restarts, the flag of the finished episode stays until the next tick recomputes it); lap counters and time
    are those of the new episode."""
    N, A = 64, 1
    rng = np.random.default_rng(5)
    sim = f110.Simulator(f110.maps.DEFAULT_PARAMS, A, 7, num_envs=N, device=dev)
    sim.set_device_map(example_map)
    wp = torch.from_numpy(f110.maps.load_waypoints()).to(dev)
    sim.env_reset(_start_poses(f110, rng, N, A))
    seen = 0
    for t in range(300):
        act = np.stack([rng.uniform(-0.4189, 0.4189, (N, A)), rng.uniform(6, 8, (N, A))], axis=2)
        obs = sim.tick(act, env_level=True, autoreset_poses=wp)
        done = cpu(sim.done).astype(bool)
        col = cpu(obs['collisions'])[:, 0] != 0
        assert np.array_equal(done, col | (cpu(sim.toggle_list).reshape(N, A) >= 4).all(axis=1))
        if done.any():
            seen += int(done.sum())
            e = int(np.flatnonzero(done)[0])
            assert float(sim.current_time[e]) == 0.0                       # the env was restarted in the same tick ...
            assert cpu(sim.state)[3, e] == 0.0                             # ... at rest on a start pose
    assert seen > 0
