"""BASELINE configs[0] on the CUDA path: the two-car pure-pursuit race of the real reference F110Env
(tests/golden/env_race2.npz, 2088 ticks) replayed through the mirrored F110Env -- lap toggles of both cars, opponent
occlusion while the ego closes in, `done` through the ego's collision.  Its CPU twin,
test_oracle_vs_golden.py::test_env_race2, pins the oracle on the same fixture."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), 'golden')
MAPS = os.path.join(os.path.dirname(__file__), '..', 'f1tenth_gym_b200', 'maps')


def test_env_race2_vs_reference():
    import torch
    import f1tenth_gym_b200 as f110
    k = np.load(os.path.join(G, 'env_race2.npz'))
    env = f110.F110Env(map=os.path.join(MAPS, 'example_map'), map_ext='.png', num_agents=2, timestep=0.01,
                       integrator=f110.Integrator.RK4, scan_noise_std=0.0, device=torch.device('cuda:0'))
    obs, rew, done, info = env.reset(k['pose0'])
    T = k['actions'].shape[0]
    worst = 0.0
    for t in range(T):
        if t > 0:
            obs, rew, done, info = env.step(k['actions'][t])
        st = env.sim.state.cpu().numpy().T                      # (A, 7)
        worst = max(worst, np.abs(st - k['states'][t]).max())
        assert np.array_equal(np.asarray(obs['collisions']), k['collisions'][t]), t
        assert np.array_equal(np.asarray(obs['lap_counts']), k['lap_counts'][t]), t
        assert np.allclose(obs['lap_times'], k['lap_times'][t], rtol=0, atol=1e-9), t
        assert np.array_equal(np.asarray(env.toggle_list), k['toggles'][t]), t
        assert bool(done) == bool(k['done'][t]), t
    assert done and obs['collisions'][0] == 1.0
    assert worst < 1e-9, worst


def test_rotated_origin_map():
    """Map origin with a yaw: the rotation terms of xy_2_rc on the literal kernel (stand-alone scan against the reference
    golden, bit-exact in fp64) and on the step path (against the oracle).  Written after the round's GPU budget was
    spent -- the oracle half is verified (test_oracle_vs_golden.py::test_scans_rotated_origin); this CUDA half first
    runs at the round-end GPU tier, which is why it sits in the last test file."""
    import torch
    import oracle
    import f1tenth_gym_b200 as f110
    dev = torch.device('cuda:0')
    k = np.load(os.path.join(G, 'scans_rotated_origin.npz'))
    hm0 = f110.maps.load_map(os.path.join(MAPS, 'example_map.yaml'), '.png')
    hm = f110.maps.HostMap(hm0.dt, float(k['resolution']), tuple(k['origin']))
    assert hm.fast_path == 0 and hm.orig_s != 0.0
    dm = f110.DeviceMap(hm, dev)
    ss = f110.ScanSimulator2D(1080, 4.7, device=dev)
    ss.set_device_map(dm)
    assert np.array_equal(ss.scan(k['poses'], out_f64=True).cpu().numpy(), k['scan_1080'])
    # step path: 3 envs x 2 agents, 25 ticks
    N, A = 3, 2
    rng = np.random.default_rng(4)
    poses = np.zeros((N, A, 3))
    th = float(k['origin'][2])
    for e in range(N):
        p = k['poses'][2 * e]
        poses[e, 0] = p
        poses[e, 1] = [p[0] - 0.9 * np.cos(p[2]), p[1] - 0.9 * np.sin(p[2]), p[2]]
    sim = f110.Simulator(f110.maps.DEFAULT_PARAMS, A, 3, num_envs=N, device=dev)
    sim.set_device_map(dm)
    sim.reset(poses)
    om = oracle.OracleMap(hm.dt, hm.resolution, tuple(k['origin']))
    osims = [oracle.OracleSim(om, num_agents=A) for _ in range(N)]
    for e in range(N):
        osims[e].reset(poses[e])
    for _ in range(25):
        act = np.stack([rng.uniform(-0.3, 0.3, (N, A)), rng.uniform(0, 5, (N, A))], axis=2)
        obs = sim.step(act)
        st = sim.state.cpu().numpy().reshape(7, N, A).transpose(1, 2, 0)
        sc = obs['scans'].cpu().numpy().astype(np.float64)
        for e in range(N):
            osims[e].step(act[e])
            assert np.abs(st[e] - osims[e].state).max() < 1e-9
            assert np.abs(sc[e] - osims[e].scans).max() < 4e-6
            assert np.array_equal(obs['collisions'].cpu().numpy()[e], osims[e].collisions)
