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
