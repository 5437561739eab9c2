#!/bin/bash
# compute-sanitizer over a small tick with auto-reset (run through gpurun from the repo root):
#   bash tools/sanitize.sh   -> gpurun_out/sanitize_{memcheck,racecheck,synccheck}.log  (copy into profiles/<round>/)
# 64 envs x 2 agents, 30 fused ticks (k_dynamics + queue build, k_march_lean, k_tail with the block-barrier hand-off and
# the auto-reset), then the TMA-tile variant of the march, then 1024 envs x 2 agents for 3 ticks: the smallest batch that takes
# the dynamic queue (claims through the shared-memory ring, ticket sizes by queue class) and the two-phase k_tail2.
set -u
O=gpurun_out
mkdir -p $O
cat > /tmp/f110_sanitize.py <<'PY'
import sys
sys.path.insert(0, '.')
import numpy as np, torch
import f1tenth_gym_b200 as f110
dev = torch.device('cuda:0')
L = f110._native.lib()
dmap = f110.DeviceMap.from_yaml(f110.maps.resolve_map_path('example_map'), '.png', dev)
wp_np = f110.maps.load_waypoints()
wp = torch.from_numpy(wp_np).to(dev)
for variant, N, T in ((0, 64, 30), (30, 64, 30), (0, 1024, 3)):
    L.f110_debug_set_variant(variant)
    A = 2
    sim = f110.Simulator(f110.maps.DEFAULT_PARAMS, A, 1, num_envs=N, device=dev)
    sim.set_device_map(dmap)
    rng = np.random.default_rng(0)
    k = rng.integers(0, wp_np.shape[0], N)
    sim.env_reset(np.stack([np.stack([wp_np[kk], wp_np[(kk - 2) % len(wp_np)]]) for kk in k]))   # 0.4 m apart: bodies overlap -> GJK contact -> auto-reset
    resets = 0
    for t in range(T):
        act = np.stack([rng.uniform(-0.4189, 0.4189, (N, A)), rng.uniform(4, 8, (N, A))], axis=2)
        sim.tick(act, env_level=True, autoreset_poses=wp, pose_gap=(2 if t < 15 else 23))
        resets += int(sim.done.sum().item())
    torch.cuda.synchronize()
    print('variant', variant, 'envs', N, 'ticks', T, 'episodes ended', resets, 'scan checksum %.3f' % float(sim.scans.double().sum()))
PY
for tool in memcheck racecheck synccheck; do
  timeout 600 compute-sanitizer --tool $tool python /tmp/f110_sanitize.py > $O/sanitize_$tool.log 2>&1
  tail -4 $O/sanitize_$tool.log
done
