#!/bin/bash
# racecheck over the dynamic queue's ring hand-off, one run per formulation (run through gpurun from the repo root):
#   bash tools/sanitize_ring.sh  -> gpurun_out/racecheck_ring_<variant>.log
# 1024 envs x 2 agents x 3 ticks is the smallest batch that takes the dynamic queue; the ticket size is forced to 4 so that the
# variants 45 / 81 / 82 (volatile + fence / shared atomics / st.release + ld.acquire) all run the same launch.
set -u
O=gpurun_out
mkdir -p $O
cat > /tmp/f110_ring.py <<'PY'
import sys
sys.path.insert(0, '.')
import numpy as np, torch
import f1tenth_gym_b200 as f110
dev = torch.device('cuda:0')
L = f110._native.lib()
variant = int(sys.argv[1])
L.f110_debug_set_variant(variant)
dmap = f110.DeviceMap.from_yaml(f110.maps.resolve_map_path('example_map'), '.png', dev)
wp_np = f110.maps.load_waypoints()
N, A, T = 1024, 2, 3
sim = f110.Simulator(f110.maps.DEFAULT_PARAMS, A, 1, num_envs=N, device=dev)
sim.set_device_map(dmap)
rng = np.random.default_rng(0)
k = rng.integers(0, wp_np.shape[0], N)
sim.env_reset(np.stack([np.stack([wp_np[kk], wp_np[(kk - 23) % len(wp_np)]]) for kk in k]))
for t in range(T):
    act = np.stack([rng.uniform(-0.4189, 0.4189, (N, A)), rng.uniform(4, 8, (N, A))], axis=2)
    sim.tick(act, env_level=True)
torch.cuda.synchronize()
print('variant', variant, 'scan checksum %.3f' % float(sim.scans.double().sum()))
PY
for v in "$@"; do
  timeout 300 compute-sanitizer --tool racecheck python /tmp/f110_ring.py $v > $O/racecheck_ring_$v.log 2>&1
  echo "variant $v: $(grep -E 'RACECHECK SUMMARY' $O/racecheck_ring_$v.log) $(grep -E '^variant' $O/racecheck_ring_$v.log)"
done
