import sys, json, ctypes as C
sys.path.insert(0, '.')
import numpy as np, torch
import bench, f1tenth_gym_b200 as f110
from f1tenth_gym_b200 import _native as nat
dev = torch.device('cuda:0')
L = nat.lib()
L.f110_debug_set_classes.argtypes = [C.c_uint, C.c_uint]
def run(N, A, variant, sorted_poses, one_class, ticks=25):
    L.f110_debug_set_variant(variant)
    L.f110_debug_set_classes(0 if one_class else 64, 0 if one_class else 24)
    sim = f110.Simulator(f110.maps.DEFAULT_PARAMS, A, 1, num_envs=N, device=dev)
    sim.set_map(f110.maps.resolve_map_path('example_map'), '.png')
    wp_np = f110.maps.load_waypoints()
    ks = np.array([np.random.default_rng(bench.SEED + e).integers(0, wp_np.shape[0]) for e in range(N)])
    if sorted_poses: ks = np.sort(ks)
    poses = np.stack([wp_np[(ks - bench.POSE_GAP * i) % wp_np.shape[0]] for i in range(A)], axis=1)
    sim.env_reset(poses)
    gen = torch.Generator(device=dev); gen.manual_seed(1)
    u = torch.rand((ticks + 8, N * A, 2), generator=gen, device=dev, dtype=torch.float64)
    u[..., 0] = -0.4189 + 0.8378 * u[..., 0]; u[..., 1] = 8.0 * u[..., 1]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    k = np.zeros(3)
    for t in range(ticks + 8):
        flush.zero_()
        d = sim.step_profile(u[t].view(N, A, 2))
        sim.env_post_step()          # no auto-reset: crashed cars stay where they are (they keep scanning)
        if t >= 8: k += np.array(d)
    torch.cuda.synchronize()
    return {'N': N, 'A': A, 'variant': variant, 'sorted': sorted_poses, 'one_class': one_class, 'march_us': 1e3 * k[1] / ticks, 'scan_sum': float(sim.scans.double().sum())}
for N, A in ((16384, 2), (4096, 2)):
    for variant in (66, 94, 0):
        for one_class in (False, True):
            for srt in (False, True):
                if variant == 94 and not one_class: continue
                print(json.dumps(run(N, A, variant, srt, one_class)), flush=True)
