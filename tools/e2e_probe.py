"""Where does the end-to-end tick spend its time?  Times, for one workload, the eager tick, the host-buffer tick with
and without the scan copy, and raw pinned D2H copies of the same sizes, all in one process.  GPU box only."""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import f1tenth_gym_b200 as f110

N, A, B = [int(x) for x in (sys.argv[1:4] if len(sys.argv) > 3 else (16384, 2, 270))]
dev = torch.device('cuda', 0)
NA = N * A
sim = f110.Simulator(f110.maps.DEFAULT_PARAMS, A, 12345, num_envs=N, num_beams=B, device=dev)
sim.set_map(f110.maps.resolve_map_path('example_map'), '.png')
wp_np = f110.maps.load_waypoints()
wp = torch.from_numpy(wp_np).to(dev)
ks = np.random.default_rng(1).integers(0, wp_np.shape[0], N)
sim.env_reset(np.stack([wp_np[(ks - 15 * i) % wp_np.shape[0]] for i in range(A)], axis=1))
act = torch.zeros((N, A, 2), dtype=torch.float64, device=dev)
act[..., 1] = 3.0


def timed(name, fn, n=30):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    ms = 1e3 * (time.perf_counter() - t) / n
    print('%-44s %8.3f ms' % (name, ms), flush=True)
    return ms


def tick():
    sim.tick(act)
    sim.autoreset(wp, 15, 1)


timed('eager f110_tick + autoreset', tick)
io0 = sim.make_host_io(with_scans=False)
io1 = sim.make_host_io(with_scans=True)
timed('f110_step_host, no scans', lambda: (sim.step_host(io0), sim.autoreset(wp, 15, 1)))
ms = timed('f110_step_host, scans', lambda: (sim.step_host(io1), sim.autoreset(wp, 15, 1)))
print('   scans %.1f MB -> %.1f GB/s if all of it were the copy' % (NA * B * 4 / 1e6, NA * B * 4 / 1e6 / ms))
h = io1['scans']
print('pinned', h.is_pinned(), 'ptr %% 4096 = %d' % (h.data_ptr() % 4096))
ms = timed('raw D2H sim.scans -> io scans (torch copy_)', lambda: (h.copy_(sim.scans.view(NA, B), non_blocking=True), torch.cuda.synchronize()))
print('   %.1f GB/s' % (NA * B * 4 / 1e6 / ms))
h2 = torch.empty((NA, B), dtype=torch.float32, pin_memory=True)
ms = timed('raw D2H sim.scans -> fresh pinned', lambda: (h2.copy_(sim.scans.view(NA, B), non_blocking=True), torch.cuda.synchronize()))
print('   %.1f GB/s' % (NA * B * 4 / 1e6 / ms))
sets = sim.make_host_pipeline(depth=2)


def pipe(n):
    for t in range(n):
        io = sets[t % 2]
        sim.wait_host(io)
        sim.step_host_async(io)
        sim.autoreset(wp, 15, 1)
    for io in sets:
        sim.wait_host(io)


pipe(6)
torch.cuda.synchronize()
t = time.perf_counter()
pipe(60)
torch.cuda.synchronize()
ms = 1e3 * (time.perf_counter() - t) / 60
print('%-44s %8.3f ms  (%.2f M agent-steps/s)' % ('pipelined step_host_async', ms, NA / ms / 1e3))
