"""Generated tracks (reference unittest/random_trackgen.py).

CPU: the centerline generator against centerlines produced by the unmodified reference
(tests/golden/trackgen_centerlines.npz, made by tests/golden/make_golden_trackgen.py), and the frame conventions.
GPU: the wall rasteriser against a numpy restatement (bit-exact), the rasterise -> EDT -> DeviceMap pipeline against
scipy, stepping on stacked generated tracks against the oracle, and a closed-loop two-lap drive.
"""
import os

import numpy as np
import pytest

from f1tenth_gym_b200 import trackgen as tg

G = os.path.join(os.path.dirname(__file__), 'golden')


def test_centerlines_match_reference():
    g = np.load(os.path.join(G, 'trackgen_centerlines.npz'))
    failures = 0
    for s in g['seeds']:
        rng = np.random.RandomState(int(s))              # same stream as the reference's np.random.seed(s)
        for c in range(int(g['calls'])):
            xy = tg.create_track(rng)
            ref = g['seed%d_call%d' % (s, c)]
            if ref.shape[0] == 0:
                assert xy is None                        # the reference returned False on this attempt
                failures += 1
            else:
                assert xy.shape == ref.shape and np.array_equal(xy, ref), (s, c)
    assert failures >= 1                                 # the failure branch is pinned too


def test_global_numpy_stream_is_the_default():
    g = np.load(os.path.join(G, 'trackgen_centerlines.npz'))
    state = np.random.get_state()
    try:
        np.random.seed(123)
        assert np.array_equal(tg.create_track(), g['seed123_call0'])
    finally:
        np.random.set_state(state)


def test_track_frames():
    a, b = tg.random_tracks(7, 2)
    c = tg.random_tracks(7, 2)
    assert np.array_equal(a.centerline, c[0].centerline) and np.array_equal(b.centerline, c[1].centerline)
    assert a.origin == b.origin == (-40.0, -50.0, 0.0)                     # one shared map frame
    # world -> pixel through the map transform (laser_models.py:75-113 with orig_c = 1) lands on the raster point
    px = (a.waypoints - np.array(a.origin[:2])) / a.resolution
    assert np.abs(px - a.pixels).max() < 1e-9
    assert a.pixels.min() > 0 and a.pixels.max() < a.canvas
    assert abs(a.half_width - 10.0 * (1600 / 600) * 0.0625) < 1e-12
    f = tg.Track(a.centerline, origin='first')                             # the reference's convention (:196-200)
    assert np.array_equal(f.waypoints[0], [0.0, 0.0])
    assert np.allclose(f.waypoints - f.waypoints[0], a.waypoints - a.waypoints[0], atol=1e-12)
    seg = a.segments()
    assert seg.shape == (a.pixels.shape[0], 5)
    assert np.allclose(seg[:-1, 0:2] + seg[:-1, 2:4], a.pixels[1:]) and np.allclose(seg[-1, 0:2] + seg[-1, 2:4], a.pixels[0])
    # consecutive centerline points are TRACK_DETAIL_STEP apart (closing segment: at most two steps)
    d = np.linalg.norm(np.diff(a.centerline, axis=0), axis=1)
    assert np.abs(d - tg.TRACK_DETAIL_STEP).max() < 1e-9
    with pytest.raises(ValueError):
        tg.Track(a.centerline, origin='nowhere')


def raster_numpy(track, H, W, lo, hi):
    """The kernel's arithmetic, operation for operation, in numpy (fp64, no contraction)."""
    seg = track.segments()
    px = (np.arange(W, dtype=np.float64) + 0.5)[None, :]
    py = (np.arange(H, dtype=np.float64) + 0.5)[:, None]
    best = np.full((H, W), 1.0e300)
    for ax, ay, bx, by, inv in seg:
        dx, dy = px - ax, py - ay
        t = np.minimum(np.maximum((dx * bx + dy * by) * inv, 0.0), 1.0)
        qx, qy = dx - t * bx, dy - t * by
        best = np.minimum(best, qx * qx + qy * qy)
    return ((best >= lo * lo) & (best <= hi * hi)).astype(np.uint8), best


@pytest.mark.gpu
def test_rasterizer_matches_numpy():
    import torch
    dev = torch.device('cuda:0')
    t = tg.Track(tg.random_tracks(123, 1)[0].centerline, canvas=400)       # quarter-size canvas keeps numpy quick
    occ, d2 = tg.rasterize(t, dev, want_dist2=True)
    lo, hi = t.wall_band()
    ref_occ, ref_d2 = raster_numpy(t, 400, 400, lo, hi)
    assert np.array_equal(d2.cpu().numpy(), ref_d2)
    assert np.array_equal(occ.cpu().numpy(), ref_occ)
    # two closed wall rings, ~ line_width wide: wall area ~ 2 * track length * line width
    length = np.linalg.norm(np.roll(t.pixels, -1, 0) - t.pixels, axis=1).sum()
    assert 0.8 < ref_occ.sum() / (2 * length * tg.LINE_WIDTH_PX) < 1.2


@pytest.mark.gpu
def test_device_map_pipeline_matches_scipy():
    import torch
    from scipy.ndimage import distance_transform_edt
    import f1tenth_gym_b200 as f110
    dev = torch.device('cuda:0')
    t = tg.random_tracks(2024, 1)[0]
    occ = tg.rasterize(t, dev)
    dm = tg.device_map(t, dev)
    ref = t.resolution * distance_transform_edt(1 - occ.cpu().numpy())
    assert np.array_equal(dm.dt.cpu().numpy(), ref)
    assert np.array_equal(dm.dt_cells.cpu().numpy(), ref / t.resolution)
    assert dm.host.fast_path == 1 and dm.host.dt_oob == ref[-1, -1]
    # on the centerline the nearest wall is half a track width minus half a line away (to within a pixel diagonal)
    px = np.floor(t.pixels).astype(int)
    on_line = ref[px[:, 1], px[:, 0]]
    expect = t.half_width - 0.5 * tg.LINE_WIDTH_PX * t.resolution
    assert np.abs(on_line - expect).max() < 2.0 * t.resolution
    # and a scan from the start pose sees walls at about that distance sideways
    ss = f110.ScanSimulator2D(1080, 4.7, device=dev)
    ss.set_device_map(dm)
    scan = ss.scan(t.start_pose(5)).cpu().numpy()[0]
    side = scan[[1080 // 2 - 361, 1080 // 2 + 361]]                        # +-pi/2 beams
    assert np.all(np.abs(side - expect) < 0.35)


@pytest.mark.gpu
def test_stacked_generated_tracks_vs_oracle():
    """Three generated tracks as one multi-map batch; every env steps exactly like an oracle Simulator built on the
    table of its own track."""
    import torch
    import oracle
    import f1tenth_gym_b200 as f110
    dev = torch.device('cuda:0')
    tracks = tg.random_tracks(1, 3)
    stacked, layers = tg.device_maps(tracks, dev)
    N, A, T = 9, 2, 40
    ids = np.arange(N) % 3
    rng = np.random.default_rng(5)
    poses = np.zeros((N, A, 3))
    for e in range(N):
        t = tracks[ids[e]]
        k = int(rng.integers(0, t.waypoints.shape[0]))
        for a in range(A):
            poses[e, a] = t.start_pose((k - 6 * a) % t.waypoints.shape[0])
    sim = f110.Simulator(f110.maps.DEFAULT_PARAMS, A, 3, num_envs=N, device=dev, count_lookups=True)
    sim.set_device_map(stacked, env_map_ids=ids)
    sim.reset(poses)
    omaps = [oracle.OracleMap(l.dt.cpu().numpy(), t.resolution, t.origin) for l, t in zip(layers, tracks)]
    osims = [oracle.OracleSim(omaps[ids[e]], num_agents=A) for e in range(N)]
    for e in range(N):
        osims[e].reset(poses[e])
    worst_state = worst_scan = 0.0
    for _ in range(T):
        act = np.stack([rng.uniform(-0.4189, 0.4189, (N, A)), rng.uniform(0, 8, (N, A))], axis=2)
        obs = sim.step(act)
        st = sim.state.cpu().numpy().reshape(7, N, A).transpose(1, 2, 0)
        sc = obs['scans'].cpu().numpy().astype(np.float64)
        for e in range(N):
            osims[e].step(act[e])
            worst_state = max(worst_state, np.abs(st[e] - osims[e].state).max())
            worst_scan = max(worst_scan, np.abs(sc[e] - osims[e].scans).max())
            assert np.array_equal(obs['collisions'].cpu().numpy()[e], osims[e].collisions)
    assert worst_state < 1e-9 and worst_scan < 4e-6, (worst_state, worst_scan)
    assert sim.lookups() == sum(o.nlook for o in osims)
    assert float(obs['scans'].min()) > 0.0 and float(obs['scans'][:, :, 540].max()) < 30.0      # walls ahead of every car


@pytest.mark.gpu
def test_closed_loop_two_laps_on_generated_track():
    """Rasterised walls + device EDT + lap logic + planner, closed loop on the device: pure pursuit on the generated
    centerline drives two clean laps."""
    import torch
    import f1tenth_gym_b200 as f110
    dev = torch.device('cuda:0')
    t = tg.random_tracks(123, 1)[0]
    dm = tg.device_map(t, dev)
    N = 4
    starts = np.stack([t.start_pose(k)[None] for k in (0, 60, 140, 220)])                 # (N, 1, 3)
    sim = f110.Simulator(f110.maps.DEFAULT_PARAMS, 1, 3, num_envs=N, device=dev)
    sim.set_device_map(dm)
    sim.env_reset(starts)
    pl = f110.PurePursuitPlanner(device=dev, waypoints=t.raceline(speed=5.0), xind=0, yind=1, vind=2)
    obs = sim.observations()
    lap_len = np.linalg.norm(np.roll(t.waypoints, -1, 0) - t.waypoints, axis=1).sum()
    max_ticks = int(2.6 * lap_len / 5.0 / 0.01)
    for k in range(max_ticks):
        obs = sim.tick(pl.plan_actions(obs, 1.2, 1.0))
        if k % 200 == 199 and bool(sim.done.all()):
            break
    assert bool(sim.done.all()), sim.lap_counts.cpu().numpy()
    assert float(sim.collisions.sum()) == 0.0
    assert np.all(sim.lap_counts.cpu().numpy() >= 2.0)
    laps = sim.lap_times.cpu().numpy()
    assert np.all(np.abs(laps - 2 * lap_len / 5.0) < 0.15 * 2 * lap_len / 5.0), laps


@pytest.mark.gpu
def test_closed_loop_on_stacked_tracks_with_per_env_waypoint_tables():
    """Domain randomisation end to end on the device: three generated tracks in one multi-map batch, every env driven
    by pure pursuit on the centerline of ITS track (f110_pure_pursuit_tables), two clean laps everywhere."""
    import torch
    import f1tenth_gym_b200 as f110
    dev = torch.device('cuda:0')
    tracks = tg.random_tracks(11, 3)
    stacked, _ = tg.device_maps(tracks, dev)
    N = 6
    ids = np.arange(N) % 3
    starts = np.stack([tracks[ids[e]].start_pose(40 * e)[None] for e in range(N)])
    sim = f110.Simulator(f110.maps.DEFAULT_PARAMS, 1, 3, num_envs=N, device=dev)
    sim.set_device_map(stacked, env_map_ids=ids)
    sim.env_reset(starts)
    pl = f110.PurePursuitPlanner(device=dev, waypoints=[t.raceline(speed=5.0) for t in tracks], xind=0, yind=1, vind=2)
    obs = sim.observations()
    # the multi-table kernel is the single-table kernel on each subset
    multi = pl.plan_actions(obs, 1.2, 1.0, table_ids=ids).clone()
    for k, t in enumerate(tracks):
        single = f110.PurePursuitPlanner(device=dev, waypoints=t.raceline(speed=5.0), xind=0, yind=1, vind=2)
        assert torch.equal(single.plan_actions(obs, 1.2, 1.0)[ids == k], multi[ids == k])
    with pytest.raises(ValueError):
        pl.plan_actions(obs, 1.2, 1.0)
    longest = max(np.linalg.norm(np.roll(t.waypoints, -1, 0) - t.waypoints, axis=1).sum() for t in tracks)
    tid = torch.as_tensor(ids, device=dev)
    for k in range(int(2.6 * longest / 5.0 / 0.01)):
        obs = sim.tick(pl.plan_actions(obs, 1.2, 1.0, table_ids=tid))
        if k % 200 == 199 and bool(sim.done.all()):
            break
    assert bool(sim.done.all()), sim.lap_counts.cpu().numpy()
    assert float(sim.collisions.sum()) == 0.0
    assert np.all(sim.lap_counts.cpu().numpy() >= 2.0)
