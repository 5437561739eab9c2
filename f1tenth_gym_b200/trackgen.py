"""Random closed tracks for domain randomisation.

Host side: the centerline generator of the reference's `unittest/random_trackgen.py:56-154` (itself adapted from
CarRacing-v0), restated so that for the same numpy seed it consumes the same random draws in the same order and
returns the same centerline bit for bit (pinned by tests/golden/trackgen_centerlines.npz, generated from the
unmodified reference).  Device side: instead of shapely offsetting + matplotlib rendering + cv2 re-encoding
(:156-208) the walls are rasterised by the C ABI `f110_rasterize_track` (distance-to-centerline level set), the
distance transform is the exact device EDT (`f110_edt`), and the result is a `DeviceMap` that can be stacked into a
multi-map batch -- no image file, no host round trip of the 20 MB table.

Geometry conventions (ours; the reference's come out of matplotlib's layout engine and are not reproducible
without it): the 1600 x 1600 canvas shows y in [-300, 300] track units at 8/3 px per unit with x in [-180, 300]
centred, like the reference's figure (:170-176); resolution 0.0625 m/px (the yaml the reference writes, :213).  World
(0, 0) is by default the centre of the track annulus (track units (0, 0), canvas pixel (640, 800), map origin
(-40, -50) m exactly), so that every generated track shares one map frame and tracks can be stacked into a
multi-map batch; `origin='first'` puts it at the first centerline point like the reference (:196-200).  The
reference scales its csv by 0.05 while its yaml says 0.0625 (:200 vs :213), so its waypoints do not sit on its own
track; we use the yaml's resolution for both.
"""
import math

import numpy as np

CHECKPOINTS = 16
SCALE = 6.0
TRACK_RAD = 900 / SCALE
TRACK_DETAIL_STEP = 21 / SCALE
TRACK_TURN_RATE = 0.31
WIDTH = 10.0                     # half track width in track units (random_trackgen.py:55)

CANVAS = 1600                    # 20 in x 80 dpi (:169, :178)
X_LIM = (-180.0, 300.0)          # (:174)
Y_LIM = (-300.0, 300.0)          # (:175)
RESOLUTION = 0.0625              # (:213)
LINE_WIDTH_PX = 3.0 * 80.0 / 72.0     # linewidth=3 pt at 80 dpi (:170-171, :178)
TWO_PI = 2 * math.pi


def _draw_checkpoints(rng):
    """(:65-77) 16 (angle, x, y) gates; every gate draws its two uniforms even where they are then overridden."""
    gates = []
    start_alpha = 0.
    last = CHECKPOINTS - 1
    for c in range(CHECKPOINTS):
        alpha = TWO_PI * c / CHECKPOINTS + rng.uniform(0, TWO_PI * 1 / CHECKPOINTS)
        rad = rng.uniform(TRACK_RAD / 3, TRACK_RAD)
        if c == 0:
            alpha, rad = 0, 1.5 * TRACK_RAD
        if c == last:
            alpha, rad = TWO_PI * c / CHECKPOINTS, 1.5 * TRACK_RAD
            start_alpha = TWO_PI * (-0.5) / CHECKPOINTS
        gates.append((alpha, rad * math.cos(alpha), rad * math.sin(alpha)))
    return gates, start_alpha


def _next_gate(gates, dest_i, alpha):
    """(:97-111) first gate at or ahead of polar angle alpha, unwinding alpha by whole turns when none is."""
    n = len(gates)
    while True:
        found = False
        while True:
            g = gates[dest_i % n]
            if alpha <= g[0]:
                found = True
                break
            dest_i += 1
            if dest_i % n == 0:
                break
        if found:
            return g, dest_i, alpha
        alpha -= TWO_PI


def _drive(gates):
    """(:80-139) steer a point from gate to gate with a bounded turn rate; one (alpha, beta, x, y) per step."""
    x, y, beta = 1.5 * TRACK_RAD, 0, 0
    dest_i, laps, budget = 0, 0, 2500
    crossed = False
    trail = []
    while True:
        alpha = math.atan2(y, x)
        if crossed and alpha > 0:
            laps += 1
            crossed = False
        if alpha < 0:
            crossed = True
            alpha += TWO_PI
        (_, gx, gy), dest_i, alpha = _next_gate(gates, dest_i, alpha)
        hx, hy = math.cos(beta), math.sin(beta)
        proj = hx * (gx - x) + hy * (gy - y)
        while beta - alpha > 1.5 * math.pi:
            beta -= TWO_PI
        while beta - alpha < -1.5 * math.pi:
            beta += TWO_PI
        before = beta
        proj *= SCALE
        if proj > 0.3:
            beta -= min(TRACK_TURN_RATE, abs(0.001 * proj))
        if proj < -0.3:
            beta += min(TRACK_TURN_RATE, abs(0.001 * proj))
        x += -hy * TRACK_DETAIL_STEP
        y += hx * TRACK_DETAIL_STEP
        trail.append((alpha, before * 0.5 + beta * 0.5, x, y))
        if laps > 4:
            break
        budget -= 1
        if budget == 0:
            break
    return trail


def _closed_loop(trail, start_alpha):
    """(:141-154) the last full lap between two crossings of start_alpha, or None if head and tail do not glue."""
    i1 = i2 = -1
    i = len(trail)
    while True:
        i -= 1
        if i == 0:
            return None
        crossing = trail[i][0] > start_alpha and trail[i - 1][0] <= start_alpha
        if crossing and i2 == -1:
            i2 = i
        elif crossing and i1 == -1:
            i1 = i
            break
    lap = trail[i1:i2 - 1]
    b0 = lap[0][1]
    gap = np.sqrt(np.square(math.cos(b0) * (lap[0][2] - lap[-1][2])) +
                  np.square(math.sin(b0) * (lap[0][3] - lap[-1][3])))
    if gap > TRACK_DETAIL_STEP:
        return None
    return np.asarray([(x, y) for (_, _, x, y) in lap], dtype=np.float64)


def create_track(rng=np.random):
    """One attempt of the reference's create_track() up to the centerline: (M, 2) fp64 track units, or None where the
    reference returns False (its caller then retries with the next draws).  `rng` is anything with numpy's
    `uniform(low, high)`: the `np.random` module after `np.random.seed(s)` (what the reference uses) or a
    `np.random.RandomState(s)`, which is the same stream."""
    gates, start_alpha = _draw_checkpoints(rng)
    return _closed_loop(_drive(gates), start_alpha)


class Track(object):
    """A generated track in the three frames it is used in: `centerline` (track units), `pixels` (canvas px, row 0 =
    bottom), `waypoints` (metres, world frame with the first point at the origin)."""

    def __init__(self, centerline, canvas=CANVAS, resolution=RESOLUTION, origin='center'):
        if origin not in ('center', 'first'):
            raise ValueError("origin must be 'center' or 'first'")
        self.centerline = np.ascontiguousarray(centerline, dtype=np.float64)
        self.canvas = int(canvas)
        self.resolution = float(resolution)
        self.px_per_unit = self.canvas / (Y_LIM[1] - Y_LIM[0])
        x_pad = 0.5 * (self.canvas - (X_LIM[1] - X_LIM[0]) * self.px_per_unit)
        px = np.empty_like(self.centerline)
        px[:, 0] = x_pad + (self.centerline[:, 0] - X_LIM[0]) * self.px_per_unit
        px[:, 1] = (self.centerline[:, 1] - Y_LIM[0]) * self.px_per_unit
        self.pixels = px
        self.half_width_px = WIDTH * self.px_per_unit
        zero = px[0] if origin == 'first' else np.array([x_pad - X_LIM[0] * self.px_per_unit, -Y_LIM[0] * self.px_per_unit])
        self.origin = (-zero[0] * self.resolution, -zero[1] * self.resolution, 0.0)
        self.waypoints = (px - zero) * self.resolution

    @property
    def half_width(self):
        """Centerline-to-wall distance in metres."""
        return self.half_width_px * self.resolution

    def segments(self):
        """[M][5] table f110_rasterize_track consumes: ax, ay, bx-ax, by-ay, 1/|b-a|^2 of the closed polyline."""
        a = self.pixels
        ab = np.roll(a, -1, axis=0) - a
        len2 = ab[:, 0] * ab[:, 0] + ab[:, 1] * ab[:, 1]
        inv = np.zeros_like(len2)
        np.divide(1.0, len2, out=inv, where=len2 > 0)
        return np.ascontiguousarray(np.concatenate([a, ab, inv[:, None]], axis=1))

    def wall_band(self, line_width_px=LINE_WIDTH_PX):
        return self.half_width_px - 0.5 * line_width_px, self.half_width_px + 0.5 * line_width_px

    def headings(self):
        """Heading of the centerline at every waypoint (towards the next one)."""
        d = np.roll(self.waypoints, -1, axis=0) - self.waypoints
        return np.arctan2(d[:, 1], d[:, 0])

    def start_pose(self, index=0):
        return np.array([self.waypoints[index, 0], self.waypoints[index, 1], self.headings()[index]])

    def raceline(self, speed=4.0):
        """(M, 3) = x, y, target speed: the waypoint table PurePursuitPlanner takes."""
        return np.ascontiguousarray(np.concatenate([self.waypoints, np.full((self.waypoints.shape[0], 1), speed)], axis=1))


def random_tracks(seed, count, max_attempts=None, origin='center'):
    """`count` tracks from numpy seed `seed`.  The reference's main loop (:226-233) makes NUM_MAPS attempts and skips
    the failures; here failed attempts are retried with the next draws of the same stream until `count` succeeded."""
    rng = np.random.RandomState(seed)
    out, attempts = [], 0
    limit = max_attempts if max_attempts is not None else 20 * count + 20
    while len(out) < count:
        if attempts >= limit:
            raise RuntimeError('track generation failed %d times in a row' % attempts)
        attempts += 1
        xy = create_track(rng)
        if xy is not None:
            out.append(Track(xy, origin=origin))
    return out


def rasterize(track, device, line_width_px=LINE_WIDTH_PX, want_dist2=False):
    """Wall bitmap of a track on the device (C ABI f110_rasterize_track): uint8 CUDA tensor [canvas, canvas], 1 = wall
    (what f110_edt takes), optionally with the squared centerline distance field."""
    import torch
    from . import _native as nat
    seg = torch.from_numpy(track.segments()).to(device)
    occ = torch.empty((track.canvas, track.canvas), dtype=torch.uint8, device=device)
    d2 = torch.empty((track.canvas, track.canvas), dtype=torch.float64, device=device) if want_dist2 else None
    lo, hi = track.wall_band(line_width_px)
    nat.check(nat.lib().f110_rasterize_track(nat.ptr(seg), seg.shape[0], float(lo), float(hi), track.canvas, track.canvas,
                                             nat.ptr(occ), nat.ptr(d2), torch.cuda.current_stream(device).cuda_stream))
    return (occ, d2) if want_dist2 else occ


def device_map(track, device, line_width_px=LINE_WIDTH_PX, **kw):
    """Track -> DeviceMap without leaving the GPU: rasterise, exact EDT, cell-unit table."""
    import torch
    from . import _native as nat
    from .simulator import DeviceMap
    occ = rasterize(track, device, line_width_px)
    H, W = occ.shape
    scratch = torch.empty((H, W), dtype=torch.int32, device=device)
    dt = torch.empty((H, W), dtype=torch.float64, device=device)
    nat.check(nat.lib().f110_edt(nat.ptr(occ), H, W, track.resolution, nat.ptr(scratch), nat.ptr(dt), None,
                                 torch.cuda.current_stream(device).cuda_stream))
    return DeviceMap.from_device_dt(dt, track.resolution, track.origin, **kw)


def device_maps(tracks, device, **kw):
    """Several tracks (origin='center', i.e. one shared map frame) -> one stacked multi-map DeviceMap for
    `Simulator.set_device_map(stacked, env_map_ids)`.  Returns (stacked_map, [DeviceMap per track])."""
    from .simulator import DeviceMap
    layers = [device_map(t, device, **kw) for t in tracks]
    return DeviceMap.stack(layers), layers
