"""Batched pure-pursuit planner — the on-device mirror of reference examples/waypoint_follow.py
(PurePursuitPlanner, :146-217), so that cfg1-style closed-loop rollouts run without a host round trip.

    planner = PurePursuitPlanner(conf, wheelbase)            # conf: wpt_path / wpt_delim / wpt_rowskip / wpt_*ind
    speed, steer = planner.plan(pose_x, pose_y, pose_theta, lookahead_distance, vgain)      # tensors (M,)
    actions = planner.plan_actions(obs, lookahead_distance, vgain)                          # (N, A, 2) for env.step
"""
import ctypes as C

import numpy as np
import torch

from . import _native as nat
from . import maps as hostmaps
from .simulator import _stream_ptr


class PurePursuitPlanner(object):
    def __init__(self, conf=None, wb=0.17145 + 0.15875, device=None, waypoints=None, xind=1, yind=2, vind=5):
        nat.lib()
        self.wheelbase = wb
        self.conf = conf
        self.max_reacquire = 20.      # waypoint_follow.py:154
        if waypoints is None:
            if conf is not None:
                waypoints = np.loadtxt(conf.wpt_path, delimiter=conf.wpt_delim, skiprows=conf.wpt_rowskip)
                xind, yind, vind = conf.wpt_xind, conf.wpt_yind, conf.wpt_vind
            else:
                import os
                waypoints = np.loadtxt(os.path.join(hostmaps.MAPS_DIR, 'example_waypoints.csv'), delimiter=';', skiprows=3)
        self.waypoints = np.asarray(waypoints, dtype=np.float64)
        self.device = torch.device(device) if device is not None else torch.device('cuda', torch.cuda.current_device())
        self.wx = torch.from_numpy(np.ascontiguousarray(self.waypoints[:, xind])).to(self.device)
        self.wy = torch.from_numpy(np.ascontiguousarray(self.waypoints[:, yind])).to(self.device)
        self.wv = torch.from_numpy(np.ascontiguousarray(self.waypoints[:, vind])).to(self.device)

    def _dev(self, t):
        if not torch.is_tensor(t):
            t = torch.as_tensor(np.asarray(t, dtype=np.float64))
        return t.to(device=self.device, dtype=torch.float64).reshape(-1).contiguous()

    def plan_into(self, pose_x, pose_y, pose_theta, lookahead_distance, vgain, actions_out):
        """Writes (steer, speed) rows into actions_out (M,2) fp64 CUDA tensor; no host synchronisation."""
        px, py, pt = self._dev(pose_x), self._dev(pose_y), self._dev(pose_theta)
        nat.check(nat.lib().f110_pure_pursuit(nat.ptr(self.wx), nat.ptr(self.wy), nat.ptr(self.wv), self.wx.shape[0],
                                              nat.ptr(px), nat.ptr(py), nat.ptr(pt), px.shape[0],
                                              float(lookahead_distance), float(vgain), float(self.wheelbase),
                                              float(self.max_reacquire), nat.ptr(actions_out), _stream_ptr(self.device)))
        return actions_out

    def plan(self, pose_x, pose_y, pose_theta, lookahead_distance, vgain):
        """Reference signature (waypoint_follow.py:204): returns (speed, steering_angle); batched tensors, or
        python floats when called with scalars."""
        scalar = not torch.is_tensor(pose_x) and np.ndim(pose_x) == 0
        M = 1 if scalar else int(np.size(pose_x) if not torch.is_tensor(pose_x) else pose_x.numel())
        out = torch.empty((M, 2), dtype=torch.float64, device=self.device)
        self.plan_into(pose_x, pose_y, pose_theta, lookahead_distance, vgain, out)
        if scalar:
            o = out[0].cpu().numpy()
            return float(o[1]), float(o[0])
        return out[:, 1], out[:, 0]

    def plan_actions(self, obs, lookahead_distance, vgain):
        """obs from a batched env/simulator -> actions (N, A, 2) = (steer, speed)."""
        px = obs['poses_x']
        out = torch.empty((px.numel(), 2), dtype=torch.float64, device=self.device)
        self.plan_into(px, obs['poses_y'], obs['poses_theta'], lookahead_distance, vgain, out)
        return out.view(tuple(px.shape) + (2,))
