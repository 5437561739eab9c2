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
        self.device = torch.device(device) if device is not None else torch.device('cuda', torch.cuda.current_device())
        # batch extension: a list of waypoint tables (one per track of a multi-map batch), concatenated on the device;
        # plan*/plan_actions then take `table_ids`, the table each pose follows
        self.table_start = None
        if isinstance(waypoints, (list, tuple)):
            tables = [np.asarray(w, dtype=np.float64) for w in waypoints]
            if not tables or any(w.ndim != 2 or w.shape[0] < 2 for w in tables):
                raise ValueError('every waypoint table needs at least two rows')
            starts = np.concatenate([[0], np.cumsum([w.shape[0] for w in tables])]).astype(np.int32)
            self.table_start = torch.from_numpy(starts).to(self.device)
            self.num_tables = len(tables)
            waypoints = np.concatenate(tables, axis=0)
        self.waypoints = np.asarray(waypoints, dtype=np.float64)
        self.wx = torch.from_numpy(np.ascontiguousarray(self.waypoints[:, xind])).to(self.device)
        self.wy = torch.from_numpy(np.ascontiguousarray(self.waypoints[:, yind])).to(self.device)
        self.wv = torch.from_numpy(np.ascontiguousarray(self.waypoints[:, vind])).to(self.device)

    def _dev(self, t):
        if not torch.is_tensor(t):
            t = torch.as_tensor(np.asarray(t, dtype=np.float64))
        return t.to(device=self.device, dtype=torch.float64).reshape(-1).contiguous()

    def plan_into(self, pose_x, pose_y, pose_theta, lookahead_distance, vgain, actions_out, table_ids=None):
        """Writes (steer, speed) rows into actions_out (M,2) fp64 CUDA tensor; no host synchronisation.
        table_ids (M,) int: with a multi-table planner, the waypoint table each pose follows."""
        px, py, pt = self._dev(pose_x), self._dev(pose_y), self._dev(pose_theta)
        if self.table_start is not None:
            if table_ids is None:
                raise ValueError('this planner holds %d waypoint tables: pass table_ids' % self.num_tables)
            ids = torch.as_tensor(table_ids).to(device=self.device, dtype=torch.int32).reshape(-1).contiguous()
            if ids.numel() != px.numel():
                raise ValueError('table_ids must hold one table index per pose')
            nat.check(nat.lib().f110_pure_pursuit_tables(
                nat.ptr(self.wx), nat.ptr(self.wy), nat.ptr(self.wv), nat.ptr(self.table_start), self.num_tables,
                nat.ptr(ids), nat.ptr(px), nat.ptr(py), nat.ptr(pt), px.shape[0], float(lookahead_distance), float(vgain),
                float(self.wheelbase), float(self.max_reacquire), nat.ptr(actions_out), _stream_ptr(self.device)))
            return actions_out
        if table_ids is not None:
            raise ValueError('table_ids given to a single-table planner')
        nat.check(nat.lib().f110_pure_pursuit(nat.ptr(self.wx), nat.ptr(self.wy), nat.ptr(self.wv), self.wx.shape[0],
                                              nat.ptr(px), nat.ptr(py), nat.ptr(pt), px.shape[0],
                                              float(lookahead_distance), float(vgain), float(self.wheelbase),
                                              float(self.max_reacquire), nat.ptr(actions_out), _stream_ptr(self.device)))
        return actions_out

    def plan(self, pose_x, pose_y, pose_theta, lookahead_distance, vgain, table_ids=None):
        """Reference signature (waypoint_follow.py:204): returns (speed, steering_angle); batched tensors, or
        python floats when called with scalars."""
        scalar = not torch.is_tensor(pose_x) and np.ndim(pose_x) == 0
        M = 1 if scalar else int(np.size(pose_x) if not torch.is_tensor(pose_x) else pose_x.numel())
        out = torch.empty((M, 2), dtype=torch.float64, device=self.device)
        self.plan_into(pose_x, pose_y, pose_theta, lookahead_distance, vgain, out, table_ids=table_ids)
        if scalar:
            o = out[0].cpu().numpy()
            return float(o[1]), float(o[0])
        return out[:, 1], out[:, 0]

    def plan_actions(self, obs, lookahead_distance, vgain, table_ids=None):
        """obs from a batched env/simulator -> actions (N, A, 2) = (steer, speed).  table_ids: (N,) per env or
        (N, A) per agent, for a multi-table planner."""
        px = obs['poses_x']
        out = torch.empty((px.numel(), 2), dtype=torch.float64, device=self.device)
        if table_ids is not None:
            table_ids = torch.as_tensor(table_ids).to(self.device)
            if table_ids.dim() == 1 and px.dim() == 2:
                table_ids = table_ids[:, None].expand(px.shape[0], px.shape[1])
            table_ids = table_ids.contiguous()
        self.plan_into(px, obs['poses_y'], obs['poses_theta'], lookahead_distance, vgain, out, table_ids=table_ids)
        return out.view(tuple(px.shape) + (2,))
