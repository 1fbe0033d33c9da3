"""Pendulum and CartPendulum problem families (parameters only; arithmetic in csrc/models.cuh).

Constructor arguments, defaults, initial states and the host-side integer truncation
`horizon - int(stay_put_time / dt)` follow envs/pendulum.py:13-54 and :115-176 of the reference.
"""
import math
from typing import Optional

import torch

from .. import _lib
from .forward import DiffEnv


class Pendulum(DiffEnv):
    model_id = _lib.MODEL_PENDULUM

    def __init__(self, reg_speed: float = 0.1, reg_ctrl: float = 1e-6, dt: float = 0.05, horizon: int = 40,
                 stay_put_time: Optional[float] = None, seed: int = 0):
        super().__init__()
        self.horizon, self.dt = horizon, dt
        self.dim_ctrl, self.dim_state = 1, 2
        self.discretization = "euler"
        self.init_state, self.init_time_iter = torch.tensor([math.pi, 0.0], dtype=torch.float64), 0
        if seed != 0:
            torch.manual_seed(seed)
            self.init_state[0] = self.init_state[0] + 1e0 * torch.randn(1, dtype=torch.float64)
        self.reg_speed, self.reg_ctrl = reg_speed, reg_ctrl
        self.stay_put_time_start_iter = horizon - int(stay_put_time / dt) if stay_put_time is not None else horizon
        self.g, self.m, self.l, self.mu = 10, 1.0, 1.0, 0.01

    def _fill_desc(self, d):
        d.reg_speed, d.reg_ctrl = float(self.reg_speed), float(self.reg_ctrl)
        d.stay_put_start = int(self.stay_put_time_start_iter)
        for i, v in enumerate((self.g, self.m, self.l, self.mu)):
            d.phys[i] = float(v)


class CartPendulum(DiffEnv):
    model_id = _lib.MODEL_CARTPOLE

    def __init__(self, reg_speed: float = 0.1, reg_ctrl: float = 1e-6, reg_barrier: float = 1.0, dt: float = 0.05,
                 horizon: int = 40, stay_put_time: Optional[float] = None, discretization: str = "euler",
                 x_limits: Optional[tuple] = None, seed: int = 0):
        super().__init__()
        self.horizon, self.dt, self.discretization = horizon, dt, discretization
        self.dim_ctrl, self.dim_state = 1, 4
        if discretization == "rk4":
            self.dim_ctrl *= 3
        self.init_state, self.init_time_iter = torch.zeros(4, dtype=torch.float64), 0
        if seed != 0:
            torch.manual_seed(seed)
            self.init_state[0] = self.init_state[0] + 1e0 * torch.randn(1, dtype=torch.float64)
        self.reg_speed, self.reg_ctrl, self.reg_barrier = reg_speed, reg_ctrl, reg_barrier
        self.stay_put_time_start_iter = horizon - int(stay_put_time / dt) if stay_put_time is not None else horizon
        self.g, self.M, self.m, self.b, self.I, self.l = 10, 0.5, 0.2, 0.1, 0.006, 0.3
        self.x_limits = x_limits

    def _fill_desc(self, d):
        d.reg_speed, d.reg_ctrl, d.reg_barrier = float(self.reg_speed), float(self.reg_ctrl), float(self.reg_barrier)
        d.stay_put_start = int(self.stay_put_time_start_iter)
        d.has_x_limits = int(self.x_limits is not None)
        if self.x_limits is not None:
            d.x_min, d.x_max = float(self.x_limits[0]), float(self.x_limits[1])
        for i, v in enumerate((self.g, self.M, self.m, self.b, self.I, self.l)):
            d.phys[i] = float(v)
