"""Car problem families: simple kinematic and bicycle/Pacejka models (parameters only).

Constructor arguments, defaults, seeding, initial state and obstacle follow envs/car.py:27-114 of the
reference; the Pacejka / drivetrain constants are those of envs/bicycle_model.json.  Dynamics, costs and
barriers are evaluated by the kernels (csrc/models.cuh).  Plotting / rendering (car.py:304-598) is out of
scope.
"""
import math

import numpy as np
import torch

from .. import _lib
from .forward import DiffEnv
from .tracks import get_track

# envs/bicycle_model.json AS THE REFERENCE READS IT: pandas.read_json(typ="series") (car.py:104-106), whose
# fast float parser is off by one ulp from the decimal literal for Cm1, Cr0, Cr2, Dr, Bf and Iz.  The digits
# below are the doubles the reference actually computes with (tools/import_track_data.py prints them).
BICYCLE_CONSTANTS = dict(
    Cm1=0.28700000000000003, Cm2=0.0545, Cr0=0.051800000000000006, Cr2=0.00035000000000000005, Br=3.3852,
    Cr=1.2691, Dr=0.17370000000000002, Bf=2.5789999999999997, Cf=1.2, Df=0.192, m=0.041,
    Iz=2.7799999999999998e-05, lf=0.029, lr=0.033, car_l=0.06, car_w=0.03,
)

_TIME_PENALTY = {"vref_squared": _lib.TP_VREF_SQUARED, "tref": _lib.TP_TREF, "dref": _lib.TP_DREF,
                 "dref_squared": _lib.TP_DREF_SQUARED}


class Car(DiffEnv):
    def __init__(self, dt: float = 0.02, discretization: str = "rk4_cst_ctrl", horizon: int = 20,
                 track: str = "simple", model: str = "simple", reg_ctrl=1e-6, reg_cont: float = 0.1,
                 reg_lag: float = 10.0, reg_speed: float = 0.1, reg_bar: float = 100.0, reg_obs: float = 0.0,
                 cost: str = "contouring", time_penalty: str = "vref_squared", vref: float = 3.0,
                 vinit: float = 1.0, constrain_angle: float = math.pi / 3, acc_bounds=(-0.1, 1.0), seed: int = 0):
        super().__init__()
        if seed != 0:
            torch.manual_seed(seed)
            vinit = vinit + 1e-2 * torch.randn(1, dtype=torch.float64).item()
            reg_cont, reg_lag, reg_speed = (
                reg_cont + 1e-3 * torch.randn(1, dtype=torch.float64).item(),
                reg_lag + 1e-3 * torch.randn(1, dtype=torch.float64).item(),
                reg_speed + 1e-3 * torch.randn(1, dtype=torch.float64).item(),
            )
        self.dt, self.discretization, self.horizon = dt, discretization, horizon
        self.model = model
        self.vref, self.constrain_angle, self.acc_bounds = vref, constrain_angle, acc_bounds
        self.reg_ctrl, self.reg_cont, self.reg_lag = reg_ctrl, reg_cont, reg_lag
        self.reg_speed, self.reg_bar, self.reg_obs = reg_speed, reg_bar, reg_obs
        self.cost_type, self.time_penalty = cost, time_penalty
        self.track_name = track
        self.track = get_track(track)
        dx_ref, dy_ref = self.track.derivative(0.0)
        phi_ref = math.atan2(dy_ref, dx_ref)
        if model == "simple":
            self.model_id = _lib.MODEL_CAR_SIMPLE
            self.init_state = torch.tensor([0.0, 0.0, phi_ref, vinit, 0.0, vinit], dtype=torch.float64)
            self.dim_ctrl, self.dim_state = 3, 6
        elif model == "real":
            self.model_id = _lib.MODEL_CAR_BICYCLE
            self.init_state = torch.tensor([0.0, 0.0, phi_ref, vinit, 0.0, 0.0, 0.0, vinit], dtype=torch.float64)
            self.dim_ctrl, self.dim_state = 3, 8
        else:
            raise NotImplementedError
        if self.discretization == "rk4":
            self.dim_ctrl *= 3
        self.dyn_csts = dict(BICYCLE_CONSTANTS)
        self.init_time_iter = 0
        self.obstacle = self.track.evaluate(self.track.length / 20) + np.array([0.2, 0.0])

    def _fill_desc(self, d):
        if self.cost_type not in ("contouring", "exact"):
            raise NotImplementedError(self.cost_type)
        d.cost = _lib.COST_CONTOURING if self.cost_type == "contouring" else _lib.COST_EXACT
        if self.time_penalty not in _TIME_PENALTY:
            raise NotImplementedError(self.time_penalty)
        d.time_penalty = _TIME_PENALTY[self.time_penalty]
        if isinstance(self.reg_ctrl, tuple):
            d.reg_ctrl_is_tuple = 1
            d.reg_ctrl, d.reg_ctrl_tref = float(self.reg_ctrl[0]), float(self.reg_ctrl[1])
        else:
            d.reg_ctrl = float(self.reg_ctrl)
        d.reg_cont, d.reg_lag, d.reg_speed = float(self.reg_cont), float(self.reg_lag), float(self.reg_speed)
        d.reg_bar, d.reg_obs, d.vref = float(self.reg_bar), float(self.reg_obs), float(self.vref)
        d.constrain_angle = float(self.constrain_angle)
        d.acc_lo, d.acc_hi = float(self.acc_bounds[0]), float(self.acc_bounds[1])
        for k, v in self.dyn_csts.items():
            setattr(d, k, float(v))
        d.obs_x, d.obs_y = float(self.obstacle[0]), float(self.obstacle[1])
        d.track_len = float(self.track.length)
        d.n_seg = int(self.track.n_seg)
        d.termination = _lib.TERM_LOOP if self.track.loop else _lib.TERM_REPEAT_END
