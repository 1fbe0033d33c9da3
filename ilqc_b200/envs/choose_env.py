"""make_env(env_cfg): string -> problem family, as envs/choose_env.py:10-32 of the reference."""
from typing import Any

from .car import Car
from .forward import DiffEnv
from .pendulum import Pendulum, CartPendulum


def make_env(env_cfg: dict[str, Any]) -> DiffEnv:
    env = env_cfg["env"]
    env_opts = {k: env_cfg[k] for k in set(env_cfg.keys()) - {"env"}}
    if env == "pendulum":
        return Pendulum(**env_opts)
    if env == "cart_pendulum":
        return CartPendulum(**env_opts)
    if env == "real_car":
        return Car(model="real", **env_opts)
    if env == "simple_car":
        return Car(model="simple", **env_opts)
    raise NotImplementedError(env)
