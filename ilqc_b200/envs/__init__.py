from .forward import DiffEnv
from .pendulum import Pendulum, CartPendulum
from .car import Car
from .choose_env import make_env
