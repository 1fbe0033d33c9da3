"""ilqc_b200: B200-native batched iLQR/DDP engine, drop-in for the data-parallel hot path of vroulet/ilqc.

    from ilqc_b200.envs import Pendulum, CartPendulum, Car, make_env
    from ilqc_b200.algorithms import run_min_algo, run_min_algo_batched
    from ilqc_b200.model_predictive_control import mpc_step, run_mpc_batched

Every numerical routine runs in hand-written sm_100a CUDA kernels (ilqc_b200/csrc) behind a C ABI
(include/ilqc_b200.h).  There is no CPU fallback.
"""
__version__ = "0.1.0"
