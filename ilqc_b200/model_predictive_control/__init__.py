from .mpc_pipeline import mpc_step, run_mpc_step, run_mpc, run_mpc_batched, check_exp_fail, MpcResult
