from .mpc_pipeline import mpc_step, run_mpc_step, run_mpc_batched, MpcResult
