from .fhc_pipeline import solve_ctrl_pb, solve_ctrl_pb_incrementally, check_exp_done
