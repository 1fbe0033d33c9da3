from .fhc_pipeline import solve_ctrl_pb
