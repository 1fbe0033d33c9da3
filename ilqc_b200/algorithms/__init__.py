from .run_min_algo import (run_min_algo, run_min_algo_batched, check_cvg_status, check_nomenclature,
                           compute_min_sing_eigval_jac)
from .algos_steps import (gd_oracle, newton_oracle, classic_oracle, ddp_oracle, min_step,
                          backtracking_line_search)
