from .run_min_algo import run_min_algo, run_min_algo_batched, check_cvg_status, check_nomenclature
