"""GPU tests (`-m gpu`) of the reference-named Python API and of the MPC / status / restart / dense-diagnostic paths ON
THE CUDA ENGINE (these were only exercised through the host simulator).  The API functions call
`ilqc_b200.engine.default_engine()`, which is libilqc_b200.so on cuda:0 here; the checks are the same functions
tests/test_host_logic.py runs on the simulator, against golden vectors of the unmodified reference."""
import numpy as np
import pytest
import torch

from tests import api_checks as ac
from tests import parity_checks as pc

pytestmark = pytest.mark.gpu
torch.set_default_dtype(torch.float64)  # the reference's callers do the same (tests/test_newton.py:12)


@pytest.fixture(autouse=True)
def _api_on_cuda(cuda_engine):
    """The drop-in API must be running on the CUDA library (no simulator, no fallback)."""
    import ilqc_b200.engine as eng
    assert eng.default_engine() is cuda_engine and cuda_engine.device.type == "cuda"
    yield cuda_engine


def test_run_min_algo_dropin_and_restart():
    ac.check_run_min_algo_dropin_and_restart()


def test_reference_style_lowlevel_api():
    ac.check_reference_style_lowlevel_api()


@pytest.mark.parametrize("name,iters", [("run_pend_h100_ddp_linquad_reg", 4), ("run_pend_h100_classic_quad_dir", 3),
                                         ("run_cart_h50_gd", 2), ("run_pend_h100_ddp_linquad_dirvar", 3),
                                         ("run_scar_euler_exact_h50_classic_linquad_reg", 2),
                                         ("run_bcar_h25_ddp_quad_reg", 2)])
def test_min_step_chain_matches_reference_trace(name, iters):
    pc.check_min_step_chain(name, iters)


def test_reference_test_newton_with_imports_changed():
    pc.check_newton_oracle()


def test_dense_diagnostics_vs_reference():
    pc.check_dense_diagnostics()


def test_solve_ctrl_pb_incrementally_and_restart_cost0():
    pc.check_incremental_and_restart()


def test_status_cases_vs_reference(cuda_engine):
    """diverged (NaN at the start, NaN by an accepted step, cost > cost0 + 1e3), stuck, converged: metrics rows and status
    of run_min_algo against the reference (run_min_algo.py:195-264)."""
    pc.check_status_cases(cuda_engine)


def test_mpc_step_dropin_against_live_reference_windows(cuda_engine):
    ac.check_mpc_step_dropin_against_live_reference_windows(cuda_engine)


def test_mpc_cached_windows_teacher_forced():
    ac.check_mpc_cached_windows_teacher_forced()


@pytest.mark.parametrize("variant", ac.MPC_VARIANTS)
def test_mpc_batched_variants_equal_chained_mpc_step(cuda_engine, variant):
    ac.check_mpc_batched_variants_equal_chained_mpc_step(cuda_engine, variant)


MPC_EPISODES = pc.golden_names("mpcb_mpc_")
MPC_VARIANT_RUNS = pc.golden_names("mpcb_mpcvar_")


@pytest.mark.parametrize("name", MPC_EPISODES + MPC_VARIANT_RUNS)
def test_mpc_windows_teacher_forced_vs_reference(name):
    """Perturbed episodes (seeded generator of BASELINE config 5) and the MPC variants (optim_on_full_window,
    keep_applying_last_ctrl=False, sliding time > 1): every window of the reference's chained mpc_step, warm-started from
    the reference's previous window."""
    pc.check_mpc_teacher_forced(name)


def test_mpc_batched_episodes_vs_reference(cuda_engine):
    """ilqc_mpc (window loop on the device) on the reference's perturbed episodes as one batch."""
    if not MPC_EPISODES:
        pytest.skip("no mpcb_mpc_* fixtures")
    pc.check_mpc_batched_vs_reference(cuda_engine, MPC_EPISODES)


@pytest.mark.parametrize("name", MPC_VARIANT_RUNS)
def test_mpc_batched_variant_vs_reference(cuda_engine, name):
    pc.check_mpc_batched_vs_reference(cuda_engine, [name])


def test_run_mpc_single_episode(cuda_engine):
    from ilqc_b200.envs import make_env
    from ilqc_b200.model_predictive_control import run_mpc, run_mpc_batched
    g = pc.golden("mpc_complex_live")
    cfg = pc.cfg_of(g)
    optim = dict(algo="ddp_linquad_reg", max_iter=10, overlap=39, window_size=40, full_horizon=41)
    cmd = run_mpc(cfg, optim)
    r = run_mpc_batched(make_env(cfg), full_horizon=41, window_size=40, overlap=39, max_iter=10, batch=1, engine=cuda_engine, expansion=0)
    assert cmd.shape == (40, 3) and torch.equal(cmd, r.cmd[0].cpu())
