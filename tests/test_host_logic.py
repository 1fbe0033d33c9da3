"""CPU tests of the host side: C-ABI exports, track fit, the reference-style Python API (through the
test-only host simulator of the kernels) and the world_size-2 sharded path on gloo."""
import os
import re
import sys

import numpy as np
import pytest
import torch

from tests import parity_checks as pc
from tests import api_checks as ac

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
torch.set_default_dtype(torch.float64)  # the reference's callers do the same (tests/test_newton.py:12)


def test_cabi_library_exports_every_declared_symbol():
    """libilqc_b200.so loads without a GPU and exports every function include/ilqc_b200.h declares."""
    import ctypes
    from ilqc_b200 import _lib, build
    path = build.build_cuda(verbose=False)
    header = open(os.path.join(ROOT, "include", "ilqc_b200.h")).read()
    declared = set(re.findall(r"\b(ilqc_[a-z0-9_]+)\s*\(", header)) - {"ilqc_model_desc", "ilqc_algo_desc"}
    assert len(declared) >= 20
    lib = ctypes.CDLL(path)
    for name in sorted(declared):
        assert hasattr(lib, name), name
    assert declared == set(_lib.PROTOTYPES), declared ^ set(_lib.PROTOTYPES)
    assert _lib.bind(path).ilqc_abi_version() == _lib.ABI_VERSION


def test_product_has_no_cpu_fallback():
    """Without a CUDA device the product engine must refuse to run (no silent CPU path)."""
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import ilqc_b200.engine as eng
    old, eng._default = eng._default, None
    try:
        with pytest.raises(RuntimeError):
            eng.default_engine()
    finally:
        eng._default = old
    src = open(os.path.join(ROOT, "ilqc_b200", "engine.py")).read() + open(os.path.join(ROOT, "ilqc_b200", "_lib.py")).read()
    assert "hostsim" not in src.replace("tests/hostsim", "") and "import oracle" not in src and "from oracle" not in src


def test_track_fit_matches_reference_tables():
    """Natural cubic spline coefficients / knots of the five tracks vs the reference's get_track()."""
    from ilqc_b200.envs.tracks import get_track
    g = pc.golden("tracks")
    for name in ("simple", "complex", "circle", "line", "bend"):
        tr = get_track(name)
        assert pc.relerr(tr.knots, g[name + "_knots"]) < 1e-14
        for s, tag in enumerate("cio"):
            for k, c in enumerate("abcd"):
                ref = g[f"{name}_{tag}_{c}"]
                assert np.max(np.abs(tr.coef[:, s, :, k] - ref)) < 1e-10 * (1 + np.max(np.abs(ref))), (name, tag, c)
        if tr.loop:
            for t, ev, de in zip(g[name + "_probe_t"], g[name + "_probe_eval"], g[name + "_probe_deriv"]):
                assert np.max(np.abs(tr.evaluate(t) - ev)) < 1e-10 and np.max(np.abs(tr.derivative(t) - de)) < 1e-9


@pytest.fixture()
def api_on_sim(sim_engine):
    """Route the reference-style API through the host simulator for the duration of a test."""
    import ilqc_b200.engine as eng
    old, eng._default = eng._default, sim_engine
    yield sim_engine
    eng._default = old


def test_run_min_algo_dropin_and_restart(api_on_sim):
    ac.check_run_min_algo_dropin_and_restart()


def test_reference_style_lowlevel_api(api_on_sim):
    ac.check_reference_style_lowlevel_api()


def test_mpc_step_dropin_against_live_reference_windows(api_on_sim):
    ac.check_mpc_step_dropin_against_live_reference_windows(api_on_sim)


@pytest.mark.parametrize("variant", ac.MPC_VARIANTS)
def test_mpc_batched_variants_equal_chained_mpc_step(api_on_sim, variant):
    ac.check_mpc_batched_variants_equal_chained_mpc_step(api_on_sim, variant)


def test_mpc_cached_windows_teacher_forced(api_on_sim):
    ac.check_mpc_cached_windows_teacher_forced()


def test_batch_rows_are_independent(sim_engine):
    """Row b of a batch == the same problem solved alone, bit for bit; L-invariance of the line search."""
    from ilqc_b200.envs import CartPendulum
    env = CartPendulum(horizon=50, dt=0.05, stay_put_time=0.6, x_limits=(-2.0, 2.0))
    g = torch.Generator().manual_seed(1234)
    x0 = torch.zeros(24, 4, dtype=torch.float64)
    x0[:, 0] = torch.randn(24, generator=g, dtype=torch.float64)
    r = sim_engine.solve(env, "ddp_linquad_reg", 5, x0=x0, n_candidates=1, record_trips=True)
    r8 = sim_engine.solve(env, "ddp_linquad_reg", 5, x0=x0, n_candidates=8, record_trips=True)
    assert torch.equal(r.cmd, r8.cmd) and torch.equal(r.cost, r8.cost) and torch.equal(r.ls_trips, r8.ls_trips)
    one = sim_engine.solve(env, "ddp_linquad_reg", 5, x0=x0[7:8])
    assert torch.equal(one.cmd[0], r.cmd[7]) and torch.equal(one.cost[0], r.cost[7])
    assert (r.cost[:, 1:] <= r.cost[:, :-1]).all()


@pytest.mark.parametrize("algo", ["ddp_linquad_reg", "classic_quad_reg", "gd"])
def test_schedulers_agree(sim_engine, algo):
    """Lock-step iterations (scheduler=1) and asynchronous per-problem state machines (scheduler=2) run the
    same per-problem arithmetic: bit-identical controls, metrics, trip counts and status words."""
    from ilqc_b200.envs import CartPendulum
    env = CartPendulum(horizon=30, dt=0.05, stay_put_time=0.6, x_limits=(-2.0, 2.0))
    g = torch.Generator().manual_seed(99)
    x0 = torch.zeros(40, 4, dtype=torch.float64)
    x0[:, 0] = torch.randn(40, generator=g, dtype=torch.float64)
    a = sim_engine.solve(env, algo, 6, x0=x0, n_candidates=4, record_trips=True, scheduler=1)
    b = sim_engine.solve(env, algo, 6, x0=x0, n_candidates=4, record_trips=True, scheduler=2)
    same = lambda p, q: torch.equal(torch.nan_to_num(p, nan=-7.0), torch.nan_to_num(q, nan=-7.0))
    assert same(a.cmd, b.cmd) and same(a.cost, b.cost) and same(a.norm_grad_obj, b.norm_grad_obj)
    assert same(a.stepsize_reported, b.stepsize_reported) and same(a.stepsize, b.stepsize)
    assert torch.equal(a.ls_trips, b.ls_trips) and torch.equal(a.status, b.status) and torch.equal(a.n_done, b.n_done)


def _gloo_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from tests.hostsim import hostsim_engine
    from ilqc_b200.envs import Pendulum
    from ilqc_b200.parallel import solve_sharded
    dist.init_process_group("gloo", rank=rank, world_size=world)
    env = Pendulum(horizon=30, dt=0.05, stay_put_time=0.5)
    g = torch.Generator().manual_seed(5)
    x0 = torch.tensor([[np.pi, 0.0]]).repeat(7, 1) + 0.3 * torch.randn(7, 2, generator=g, dtype=torch.float64)
    res = solve_sharded(hostsim_engine(), env, "ddp_linquad_reg", 4, x0)
    if rank == 0:
        q.put((res.cost.numpy(), res.cmd.numpy(), res.status.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_solve_world_size_2_gloo(sim_engine):
    """N>1 path on CPU: 2 ranks (gloo), uneven shards (7 problems), no data-path collective, final all-gather
    of the results == the single-process solve of the whole batch."""
    import torch.multiprocessing as mp
    from ilqc_b200.envs import Pendulum
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    cost, cmd, status = q.get(timeout=300)
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    env = Pendulum(horizon=30, dt=0.05, stay_put_time=0.5)
    g = torch.Generator().manual_seed(5)
    x0 = torch.tensor([[np.pi, 0.0]]).repeat(7, 1) + 0.3 * torch.randn(7, 2, generator=g, dtype=torch.float64)
    ref = sim_engine.solve(env, "ddp_linquad_reg", 4, x0=x0)
    assert np.array_equal(cost, ref.cost.numpy()) and np.array_equal(cmd, ref.cmd.numpy())
    assert np.array_equal(status, ref.status.numpy())


# ---- lower-level reference API (algorithms/algos_steps.py of the reference) ----------------------------------------
@pytest.mark.parametrize("name,iters", [("run_pend_h100_ddp_linquad_reg", 4), ("run_pend_h100_classic_quad_dir", 3),
                                         ("run_cart_h50_gd", 2), ("run_pend_h100_ddp_linquad_dirvar", 3),
                                         ("run_scar_euler_exact_h50_classic_linquad_reg", 2)])
def test_min_step_chain_matches_reference_trace(api_on_sim, name, iters):
    pc.check_min_step_chain(name, iters)


def test_reference_test_newton_with_imports_changed(api_on_sim):
    pc.check_newton_oracle()


def test_solve_ctrl_pb_incrementally_and_restart_cost0(api_on_sim):
    pc.check_incremental_and_restart()


def test_status_cases_vs_reference(api_on_sim):
    pc.check_status_cases(api_on_sim)


def test_run_mpc_single_episode(api_on_sim):
    """run_mpc(env_cfg, optim_cfg) (mpc_pipeline.py:87-115) == the batched ilqc_mpc episode, and its first window is the
    reference's (live golden)."""
    from ilqc_b200.envs import make_env
    from ilqc_b200.model_predictive_control import run_mpc, run_mpc_batched
    g = pc.golden("mpc_complex_live")
    cfg = pc.cfg_of(g)
    optim = dict(algo="ddp_linquad_reg", max_iter=10, overlap=39, window_size=40, full_horizon=41)
    cmd = run_mpc(cfg, optim)
    r = run_mpc_batched(make_env(cfg), full_horizon=41, window_size=40, overlap=39, max_iter=10, batch=1, engine=api_on_sim)
    assert cmd.shape == (40, 3) and torch.equal(cmd, r.cmd[0].cpu())
    with pytest.raises(Exception):
        run_mpc_batched(make_env(cfg), full_horizon=41, window_size=40, overlap=40, batch=1, engine=api_on_sim)   # overlap >= window


def test_dense_diagnostics_vs_reference(api_on_sim):
    pc.check_dense_diagnostics()


MPC_EPISODES = pc.golden_names("mpcb_mpc_")
MPC_VARIANT_RUNS = pc.golden_names("mpcb_mpcvar_")


@pytest.mark.parametrize("name", MPC_EPISODES + MPC_VARIANT_RUNS)
def test_mpc_windows_teacher_forced_vs_reference(api_on_sim, name):
    """Perturbed episodes of the config-5 generator and the MPC variants (optim_on_full_window, keep_applying_last_ctrl =
    False, sliding time > 1) against the REFERENCE's chained mpc_step windows (teacher forcing per window)."""
    pc.check_mpc_teacher_forced(name)


def test_mpc_batched_episodes_vs_reference(api_on_sim):
    pc.check_mpc_batched_vs_reference(api_on_sim, MPC_EPISODES)


@pytest.mark.parametrize("name", MPC_VARIANT_RUNS)
def test_mpc_batched_variant_vs_reference(api_on_sim, name):
    pc.check_mpc_batched_vs_reference(api_on_sim, [name])
