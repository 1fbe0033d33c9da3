"""GPU parity tests (`-m gpu`): the CUDA library through its C ABI against (1) golden vectors of the
unmodified reference, (2) the oracle on seeded inputs, (3) size-independent properties at full batch size."""
import numpy as np
import pytest
import torch

from tests import parity_checks as pc

pytestmark = pytest.mark.gpu

LIN = [n for n in pc.golden_names("lin_") if pc.is_supported(pc.golden(n))]
RUNS = [n for n in pc.golden_names("run_") if pc.is_supported(pc.golden(n))]
CHAOTIC = {
    "run_scar_euler_exact_h50_ddp_quad_reg",
    # the reference's own line search fails here (330-770 trips down to s ~ 1e-32, "line-search failed"): whether
    # two consecutive costs come out bit-identical ('converged') is rounding noise
    "run_bcar_h25_ddp_linquad_dir",
}


def test_cuda_library_loaded(cuda_engine):
    from ilqc_b200 import _lib
    assert cuda_engine.device.type == "cuda"
    assert cuda_engine.lib.ilqc_abi_version() == _lib.ABI_VERSION
    assert not hasattr(cuda_engine.lib, "ilqc_is_hostsim_marker")
    assert _lib.CUDA_LIB_PATH.endswith("libilqc_b200.so")


@pytest.mark.parametrize("name", LIN)
def test_expansion(cuda_engine, name):
    pc.check_expansion(cuda_engine, name)


@pytest.mark.parametrize("name", LIN)
def test_backward_rollout(cuda_engine, name):
    pc.check_backward_rollout(cuda_engine, name)


@pytest.mark.parametrize("name", RUNS)
def test_run_teacher_forced(cuda_engine, name):
    pc.check_run_teacher_forced(cuda_engine, name)


@pytest.mark.parametrize("name", [n for n in RUNS if n not in CHAOTIC])
def test_run_free(cuda_engine, name):
    pc.check_run_free(cuda_engine, name)


def test_baseline_config1_cached_file(cuda_engine):
    """BASELINE config 1: results/dt_2.00e-02_env_pendulum_horizon_100/algo_ddp_linquad_reg_max_iter_20.torch"""
    from ilqc_b200.envs import Pendulum
    g = pc.golden("cached_pend_h100_ddp_linquad_reg_20")
    r = cuda_engine.solve(Pendulum(horizon=100, dt=0.02), "ddp_linquad_reg", 20)
    assert pc.relerr(r.cost[0], g["cost"]) < 1e-9
    assert pc.relerr(r.cmd[0], g["cmd_opt"]) < 1e-9


def _bicycle_batch(B, H=100, dt=0.01, seed=1236):
    from ilqc_b200.envs import make_env
    env = make_env(dict(env="real_car", track="simple", cost="contouring", discretization="rk4_cst_ctrl",
                        horizon=H, dt=dt))
    g = torch.Generator().manual_seed(seed)
    xi = torch.randn(B, 4, generator=g, dtype=torch.float64)
    x0 = env.init_state.reshape(1, -1).repeat(B, 1)
    x0[:, 3] = 1.0 + 0.01 * xi[:, 0]
    x0[:, 7] = x0[:, 3]
    w = torch.tensor([0.1, 10.0, 0.1], dtype=torch.float64).reshape(1, 3) + 1e-3 * xi[:, 1:]
    return env, x0, w


def test_batch_vs_oracle_seeded(cuda_engine):
    """Seeded per-problem inputs (SURVEY.md 8d cfg4 generator) at H=25: oracle on 2 problems vs the same rows
    of a 4096-problem batch, 2 iterations."""
    from oracle import ilqc_oracle as orc
    env, x0, w = _bicycle_batch(4096, H=25, dt=0.04)
    r = cuda_engine.solve(env, "ddp_linquad_reg", 2, x0=x0, weights=w)
    for b in (0, 4095):
        oenv = orc.make_env(dict(env="real_car", track="simple", cost="contouring", discretization="rk4_cst_ctrl",
                                 horizon=25, dt=0.04, reg_cont=float(w[b, 0]), reg_lag=float(w[b, 1]),
                                 reg_speed=float(w[b, 2])))
        oenv.init_state = x0[b].clone()
        cmd, _, m = orc.run_min_algo(oenv, max_iter=2, algo="ddp_linquad_reg")
        assert pc.relerr(r.cost[b], np.array(m["cost"])) < 1e-9
        assert pc.relerr(r.cmd[b], cmd.numpy()) < 1e-9


@pytest.mark.parametrize("algo", ["ddp_linquad_reg", "classic_linquad_reg", "ddp_quad_reg", "ddp_linquad_dir", "gd"])
def test_candidate_count_invariance(cuda_engine, algo):
    """The L parallel line-search candidates must select exactly what the sequential search selects: results
    are bit-identical for L = 1, 3, 8 (size-independent property of the line-search state machine)."""
    env, x0, w = _bicycle_batch(512, H=25, dt=0.04)
    ref = cuda_engine.solve(env, algo, 3, x0=x0, weights=w, n_candidates=1, record_trips=True)
    for L in (3, 8):
        r = cuda_engine.solve(env, algo, 3, x0=x0, weights=w, n_candidates=L, record_trips=True)
        # (cost rows of problems that stopped early are NaN padding: compare with NaN == NaN)
        same = lambda a, b: torch.equal(torch.nan_to_num(a, nan=-7.0), torch.nan_to_num(b, nan=-7.0))
        assert same(r.cmd, ref.cmd) and same(r.cost, ref.cost) and same(r.stepsize, ref.stepsize)
        assert torch.equal(r.ls_trips, ref.ls_trips) and torch.equal(r.status, ref.status)


@pytest.mark.parametrize("algo", ["ddp_linquad_reg", "classic_linquad_reg", "ddp_quad_reg", "gd"])
def test_schedulers_agree(cuda_engine, algo):
    """Lock-step (scheduler=1) and asynchronous (scheduler=2, two CUDA streams) drivers: bit-identical results."""
    env, x0, w = _bicycle_batch(2048, H=25, dt=0.04)
    a = cuda_engine.solve(env, algo, 4, x0=x0, weights=w, record_trips=True, scheduler=1)
    b = cuda_engine.solve(env, algo, 4, x0=x0, weights=w, record_trips=True, scheduler=2)
    same = lambda p, q: torch.equal(torch.nan_to_num(p, nan=-7.0), torch.nan_to_num(q, nan=-7.0))
    assert same(a.cmd, b.cmd) and same(a.cost, b.cost) and same(a.norm_grad_obj, b.norm_grad_obj)
    assert same(a.stepsize_reported, b.stepsize_reported) and same(a.stepsize, b.stepsize)
    assert torch.equal(a.ls_trips, b.ls_trips) and torch.equal(a.status, b.status) and torch.equal(a.n_done, b.n_done)


def test_full_size_properties(cuda_engine):
    """BASELINE config 4 shape on one GPU (B=32768, H=100, RK4, contouring, ddp_linquad_reg): monotone costs
    (accepted steps satisfy the line-search inequality), batch-composition independence (row b of the big
    batch == the same problem solved in a batch of 8), finite outputs."""
    env, x0, w = _bicycle_batch(32768)
    r = cuda_engine.solve(env, "ddp_linquad_reg", 3, x0=x0, weights=w)
    c = r.cost
    assert torch.isfinite(c).all() and torch.isfinite(r.cmd).all()
    # (not monotone in general: linquad sub-problems may be non-convex and are accepted, backward.py:235-237)
    assert (c[:, -1] < c[:, 0]).float().mean() > 0.99
    idx = torch.tensor([0, 1, 77, 4096, 12345, 20000, 32766, 32767])
    r8 = cuda_engine.solve(env, "ddp_linquad_reg", 3, x0=x0[idx], weights=w[idx])
    assert torch.equal(r8.cost.cpu(), c[idx.to(c.device)].cpu())
    assert torch.equal(r8.cmd.cpu(), r.cmd[idx.to(c.device)].cpu())


def test_empty_and_edge_inputs(cuda_engine):
    """Error behaviour of the C ABI: bad arguments are reported as codes, nothing is launched."""
    from ilqc_b200 import _lib
    from ilqc_b200.envs import Pendulum, Car
    env = Pendulum(horizon=5)
    with pytest.raises(_lib.IlqcError):
        cuda_engine.solve(env, "ddp_linquad_reg", 1, n_candidates=0)           # ILQC_E_ARG, nothing launched
    with pytest.raises(NotImplementedError):
        cuda_engine.solve(env, "ddp_linquad_reg", 1, handle_bad_dir="modify_hess")   # dead in the reference too
    with pytest.raises(_lib.IlqcError):
        cuda_engine.solve(Car(model="real"), "ddp_linquad_reg", 1, scheduler=2, handle_bad_dir="flag", n_candidates=-3)
    r = cuda_engine.solve(env, "ddp_linquad_reg", 0)       # zero iterations: metrics row 0 only
    assert r.cost.shape == (1, 1) and torch.isfinite(r.cost).all()
    r = cuda_engine.solve(Pendulum(horizon=1), "ddp_linquad_reg", 2)   # shortest horizon
    done = int(r.n_done[0])            # (one step cannot move the pendulum: the cost repeats, status 'converged',
    assert torch.isfinite(r.cost[0, :done + 1]).all()   # later rows are NaN padding)
    assert torch.isnan(r.cost[0, done + 1:]).all()


def test_solve_host_matches_device_path(cuda_engine):
    """ilqc_solve_host (HOST buffers in the reference's layout) == device-resident ilqc_solve."""
    env, x0, w = _bicycle_batch(256, H=25, dt=0.04)
    r = cuda_engine.solve(env, "ddp_linquad_reg", 2, x0=x0, weights=w)
    x0h = x0.clone().pin_memory()
    cmdh = torch.zeros(256, 25, 3, dtype=torch.float64).pin_memory()
    steph = torch.full((256,), 100.0, dtype=torch.float64).pin_memory()
    rh = cuda_engine.solve_host(env, "ddp_linquad_reg", 2, x0h, cmdh, steph, weights=w)
    assert torch.equal(rh.cost, r.cost.cpu()) and torch.equal(cmdh, r.cmd.cpu()) and torch.equal(steph, r.stepsize.cpu())
    # cmd0_host = NULL: zeros created on the device (run_min_algo(prev_cmd=None)), separate output buffer
    out = torch.full((256, 25, 3), 7.0, dtype=torch.float64).pin_memory()
    steph.fill_(100.0)
    rz = cuda_engine.solve_host(env, "ddp_linquad_reg", 2, x0h, out, steph, weights=w, cmd0_host=None)
    assert torch.equal(rz.cost, r.cost.cpu()) and torch.equal(out, r.cmd.cpu())


def test_lq_synth_dense_riccati_vs_dense_newton(cuda_engine):
    pc.check_lq_synth(cuda_engine)


# ---- BASELINE.json configs 2, 3, 5 and the nominal algorithm of config 4 at their full per-GPU sizes ------------------
def _rows_match_alone(eng, env, algo, n_iter, x0, idx, weights=None, **kw):
    """Batch-composition independence: rows `idx` of the big batch == the same problems solved in a small batch."""
    r = eng.solve(env, algo, n_iter, x0=x0, weights=weights, **kw)
    rs = eng.solve(env, algo, n_iter, x0=x0[idx], weights=None if weights is None else weights[idx], **kw)
    dev_idx = idx.to(r.cost.device)
    same = lambda a, b: torch.equal(torch.nan_to_num(a, nan=-7.0), torch.nan_to_num(b, nan=-7.0))
    assert same(rs.cost, r.cost[dev_idx]) and same(rs.cmd, r.cmd[dev_idx]) and torch.equal(rs.status, r.status[dev_idx])
    return r


@pytest.mark.parametrize("algo", ["classic_linquad_reg", "ddp_linquad_reg"])
def test_config2_cartpole_full_batch(cuda_engine, algo):
    """BASELINE config 2: CartPendulum swing-up H=50, 16384 random initial states (pendulum.py:150-152), 20
    iterations.  Oracle parity on two rows (3 iterations), row independence, finite decreasing costs."""
    from ilqc_b200.envs import CartPendulum
    from oracle import ilqc_oracle as orc
    B = 16384
    env = CartPendulum(horizon=50, dt=0.05, stay_put_time=0.6, x_limits=(-2.0, 2.0))
    g = torch.Generator().manual_seed(1234)
    x0 = torch.zeros(B, 4, dtype=torch.float64)
    x0[:, 0] = torch.randn(B, generator=g, dtype=torch.float64)
    idx = torch.tensor([0, 1, 5000, 16383])
    r = _rows_match_alone(cuda_engine, env, algo, 20, x0, idx)
    c = r.cost
    first, done = c[:, 0], r.n_done.long()
    last = c.gather(1, done.clamp(min=0).reshape(-1, 1)).reshape(-1)
    assert torch.isfinite(first).all() and torch.isfinite(r.cmd).all()
    assert (last <= first).float().mean() > 0.99
    for b in (0, 16383):
        oenv = orc.make_env(dict(env="cart_pendulum", horizon=50, dt=0.05, stay_put_time=0.6, x_limits=(-2.0, 2.0)))
        oenv.init_state = x0[b].clone()
        cmd, _, m = orc.run_min_algo(oenv, max_iter=3, algo=algo)
        r3 = cuda_engine.solve(env, algo, 3, x0=x0[b:b + 1])
        assert pc.relerr(r3.cost[0], np.array(m["cost"])) < 1e-9 and pc.relerr(r3.cmd[0], cmd.numpy()) < 1e-9


def test_config3_simple_car_exact_full_batch(cuda_engine):
    """BASELINE config 3: simple car, exact cost, Euler H=50, ddp_linquad_reg with 8 parallel line-search
    candidates, batch 65536 (vinit_b = 1 + 0.01 xi on x0[3], x0[5], car.py:49-51, :88-90)."""
    from ilqc_b200.envs import Car
    from oracle import ilqc_oracle as orc
    B = 65536
    cfg = dict(env="simple_car", track="simple", cost="exact", discretization="euler", reg_bar=0.0, horizon=50, dt=0.04)
    env = Car(model="simple", track="simple", cost="exact", discretization="euler", reg_bar=0.0, horizon=50, dt=0.04)
    g = torch.Generator().manual_seed(1235)
    x0 = env.init_state.reshape(1, -1).repeat(B, 1)
    x0[:, 3] = 1.0 + 0.01 * torch.randn(B, generator=g, dtype=torch.float64)
    x0[:, 5] = x0[:, 3]
    idx = torch.tensor([0, 7, 40000, 65535])
    r = _rows_match_alone(cuda_engine, env, "ddp_linquad_reg", 10, x0, idx, n_candidates=8)
    assert torch.isfinite(r.cost[:, 0]).all() and torch.isfinite(r.cmd).all()
    assert (r.cost[:, 3] < r.cost[:, 0]).float().mean() > 0.99
    b = 65535
    oenv = orc.make_env(dict(cfg, vinit=float(x0[b, 3])))
    assert pc.relerr(oenv.init_state, x0[b].numpy()) < 1e-15
    cmd, _, m = orc.run_min_algo(oenv, max_iter=2, algo="ddp_linquad_reg")
    r2 = cuda_engine.solve(env, "ddp_linquad_reg", 2, x0=x0[b:b + 1], n_candidates=8)
    assert pc.relerr(r2.cost[0], np.array(m["cost"])) < 1e-9 and pc.relerr(r2.cmd[0], cmd.numpy()) < 1e-9


def test_config4_nominal_ddp_quad_reg_full_batch(cuda_engine):
    """BASELINE config 4 with its nominal algorithm (ddp_quad_reg = Newton-type DDP, second-order jets in the
    backward pass) at the per-GPU slice B = 32768, H = 100: one iteration, row independence, and the
    reference's own first iteration for the unperturbed problem (run_bcar_h100_ddp_quad_reg golden)."""
    env, x0, w = _bicycle_batch(32768)
    g = pc.golden("run_bcar_h100_ddp_quad_reg")
    x0[0] = pc.t64(g["x0"])
    w[0] = torch.tensor([0.1, 10.0, 0.1], dtype=torch.float64)
    idx = torch.tensor([0, 3, 20000, 32767])
    r = _rows_match_alone(cuda_engine, env, "ddp_quad_reg", 1, x0, idx, weights=w)
    assert torch.isfinite(r.cost).all()
    assert pc.relerr(r.cost[0], g["cost"][:2]) < 1e-9 and pc.relerr(r.cmd[0], g["cmds"][1]) < 1e-9


def test_config5_mpc_concurrent_episodes(cuda_engine):
    """BASELINE config 5 per-GPU slice: 4096 concurrent closed-loop MPC episodes on the complex track (bicycle
    model, window 40, overlap 39, 10 iterations per window), first 3 windows.  Episode 0 is the reference's
    episode (live golden window 0), the others carry per-episode vinit / weights; episodes are independent."""
    from ilqc_b200.envs import make_env
    from ilqc_b200.model_predictive_control import run_mpc_batched
    g = pc.golden("mpc_complex_live")
    cfg = pc.cfg_of(g)
    env = make_env(cfg)
    B = 4096
    gen = torch.Generator().manual_seed(1237)
    xi = torch.randn(B, 4, generator=gen, dtype=torch.float64)
    xi[0] = 0.0
    x0 = env.init_state.reshape(1, -1).repeat(B, 1)
    x0[:, 3] = x0[:, 3] + 0.01 * xi[:, 0]
    x0[:, 7] = x0[:, 3]
    w = torch.tensor([1.0, 1.0, 1.0], dtype=torch.float64).reshape(1, 3) + 1e-3 * xi[:, 1:]
    r = run_mpc_batched(env, full_horizon=42, window_size=40, overlap=39, max_iter=10, x0=x0, weights=w, engine=cuda_engine)
    assert r.horizons == [40, 41, 42] and r.cmd.shape == (B, 42, 3) and torch.isfinite(r.cmd).all()
    assert pc.relerr(r.cost[0, 0, :6], g["cost_0"][:6]) < 1e-9
    idx = torch.tensor([0, 9, 4095])
    rs = run_mpc_batched(env, full_horizon=42, window_size=40, overlap=39, max_iter=10, x0=x0[idx], weights=w[idx],
                         engine=cuda_engine)
    assert torch.equal(rs.cmd, r.cmd[idx.to(r.cmd.device)])


@pytest.mark.parametrize("seed", range(16))
def test_random_config_vs_oracle(cuda_engine, seed):
    """16 random environment configurations (4 per model family) against the oracle's autograd."""
    pc.check_random_config_vs_oracle(cuda_engine, seed)


def test_riccati_vs_dense_newton_batch(cuda_engine):
    """Dense cross-check (the idea of the reference's tests/test_auto_lin_quad.py) on a batch of bicycle problems."""
    pc.check_riccati_vs_dense_newton(cuda_engine, B=256)


def test_time_parallel_expansion(cuda_engine):
    pc.check_time_parallel_expansion(cuda_engine)


# ---- batched reference traces at SURVEY 8(d)'s sampling plan (tests/golden/make_golden_batch.py) -------------------
BATCHES = pc.golden_names("batch_")


@pytest.mark.parametrize("name", BATCHES)
def test_batch_teacher_forced(cuda_engine, name):
    """EVERY stored iterate of every problem of the fixture as one batch of one-iteration solves: the 8 seeded headline
    problems (bicycle H=100 ddp_linquad_reg, rows 0..7 of bench.py's batch) x 20 iterations, 3 x 2 of ddp_quad_reg at H=100,
    64 x 8 / 16 x 2 at H=25, 64 x 20 cart-pendulum (both algorithms) and simple car: new cost, new controls, step-size
    state, gradient norms, reported step within 1e-9 relative (a measured, reference-derived tolerance on the few rows
    where the reference itself amplifies last-bit noise beyond that, tests/golden/make_sensitivity.py) and identical
    line-search trip counts."""
    worst = pc.check_batch_teacher_forced(cuda_engine, name)
    print(name, worst)


@pytest.mark.parametrize("name", BATCHES)
def test_batch_free_running(cuda_engine, name):
    """Free-running solves of the same problems: every problem follows the reference for at least its first 3 iterations
    at 1e-9; the first diverging iteration (SURVEY section 4 statistic) is printed."""
    first = pc.check_batch_free(cuda_engine, name)
    print(name, "first diverging iteration per problem:", first.tolist())


def test_headline_rows_inside_full_batch(cuda_engine):
    """The 8 reference problems are rows 0..7 of the benchmark's own seeded batch (bench.py make_batch(32768, 1236)):
    solved INSIDE the full 32768-problem batch, 20 iterations, they are bit-identical to the same rows solved alone and
    follow the reference trace (per-iteration cost 1e-9 until the chain passes an iteration the reference itself amplifies
    rounding on; final cost 1e-6 for the chains that pass none)."""
    from bench import make_batch
    g = pc.golden("batch_bcar100")
    env, x0, w = make_batch(32768, 1236)
    rows = torch.as_tensor(g["rows"])
    assert pc.relerr(x0[rows], g["x0"]) < 1e-15 and pc.relerr(w[rows], g["weights"]) < 1e-15
    r = cuda_engine.solve(env, "ddp_linquad_reg", 20, x0=x0, weights=w)
    ra = cuda_engine.solve(env, "ddp_linquad_reg", 20, x0=x0[rows], weights=w[rows])
    dev = rows.to(r.cost.device)
    assert torch.equal(r.cost[dev], ra.cost) and torch.equal(r.cmd[dev], ra.cmd)
    err = np.abs(r.cost[dev].cpu().numpy() - g["cost"]) / np.abs(g["cost"])
    first = [next((k for k in range(21) if not err[p, k] < 1e-9), 21) for p in range(len(rows))]
    print("headline rows: first diverging iteration", first, "final cost rel err", err[:, -1])
    assert min(first) > 3
    # Free chains compound the per-iteration amplification of rounding (the teacher-forced test covers every iteration of
    # every row independently of that drift).  A chain leaves the reference's for good only through an iteration on
    # which the REFERENCE itself amplifies last-bit noise beyond the tolerance (rows of tests/golden/sensitivity.npz,
    # measured with the unmodified reference): chains that pass no such iteration end at the reference's final cost.
    sens = pc.sensitivity_rows("batch_bcar100")
    for p in range(len(rows)):
        noisy = [k for (pp, k), (sm, unstable) in sens.items() if pp == p and (unstable or sm["cost1"] > 1e-9)]
        if not noisy:
            assert err[p, -1] < 1e-6, (p, err[p])
    assert sum(f == 21 for f in first) >= len(rows) // 2


def test_coop_backward_identical_to_thread_per_candidate(cuda_engine):
    """Warp-cooperative backward pass (csrc/coop.cuh, 8 lanes per candidate) against the thread-per-candidate kernel on
    the bicycle model: identical gains, model decrease and feasibility for 300 candidates of 60 seeded problems (ragged
    last block), so results do not depend on which kernel a line-search round uses.  Then a whole solve with the
    cooperative kernel on every small round against the default solve."""
    from bench import make_batch
    eng = cuda_engine
    env, x0, w = make_batch(60, 99)
    g = torch.Generator().manual_seed(5)
    cmd = 0.3 * torch.randn(60, env.horizon, env.dim_ctrl, generator=g, dtype=torch.float64)
    ex = eng.linearize(env, x0, cmd, weights=w)
    reg = (10.0 ** torch.linspace(-3, 6, 5, dtype=torch.float64)).reshape(1, 5).repeat(60, 1)
    set_t = lambda v: _check_rc(eng.lib.ilqc_set_tuning(b"coop_max", v))
    try:
        set_t(0)
        a = eng.backward(ex, 0, reg, L=5)
        set_t(1 << 40)
        b = eng.backward(ex, 0, reg, L=5)
        assert torch.equal(a["gains"], b["gains"]) and torch.equal(a["jcst"], b["jcst"]) and torch.equal(a["feasible"], b["feasible"])
        assert torch.isfinite(a["gains"]).all() and float(a["gains"].abs().max()) > 0
        set_t(0)
        r0 = eng.solve(env, "ddp_linquad_reg", 6, x0=x0, weights=w, record_trips=True)
        set_t(1 << 40)
        r1 = eng.solve(env, "ddp_linquad_reg", 6, x0=x0, weights=w, record_trips=True)
        assert torch.equal(r0.cost, r1.cost) and torch.equal(r0.cmd, r1.cmd) and torch.equal(r0.ls_trips, r1.ls_trips)
    finally:
        _check_rc(eng.lib.ilqc_set_tuning(b"coop_max", -1))


def _check_rc(rc):
    assert rc == 0, rc
