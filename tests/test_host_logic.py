"""CPU tests of the host-side mirror (no GPU, no libdlra compute calls): the driver loop semantics of
src/primitives.jl:68-104 and the data-feed protocol (update_data!, one-snapshot lookahead, Strang's two increments)."""
import numpy as np
import pytest

import lowrankintegrators.jl_b200 as lri
from lowrankintegrators.jl_b200 import api


class FakeEngine:
    """Records the C-ABI calls the mirror would make."""
    log = []

    def __init__(self, n, m, r0, rmax=None, rank_adaptive=False, device=None, force_generic=False):
        self.n, self.m, self.r = n, m, r0
        self.calls = []
        FakeEngine.log.append(self)

    def set_factors(self, U, S, V):
        self.calls.append(("set_factors", S.shape[0]))

    def get_factors(self):
        return np.zeros((self.n, self.r)), np.eye(self.r), np.zeros((self.m, self.r))

    def data_init(self, A0):
        self.calls.append(("init", float(A0[0, 0])))

    def data_push(self, A, kind=0):
        self.calls.append(("push", float(A[0, 0])))

    def step_bug(self, t=0.0, dt=1.0):
        self.calls.append(("bug", t, dt))

    def step_ksl(self, order, t=0.0, dt=1.0):
        self.calls.append(("ksl", order, t, dt))

    def step_rabug(self, tol, rmax, t=0.0, dt=1.0):
        self.calls.append(("rabug", t, dt))
        return self.r, False

    def step_greedy(self, t=0.0, dt=1.0):
        self.calls.append(("greedy", t, dt))

    def step_greedy_two_factor(self, mode, t=0.0, dt=1.0, carry_fsal=True):
        self.calls.append(("greedy2", mode, t, dt, carry_fsal))

    def rhs_set(self, *a, **k):
        self.calls.append(("rhs_set",))

    def set_substepper(self, flow, ode, nsub=1, abstol=0.0, reltol=0.0, maxiters=None):
        self.calls.append(("substepper", flow, ode, nsub))

    def sync(self):
        pass


@pytest.fixture
def fake(monkeypatch):
    FakeEngine.log = []
    monkeypatch.setattr(api, "Engine", FakeEngine)
    return FakeEngine


def _u0(n=6, m=5, r=2):
    return lri.SVDLikeRepresentation(np.eye(n)[:, :r], np.eye(r), np.eye(m)[:, :r])


def _snaps(k, n=6, m=5):
    return [np.full((n, m), float(i)) for i in range(k)]   # snapshot i is tagged by its value


def test_discrete_solve_loop_and_lookahead_pushes(fake):
    y = _snaps(5)
    sol = lri.solve(lri.MatrixDataProblem(y, _u0()), lri.UnconventionalAlgorithm())
    assert sol.t == [1, 2, 3, 4, 5] and len(sol.Y) == 5                      # t0:dt:tf with dt = 1 (primitives.jl:77-80,100-104)
    calls = fake.log[0].calls
    assert calls[1] == ("init", 0.0)                                          # yprev = y[1]
    pushes = [c[1] for c in calls if c[0] == "push"]
    assert pushes == [1.0, 2.0, 3.0, 4.0]                                     # every snapshot exactly once, in order
    # lookahead: snapshot k+2 is pushed before step k runs, except at the end of the stream
    seq = [c[0] if c[0] != "push" else ("p", c[1]) for c in calls[2:]]
    assert seq == [("p", 1.0), ("p", 2.0), "bug", ("p", 3.0), "bug", ("p", 4.0), "bug", "bug"]


def test_lookahead_can_be_disabled_and_other_integrators_push_once(fake):
    y = _snaps(4)
    lri.solve(lri.MatrixDataProblem(y, _u0()), lri.UnconventionalAlgorithm(), lookahead=False)
    seq = [c[0] for c in fake.log[0].calls[2:]]
    assert seq == ["push", "bug", "push", "bug", "push", "bug"]
    lri.solve(lri.MatrixDataProblem(y, _u0()), lri.ProjectorSplitting(lri.DualLieTrotter()))
    seq = [c[0] for c in fake.log[1].calls[2:]]
    assert seq == ["push", "ksl", "push", "ksl", "push", "ksl"]
    assert all(c[1] == lri._lib.KSL_DUAL for c in fake.log[1].calls if c[0] == "ksl")


def test_continuous_stream_step_count_and_strang_two_increments(fake):
    seen = []

    def y(t):
        seen.append(round(t, 10))
        return np.full((6, 5), t)

    sol = lri.solve(lri.MatrixDataProblem(y, _u0(), (0.0, 0.3)), lri.ProjectorSplitting(lri.Strang()), 0.1)
    assert len(sol.Y) == 4 and abs(sol.t[-1] - 0.3) < 1e-12                   # floor((tf-t0)/dt)+1 slots (primitives.jl:92-98)
    assert seen == [0.0, 0.05, 0.1, 0.15, 0.2, 0.25, 0.3]                     # y(t0), then y(t+dt/2), y(t+dt) per step
    kinds = [c[1] for c in fake.log[0].calls if c[0] == "ksl"]
    assert kinds == [lri._lib.KSL_PRIMAL, lri._lib.KSL_DUAL] * 3


def test_errors_match_the_reference(fake):
    with pytest.raises(AssertionError):   # reverse time span (projector_splitting.jl:109)
        lri.solve(lri.MatrixDataProblem(lambda t: np.zeros((6, 5)), _u0(), (1.0, 0.0)), lri.UnconventionalAlgorithm(), 0.1)
    with pytest.raises(AssertionError):   # function data needs dt (primitives.jl:78)
        lri.solve(lri.MatrixDataProblem(lambda t: np.zeros((6, 5)), _u0(), (0.0, 1.0)), lri.UnconventionalAlgorithm())
    with pytest.raises(TypeError):        # Strang on a snapshot vector: update_data! needs Int t/dt (data_integrator.jl:26)
        lri.solve(lri.MatrixDataProblem(_snaps(3), _u0()), lri.ProjectorSplitting(lri.Strang()))


def test_truncated_svd_and_rank_rule():
    rng = np.random.default_rng(0)
    A = rng.standard_normal((30, 8)) @ np.diag(10.0 ** -np.arange(8)) @ rng.standard_normal((8, 20))
    u = lri.truncated_svd(A, 3)
    assert u.rank == 3 and np.allclose(u.U.T @ u.U, np.eye(3))
    s = np.linalg.svd(A, compute_uv=False)
    r = lri.truncate_to_tolerance(s, 1e-4)
    assert np.sqrt(np.sum(s[r:] ** 2)) <= 1e-4 < np.sqrt(np.sum(s[r - 1:] ** 2))
    assert lri.truncated_svd(A, tol=1e-4).rank == r


def test_checkpoint_resume_restarts_the_stream_at_the_saved_time(fake):
    y = _snaps(6)
    prob = lri.MatrixDataProblem(y, _u0())
    integ = lri.init(prob, lri.UnconventionalAlgorithm(), 1)
    for _ in range(2):
        lri.step(integ)
    state = integ.checkpoint()
    assert state["t"] == 3 and state["iter"] == 2
    integ2 = lri.init(prob, lri.UnconventionalAlgorithm(), 1, resume=state)
    assert integ2.t == 3
    calls = fake.log[1].calls
    assert calls[1] == ("init", 2.0)          # yprev = y[3] (value 2.0), the snapshot at the checkpoint time
    lri.step(integ2)
    assert [c[1] for c in calls if c[0] == "push"][:2] == [3.0, 4.0]


def test_greedy_dispatch_on_representation_and_problem_type(fake):
    # greedy_step! has three methods (greedy_integrator.jl:72-104): (TwoFactor, Hybrid), (TwoFactor, Data), (SVDLike, Data)
    L = lri._lib
    y = _snaps(3)
    u2 = lri.TwoFactorRepresentation(np.eye(6)[:, :2], np.eye(5)[:, :2])
    sol = lri.solve(lri.MatrixDataProblem(y, u2), lri.GreedyIntegrator())
    calls = fake.log[0].calls
    assert calls[0] == ("set_factors", 2)                                     # (U, I, Z)
    assert [c[:2] for c in calls if c[0] == "greedy2"] == [("greedy2", L.GREEDY_DATA)] * 2
    assert isinstance(sol.Y[-1], lri.TwoFactorRepresentation)
    lri.solve(lri.MatrixDataProblem(y, _u0()), lri.GreedyIntegrator())
    assert [c[0] for c in fake.log[1].calls if c[0].startswith("greedy")] == ["greedy", "greedy"]
    rhs = lri.LinearRHS(A=1.0)
    prob = lri.MatrixHybridProblem(lambda t: np.full((6, 5), t), rhs, u2, (0.0, 0.3))
    lri.solve(prob, lri.GreedyIntegrator(Z_alg=lri.SubStepper("rk4", 2), fsal_carry=False), 0.1)
    calls = fake.log[2].calls
    assert ("rhs_set",) in calls and ("substepper", L.FLOW_L, L.ODE_RK4, 2) in calls
    g = [c for c in calls if c[0] == "greedy2"]
    assert len(g) == 3 and all(c[1] == L.GREEDY_HYBRID and c[4] is False for c in g)
    pushes = [c[1] for c in calls if c[0] == "push"]
    assert np.allclose(pushes, [0.1, 0.2, 0.3])                               # X = y(t + dt) (update_data!, :77)


def test_two_factor_needs_the_greedy_integrator(fake):
    u2 = lri.TwoFactorRepresentation(np.eye(6)[:, :2], np.eye(5)[:, :2])
    with pytest.raises(TypeError):
        lri.init(lri.MatrixDataProblem(_snaps(3), u2), lri.UnconventionalAlgorithm(), 1)
    with pytest.raises(TypeError):
        lri.init(lri.MatrixHybridProblem(lambda t: np.zeros((6, 5)), lri.LinearRHS(A=1.0), _u0(), (0.0, 1.0)), lri.GreedyIntegrator(), 0.1)


@pytest.mark.parametrize("tf,dt", [(1.0, 0.25), (1.0, 0.3), (0.5, 0.1), (1.0, 1e-2), (2.0, 0.7)])
def test_driver_time_grid_matches_the_oracle_restatement(fake, tf, dt):
    # primitives.jl:68-104: init_sol pre-sizes floor((tf-t0)/dt)+1 entries on a linspace, the loop runs while (tf-t)/T > 1e-8 and
    # update_sol! push!es once the pre-sized vector is full (step sizes that do not divide the span overshoot tf by one step)
    from oracle import dlra_oracle as O
    y = lambda t: np.full((6, 5), t)
    sol = lri.solve(lri.MatrixDataProblem(y, _u0(), (0.0, tf)), lri.ProjectorSplitting(lri.PrimalLieTrotter()), dt)
    X0 = O.SVDLikeRepresentation(np.eye(6)[:, :2], np.eye(2), np.eye(5)[:, :2])
    osol = O.solve(O.MatrixDataProblem(lambda t: np.zeros((6, 5)), X0, (0.0, tf)), O.ProjectorSplitting(O.PrimalLieTrotter()), dt)
    assert len(sol.Y) == len(osol.Y) and np.allclose(sol.t, osol.t, rtol=0, atol=1e-14)
    steps = [c for c in fake.log[0].calls if c[0] == "ksl"]
    assert len(steps) == len(osol.Y) - 1
    assert np.allclose([c[2] for c in steps], [dt * k for k in range(len(steps))], atol=1e-12)      # t passed to every step


def test_switching_algorithms_after_a_lookahead_step_pushes_each_snapshot_once(fake):
    # the BUG lookahead leaves y(t + 2dt)... y(t + dt) of the NEXT step already queued in the engine; a following
    # step with another algorithm must consume it instead of pushing the same snapshot again
    y = _snaps(6)
    integ = lri.init(lri.MatrixDataProblem(y, _u0()), lri.UnconventionalAlgorithm(), 1)
    lri.step(integ)                                                     # pushes y[2], y[3] (lookahead), consumes y[2]
    lri.step(integ, lri.ProjectorSplitting(lri.PrimalLieTrotter()))     # y[3] is already there
    lri.step(integ, lri.RankAdaptiveUnconventionalAlgorithm(1e-3, rmax=2))
    pushes = [c[1] for c in fake.log[0].calls if c[0] == "push"]
    assert pushes == [1.0, 2.0, 3.0]
    kinds = [c[0] for c in fake.log[0].calls if c[0] in ("bug", "ksl", "rabug")]
    assert kinds == ["bug", "ksl", "rabug"]


def test_strang_step_after_a_lookahead_step_is_refused(fake):
    # the queued snapshot is y(t + dt), the Strang step needs y(t + dt/2) first: refuse instead of silently feeding the wrong increment
    y = lambda t: np.full((6, 5), float(t))
    integ = lri.init(lri.MatrixDataProblem(y, _u0(), tspan=(0.0, 4.0)), lri.UnconventionalAlgorithm(), 1.0)
    lri.step(integ)
    with pytest.raises(RuntimeError, match="lookahead"):
        lri.step(integ, lri.ProjectorSplitting(lri.Strang()))
