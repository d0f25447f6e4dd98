"""TEST INFRASTRUCTURE: the trajectory QP of TrajOptimizer::populatebyrow (reference src/traj_optimizer.cpp:225-527)
restated a SECOND time, independently of oracle/dlsc_oracle.cpp and of the kernels -- in x-space, with every equality
row kept as a row (no null-space elimination, no tables) -- and solved by an independent solver (HiGHS' convex QP
solver through the binding scipy bundles).  SURVEY s8(c): CPLEX is absent, so this is the cross-check that pins the
oracle's (and the kernels') objective values.

build_qp(...) -> dict(Q, c, c0, A, lo, hi, xlo, xhi);  solve_highs(qp) -> (x, objective)
Variable order (traj_optimizer.cpp:230-232): x[k * M * P + m * P + i], k = axis, m = segment, i = control point.
"""
import math

import numpy as np

INF = float("inf")


def q_base(n, dt):
    """traj_optimizer.cpp:172-187 with phi = 3, phi_n = 1; B from include/polynomial.hpp:280-293."""
    P = n + 1
    B = np.zeros((P, P))
    for i in range(P):
        for j in range(i, P):
            B[i, j] = math.comb(n, i) * math.comb(n - i, n - j) * (-1) ** (j - i)
    cd = lambda a: a * (a - 1) * (a - 2) if a >= 3 else 0
    Z = np.zeros((P, P))
    for i in range(P):
        for j in range(P):
            if i + j - 5 > 0:
                Z[i, j] = cd(i) * cd(j) / (i + j - 5)
    return B @ Z @ B.T * dt ** -5


def terminal_segments(M, dt, goal, pos, nominal_vel):
    """getTerminalSegments_old (:543-551); (goal - pos).norm() is octomath float arithmetic widened to double."""
    d = (np.asarray(goal, np.float32) - np.asarray(pos, np.float32)).astype(np.float32)
    nsq = np.float32(np.float32(d[0] * d[0] + d[1] * d[1]) + d[2] * d[2])
    ideal = math.sqrt(float(nsq)) / nominal_vel
    return max(int((M * dt - ideal + 1e-9) / dt), 1)


def build_qp(M, n, dim, dt, w_control, w_terminal, world_min, world_max, comm_range, pos, vel, acc, goal, waypoint,
             radius, max_vel, max_acc, nominal_vel, sfc=None, lsc_normal=None, lsc_anchor=None, lsc_d=None,
             n_dyn=0, slack_weight=0.0):
    """sfc [M][6] or None; lsc_* [K][M][3] / [K][M][P][3] / [K][M][P] (only real neighbours).
    The first n_dyn obstacles are dynamic (non-agent) obstacles: M slack variables each, appended after the control
    points, epsilon <= 0 (:272-283), cost slack_weight (M - m)/M epsilon^2 (:317-331), rows n.(x - p) - d - epsilon >= 0
    (:436-448)."""
    P, phi = n + 1, 3
    nxc = dim * M * P                      # control-point variables
    nx = nxc + n_dyn * M                   # + slack variables
    vid = lambda k, m, i: k * M * P + m * P + i
    f64 = lambda a: np.asarray(a, np.float32).astype(np.float64)
    pos, vel, acc, goal, waypoint = f64(pos), f64(vel), f64(acc), f64(goal), f64(waypoint)
    Qb = q_base(n, dt)
    Q = np.zeros((nx, nx)); c = np.zeros(nx); c0 = 0.0
    for k in range(dim):                                            # :285-298
        for m in range(M):
            s = vid(k, m, 0)
            Q[s:s + P, s:s + P] += 2.0 * w_control * Qb
    ts = terminal_segments(M, dt, goal, pos, nominal_vel)
    for m in range(M - ts, M):                                       # :300-315
        for k in range(dim):
            j = vid(k, m, n)
            Q[j, j] += 2.0 * w_terminal
            c[j] += -2.0 * w_terminal * goal[k]
            c0 += w_terminal * goal[k] * goal[k]
    for oi in range(n_dyn):
        for m in range(M):
            Q[nxc + M * oi + m, nxc + M * oi + m] = 2.0 * slack_weight * (M - m) / M
    xlo = np.full(nx, -INF); xhi = np.full(nx, INF)                  # :251-265
    xhi[nxc:] = 0.0
    for k in range(dim):
        for m in range(M):
            for i in range(P):
                if not (m == 0 and i < 3):
                    xlo[vid(k, m, i)] = float(np.float32(world_min[k]))
                    xhi[vid(k, m, i)] = float(np.float32(world_max[k]))
    rows, lo, hi = [], [], []

    def add(terms, l, h):
        r = np.zeros(nx)
        for j, v in terms:
            r[j] += v
        rows.append(r); lo.append(l); hi.append(h)

    for k in range(dim):                                             # :335-366
        add([(vid(k, 0, 0), 1.0)], pos[k], pos[k])
        add([(vid(k, 0, n), 1.0), (vid(k, 1, 0), -1.0)], 0.0, 0.0)
        add([(vid(k, 0, 1), n / dt), (vid(k, 0, 0), -n / dt)], vel[k], vel[k])
        a2 = n * (n - 1) / dt ** 2
        add([(vid(k, 0, 2), a2), (vid(k, 0, 1), -2 * a2), (vid(k, 0, 0), a2)], acc[k], acc[k])
        add([(vid(k, 1, 1), 1.0), (vid(k, 1, 0), -1.0), (vid(k, 0, n), -1.0), (vid(k, 0, n - 1), 1.0)], 0.0, 0.0)
        add([(vid(k, 1, 2), 1.0), (vid(k, 1, 1), -2.0), (vid(k, 1, 0), 1.0), (vid(k, 0, n), -1.0), (vid(k, 0, n - 1), 2.0),
             (vid(k, 0, n - 2), -1.0)], 0.0, 0.0)
    A0 = np.array([[1, 0, 0, 0, 0, 0], [-1, 1, 0, 0, 0, 0], [1, -2, 1, 0, 0, 0]], float)      # buildAeqBase :189-223
    AT = np.array([[0, 0, 0, 0, 0, 1], [0, 0, 0, 0, -1, 1], [0, 0, 0, 1, -2, 1]], float)
    for k in range(dim):                                             # :369-381
        for m in range(2, M):
            nn = 1
            for j in range(phi):
                t = [(vid(k, m - 1, i), dt ** -j * nn * AT[j, i]) for i in range(P) if AT[j, i] != 0]
                t += [(vid(k, m, i), -dt ** -j * nn * A0[j, i]) for i in range(P) if A0[j, i] != 0]
                add(t, 0.0, 0.0)
                nn *= (n - j)
    if sfc is not None:                                              # :384-410
        box = f64(sfc)
        for m in range(M):
            for k in range(dim):
                for j in range(P):
                    if m == 0 and j < phi:
                        continue
                    add([(vid(k, m, j), 1.0)], box[m][k], INF)
                    add([(vid(k, m, j), -1.0)], -box[m][3 + k], INF)
    if lsc_normal is not None:                                       # :412-450
        nv, an, dd = f64(lsc_normal), f64(lsc_anchor), np.asarray(lsc_d, np.float64)
        f32n = np.asarray(lsc_normal, np.float32)
        for oi in range(nv.shape[0]):
            for m in range(M):
                # LSC::normal_vector.norm(): float sum of squares, double sqrt (octomath)
                q = f32n[oi, m]
                nsq = np.float32(np.float32(q[0] * q[0] + q[1] * q[1]) + q[2] * q[2])
                if math.sqrt(float(nsq)) < 1e-5:
                    continue
                for i in range(P):
                    if m == 0 and i < phi:
                        continue
                    rhs = dd[oi, m, i] + sum(nv[oi, m, k] * an[oi, m, i, k] for k in range(dim))
                    slack = [(nxc + M * oi + m, -1.0)] if oi < n_dyn else []
                    add([(vid(k, m, i), nv[oi, m, k]) for k in range(dim)] + slack, rhs, INF)
    for k in range(dim):                                             # :452-487
        for m in range(M):
            for i in range(n):
                if m == 0 and i in (0, 1):
                    continue
                add([(vid(k, m, i + 1), n / dt), (vid(k, m, i), -n / dt)], -max_vel, max_vel)
            for i in range(n - 1):
                if m == 0 and i == 0:
                    continue
                a2 = n * (n - 1) / dt ** 2
                add([(vid(k, m, i + 2), a2), (vid(k, m, i + 1), -2 * a2), (vid(k, m, i), a2)], -max_acc, max_acc)
    if comm_range > 0:                                               # :490-513
        for k in range(dim):
            for mi in range(M):
                for m in range(mi, M):
                    b = 0.5 * comm_range - radius
                    add([(vid(k, m, n), 1.0), (vid(k, mi, 0), -1.0)], -b, b)
        for k in range(dim):
            for m in range(M):
                b = 0.5 * comm_range - 1e-5
                add([(vid(k, m, n), 1.0)], waypoint[k] - b, waypoint[k] + b)
    for k in range(dim):                                             # :515-524
        for i in range(1, phi):
            add([(vid(k, M - 1, n), 1.0), (vid(k, M - 1, n - i), -1.0)], 0.0, 0.0)
    return dict(Q=Q, c=c, c0=c0, A=np.array(rows), lo=np.array(lo), hi=np.array(hi), xlo=xlo, xhi=xhi, ts=ts)


def solve_highs(qp, tol=None, time_limit=20.0):
    """-> (x, objective incl. the constant, model status string).  tol: HiGHS feasibility tolerances (default: HiGHS'
    own 1e-7; its active-set QP solver stalls on these problems when asked for much more)."""
    from scipy.optimize._highspy import _core as hs
    import scipy.sparse as sp
    nx = len(qp["c"])
    h = hs._Highs()
    h.setOptionValue("output_flag", False)
    h.setOptionValue("time_limit", float(time_limit))
    if tol is not None:
        h.setOptionValue("primal_feasibility_tolerance", tol)
        h.setOptionValue("dual_feasibility_tolerance", tol)
    lp = hs.HighsLp()
    lp.num_col_ = nx
    lp.num_row_ = qp["A"].shape[0]
    big = hs.kHighsInf
    clip = lambda a: np.clip(a, -big, big)
    lp.col_cost_ = qp["c"]
    lp.col_lower_ = clip(qp["xlo"]); lp.col_upper_ = clip(qp["xhi"])
    lp.row_lower_ = clip(qp["lo"]); lp.row_upper_ = clip(qp["hi"])
    lp.offset_ = qp["c0"]
    A = sp.csc_matrix(qp["A"])
    lp.a_matrix_.format_ = hs.MatrixFormat.kColwise
    lp.a_matrix_.num_col_ = nx
    lp.a_matrix_.num_row_ = qp["A"].shape[0]
    lp.a_matrix_.start_ = A.indptr
    lp.a_matrix_.index_ = A.indices
    lp.a_matrix_.value_ = A.data
    assert h.passModel(lp) == hs.HighsStatus.kOk
    L = sp.csc_matrix(np.tril(qp["Q"]))
    hess = hs.HighsHessian()
    hess.dim_ = nx
    hess.format_ = hs.HessianFormat.kTriangular
    hess.start_ = L.indptr
    hess.index_ = L.indices
    hess.value_ = L.data
    assert h.passHessian(hess) == hs.HighsStatus.kOk
    h.run()
    x = np.array(h.getSolution().col_value)
    obj = 0.5 * x @ qp["Q"] @ x + qp["c"] @ x + qp["c0"]
    return x, obj, h.modelStatusToString(h.getModelStatus())


def violation(qp, x):
    ax = qp["A"] @ x
    v = max(np.max(qp["lo"] - ax), np.max(ax - qp["hi"]), np.max(qp["xlo"] - x), np.max(x - qp["xhi"]))
    return float(v)


def kkt_certificate(qp, x, act_tol=1e-7):
    """Optimality certificate of x for the convex QP, independent of any solver: multipliers of the rows / bounds
    active at x by bounded least squares on the stationarity condition Q x + c = sum_i lambda_i a_i (lambda >= 0 for a
    row active at its lower side, <= 0 at its upper side, free for equalities).
    -> (stationarity residual |Qx + c - A' lambda|_inf / (1 + |c|_inf), complementarity sum |lambda_i slack_i|, violation)"""
    from scipy.optimize import lsq_linear
    A, lo, hi = qp["A"], qp["lo"], qp["hi"]
    nx = len(x)
    ax = A @ x
    cols, lb, ub, slack = [], [], [], []
    for r in range(A.shape[0]):
        eq = lo[r] == hi[r]
        at_lo = abs(ax[r] - lo[r]) <= act_tol
        at_hi = abs(hi[r] - ax[r]) <= act_tol
        if eq or at_lo or at_hi:
            cols.append(A[r])
            lb.append(-INF if (eq or at_hi) else 0.0)
            ub.append(INF if (eq or at_lo) else 0.0)
            slack.append(0.0 if eq else (ax[r] - lo[r] if at_lo else hi[r] - ax[r]))
    for j in range(nx):
        at_lo = abs(x[j] - qp["xlo"][j]) <= act_tol
        at_hi = abs(qp["xhi"][j] - x[j]) <= act_tol
        if at_lo or at_hi:
            e = np.zeros(nx); e[j] = 1.0
            cols.append(e)
            lb.append(-INF if at_hi else 0.0)
            ub.append(INF if at_lo else 0.0)
            slack.append(x[j] - qp["xlo"][j] if at_lo else qp["xhi"][j] - x[j])
    g = qp["Q"] @ x + qp["c"]
    At = np.array(cols).T
    res = lsq_linear(At, g, bounds=(np.array(lb), np.array(ub)), method="bvls", tol=1e-14, max_iter=2000)
    lam = res.x
    r = np.abs(At @ lam - g).max() / (1.0 + np.abs(qp["c"]).max())
    comp = float(np.abs(lam * np.array(slack)).sum())
    return float(r), comp, violation(qp, x)
