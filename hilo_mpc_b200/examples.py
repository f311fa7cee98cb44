"""Benchmark / test problem builders (BASELINE.json ``configs``), written against the public API only."""
from __future__ import annotations

import numpy as np

from .library import CSTR_BOUNDS, CSTR_SCALING, cstr_model, cstr_steady_state
from .nmpc import NMPC

# weights act on the scaled variables (mpc.py:207-263): c/1, T/100, F/10, Q/1000
CSTR_Q = [1., 10., 1., 1.]
CSTR_R = [.1, .1]


def cstr_nmpc(N: int = 50, dt: float = 0.005, integration_method: str = 'collocation', solver_options=None,
              setup: bool = True) -> NMPC:
    """Config 2: reference-tracking NMPC for the Klatt-Engell CSTR (nx=4, nu=2), default Radau-3 collocation."""
    model = cstr_model(dt)
    x_ss, u_ss = cstr_steady_state()
    nmpc = NMPC(model)
    nmpc.horizon = N
    nmpc.quad_stage_cost.add_states(names=['c_A', 'c_B', 'T_R', 'T_K'], ref=list(x_ss), weights=CSTR_Q)
    nmpc.quad_stage_cost.add_inputs(names=['F', 'Q_dot'], ref=list(u_ss), weights=CSTR_R)
    nmpc.quad_terminal_cost.add_states(names=['c_A', 'c_B', 'T_R', 'T_K'], ref=list(x_ss), weights=CSTR_Q)
    nmpc.set_box_constraints(**CSTR_BOUNDS)
    nmpc.set_scaling(**CSTR_SCALING)
    nmpc.set_initial_guess(x_guess=x_ss, u_guess=u_ss)
    nmpc.set_nlp_options(integration_method=integration_method)
    if solver_options:
        nmpc.set_solver_opts(solver_options)
    if setup:
        nmpc.setup()
    return nmpc


def cstr_side_reaction_gp(n_train: int = 24, seed: int = 1):
    """Config 3, the learned component: a GP (SE kernel with ARD length scales, fixed hyper-parameters, constant prior
    mean) of the side-reaction rate ``k_3(T_R) c_A^2`` over the feature box ``c_A in [0.6, 2], T_R in [110, 140]``,
    trained on ``n_train`` uniformly drawn points with N(0, 1e-4) noise and fitted on the GPU (``hb_gp_fit``)."""
    from .gp import GaussianProcess, Kernel
    from .library import cstr_side_reaction
    rng = np.random.default_rng(seed)
    X = np.vstack([rng.uniform(.6, 2., n_train), rng.uniform(110., 140., n_train)])
    y = cstr_side_reaction(X[0], X[1]) + 1e-2 * rng.standard_normal(n_train)
    gp = GaussianProcess(['c_A', 'T_R'], 'r_3', kernel=Kernel.squared_exponential(signal_variance=25., ard=True,
                                                                                    length_scales=[.7, 15.]),
                         noise_variance=1e-4, mean=float(y.mean()))
    gp.set_training_data(X, y.reshape(1, -1))
    gp.setup()
    return gp


def cstr_gp_nmpc(gp, N: int = 10, dt: float = 0.005, solver_options=None, setup: bool = True) -> NMPC:
    """Config 3 (GP-output-feedback NMPC): the CSTR of config 2 with the side-reaction rate replaced by the posterior
    mean of ``gp`` (symbolic query, the way the reference embeds learned components: dynamic_model.py:3040-3080)."""
    from . import symbolic as sx

    def side_reaction(cA, TR):
        mean, _ = gp.predict(sx.as_vector([cA, TR]))
        return mean[0]
    model = cstr_model(dt, side_reaction=side_reaction, name=f'cstr_gp{gp.X_train.shape[1]}')
    x_ss, u_ss = cstr_steady_state()
    nmpc = NMPC(model)
    nmpc.horizon = N
    nmpc.quad_stage_cost.add_states(names=['c_A', 'c_B', 'T_R', 'T_K'], ref=list(x_ss), weights=CSTR_Q)
    nmpc.quad_stage_cost.add_inputs(names=['F', 'Q_dot'], ref=list(u_ss), weights=CSTR_R)
    nmpc.quad_terminal_cost.add_states(names=['c_A', 'c_B', 'T_R', 'T_K'], ref=list(x_ss), weights=CSTR_Q)
    nmpc.set_box_constraints(**CSTR_BOUNDS)
    nmpc.set_scaling(**CSTR_SCALING)
    nmpc.set_initial_guess(x_guess=x_ss, u_guess=u_ss)
    if solver_options:
        nmpc.set_solver_opts(solver_options)
    if setup:
        nmpc.setup()
    return nmpc


def cstr_initial_states(B: int, seed: int = 0) -> np.ndarray:
    """``x0 = x_ss + diag([.2,.2,5,5]) U(-1,1)`` with ``np.random.default_rng(seed)`` (cfg 2)."""
    x_ss, _ = cstr_steady_state()
    rng = np.random.default_rng(seed)
    return x_ss + np.array([.2, .2, 5., 5.]) * rng.uniform(-1., 1., size=(B, 4))


# ---------------------------------------------------------------------------------------------------------------------
_ANN_CACHE = {}


def bicycle_ann(hidden: int = 32, seed: int = 0):
    """MLP (1 x ``hidden`` tanh, z-scored inputs/outputs, the graph of ``machine_learning.py:521-578``) of the
    kinematic-bicycle right-hand side as a function of the features ``(v, sin(phi), cos(phi), a, delta)`` - the heading
    enters through its sine and cosine so that the learned dynamics are periodic in ``phi`` like the true ones (the path
    of config 5 is a closed curve: the heading winds up by 2 pi per lap).  Trained offline, deterministically: full-batch
    L-BFGS on samples of the true dynamics (stand-in for the offline-trained network of BASELINE config 5)."""
    from .ann import ANN, Layer
    from .library import bicycle_rhs
    rng = np.random.default_rng(seed)
    n = 8000
    v, phi = rng.uniform(-2., 10., n), rng.uniform(-np.pi, np.pi, n)
    a, delta = rng.uniform(-1., 1., n), rng.uniform(-1., 1., n)
    Z = np.column_stack([v, np.sin(phi), np.cos(phi), a, delta])
    F = np.array([bicycle_rhs([0., 0., v[i], phi[i]], [a[i], delta[i]], sin=np.sin, cos=np.cos, tan=np.tan,
                              atan=np.arctan) for i in range(n)])
    zm, zs = Z.mean(0), Z.std(0)
    fm, fs = F.mean(0), F.std(0)
    fs[fs == 0.] = 1.
    key = (hidden, seed)
    if key not in _ANN_CACHE:
        import torch
        torch.manual_seed(seed)
        Zn = torch.tensor((Z - zm) / zs)
        Fn = torch.tensor((F - fm) / fs)
        lin1 = torch.nn.Linear(5, hidden).double()
        lin2 = torch.nn.Linear(hidden, 4).double()
        params = list(lin1.parameters()) + list(lin2.parameters())
        opt = torch.optim.LBFGS(params, lr=1., max_iter=600, history_size=30, line_search_fn='strong_wolfe')

        def closure():
            opt.zero_grad()
            loss = ((lin2(torch.tanh(lin1(Zn))) - Fn) ** 2).mean()
            loss.backward()
            return loss
        opt.step(closure)
        _ANN_CACHE[key] = tuple(t.detach().numpy().copy() for t in (lin1.weight, lin1.bias, lin2.weight, lin2.bias))
    W1, b1, W2, b2 = _ANN_CACHE[key]
    net = ANN(['v', 'sin_phi', 'cos_phi', 'a', 'delta'], ['dpx', 'dpy', 'dv', 'dphi'], name='bicycle_mlp')
    net.add_layers(Layer.dense(hidden, 'tanh'))
    net.setup()
    net.set_weights([W1, W2], [b1, b2], input_scaling=(zm, zs), output_scaling=(fm, fs))
    return net


def bicycle_features(x, u, sin=None, cos=None):
    """Feature vector of ``bicycle_ann`` from the robot state ``[px, py, v, phi]`` and input ``[a, delta]``."""
    from . import symbolic as sx
    sin, cos = sin or sx.sin, cos or sx.cos
    return np.array([x[2], sin(x[3]), cos(x[3]), u[0], u[1]], dtype=object)


def robot_nn_nmpc(N: int = 30, dt: float = 1., integration_method: str = 'collocation', setup: bool = True,
                  hidden: int = 32) -> NMPC:
    """Config 5: path-following NMPC of a mobile robot whose right-hand side is an MLP
    (docs/docsource/examples/mpc_formula1.ipynb: path ``(30 - 14 cos(theta), 30 - 16 sin(theta))``, ``u_pf in [0.1, 1]``)."""
    from . import symbolic as sx
    from .model import Model
    net = bicycle_ann(hidden)
    model = Model(name=f'robot_mlp{hidden}')
    x = model.set_dynamical_states(['px', 'py', 'v', 'phi'])
    u = model.set_inputs(['a', 'delta'])
    model.set_dynamical_equations(net.predict(bicycle_features(x, u)))
    model.setup(dt=dt)
    nmpc = NMPC(model)
    nmpc.horizon = N
    theta = nmpc.create_path_variable(name='theta', u_pf_lb=.1, u_pf_ub=1., u_pf_ref=.5, u_pf_weight=10.,
                                      theta_guess=0., theta_lb=0.)   # theta_ub = inf: the reference's default
    nmpc.quad_stage_cost.add_states(names=['px', 'py'], weights=[1., 1.],
                                    ref=np.array([30. - 14. * sx.cos(theta), 30. - 16. * sx.sin(theta)], dtype=object),
                                    path_following=True)
    nmpc.quad_stage_cost.add_inputs(names=['a', 'delta'], weights=[.1, .1])
    nmpc.quad_terminal_cost.add_states(names=['px', 'py'], weights=[1., 1.],
                                       ref=np.array([30. - 14. * sx.cos(theta), 30. - 16. * sx.sin(theta)], dtype=object),
                                       path_following=True)
    nmpc.set_box_constraints(x_ub=[100., 100., 10., 100.], x_lb=[-100., -100., -10., -100.], u_ub=[1., 1.],
                             u_lb=[-1., -1.])
    nmpc.set_initial_guess(x_guess=[15., 30., 0., 0.], u_guess=[0., 0.])
    nmpc.set_nlp_options(integration_method=integration_method)
    if setup:
        nmpc.setup()
    return nmpc


def robot_initial_states(B: int, seed: int = 3) -> np.ndarray:
    """``x0 = [15, 30, 0, 0] + diag([2, 2, .5, .2]) U(-1, 1)`` (cfg 5)."""
    rng = np.random.default_rng(seed)
    return np.array([15., 30., 0., 0.]) + np.array([2., 2., .5, .2]) * rng.uniform(-1., 1., size=(B, 4))


# ---------------------------------------------------------------------------------------------------------------------
def cstr_mhe(N: int = 20, dt: float = 0.005, solver_options=None, setup: bool = True):
    """Config 4, the optimisation half: moving-horizon estimation of the CSTR states and of the scale of the rate
    constant ``k_10`` from measurements of ``c_B`` and ``T_R`` (horizon 20, Radau-3 collocation, state noise on all four
    states).  Stage structure seen by the engine: nx = 4 + 1 parameter state, nu = 4 noise inputs,
    free x_0 with arrival cost, discrete-type objective."""
    from .mhe import MHE
    model = cstr_model(dt, name='cstr_k1', with_parameter=True)
    x_ss, _ = cstr_steady_state()
    mhe = MHE(model)
    mhe.horizon = N
    mhe.quad_arrival_cost.add_states(weights=[10., 10., .01, .01], guess=list(x_ss))
    mhe.quad_arrival_cost.add_parameters(weights=[1.], guess=[1.])
    mhe.quad_stage_cost.add_measurements(weights=[1e4, 1e2])        # 1 / sigma^2 of the measurement noise
    mhe.quad_stage_cost.add_state_noise(weights=[1e4, 1e4, 1., 1.])
    mhe.set_box_constraints(x_lb=CSTR_BOUNDS['x_lb'], x_ub=CSTR_BOUNDS['x_ub'], p_lb=[.5], p_ub=[2.])
    mhe.set_scaling(x_scaling=CSTR_SCALING['x_scaling'], w_scaling=[.01, .01, 1., 1.])
    if solver_options:
        mhe.set_solver_opts(solver_options)
    if setup:
        mhe.setup()
    return mhe


def cstr_mhe_data(B: int, steps: int, dt: float = 0.005, seed: int = 2):
    """Synthetic closed-loop data for config 4: ``B`` plants with ``k1_scale ~ U(.9, 1.1)``, inputs wandering around the
    steady-state input, process noise ``N(0, diag(1e-4 x_ss^2)) dt``-scaled, measurement noise ``sigma = (.01, .1)``.
    Returns ``u [steps, B, 2]``, ``y [steps, B, 2]``, ``x [steps + 1, B, 4]``, ``k1_scale [B]`` (NumPy, RK4 with 4 sub-steps)."""
    from .library import cstr_rhs
    rng = np.random.default_rng(seed)
    x_ss, u_ss = cstr_steady_state()
    ks = rng.uniform(.9, 1.1, size=B)
    x = np.tile(x_ss, (B, 1)) * (1. + .02 * rng.uniform(-1., 1., size=(B, 4)))
    xs, us, ys = [x.copy()], [], []
    sig_w = 1e-2 * x_ss * np.sqrt(dt)
    for k in range(steps):
        u = u_ss * (1. + .1 * np.sin(.3 * k + rng.uniform(0., 6.28, size=(B, 1)) * np.array([1., .5])))
        f = lambda z: np.stack(cstr_rhs(z.T, u.T, exp=np.exp, k1_scale=ks), axis=1)
        h = dt / 4
        for _ in range(4):
            k1 = f(x); k2 = f(x + .5 * h * k1); k3 = f(x + .5 * h * k2); k4 = f(x + h * k3)
            x = x + h / 6. * (k1 + 2 * k2 + 2 * k3 + k4)
        x = x + sig_w * rng.standard_normal((B, 4))
        y = x[:, 1:3] + np.array([.01, .1]) * rng.standard_normal((B, 2))
        us.append(u); ys.append(y); xs.append(x.copy())
    return np.stack(us), np.stack(ys), np.stack(xs), ks
