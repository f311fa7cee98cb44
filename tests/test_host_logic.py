"""Host-side logic that needs no GPU: kernel objects against the reference's Gram goldens, argument checks of the control
loop and the GP trainer, the loud failure of product paths without a CUDA device."""
import numpy as np
import pytest

from hilo_mpc_b200 import GP, Kernel, SimpleControlLoop
from hilo_mpc_b200._lib import EngineUnavailable


def test_kernel_objects_reproduce_reference_grams():
    x = np.array([[1., 2., 3., 4., 5.]])
    # tests/test_kernels.py:1232-1247, :1826-1845, :856-875 and the SE golden of :262-277
    np.testing.assert_allclose(Kernel.matern_32()(x)[0], [1., .48335772, .13973135, .03431324, .007767734], rtol=2e-7)
    np.testing.assert_allclose(Kernel.matern_52()(x, x / 2.)[0], [.82864914, 1., .82864914, .52399411, .28316327],
                               rtol=2e-7)
    np.testing.assert_allclose(Kernel.exponential()(x, x / 2.)[0], [.60653066, 1., .60653066, .36787944, .22313016],
                               rtol=2e-7)
    np.testing.assert_allclose(Kernel.squared_exponential()(x)[0], np.exp(-.5 * np.arange(5.) ** 2), rtol=1e-14)
    with pytest.raises(NotImplementedError):
        Kernel.rational_quadratic()
    k = Kernel.matern_52(length_scales=[1., 2.], ard=True)
    with pytest.raises(ValueError):
        k.length_scale_vector(3)


def test_stationary_kernel_derivative_factors_match_finite_differences():
    """g_h returns h with dK/dlog l = sigma_f^2 h(r) r^2 for an isotropic kernel: check against central differences."""
    r2 = np.linspace(.01, 9., 50)
    for factory in (Kernel.squared_exponential, Kernel.exponential, Kernel.matern_32, Kernel.matern_52):
        k = factory()
        g, h = k.g_h(r2)
        eps = 1e-6   # log l -> log l + eps scales r^2 by exp(-2 eps)
        gp_, _ = k.g_h(r2 * np.exp(-2 * eps))
        gm_, _ = k.g_h(r2 * np.exp(2 * eps))
        np.testing.assert_allclose((gp_ - gm_) / (2 * eps), h * r2, rtol=1e-6, atol=1e-9)


def test_gp_argument_checks_without_gpu():
    gp = GP(['x'], ['y'])
    with pytest.raises(RuntimeError):
        gp.fit_model()
    with pytest.raises(RuntimeError):
        gp.setup()
    with pytest.raises(ValueError):
        gp.set_training_data(np.zeros((2, 5)), np.zeros((1, 5)))
    with pytest.raises(ValueError):
        GP(['x'], ['y', 'z'])
    assert [p.name for p in gp.hyperparameters] == ['GP.noise_variance', 'SE.length_scales', 'SE.signal_variance']


def test_control_loop_checks_and_fails_loudly_without_gpu():
    import torch
    from hilo_mpc_b200.examples import cstr_nmpc
    from hilo_mpc_b200.library import cstr_model
    from hilo_mpc_b200 import Model
    nmpc = cstr_nmpc(N=10, setup=False)
    nmpc._build_module = False
    nmpc.setup()
    plant = cstr_model(0.005)
    with pytest.raises(ValueError):
        SimpleControlLoop(plant, nmpc, feedback='observer')          # no observer given
    with pytest.raises(ValueError):
        SimpleControlLoop(plant, nmpc, feedback='nonsense')
    other = Model(name='other')
    other.set_dynamical_states(['a', 'b'])
    other.set_inputs(['u'])
    other.set_dynamical_equations(['-a + u', 'a - b'])
    other.setup(dt=.1)
    with pytest.raises(ValueError):
        SimpleControlLoop(other, nmpc)                               # controller states are not plant states
    with pytest.raises(TypeError):
        SimpleControlLoop(plant, object())
    loop = SimpleControlLoop(plant, nmpc)
    if not torch.cuda.is_available():
        with pytest.raises(EngineUnavailable):
            loop.run(1, x0=np.zeros((2, 4)))


def test_control_loop_hands_only_constant_parameters_to_the_controller():
    """Plant parameters that the controller treats as time-varying must not reach ``cp`` (reference
    control_loop.py:357-362); a ``cp`` of the wrong length is an error, not a silent batch."""
    import torch
    from hilo_mpc_b200 import Model

    plant = Model(name='two_par')
    x = plant.set_dynamical_states(['a'])
    u = plant.set_inputs(['u'])
    p = plant.set_parameters(['k', 'w'])
    plant.set_dynamical_equations([-p[0] * x[0] + p[1] * u[0]])
    plant.setup(dt=.1)

    class Recorder:
        _model = plant
        _tvp_names = ['w']

        def optimize(self, x0, cp=None, **kwargs):
            self.cp = cp
            return torch.zeros(x0.shape[0], 1, dtype=torch.float64)

    ctrl = Recorder()
    loop = SimpleControlLoop(plant, ctrl)
    loop._control(torch.zeros(3, 1, dtype=torch.float64), [2., 5.], {})
    assert [float(v) for v in ctrl.cp] == [2.]
    ctrl._tvp_names = []
    loop._control(torch.zeros(3, 1, dtype=torch.float64), [2., 5.], {})
    assert [float(v) for v in np.asarray(ctrl.cp).reshape(-1)] == [2., 5.]

    # NMPC side: the alias the reference's loop reads exists, and a cp of the wrong size is rejected
    from hilo_mpc_b200 import NMPC
    nmpc = NMPC(plant)
    nmpc.horizon = 4
    nmpc.quad_stage_cost.add_states(names=['a'], weights=[1.], ref=[0.])
    nmpc.set_time_varying_parameters(['w'], values={'w': [1., 2., 3., 4., 5.]})
    nmpc._build_module = False
    nmpc.setup()
    assert list(nmpc._time_varying_parameters) == ['w'] == list(nmpc._tvp_names)
    with pytest.raises(ValueError):
        nmpc._stage_parameters([1., 2., 3.], None, None, torch, torch.device('cpu'), 2)
    sp = nmpc._stage_parameters([7.], None, None, torch, torch.device('cpu'), 2)
    assert sp.shape[-1] == 2 and sp.shape[-2] == 4                    # per stage: [w_k, k]
    assert np.allclose(sp.reshape(-1, 4, 2)[0].numpy(), [[1., 7.], [2., 7.], [3., 7.], [4., 7.]])
