"""hilo_mpc_b200 — a B200-native batched MPC / estimation solve engine behind the HILO-MPC API.

Public names follow ``hilo_mpc/__init__.py:43-99`` of the reference for the classes on the hot path.
"""
from . import symbolic
from . import symbolic as ca  # CasADi-flavoured helper namespace: ca.cos, ca.sin, ca.exp, ca.sym ...
from .model import Model
from .nmpc import NMPC
from .lmpc import LMPC
from .mhe import MHE, MovingHorizonEstimator
from .gp import GP, GaussianProcess, Kernel, Mean
from .ann import ANN, Layer
from .control_loop import SimpleControlLoop
from .estimators import EKF, KF, PF, UKF, ExtendedKalmanFilter, KalmanFilter, ParticleFilter, UnscentedKalmanFilter

pi = 3.141592653589793

__all__ = ['Model', 'NMPC', 'LMPC', 'MHE', 'MovingHorizonEstimator', 'GP', 'GaussianProcess', 'Kernel', 'Mean', 'ANN', 'Layer',
           'KF', 'EKF', 'UKF', 'PF', 'KalmanFilter', 'ExtendedKalmanFilter', 'UnscentedKalmanFilter',
           'ParticleFilter', 'SimpleControlLoop', 'symbolic', 'ca']
