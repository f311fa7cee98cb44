"""CPU oracle — TEST INFRASTRUCTURE ONLY.

A NumPy/SciPy restatement of the HILO-MPC algorithms on the hot path (NMPC/LMPC/MHE transcription and solve,
KF/EKF/UKF, PF, GP, finite-horizon LQR).  Every function cites the reference file:line it follows
(paths relative to ``/root/reference``).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may
import this package.  The product (``hilo_mpc_b200``) never does; it fails loudly if the CUDA library is missing.

Parity status:
* KF / EKF / UKF (one step and the 3-state bioreactor sigma-point matrices) / LQR-Riccati / SE, Matern and exponential
  Gram matrices / GP-LML / fitted GP hyper-parameters (gpml values, with and without a trainable mean): pinned by the
  golden numbers in the reference's own tests (``tests/test_KFs.py``, ``tests/test_LQR.py``, ``tests/test_kernels.py``,
  ``tests/test_GPs.py``) — reproduced in ``tests/test_oracle_golden.py``.
* Pinned by EXECUTING reference source under a stub ``casadi`` (``tools/make_reference_goldens.py`` ->
  ``tests/golden/reference_numeric.json``, ``tests/test_reference_numeric_goldens.py``): the collocation constants
  (``modeling.py:1092-1127``), one interval of the collocation and of the explicit Runge-Kutta transcription
  (``RungeKutta._collocation`` / ``_explicit``, ``modeling.py:1129-1281``) against the stage functions of ``FlatNlp``,
  the quadratic costs incl. scaled references, the input-change term and the MHE flavour on physical quantities
  (``QuadraticCost`` / ``MHEQuadraticCost``, ``modeling.py:38-818``),
  ``lhsnorm`` and the PF resampling / roughening / moment block (``pf.py:405-447``), the horizon of time-varying parameter
  values (``mpc.py:293-333``), the equation-string parser (``util/parsing.py``), the export of trained networks
  (``util/machine_learning.py:521-578``), the QP matrices of ``LMPC.setup`` (``mpc.py:2179-2266``) against
  ``oracle/lmpc.py`` and the product's condensed form, the horizon loops of ``NMPC._setup`` and ``MHE.setup`` (objective and
  constraint vector of the whole NLP) against ``FlatNlp`` / ``oracle/independent.py::ReactionMheNlp``, one predict + update
  step of the UKF (``kf.py:486-604``) against ``ukf_predict`` / ``ukf_update`` and the kernel body.
* NMPC / LMPC / MHE solves: the reference's tests hold NO numeric assertion for these and IPOPT / qpOASES cannot be run
  here (``casadi`` is not installed) -> "parity unpinned" at the IPOPT/qpOASES boundary only.  The
  oracle restates the transcription exactly and is validated by its own KKT residual, by SciPy's independent
  solvers (SLSQP on the flat NMPC and MHE problems, trust-constr on the condensed LMPC QP:
  ``tests/test_oracle_crosscheck.py``, ``tests/test_hostsim_lmpc_mhe.py``) and by the LQR golden in the linear-quadratic
  special case.  ``tools/run_reference.py`` produces golden vectors with the real reference where CasADi exists.
"""
