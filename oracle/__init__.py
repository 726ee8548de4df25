"""CPU oracle for the modular_rl TRPO update hot path.

THIS PACKAGE IS TEST INFRASTRUCTURE.  It is a numpy/scipy restatement of the
reference's arithmetic (ddlau/modular_rl, citations are file:line relative to
/root/reference) and exists only so that the CUDA path can be checked against
it.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import it.  Nothing under
``modular_rl_b200/`` imports it, and the product path raises when the CUDA
extension is missing rather than falling back to this code.

Pinning status (see DESIGN.md "Oracle"):

* pinned against the reference's OWN code executed in the authoring container
  (``tests/golden/make_golden.py`` loads the reference's numpy-only modules and
  runs its Theano-symbolic formulas through a numpy/torch shim):
  ``discount`` (misc_utils.py:9-27), ``compute_advantage`` (core.py:63-105),
  ``RunningStat``/``ZFilter`` (running_stat.py, filters.py:17-40),
  ``cg``/``linesearch`` (trpo.py:143-200), ``DiagGauss``/``Categorical``
  loglikelihood/kl/entropy (core.py:339-438), ``categorical_sample``
  (distributions.py:3-13), ``explained_variance_2d`` (misc_utils.py:44-49),
  and the surrogate / policy gradient / Fisher-vector product obtained by
  differentiating the reference's own loss expressions (trpo.py:37-61) with
  torch autograd (reverse-over-reverse, as trpo.py:45-58 does with Theano).
* pinned against the reference's own in-file tests: running_stat.py:35-46,
  core.py:441-483 (Monte-Carlo identities), the discount KAT of x.py:762.
* parity unpinned (no reference-owned number exists and Theano/Keras cannot be
  imported here): the Keras glorot initialiser's random stream, and the exact
  floating-point rounding of Theano's compiled fp32 graphs.  Everything else the
  oracle states is pinned by the two items above.
"""

from .policy_math import (NetSpec, num_params, param_slices, split_params,  # noqa: F401
                          init_params, forward, head_prob)
