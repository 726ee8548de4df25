from modular_rl_b200.distributions import *  # noqa: F401,F403
from modular_rl_b200 import distributions as _impl
globals().update({k: v for k, v in vars(_impl).items() if not k.startswith('__')})
