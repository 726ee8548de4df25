from modular_rl_b200.running_stat import *  # noqa: F401,F403
from modular_rl_b200 import running_stat as _impl
globals().update({k: v for k, v in vars(_impl).items() if not k.startswith('__')})
