from modular_rl_b200.cem import *  # noqa: F401,F403
from modular_rl_b200.cem import CEM_OPTIONS, cem, run_cem_algorithm  # noqa: F401
