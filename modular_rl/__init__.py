"""Drop-in alias: `import modular_rl`, `modular_rl.agentzoo.TrpoAgent`, `from modular_rl import *`
resolve to the B200-native implementation (modular_rl_b200), with the re-export surface of the
reference's modular_rl/__init__.py:1-6 (core, distributions, filters)."""
from modular_rl_b200.core import *           # noqa: F401,F403
from modular_rl_b200.distributions import *  # noqa: F401,F403
from modular_rl_b200.filters import *        # noqa: F401,F403
from modular_rl_b200.cem import *            # noqa: F401,F403
