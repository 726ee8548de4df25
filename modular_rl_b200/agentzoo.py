"""Agents = containers with policy, value function, filters and updater - the reference's
agentzoo.py API unchanged (TrpoAgent, PpoLbfgsAgent, the option tables and therefore the run_pg.py
flags), built on device-resident networks instead of Keras models."""
from . import _lib as L
from .core import (Categorical, DiagGauss, NnVf, PG_OPTIONS, StochPolicyMLP, make_value_net)
from .filters import ZFilter
from .misc_utils import IDENTITY, comma_sep_ints, update_default_config
from .ppo import PpoLbfgsUpdater, PpoSgdUpdater
from .spaces import is_box, is_discrete
from .trpo import TrpoUpdater

MLP_OPTIONS = [
    ("hid_sizes", comma_sep_ints, [64, 64], "Sizes of hidden layers of MLP"),
    ("activation", str, "tanh", "nonlinearity"),
]


def make_mlps(ob_space, ac_space, cfg):
    """(policy, baseline) as agentzoo.py:25-61: Dense(h, act) per hidden size; Box actions ->
    Dense(d) with kernel*0.1 + state-independent logstd, Discrete -> Dense(K, softmax) with
    kernel*0.1; value net on [ob, t/timestep_limit] -> Dense(1); NnVf(mixfrac=0.1)."""
    assert is_box(ob_space)
    hid_sizes = list(cfg["hid_sizes"])
    if is_box(ac_space):
        probtype = DiagGauss(ac_space.shape[0])
    elif is_discrete(ac_space):
        probtype = Categorical(ac_space.n)
    else:
        raise NotImplementedError("action space %r" % (ac_space,))
    policy = StochPolicyMLP(ob_space.shape[0], hid_sizes, probtype, cfg["activation"])
    vfnet = make_value_net(ob_space.shape[0], hid_sizes, cfg["activation"])
    baseline = NnVf(vfnet, cfg["timestep_limit"], dict(mixfrac=0.1))
    return policy, baseline


def make_deterministic_mlp(ob_space, ac_space, cfg):
    raise NotImplementedError("DeterministicAgent / CEM is outside the accelerated path (SURVEY 2.1 #16)")


FILTER_OPTIONS = [
    ("filter", int, 1, "Whether to do a running average filter of the incoming observations and rewards"),
]


def make_filters(cfg, ob_space):
    if cfg["filter"]:
        obfilter = ZFilter(ob_space.shape, clip=5)
        rewfilter = ZFilter((), demean=False, clip=10)
    else:
        obfilter = IDENTITY
        rewfilter = IDENTITY
    return obfilter, rewfilter


class AgentWithPolicy(object):
    def __init__(self, policy, obfilter, rewfilter):
        self.policy = policy
        self.obfilter = obfilter
        self.rewfilter = rewfilter
        self.stochastic = True

    def set_stochastic(self, stochastic):
        self.stochastic = stochastic

    def act(self, ob_no):
        return self.policy.act(ob_no, stochastic=self.stochastic)

    def get_flat(self):
        return self.policy.get_flat()

    def set_from_flat(self, th):
        return self.policy.set_from_flat(th)

    def obfilt(self, ob):
        return self.obfilter(ob)

    def rewfilt(self, rew):
        return self.rewfilter(rew)


class DeterministicAgent(AgentWithPolicy):
    options = MLP_OPTIONS + FILTER_OPTIONS

    def __init__(self, ob_space, ac_space, usercfg):
        make_deterministic_mlp(ob_space, ac_space, usercfg)


class TrpoAgent(AgentWithPolicy):
    options = MLP_OPTIONS + PG_OPTIONS + TrpoUpdater.options + FILTER_OPTIONS

    def __init__(self, ob_space, ac_space, usercfg):
        cfg = update_default_config(self.options, usercfg)
        policy, self.baseline = make_mlps(ob_space, ac_space, cfg)
        obfilter, rewfilter = make_filters(cfg, ob_space)
        self.updater = TrpoUpdater(policy, cfg)
        AgentWithPolicy.__init__(self, policy, obfilter, rewfilter)


class PpoLbfgsAgent(AgentWithPolicy):
    options = MLP_OPTIONS + PG_OPTIONS + PpoLbfgsUpdater.options + FILTER_OPTIONS

    def __init__(self, ob_space, ac_space, usercfg):
        cfg = update_default_config(self.options, usercfg)
        policy, self.baseline = make_mlps(ob_space, ac_space, cfg)
        obfilter, rewfilter = make_filters(cfg, ob_space)
        self.updater = PpoLbfgsUpdater(policy, cfg)
        AgentWithPolicy.__init__(self, policy, obfilter, rewfilter)


class PpoSgdAgent(AgentWithPolicy):
    options = MLP_OPTIONS + PG_OPTIONS + PpoSgdUpdater.options + FILTER_OPTIONS

    def __init__(self, ob_space, ac_space, usercfg):
        cfg = update_default_config(self.options, usercfg)
        policy, self.baseline = make_mlps(ob_space, ac_space, cfg)
        obfilter, rewfilter = make_filters(cfg, ob_space)
        self.updater = PpoSgdUpdater(policy, cfg)
        AgentWithPolicy.__init__(self, policy, obfilter, rewfilter)
