"""Utilities that keep the reference's names and meaning (misc_utils.py): the option system the
agents and run_pg.py are built on, EzPickle, flatten/unflatten, explained variance - plus
``discount``, which runs the segmented reverse scan on the device (one segment)."""
from __future__ import print_function

import atexit
import os
import os.path as osp
import sys
from collections import defaultdict

import numpy as np


# ================================================================
# Math utilities
# ================================================================

def discount(x, gamma):
    """y[t] = x[t] + gamma*x[t+1] + gamma^2*x[t+2] + ...  along axis 0 (misc_utils.py:9-27).

    Evaluated by the device scan kernel (mrl_gae with a zero baseline and lam=1 reduces to the
    plain discounted sum); columns of a 2-D input are scanned as independent segments."""
    from .device import gae_flat
    x = np.asarray(x, np.float64)
    assert x.ndim >= 1
    T = x.shape[0]
    if T == 0:
        return x.copy()
    cols = x.reshape(T, -1)
    k = cols.shape[1]
    flat = np.ascontiguousarray(cols.T).reshape(-1)          # k segments of length T
    offsets = np.arange(k + 1, dtype=np.int64) * T
    ret, _ = gae_flat(flat, np.zeros_like(flat), offsets, np.ones(k, np.uint8), gamma, 1.0)
    return ret.reshape(k, T).T.reshape(x.shape)


def explained_variance(ypred, y):
    """1 - Var[y-ypred] / Var[y]   (misc_utils.py:29-42)"""
    assert y.ndim == 1 and ypred.ndim == 1
    vary = np.var(y)
    return np.nan if vary == 0 else 1 - np.var(y - ypred) / vary


def explained_variance_2d(ypred, y):
    assert y.ndim == 2 and ypred.ndim == 2
    vary = np.var(y, axis=0)
    out = 1 - np.var(y - ypred) / vary        # numerator is un-axised in the reference too
    out[vary < 1e-10] = 0
    return out


# ================================================================
# Configuration
# ================================================================

def update_default_config(tuples, usercfg):
    """tuples: (name, type, default, description); usercfg overrides known keys only."""
    out = dict2()
    for (name, _, defval, _) in tuples:
        out[name] = defval
    if usercfg:
        for (k, v) in usercfg.items():
            if k in out:
                out[k] = v
    return out


def update_argument_parser(parser, options, **kwargs):
    kwargs = kwargs.copy()
    for (name, typ, default, desc) in options:
        flag = "--" + name
        if flag in parser._option_string_actions.keys():  # pylint: disable=W0212
            print("warning: already have option %s. skipping" % name)
        else:
            parser.add_argument(flag, type=typ, default=kwargs.pop(name, default), help=desc or " ")
    if kwargs:
        raise ValueError("options %s ignored" % kwargs)


def comma_sep_ints(s):
    # a list, not a one-shot `map`: the reference's py3 `map` is exhausted by the policy loop and
    # leaves the value net without hidden layers (SURVEY section 5, config hazard)
    return [int(tok) for tok in s.split(",")] if s else []


def IDENTITY(x):
    return x


GENERAL_OPTIONS = [
    ("seed", int, 0, "random seed"),
    ("metadata", str, "", "metadata about experiment"),
    ("outfile", str, "./tmp/a.h5", "output file"),
    ("use_hdf", int, 0, "whether to make an hdf5 file with results and snapshots"),
    ("snapshot_every", int, 0, "how often to snapshot"),
    ("load_snapshot", str, "", "path to snapshot"),
    ("video", int, 1, "whether to record video"),
]


# ================================================================
# Load/save
# ================================================================

def prepare_h5_file(args):
    outfile_default = "/tmp/a.h5"
    fname = args.outfile or outfile_default
    if osp.exists(fname) and fname != outfile_default:
        input("output file %s already exists. press enter to continue. (exit with ctrl-C)" % fname)
    import h5py  # optional dependency, only with --use_hdf
    hdf = h5py.File(fname, "w")
    hdf.create_group('params')
    for (param, val) in args.__dict__.items():
        try:
            hdf['params'][param] = val
        except (ValueError, TypeError):
            print("not storing parameter", param)
    diagnostics = defaultdict(list)
    print("Saving results to %s" % fname)

    def save():
        hdf.create_group("diagnostics")
        for (diagname, val) in diagnostics.items():
            hdf["diagnostics"][diagname] = val

    hdf["cmd"] = " ".join(sys.argv)
    atexit.register(save)
    return hdf, diagnostics


def save_agent_snapshot(agent, dirname, counter, env_id=None):
    """Pickled agent written as <dirname>/agent_snapshots/<%04i>.pkl - the same bytes run_pg.py:141-142
    stores under /agent_snapshots/%0.4i of the hdf5 file, for boxes without h5py.  env_id (run_pg.py:147's
    hdf['env_id']) goes to <dirname>/env_id.txt for sim_agent.py."""
    import pickle
    d = osp.join(dirname, "agent_snapshots")
    os.makedirs(d, exist_ok=True)
    if env_id is not None:
        with open(osp.join(dirname, "env_id.txt"), "w") as f:
            f.write(str(env_id))
    fname = osp.join(d, "%0.4i.pkl" % counter)
    with open(fname, "wb") as f:
        f.write(pickle.dumps(agent, -1))
    return fname


def snapshot_env_id(path):
    """The environment id stored next to the snapshots (hdf['env_id'], sim_agent.py:50), or None."""
    if osp.isdir(path) or path.endswith(".pkl"):
        d = path if osp.isdir(path) else osp.dirname(path)
        for cand in (d, osp.dirname(d.rstrip("/"))):
            f = osp.join(cand, "env_id.txt")
            if osp.exists(f):
                return open(f).read().strip()
        return None
    import h5py  # optional dependency
    with h5py.File(path, "r") as hdf:
        v = hdf["env_id"][()]
        return v.decode() if isinstance(v, bytes) else str(v)


def load_agent_snapshot(path, snapname=None):
    """Agent from a snapshot: a .pkl written by save_agent_snapshot, a directory of them, or an hdf5
    results file with /agent_snapshots (sim_agent.py:41-52; snapname None = the last one)."""
    import pickle
    if osp.isdir(path):
        d = osp.join(path, "agent_snapshots") if osp.isdir(osp.join(path, "agent_snapshots")) else path
        names = sorted(n for n in os.listdir(d) if n.endswith(".pkl"))
        if not names:
            raise ValueError("no snapshots under %s" % d)
        name = (snapname + ".pkl") if snapname else names[-1]
        if name not in names:
            raise ValueError("Invalid snapshot name %s" % snapname)
        path = osp.join(d, name)
    if path.endswith(".pkl"):
        with open(path, "rb") as f:
            return pickle.loads(f.read())
    import h5py  # optional dependency
    with h5py.File(path, "r") as hdf:
        snapnames = sorted(hdf["agent_snapshots"].keys())
        if snapname is None:
            snapname = snapnames[-1]
        elif snapname not in snapnames:
            raise ValueError("Invalid snapshot name %s" % snapname)
        return pickle.loads(hdf["agent_snapshots"][snapname][()].tobytes())


# ================================================================
# Misc
# ================================================================

class dict2(dict):
    "dictionary-like object that exposes its keys as attributes"

    def __init__(self, **kwargs):
        dict.__init__(self, kwargs)
        self.__dict__ = self


def zipsame(*seqs):
    L = len(seqs[0])
    assert all(len(seq) == L for seq in seqs[1:])
    return zip(*seqs)


def flatten(arrs):
    return np.concatenate([arr.flat for arr in arrs])


def unflatten(vec, shapes):
    i = 0
    arrs = []
    for shape in shapes:
        size = int(np.prod(shape))
        arrs.append(vec[i:i + size].reshape(shape))
        i += size
    return arrs


class EzPickle(object):
    """Objects that are pickled and unpickled via their constructor arguments
    (misc_utils.py:163-189): device handles are rebuilt by re-running __init__."""

    def __init__(self, *args, **kwargs):
        self._ezpickle_args = args
        self._ezpickle_kwargs = kwargs

    def __getstate__(self):
        return {"_ezpickle_args": self._ezpickle_args, "_ezpickle_kwargs": self._ezpickle_kwargs}

    def __setstate__(self, d):
        out = type(self)(*d["_ezpickle_args"], **d["_ezpickle_kwargs"])
        self.__dict__.update(out.__dict__)


def fmt_row(width, row, header=False):
    out = " | ".join(fmt_item(x, width) for x in row)
    if header:
        out = out + "\n" + "-" * len(out)
    return out


def fmt_item(x, l):
    if isinstance(x, np.ndarray):
        assert x.ndim == 0
        x = x.item()
    if isinstance(x, float):
        rep = "%g" % x
    else:
        rep = str(x)
    return " " * (l - len(rep)) + rep
