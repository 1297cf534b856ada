"""jointforces material-simulation hot path, B200-native.

Same module layout as the reference for the path that is in scope
(``/root/reference/jointforces/__init__.py``): ``materials``, ``mesh``, ``simulation``; plus
``solver`` (the saenopy object protocol jointforces calls), ``sweep`` (multi-GPU pressure
sweeps) and the data path of ``force`` (table interpolants, ``infer_pressure``, ``reconstruct``).  All arithmetic runs in ``libjf_b200.so`` (hand-written sm_100a CUDA behind the C ABI
in ``include/jf_b200.h``); there is no CPU fallback.
"""
from . import force, materials, mesh, simulation, solver, sweep  # noqa: F401


def load(file):
    """``jf.load`` of the reference (``utils.py:4``): a pickled dict stored in an ``.npy``."""
    import numpy as np
    return np.load(file, allow_pickle=True).item()
