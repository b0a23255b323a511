"""TEST INFRASTRUCTURE ONLY -- load the UNMODIFIED reference modules as the parity oracle.

Only usable in the authoring container (``/root/reference`` does not exist on the GPU box); it is
used by ``oracle/gen_golden.py`` to write the committed fixtures under ``tests/golden/`` and by the
``-m "not gpu"`` tests that pin ``oracle/restate.py`` against the reference itself.

What it does (nothing in /root/reference is edited or copied):
  * puts ``oracle/dgl_shim`` (pure-torch DGL 0.4 stand-in) and ``/root/reference`` on ``sys.path``
    and imports ``utils``, ``RGCN``, ``Aggregator``, ``model`` under private module names;
  * neutralises the 32 unconditional ``.cuda()`` calls (e.g. model.py:80,88,96, Aggregator.py:144,146,
    utils.py:212,236-237,242) with ``torch.Tensor.cuda = identity`` and
    ``torch.cuda.current_device = lambda: 0`` while the reference code runs (context manager);
  * nothing else: torch 2.11 runs the rest of the reference unchanged on CPU.
"""
import contextlib
import importlib
import os
import sys

import torch

REFERENCE_DIR = os.environ.get('RENET_REFERENCE_DIR', '/root/reference')
_SHIM_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'dgl_shim')
_REF_MODULES = ('utils', 'RGCN', 'Aggregator', 'model', 'global_model')


def available():
    return os.path.isfile(os.path.join(REFERENCE_DIR, 'RGCN.py'))


@contextlib.contextmanager
def cpu_patches():
    """Make the reference's hard-coded ``.cuda()`` sites no-ops for the duration of the block."""
    saved = (torch.Tensor.cuda, torch.cuda.current_device)
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.cuda.current_device = lambda: 0
    try:
        yield
    finally:
        torch.Tensor.cuda, torch.cuda.current_device = saved


_cache = {}


def load():
    """Returns a namespace with ``.utils .RGCN .Aggregator .model .dgl`` (reference modules)."""
    if 'ns' in _cache:
        return _cache['ns']
    if not available():
        raise RuntimeError('reference tree not present at %s' % REFERENCE_DIR)
    saved_path = list(sys.path)
    saved_mods = {k: sys.modules.get(k) for k in _REF_MODULES + ('dgl', 'dgl.function')}
    for k in saved_mods:
        sys.modules.pop(k, None)
    sys.path.insert(0, REFERENCE_DIR)
    sys.path.insert(0, _SHIM_DIR)
    try:
        class NS:
            pass
        ns = NS()
        ns.dgl = importlib.import_module('dgl')
        for name in _REF_MODULES:
            setattr(ns, name, importlib.import_module(name))
    finally:
        sys.path[:] = saved_path
        # keep the reference modules alive under private names only; the product package has its
        # own ``utils``-like modules and must never see these through sys.modules.
        for k in _REF_MODULES + ('dgl', 'dgl.function'):
            mod = sys.modules.pop(k, None)
            if mod is not None:
                sys.modules['_renet_reference_.' + k] = mod
            if saved_mods[k] is not None:
                sys.modules[k] = saved_mods[k]
    _cache['ns'] = ns
    return ns
