"""Bootstrap for importing the UNMODIFIED reference (dynaroars/neuralsat) from /root/reference.

TEST INFRASTRUCTURE ONLY.  Used in the build container to (a) validate the CPU oracle
(`oracle/crown_oracle.py`) against the reference's own auto_LiRPA path and (b) generate the
golden fixtures under `tests/golden/`.  `/root/reference` does not exist on the GPU box, so
nothing in `-m gpu` tests, `smoke()` or `bench.py` imports this module.

Follows SURVEY.md Appendix B:
  * torch>=2.9 moved the TorchScript-ONNX internals that auto_LiRPA/parse_graph.py:4-6 imports,
  * gurobipy / termcolor / onnx2pytorch are absent and are stubbed (type-check only).
Nothing under /root/reference is modified or copied.
"""
import importlib
import importlib.machinery
import os
import sys
import types

REF_ROOT = os.environ.get('NEURALSAT_REFERENCE', '/root/reference/neuralsat-pt201')


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, 'auto_LiRPA'))


_done = False


def bootstrap():
    """Make `auto_LiRPA`, `abstractor`, `heuristic`, `setting` importable. Idempotent."""
    global _done
    if _done:
        return
    if not available():
        raise RuntimeError(f'reference tree not found at {REF_ROOT}')
    import torch
    import torch.nn as nn
    import torch.onnx.utils as u
    import torch.onnx.symbolic_helper as sh
    T = 'torch.onnx._internal.torchscript_exporter.'
    try:
        g = importlib.import_module(T + '_globals')
        sys.modules['torch.onnx._globals'] = g
        torch.onnx._globals = g
        if not hasattr(u, '_optimize_graph'):
            u._optimize_graph = importlib.import_module(T + 'utils')._optimize_graph
        if not hasattr(sh, '_node_get'):
            sh._node_get = importlib.import_module(T + 'symbolic_helper')._node_get
    except ModuleNotFoundError:
        pass  # older torch: the reference's imports work as they are

    def _mod(name, **attrs):
        m = types.ModuleType(name)
        m.__spec__ = importlib.machinery.ModuleSpec(name, None)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
        return m

    class ConvertModel(nn.Module):
        """Stands in for NS/onnx2pytorch ConvertModel (beartype isinstance check only)."""

        def __init__(self, net):
            super().__init__()
            self.net = net

        def forward(self, x):
            return self.net(x)

    if 'gurobipy' not in sys.modules:
        _mod('gurobipy')
    if 'termcolor' not in sys.modules:
        _mod('termcolor', cprint=print, colored=lambda s, *a, **k: s)
    cm = _mod('onnx2pytorch.convert.model', ConvertModel=ConvertModel)
    cv = _mod('onnx2pytorch.convert', model=cm, ConvertModel=ConvertModel)
    _mod('onnx2pytorch', convert=cv, ConvertModel=ConvertModel)

    sys.dont_write_bytecode = True
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    from setting import Settings
    Settings.setup(None)
    _done = True


class Objective:
    """Duck-type of NS/verifier/objective.py:99-129 (attributes only)."""

    def __init__(self, lower_bounds, upper_bounds, cs, rhs, ids=None):
        import torch
        self.lower_bounds = lower_bounds
        self.upper_bounds = upper_bounds
        self.cs = cs
        self.rhs = rhs
        self.ids = ids if ids is not None else torch.arange(len(cs))
