"""neuralsat_b200 — B200-native batched backward bound propagation (CROWN / alpha-beta-CROWN) behind
NeuralSAT's own BoundedModule / NetworkAbstractor interface.  See DESIGN.md."""
from .bounded_module import BoundedModule, BoundedTensor, PerturbationLpNorm, SparseBeta, stop_criterion_batch_any  # noqa: F401
