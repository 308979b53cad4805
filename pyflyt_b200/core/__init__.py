"""Host-side mirror of ``PyFlyt.core`` for the batched stepper."""
from .aviary import AviaryInitException, BatchedAviary  # noqa: F401
