"""TEST INFRASTRUCTURE stub."""
from . import Env, ObservationWrapper, Wrapper  # noqa: F401
