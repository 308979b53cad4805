"""Vehicle descriptions (host side): link tables + coefficient tables → ``PfbModel``."""
from .tables import PfbEnvConfig, PfbModel, build_model, load_vehicle  # noqa: F401
