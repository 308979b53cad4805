"""TEST INFRASTRUCTURE stub."""
REGISTRY = {}


def register(id, entry_point=None, **kwargs):
    REGISTRY[id] = (entry_point, kwargs)
