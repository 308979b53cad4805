"""TEST INFRASTRUCTURE — stand-in for ``pybullet_utils.bullet_client.BulletClient``.

The reference's ``Aviary`` subclasses BulletClient (PyFlyt/core/aviary.py:47,104) and calls the
engine through ``self.<pybullet function>``; here every such call lands on one ``World``."""
import pybullet as _pb


class BulletClient:
    def __init__(self, connection_mode=None, options=""):
        self._world = _pb.World()
        self._client = 0

    def __getattr__(self, name):
        # only reached for names that are not instance/class attributes
        world = self.__dict__.get("_world")
        if world is not None and hasattr(world, name):
            return getattr(world, name)
        if hasattr(_pb, name):
            return getattr(_pb, name)
        raise AttributeError(name)
