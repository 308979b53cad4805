"""TEST INFRASTRUCTURE — stand-in for the ``pybullet_data`` package (only ``plane.urdf`` is used:
/root/reference/PyFlyt/core/aviary.py:207,240)."""
import os


def getDataPath() -> str:
    return os.path.dirname(os.path.realpath(__file__))
