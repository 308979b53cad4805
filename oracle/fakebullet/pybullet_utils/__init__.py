"""TEST INFRASTRUCTURE — stand-in for ``pybullet_utils`` (see ../pybullet.py)."""
