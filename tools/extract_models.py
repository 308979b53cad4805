#!/usr/bin/env python
"""Generates pyflyt_b200/models/vehicles/<name>.yaml from the reference's URDF + YAML model files.

Run in the build container (needs /root/reference).  The output is a flat link table in the base
inertial frame plus the vehicle's coefficient dictionary — numbers only, our own layout; the
reference files are parsed, not copied.  Sources (under /root/reference/PyFlyt/models/vehicles):
cf2x/, primitive_drone/, fixedwing/, acrowing/, rocket/  (SURVEY.md §A.2).
"""
import os
import sys

import yaml

sys.path.insert(0, os.path.join(os.path.dirname(os.path.realpath(__file__)), ".."))
from pyflyt_b200.models.urdf import load_urdf_links  # noqa: E402

REF = os.environ.get("PYFLYT_REFERENCE_ROOT", "/root/reference")
SRC = os.path.join(REF, "PyFlyt/models/vehicles")
DST = os.path.join(os.path.dirname(os.path.realpath(__file__)), "../pyflyt_b200/models/vehicles")


def _clean(o):
    """Drop free-text description strings; keep numbers/bools."""
    if isinstance(o, dict):
        return {k: _clean(v) for k, v in o.items() if k != "description"}
    if isinstance(o, list):
        return [_clean(v) for v in o]
    return o


def main():
    os.makedirs(DST, exist_ok=True)
    for name in ["cf2x", "primitive_drone", "fixedwing", "acrowing", "rocket"]:
        links = load_urdf_links(os.path.join(SRC, name, f"{name}.urdf"))
        with open(os.path.join(SRC, name, f"{name}.yaml"), "rb") as fh:
            params = _clean(yaml.safe_load(fh))
        doc = {
            "name": name,
            "generated_by": "tools/extract_models.py",
            "source": f"PyFlyt/models/vehicles/{name}/{name}.urdf + {name}.yaml (reference bd5ad15)",
            "frame": "base link inertial frame; link index -1 = base, i = i-th URDF joint",
            "links": [lk.to_dict() for lk in links],
            "params": params,
        }
        out = os.path.join(DST, f"{name}.yaml")
        with open(out, "w", encoding="utf-8") as fh:
            yaml.safe_dump(doc, fh, sort_keys=False, default_flow_style=None, width=140)
        print("wrote", out, len(links), "links")


if __name__ == "__main__":
    main()
