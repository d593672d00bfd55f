"""Swap a reference class for its B200 counterpart in every module that already holds it.

``from ocelot import *`` / ``from ocelot.cpbd.sc import *`` copy the class object into the importing
module's namespace (``ocelot.utils.section_track`` does that, and compares ``__class__ == SpaceCharge``,
section_track.py:365,383), so rebinding ``ocelot.cpbd.sc.SpaceCharge`` alone leaves stale bindings
behind.  ``swap`` walks ``sys.modules`` and rebinds every module-level name that *is* the reference
class; ``restore`` undoes exactly those rebinds.
"""
from __future__ import annotations

import sys

_undo: dict[str, list] = {}          # class name -> [(module, attribute, original object)]


def swap(name: str, new_cls):
    import ocelot.cpbd.sc as ref_sc
    old = getattr(ref_sc, name)
    if old is new_cls:
        return new_cls
    done = _undo.setdefault(name, [])
    for mod in list(sys.modules.values()):
        d = getattr(mod, "__dict__", None)
        if not isinstance(d, dict):
            continue
        for attr, val in list(d.items()):
            if val is old:
                setattr(mod, attr, new_cls)
                done.append((mod, attr, old))
    return new_cls


def restore(name: str):
    for mod, attr, old in reversed(_undo.pop(name, [])):
        setattr(mod, attr, old)
