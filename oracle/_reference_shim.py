"""Import the UNMODIFIED reference modules from /root/reference (authoring container only).

The reference package cannot be imported normally here: ``stereo/modeling/__init__.py:4-31``
imports every trainer, which pulls ``easydict``, ``timm``, ``matplotlib`` (absent).  Registering
empty namespace modules for the package levels lets the sub-modules we need import without
running those ``__init__`` files.  Nothing here is available on the GPU box;
only ``tools/make_golden.py`` and ``tests/test_oracle_pins_reference.py`` use it.
"""
import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("OPENSTEREO_REFERENCE", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "stereo", "modeling"))


def _namespace(name, path):
    if name not in sys.modules:
        mod = types.ModuleType(name)
        mod.__path__ = [path]
        sys.modules[name] = mod


def load(dotted):
    """load('stereo.modeling.cost_volume.cost_volume') -> module object of the reference."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    root = os.path.join(REFERENCE_ROOT, "stereo")
    _namespace("stereo", root)
    _namespace("stereo.modeling", os.path.join(root, "modeling"))
    _namespace("stereo.modeling.models", os.path.join(root, "modeling", "models"))
    if "timm" not in sys.modules:       # only dereferenced inside backbone constructors we never build
        try:
            importlib.import_module("timm")
        except Exception:
            sys.modules["timm"] = types.ModuleType("timm")
    return importlib.import_module(dotted)


class AttrDict(dict):
    """Five-line stand-in for easydict (absent here): cfgs.MAX_DISP and cfgs.get(...)."""

    def __getattr__(self, key):
        try:
            val = self[key]
        except KeyError as exc:
            raise AttributeError(key) from exc
        return AttrDict(val) if isinstance(val, dict) else val


def load_cfg(relpath):
    import yaml
    with open(os.path.join(REFERENCE_ROOT, relpath)) as f:
        return AttrDict(yaml.safe_load(f))
