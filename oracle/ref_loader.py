"""Import the UNMODIFIED reference (read-only, /root/reference) in the build container.

TEST INFRASTRUCTURE ONLY.  Used by ``oracle/gen_golden.py`` (fixture generation) and by the
``ref``-marked tests that cross-check the oracle against the live reference when it is present.
/root/reference does not exist on the GPU box; there the same modules come from ``baseline/_ref`` (the unmodified
package installed by ``baseline/install_ref.sh``), used only by ``bench.py``'s baseline legs (``--impl reference``,
``cpu_baseline``, ``gpu_baseline``): the reference is the thing TIMED there, never part of the product path.

Bypass: ``src/contrastors/__init__.py:1`` star-imports flash-attn-only
modules, so we register an empty ``contrastors`` package whose ``__path__`` points at the
reference tree and import the leaf modules we need (``loss``, ``distributed``, ``rand_state``,
``models.huggingface.modeling_hf_nomic_bert``) untouched.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
# the read-only tree in the build container, else the unmodified install that travels to the GPU box
# (baseline/install_ref.sh -> baseline/_ref, git-ignored)
_CANDIDATES = ["/root/reference/src/contrastors", os.path.join(os.path.dirname(_HERE), "baseline", "_ref", "contrastors")]
REF_ROOT = next((p for p in _CANDIDATES if os.path.isdir(p)), _CANDIDATES[0])


def available() -> bool:
    return os.path.isdir(REF_ROOT)


def source() -> str:
    return "tree" if REF_ROOT == _CANDIDATES[0] else "baseline/_ref"


def _pkg(name: str, path: str):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    m.__path__ = [path]
    sys.modules[name] = m
    return m


def load():
    """Returns a namespace with the reference's loss / distributed / rand_state / HF encoder modules."""
    if not available():
        raise RuntimeError("reference tree not present (expected only in the build container)")
    _pkg("contrastors", REF_ROOT)
    _pkg("contrastors.models", os.path.join(REF_ROOT, "models"))
    _pkg("contrastors.models.huggingface", os.path.join(REF_ROOT, "models", "huggingface"))
    ns = types.SimpleNamespace()
    ns.loss = importlib.import_module("contrastors.loss")
    ns.distributed = importlib.import_module("contrastors.distributed")
    ns.rand_state = importlib.import_module("contrastors.rand_state")
    ns.hf_cfg = importlib.import_module("contrastors.models.huggingface.configuration_hf_nomic_bert")
    ns.hf = importlib.import_module("contrastors.models.huggingface.modeling_hf_nomic_bert")
    return ns
