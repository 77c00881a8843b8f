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


# ---------------------------------------------------------------------------------------------------------------------------
# The BiEncoder tail (poolers, projection) lives in models/biencoder/modeling_biencoder.py, whose import chain pulls in optional
# third-party extensions this image does not have (flash-attn's fused dropout_layer_norm / fused_dense_lib, megablocks, ...).  None
# of them is CALLED by the pooler path; they are satisfied with empty stand-in modules so that the reference's own classes
# (MultiHeadAttentionPooling, FlashAttentionPooling, ClsSelector, MLP / GatedMLP) import untouched.
_OPTIONAL_THIRD_PARTY = ("dropout_layer_norm", "fused_dense_lib", "rotary_emb", "xentropy_cuda_lib", "megablocks", "stk",
                         "grouped_gemm", "deepspeed", "wandb", "webdataset", "open_clip", "timm")


def load_biencoder_tail():
    """Returns (modeling_biencoder module, layers.attention module) of the unmodified reference."""
    import importlib.abc
    import importlib.machinery
    load()

    class _Stub(types.ModuleType):
        __path__: list = []

        def __getattr__(self, name):
            if name.startswith("__"):
                raise AttributeError(name)
            return type(name, (), {})

    class _OptionalStubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
        def find_spec(self, fullname, path, target=None):
            if fullname.split(".")[0] in _OPTIONAL_THIRD_PARTY:
                try:  # a real installation wins
                    for f in sys.meta_path:
                        if f is not self and hasattr(f, "find_spec") and f.find_spec(fullname, path, target) is not None:
                            return None
                except Exception:
                    pass
                return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
            return None

        def create_module(self, spec):
            return _Stub(spec.name)

        def exec_module(self, module):
            pass

    if not any(type(f).__name__ == "_OptionalStubFinder" for f in sys.meta_path):
        sys.meta_path.append(_OptionalStubFinder())
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", SyntaxWarning)
        mb = importlib.import_module("contrastors.models.biencoder.modeling_biencoder")
        att = importlib.import_module("contrastors.layers.attention")
    return mb, att
