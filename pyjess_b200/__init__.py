"""pyjess_b200 -- B200-native implementation of PyJess's template-query hot path.

Drop-in for the reference's public classes (``src/pyjess/__init__.py:3,6``):
`Atom`, `Molecule`, `TemplateAtom`, `Template`, `Jess`, `Query`, `Hit`, `Mat4`.
The search itself runs ONLY on the GPU through the C ABI of
``include/jess_b200.h`` (``pyjess_b200/csrc``); there is no CPU fallback.
"""
from ._model import Atom, Mat4, Molecule, Template, TemplateAtom

__version__ = "0.1.0"
__all__ = ["Atom", "Mat4", "Molecule", "Template", "TemplateAtom", "Jess", "Query", "Hit"]


def __getattr__(name):
    if name in ("Jess", "Query", "Hit"):
        from . import _jess
        return getattr(_jess, name)
    raise AttributeError(name)


def __dir__():
    return sorted(set(globals()) | set(__all__))
