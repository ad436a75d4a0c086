"""Drop-in for ``models_clevr/nmn3_modules.py``: ``Modules`` backed by the CUDA C ABI."""
from ..modules import ModulesCLEVR as Modules  # noqa: F401
