"""Drop-in for ``models_shapes/nmn3_modules.py``: ``Modules`` backed by the CUDA C ABI."""
from ..modules import ModulesSHAPES as Modules  # noqa: F401
