"""Drop-in for ``models_vqa/nmn3_modules.py``: ``Modules`` backed by the CUDA C ABI."""
from ..modules import ModulesVQA as Modules  # noqa: F401
