"""Drop-in for ``models_shapes/nmn3_assembler.py`` — see n2nmn_b200/assembler.py."""
from ..assembler import (Assembler, INVALID_EXPR, MODULE_SIGNATURE, _module_input_num,  # noqa: F401
                         _module_output_type, build_validity_mats as _build_validity_mats)
