"""Static shape/vocabulary configuration of the three reference model families.

The reference hard-codes these numbers in three near-identical packages:
  * CLEVR : models_clevr/nmn3_modules.py:74,185 (map_dim=250, kernel 5), 14 modules
  * SHAPES: models_shapes/nmn3_modules.py:28,71  (map_dim=500, kernel 3), 4 modules
  * VQA   : models_vqa/nmn3_modules.py:84,123,194 (map_dim=1024, +2 coord channels)
Here one table drives the host logic and the CUDA context.
"""
from __future__ import annotations

from dataclasses import dataclass, field

# Module opcodes shared by the host scheduler (csrc/schedule.cpp), the kernels and Python.
# Order is fixed by include/n2nmn_b200.h (enum n2nmn_op).
OP_SCENE = 0
OP_FIND = 1
OP_FILTER = 2
OP_FIND_SAME_PROPERTY = 3   # also the VQA "_Transform" (pooled variant)
OP_TRANSFORM = 4            # conv variant (CLEVR/SHAPES)
OP_AND = 5
OP_OR = 6
OP_EXIST = 7                # also the SHAPES "_Answer"
OP_COUNT = 8
OP_EQUAL_NUM = 9
OP_MORE_NUM = 10
OP_LESS_NUM = 11
OP_SAME_PROPERTY = 12
OP_DESCRIBE = 13
NUM_OPS = 14

OP_NAMES = ['Scene', 'Find', 'Filter', 'FindSameProperty', 'Transform', 'And', 'Or', 'Exist',
            'Count', 'EqualNum', 'MoreNum', 'LessNum', 'SameProperty', 'Describe']

# arity (# attention inputs) and whether the op yields an answer, indexed by opcode
OP_ARITY = [0, 0, 1, 1, 1, 2, 2, 1, 1, 2, 2, 2, 2, 1]
OP_IS_ANS = [0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 1, 1, 1]


@dataclass(frozen=True)
class FamilyConfig:
    """One reference model family (CLEVR / SHAPES / VQA)."""
    name: str
    map_dim: int
    kernel_size: int
    text_dim: int = 300
    add_coords: bool = False
    # layout-token name -> (opcode, TF variable scope that holds its weights)
    token_ops: dict = field(default_factory=dict)


CLEVR = FamilyConfig(
    name='clevr', map_dim=250, kernel_size=5,
    token_ops={
        '_Scene': OP_SCENE, '_Find': OP_FIND, '_Filter': OP_FILTER,
        '_FindSameProperty': OP_FIND_SAME_PROPERTY, '_Transform': OP_TRANSFORM,
        '_And': OP_AND, '_Or': OP_OR, '_Exist': OP_EXIST, '_Count': OP_COUNT,
        '_EqualNum': OP_EQUAL_NUM, '_MoreNum': OP_MORE_NUM, '_LessNum': OP_LESS_NUM,
        '_SameProperty': OP_SAME_PROPERTY, '_Describe': OP_DESCRIBE,
    })

SHAPES = FamilyConfig(
    name='shapes', map_dim=500, kernel_size=3,
    token_ops={'_Find': OP_FIND, '_Transform': OP_TRANSFORM, '_And': OP_AND,
               '_Answer': OP_EXIST})

VQA = FamilyConfig(
    name='vqa', map_dim=1024, kernel_size=5, add_coords=True,
    token_ops={'_Find': OP_FIND, '_Transform': OP_FIND_SAME_PROPERTY, '_And': OP_AND,
               '_Describe': OP_DESCRIBE})

FAMILIES = {'clevr': CLEVR, 'shapes': SHAPES, 'vqa': VQA}
