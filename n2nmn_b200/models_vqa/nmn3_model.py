"""Part 2 (layout execution) of ``models_vqa/nmn3_model.py``. The seq2seq layout generator
(Part 1) is off the hot path: callers supply ``word_vecs`` and layout tokens."""
from ..executor import LayoutExecutor


class NMN3Model(LayoutExecutor):
    def __init__(self, image_feat_grid, word_vecs, num_choices, assembler, weights=None, **kw):
        super().__init__('vqa', image_feat_grid, word_vecs, num_choices, assembler,
                         weights=weights, **kw)
