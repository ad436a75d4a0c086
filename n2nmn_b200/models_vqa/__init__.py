"""Mirror of the reference package ``models_vqa/`` (hot-path files only)."""
