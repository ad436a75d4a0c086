"""Mirror of the reference package ``models_clevr/`` (hot-path files only)."""
