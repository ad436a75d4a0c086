"""Mirror of the reference package ``models_shapes/`` (hot-path files only)."""
