"""Mirror of the reference package of the same name (see module docstrings)."""
