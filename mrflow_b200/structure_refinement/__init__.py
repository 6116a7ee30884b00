"""Drop-in replacement for ``structure_refinement/interpolate_through_H.py`` (sampling part)."""
