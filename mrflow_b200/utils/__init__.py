"""Drop-in replacements for ``utils/flow_homography.py`` and ``utils/robust_functions.py``."""
