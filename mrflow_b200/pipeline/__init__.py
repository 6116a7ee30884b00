"""Drop-in replacements for the reference's ``pipeline`` stages on the plane+parallax path:
``compute_structure``, ``structure2flow`` and ``build_cost_volume`` with the reference's function
signatures (numpy in, numpy out) running on the B200."""
