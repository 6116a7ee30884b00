"""Drop-in replacements for the reference's ``rigidity`` modules next to the plane+parallax path:
``occlusions.computeOcclusionsFromConsistency`` (producer of the path's occlusion input) and the device
front-end of ``inference.infer_mrf`` (edge weights + int32 unaries for the unchanged C++ solver)."""
