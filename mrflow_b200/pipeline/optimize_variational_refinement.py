"""``pipeline/optimize_variational_refinement.py`` of the reference on the B200: the stage between
``compute_structure`` and ``combine_flow``.  With ``params.override_optimization`` (the reference's
fast mode, README.md:68 and BASELINE.json configs[0]) the stage is the occlusion-aware merge of the two
structures (:20-73) and runs as one kernel.  Without it the stage continues into the variational solver
(structure_refinement/optimize_structure_multi_scale.py -> optimize_structure_single_scale.py:310-652), which
runs on the GPU as well (mrflow_b200.variational)."""
import numpy as np

from .. import device
from .._host import down, flag, require_cuda, up


def optimize(images, structures, rigidity, epipoles, homographies, B, mu, occlusions, params):
    """pipeline/optimize_variational_refinement.py:20-101.  Returns ``[structure]`` like the reference."""
    require_cuda()
    reasoning = int(getattr(params, "occlusion_reasoning", 1))
    merged, occ_b, occ_f = device.merge_structures(up(structures[0], np.float64), up(structures[1], np.float64),
                                                   flag(np.asarray(occlusions[0]) != 0),
                                                   flag(np.asarray(occlusions[1]) != 0), mu[0], mu[1], reasoning,
                                                   want_blurred=True)
    structure_optimized = down(merged)
    if getattr(params, "override_optimization", 0):
        return [structure_optimized, ]                                   # :72-73
    from ..structure_refinement import optimize_structure_multi_scale
    occ_b, occ_f = occ_b.cpu().numpy() > 0, occ_f.cpu().numpy() > 0        # :23-24, from the kernel
    S0 = np.asarray(structures[0]) * mu[1] / mu[0]                        # :41
    A_optimized, _ = optimize_structure_multi_scale.optimizeStructureMultiScale(
        structure_optimized, S0, structures[1], occ_b, occ_f, rigidity, images, homographies, B, epipoles, mu, params)   # :82-89
    return [A_optimized, ]
