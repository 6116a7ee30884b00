"""Inputs of the variational refinement test case, built from a golden fixture the way stage 5 hands them over
(pipeline/optimize_variational_refinement.py:79-87): the merged structure as A_init, the two per-frame
structures as references, exclusive occlusion maps, the refined rigidity."""
import numpy as np

LAMBDA_1, LAMBDA_2 = 0.5, 5.0        # parameters.py:38-39
# (name, lambda_consistency, N_outer, with_nonrigid): the fixtures' own rigidity maps are all-rigid, so one case
# carves a non-rigid block and speckles into the map (Z masking, eroded edge weights, zeroed data term)
CASES = (("lc0", 0.0, 3, False), ("lc01", 0.1, 2, False), ("mask", 0.05, 2, True))


def with_nonrigid(vin):
    out = dict(vin)
    rig = np.array(vin["rigidity"], dtype=bool, copy=True)
    h, w = rig.shape
    rig[h // 4:h // 2 + 5, w // 3:2 * w // 3] = False
    rs = np.random.RandomState(77)
    rig[rs.rand(h, w) < 0.03] = False
    out["rigidity"] = rig
    return out


def variational_inputs(t, g):
    occ = g["full_occ"].astype(bool)
    occ_b = occ[0] & ~occ[1]
    occ_f = occ[1] & ~occ[0]
    return dict(A_init=np.nan_to_num(g["full_merged1"], posinf=0.0, neginf=0.0),
                A_bwd_reference=np.nan_to_num(g["full_S"][0], posinf=0.0, neginf=0.0),
                A_fwd_reference=np.nan_to_num(g["full_S"][1], posinf=0.0, neginf=0.0),
                occlusions_bwd=occ_b, occlusions_fwd=occ_f, occlusions_both=occ[0] & occ[1],
                rigidity=g["full_rigidity"].astype(bool), images=[np.asarray(I, dtype=np.float64) for I in t["images"]],
                H_ar=[np.asarray(H, dtype=np.float64) for H in t["homographies"]], B_ar=[float(b) for b in g["full_B"]],
                q_ar=[np.asarray(q, dtype=np.float64) for q in g["full_q"]], mu_ar=[float(m) for m in g["full_mu"]])


def stage5_inputs(t, g):
    """What pipeline/optimize_variational_refinement.py:20-89 hands to the solver for the fixture's full-stage
    outputs (median-blurred occlusion maps :23-24, rescaled backward structure :41, merged initialisation)."""
    import cv2
    occ_b = cv2.medianBlur(g["full_occ"][0].astype("uint8"), ksize=3) > 0
    occ_f = cv2.medianBlur(g["full_occ"][1].astype("uint8"), ksize=3) > 0
    both = occ_f * occ_b
    occ_f = occ_f.copy(); occ_b = occ_b.copy()
    occ_f[both] = 0; occ_b[both] = 0            # optimize_structure_multi_scale.py:80-83
    mu = [float(m) for m in g["full_mu"]]
    return dict(A_init=g["full_merged1"], A_bwd_reference=g["full_S"][0] * mu[1] / mu[0], A_fwd_reference=g["full_S"][1],
                occlusions_bwd=occ_b, occlusions_fwd=occ_f, occlusions_both=both, rigidity=g["full_rigidity"].astype(bool),
                images=[np.asarray(I, dtype=np.float64) for I in t["images"]],
                H_ar=[np.asarray(H, dtype=np.float64) for H in t["homographies"]], B_ar=[float(b) for b in g["full_B"]],
                q_ar=[np.asarray(q, dtype=np.float64) for q in g["full_q"]], mu_ar=mu)
