"""Inputs of the variational refinement test case, built from a golden fixture the way stage 5 hands them over
(pipeline/optimize_variational_refinement.py:79-87): the merged structure as A_init, the two per-frame
structures as references, exclusive occlusion maps, the refined rigidity."""
import numpy as np

LAMBDA_1, LAMBDA_2 = 0.5, 5.0        # parameters.py:38-39
CASES = (("lc0", 0.0, 3), ("lc01", 0.1, 2))     # (name, lambda_consistency, N_outer)


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
