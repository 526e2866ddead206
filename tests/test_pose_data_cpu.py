"""CPU: the host side of the device data path.  The oracle restatement (oracle/pose_data.py) is pinned
against the REFERENCE's own pose_masks / affine_transforms / cords_to_map running unmodified on top of the two restated
scikit-image functions (the library itself is absent: parity against scikit-image is unpinned and says so), and the
product's host function pose_geometry.affine_transforms against the oracle."""
import numpy as np
import pytest

from oracle import pose_data as od, ref_import, synth
from pose_transfer_b200.utils import pose_geometry as pg

CASES = [(18, 256, 256, 3), (16, 224, 224, 4), (18, 128, 64, 5)]


@pytest.mark.skipif(not ref_import.available(), reason="reference tree not available")
@pytest.mark.parametrize("P,H,W,seed", CASES)
def test_oracle_matches_reference_functions(P, H, W, seed):
    ns = od.install_into_reference(ref_import.load())
    kp = synth.make_keypoints(5, H, W, P, seed=seed).numpy()
    for n in range(4):
        assert np.array_equal(ns.pose_transform.pose_masks(kp[n + 1], (H, W), P), od.pose_masks(kp[n + 1], (H, W), P))
        assert np.array_equal(ns.pose_utils.cords_to_map(kp[n], (H, W)), od.cords_to_map(kp[n], (H, W)))
        a, b = ns.pose_transform.affine_transforms(kp[n], kp[n + 1], P), od.affine_transforms(kp[n], kp[n + 1], P)
        assert a.shape == (10, 8) and np.abs(a - b).max() <= 1e-9


@pytest.mark.parametrize("P,H,W,seed", CASES)
def test_product_affine_transforms_match_oracle(P, H, W, seed):
    kp = synth.make_keypoints(6, H, W, P, seed=seed + 10, missing=0.25).numpy()
    saw_no_point = False
    for n in range(5):
        got, want = pg.affine_transforms(kp[n], kp[n + 1], P), od.affine_transforms(kp[n], kp[n + 1], P)
        assert got.shape == want.shape == (10, 8)
        assert np.abs(got - want).max() <= 1e-9 * max(1.0, np.abs(want).max())
        saw_no_point |= bool((want[:, 2] == 1000).any())
    assert saw_no_point or P == 18          # missing limbs produce no_point_tr rows (pose_transform.py:221)
    with pytest.raises(KeyError):           # the reference raises on a missing torso key-point
        bad = kp[0].copy()
        bad[pg.LABELS_PAF.index('Rhip') if P == 18 else pg.LABELS.index('Rhip')] = -1
        pg.affine_transforms(bad, kp[1], P)


def test_restated_point_in_polygon_known_answers():
    sq = np.array([[1.0, 1.0], [1.0, 4.0], [4.0, 4.0], [4.0, 1.0]])          # rows, cols
    m = od.grid_points_in_poly((6, 6), sq)
    assert m[2, 2] and m[3, 3] and not m[0, 0] and not m[5, 5]
    assert m.sum() == 9                                                        # half-open: [1, 4) x [1, 4)
    tri = np.array([[0.5, 0.5], [0.5, 5.5], [5.5, 0.5]])
    t = od.grid_points_in_poly((6, 6), tri)
    assert t[1, 1] and t[1, 4] and not t[4, 4] and not t[5, 5]
