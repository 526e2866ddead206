"""Host side of the device data path: the tiny per-sample pose geometry that stays on the CPU --
``affine_transforms`` (src_deformable/utils/pose_transform.py:216-289: ten 3x3 inverse affine maps per sample,
80 numbers) -- and the key-point name tables shared with the CUDA kernels (csrc/pose_data.cu).

``skimage.transform.estimate_transform('affine', src, dst)`` is a third-party call of the reference (scikit-image is not
pinned by the reference and absent from this image): ``estimate_affine`` restates its published algorithm
(ProjectiveTransform.estimate restricted to the six affine coefficients, scikit-image >= 0.14: Hartley normalisation of
both point sets, total-least-squares solution = right singular vector of the smallest singular value, de-normalisation).
"""
import numpy as np

MISSING_VALUE = -1
# utils/pose_utils.py:27,36-37
LABELS = ['Rank', 'Rknee', 'Rhip', 'Lhip', 'Lknee', 'Lank', 'pelv', 'spine', 'neck', 'head', 'Rwri', 'Relb', 'Rsho', 'Lsho',
          'Lelb', 'Lwri']
LABELS_PAF = ['nose', 'neck', 'Rsho', 'Relb', 'Rwri', 'Lsho', 'Lelb', 'Lwri', 'Rhip', 'Rkne', 'Rank', 'Lhip', 'Lkne', 'Lank',
              'Leye', 'Reye', 'Lear', 'Rear']
NO_POINT_TR = np.array([[1, 0, 1000], [0, 1, 1000], [0, 0, 1]], dtype=np.float64)      # pose_transform.py:221
JOINTS = (('Rhip', 'Rkne'), ('Lhip', 'Lkne'), ('Rkne', 'Rank'), ('Lkne', 'Lank'), ('Rsho', 'Relb'), ('Lsho', 'Lelb'),
          ('Relb', 'Rwri'), ('Lelb', 'Lwri'))
INC_TO_TRANSFORM = (0.1, 0.1, 0.3, 0.3, 0.1, 0.1, 0.3, 0.3)                              # pose_transform.py:278-288


def named_keypoints(array, pose_dim):
    """give_name_to_keypoints (pose_transform.py:94-104): name -> (x, y) float64 for the key-points that are present."""
    labels = LABELS if pose_dim == 16 else LABELS_PAF
    return {name: np.asarray(array[i][::-1], dtype=np.float64) for i, name in enumerate(labels)
            if array[i][0] != MISSING_VALUE and array[i][1] != MISSING_VALUE}


def require_torso(array, pose_dim):
    """The reference dereferences Rhip / Lhip / Rsho / Lsho unconditionally (compute_st_distance, :121-124) and raises
    KeyError when one is missing; the device kernels cannot, so the host checks first."""
    kp = named_keypoints(array, pose_dim)
    for name in ('Rhip', 'Rsho', 'Lhip', 'Lsho'):
        if name not in kp:
            raise KeyError(name)
    return kp


def st_distance(kp):
    return np.sqrt((np.sum((kp['Rhip'] - kp['Rsho']) ** 2) + np.sum((kp['Lhip'] - kp['Lsho']) ** 2)) / 2.0)


def estimate_polygon(fr, to, st, inc_to, inc_from=0.1, p_to=0.2, p_from=0.2):
    """pose_transform.py:186-210."""
    fr = fr + (fr - to) * inc_from
    to = to + (to - fr) * inc_to
    nv = fr - to
    nv = np.array([-nv[1], nv[0]])
    norm = np.linalg.norm(nv)
    if norm == 0:
        return np.array([fr + 1, fr - 1, to - 1, to + 1])
    nv = nv / norm
    return np.array([fr + st * p_from * nv, fr - st * p_from * nv, to - st * p_to * nv, to + st * p_to * nv])


def _center_and_normalize(points):
    centroid = np.mean(points, axis=0)
    rms = np.sqrt(np.sum((points - centroid) ** 2) / points.shape[0])
    if rms == 0:
        raise ZeroDivisionError
    nf = np.sqrt(2.0) / rms
    m = np.array([[nf, 0, -nf * centroid[0]], [0, nf, -nf * centroid[1]], [0, 0, 1]])
    ph = np.vstack([points.T, np.ones(points.shape[0])])
    q = m @ ph
    return m, (q[:2] / q[2]).T


def estimate_affine(src, dst):
    """3x3 matrix of skimage.transform.estimate_transform('affine', src=src, dst=dst).params (see module docstring)."""
    src, dst = np.asarray(src, dtype=np.float64), np.asarray(dst, dtype=np.float64)
    try:
        sm, s = _center_and_normalize(src)
        dm, d = _center_and_normalize(dst)
    except ZeroDivisionError:
        return np.full((3, 3), np.nan)
    n = s.shape[0]
    A = np.zeros((2 * n, 7))
    A[:n, 0], A[:n, 1], A[:n, 2] = s[:, 0], s[:, 1], 1
    A[n:, 3], A[n:, 4], A[n:, 5] = s[:, 0], s[:, 1], 1
    A[:n, 6], A[n:, 6] = d[:, 0], d[:, 1]
    _, _, V = np.linalg.svd(A)
    if np.isclose(V[-1, -1], 0):
        return np.full((3, 3), np.nan)
    Hm = np.zeros((3, 3))
    Hm.flat[[0, 1, 2, 3, 4, 5]] = -V[-1, :-1] / V[-1, -1]
    Hm[2, 2] = 1
    return np.linalg.inv(dm) @ Hm @ sm


def affine_transforms(array1, array2, pose_dim):
    """pose_transform.py:216-289: [10, 8] = the first 8 entries of ten row-major 3x3 maps from the TARGET pose (array2) to
    the SOURCE pose (array1) -- body, head, eight limbs; ``no_point_tr`` where a part cannot be estimated."""
    kp1, kp2 = require_torso(array1, pose_dim), require_torso(array2, pose_dim)
    st1, st2 = st_distance(kp1), st_distance(kp2)
    out = []

    def push(tr):
        tr = np.asarray(tr, dtype=np.float64)
        try:
            np.linalg.inv(tr)
            out.append(tr)
        except np.linalg.LinAlgError:
            out.append(NO_POINT_TR)

    torso = ['Rhip', 'Lhip', 'Lsho', 'Rsho']
    push(estimate_affine([kp2[k] for k in torso], [kp1[k] for k in torso]))
    head = [k for k in ('Leye', 'Reye', 'Lear', 'Rear', 'nose') if k in kp1 and k in kp2]
    if head:
        # the reference iterates a Python set (arbitrary but identical order for both point lists): the least-squares
        # solution does not depend on the order of the correspondences
        names = sorted(set(head) | {'Lsho', 'Rsho'})
        push(estimate_affine([kp2[k] for k in names], [kp1[k] for k in names]))
    else:
        push(NO_POINT_TR)
    for (fr, to), inc_to in zip(JOINTS, INC_TO_TRANSFORM):
        if not (fr in kp2 and to in kp2):
            push(NO_POINT_TR)
            continue
        poly2 = estimate_polygon(kp2[fr], kp2[to], st2, inc_to)
        f1, t1 = fr, to
        if not (f1 in kp1 and t1 in kp1):
            swap = {'R': 'L', 'L': 'R'}
            f1, t1 = swap[fr[0]] + fr[1:], swap[to[0]] + to[1:]       # the mirrored limb of the source pose (:262-270)
            if not (f1 in kp1 and t1 in kp1):
                push(NO_POINT_TR)
                continue
        poly1 = estimate_polygon(kp1[f1], kp1[t1], st1, inc_to)
        push(estimate_affine(poly2, poly1))
    return np.array(out).reshape((-1, 9))[..., :-1]
