"""TEST INFRASTRUCTURE ONLY.  numpy restatement of the reference's host data path:

  cords_to_map        src_deformable/utils/pose_utils.py:79-86
  pose_masks          src_deformable/utils/pose_transform.py:143-183   (+ mask_from_kp_array :127-138, estimate_polygon :186-210)
  affine_transforms   src_deformable/utils/pose_transform.py:216-289

and of the two scikit-image calls the reference makes there (the library is NOT pinned by the reference and is absent from
this image, so parity against scikit-image itself is UNPINNED; what is pinned is the reference's own code running on top
of these restatements, see ``install_into_reference``):

  skimage.measure.grid_points_in_poly(shape, verts)       -- W. R. Franklin's pnpoly as shipped in scikit-image 0.13 / 0.14
  skimage.transform.estimate_transform('affine', src, dst) -- ProjectiveTransform.estimate restricted to the affine
                                                              coefficients (scikit-image >= 0.14): normalise, SVD, de-normalise
"""
import types

import numpy as np

MISSING_VALUE = -1
LABELS = ['Rank', 'Rknee', 'Rhip', 'Lhip', 'Lknee', 'Lank', 'pelv', 'spine', 'neck', 'head', 'Rwri', 'Relb', 'Rsho', 'Lsho',
          'Lelb', 'Lwri']
LABELS_PAF = ['nose', 'neck', 'Rsho', 'Relb', 'Rwri', 'Lsho', 'Lelb', 'Lwri', 'Rhip', 'Rkne', 'Rank', 'Lhip', 'Lkne', 'Lank',
              'Leye', 'Reye', 'Lear', 'Rear']


# ----------------------------------------------------------------------------- scikit-image restatements
def grid_points_in_poly(shape, verts):
    """out[r, c] = point (r, c) inside the polygon verts[:, (row, col)] (crossing-number test, half-open edges)."""
    verts = np.asarray(verts, dtype=np.float64)
    vx, vy = verts[:, 0], verts[:, 1]
    x = np.arange(shape[0], dtype=np.float64)[:, None]
    y = np.arange(shape[1], dtype=np.float64)[None, :]
    c = np.zeros(shape, dtype=bool)
    j = len(vx) - 1
    for i in range(len(vx)):
        with np.errstate(divide="ignore", invalid="ignore"):
            cond = (((vy[i] <= y) & (y < vy[j])) | ((vy[j] <= y) & (y < vy[i]))) & \
                   (x < (vx[j] - vx[i]) * (y - vy[i]) / (vy[j] - vy[i]) + vx[i])
        c ^= cond
        j = i
    return c


class _AffineResult:
    def __init__(self, params):
        self.params = params


def estimate_transform(ttype, src, dst):
    assert ttype == 'affine'
    src, dst = np.asarray(src, dtype=np.float64), np.asarray(dst, dtype=np.float64)

    def normalise(pts):
        c = pts.mean(axis=0)
        rms = np.sqrt(((pts - c) ** 2).sum() / len(pts))
        f = np.sqrt(2) / rms
        M = np.array([[f, 0, -f * c[0]], [0, f, -f * c[1]], [0, 0, 1]])
        return M, (pts - c) * f

    Ms, s = normalise(src)
    Md, d = normalise(dst)
    rows = len(s)
    A = np.zeros((rows * 2, 9))
    A[:rows, 0], A[:rows, 1], A[:rows, 2] = s[:, 0], s[:, 1], 1
    A[:rows, 6], A[:rows, 7] = -d[:, 0] * s[:, 0], -d[:, 0] * s[:, 1]
    A[rows:, 3], A[rows:, 4], A[rows:, 5] = s[:, 0], s[:, 1], 1
    A[rows:, 6], A[rows:, 7] = -d[:, 1] * s[:, 0], -d[:, 1] * s[:, 1]
    A[:rows, 8], A[rows:, 8] = d[:, 0], d[:, 1]
    A = A[:, [0, 1, 2, 3, 4, 5, 8]]                     # affine: coefficients 0..5 only
    _, _, V = np.linalg.svd(A)
    Hm = np.zeros((3, 3))
    Hm.flat[[0, 1, 2, 3, 4, 5, 8]] = -V[-1, :-1] / V[-1, -1]
    Hm[2, 2] = 1
    return _AffineResult(np.linalg.inv(Md) @ Hm @ Ms)


def install_into_reference(ns):
    """Give the reference's pose_transform module (imported with a stubbed skimage, oracle/ref_import.py) the two
    restated functions, so that ITS pose_masks / affine_transforms run unmodified."""
    sk = ns.pose_transform.skimage
    sk.measure = types.SimpleNamespace(grid_points_in_poly=grid_points_in_poly)
    sk.transform = types.SimpleNamespace(estimate_transform=estimate_transform)
    return ns


# ----------------------------------------------------------------------------- reference functions restated
def cords_to_map(cords, img_size, sigma=6):
    result = np.zeros(tuple(img_size) + cords.shape[0:1], dtype='float32')
    for i, point in enumerate(cords):
        if point[0] == MISSING_VALUE or point[1] == MISSING_VALUE:
            continue
        xx, yy = np.meshgrid(np.arange(img_size[1]), np.arange(img_size[0]))
        result[..., i] = np.exp(-((yy - point[0]) ** 2 + (xx - point[1]) ** 2) / (2 * sigma ** 2))
    return result


def give_name_to_keypoints(array, pose_dim):
    names = LABELS if pose_dim == 16 else LABELS_PAF
    return {name: array[i][::-1] for i, name in enumerate(names)
            if array[i][0] != MISSING_VALUE and array[i][1] != MISSING_VALUE}


def compute_st_distance(kp):
    return np.sqrt((np.sum((kp['Rhip'] - kp['Rsho']) ** 2) + np.sum((kp['Lhip'] - kp['Lsho']) ** 2)) / 2.0)


def mask_from_kp_array(kp_array, border_inc, img_size):
    mn, mx = np.min(kp_array, axis=0), np.max(kp_array, axis=0)
    mn -= int(border_inc)
    mx += int(border_inc)
    mn = np.maximum(mn, 0)
    mx = np.minimum(mx, img_size[::-1])
    mask = np.zeros(img_size)
    mask[mn[1]:mx[1], mn[0]:mx[0]] = 1
    return mask


def estimate_polygon(fr, to, st, inc_to, inc_from, p_to, p_from):
    fr = fr + (fr - to) * inc_from
    to = to + (to - fr) * inc_to
    norm_vec = fr - to
    norm_vec = np.array([-norm_vec[1], norm_vec[0]])
    norm = np.linalg.norm(norm_vec)
    if norm == 0:
        return np.array([fr + 1, fr - 1, to - 1, to + 1])
    norm_vec = norm_vec / norm
    return np.array([fr + st * p_from * norm_vec, fr - st * p_from * norm_vec, to - st * p_to * norm_vec, to + st * p_to * norm_vec])


def pose_masks(array2, img_size, pose_dim):
    kp2 = give_name_to_keypoints(array2, pose_dim)
    st2 = compute_st_distance(kp2)
    empty = np.zeros(img_size)
    masks = [np.ones(img_size)]
    head = [n for n in ('Leye', 'Reye', 'Lear', 'Rear', 'nose') if n in kp2]
    if head:
        com = np.mean(np.array([kp2[n] for n in head]), axis=0, keepdims=True).astype(int)
        masks.append(mask_from_kp_array(com, 0.40 * st2, img_size))
    else:
        masks.append(empty)
    for (fr, to), inc_to in zip((('Rhip', 'Rkne'), ('Lhip', 'Lkne'), ('Rkne', 'Rank'), ('Lkne', 'Lank'), ('Rsho', 'Relb'),
                                 ('Lsho', 'Lelb'), ('Relb', 'Rwri'), ('Lelb', 'Lwri')), (0.1, 0.1, 0.5, 0.5, 0.1, 0.1, 0.5, 0.5)):
        if fr in kp2 and to in kp2:
            masks.append(grid_points_in_poly(img_size, estimate_polygon(kp2[fr], kp2[to], st2, inc_to, 0.1, 0.2, 0.2)[:, ::-1]))
        else:
            masks.append(empty)
    return np.array(masks)


def affine_transforms(array1, array2, pose_dim):
    kp1, kp2 = give_name_to_keypoints(array1, pose_dim), give_name_to_keypoints(array2, pose_dim)
    st1, st2 = compute_st_distance(kp1), compute_st_distance(kp2)
    no_point_tr = np.array([[1, 0, 1000], [0, 1, 1000], [0, 0, 1]])
    transforms = []

    def to_transforms(tr):
        try:
            np.linalg.inv(tr)
            transforms.append(tr)
        except np.linalg.LinAlgError:
            transforms.append(no_point_tr)

    torso = ['Rhip', 'Lhip', 'Lsho', 'Rsho']
    to_transforms(estimate_transform('affine', src=np.array([kp2[n] for n in torso]), dst=np.array([kp1[n] for n in torso])).params)
    head = set(n for n in ('Leye', 'Reye', 'Lear', 'Rear', 'nose') if n in kp1 and n in kp2)
    if head:
        head |= {'Lsho', 'Rsho'}
        names = list(head)
        to_transforms(estimate_transform('affine', src=np.array([kp2[n] for n in names]), dst=np.array([kp1[n] for n in names])).params)
    else:
        to_transforms(no_point_tr)

    def estimate_join(fr, to, inc_to):
        if not (fr in kp2 and to in kp2):
            return no_point_tr
        poly_2 = estimate_polygon(kp2[fr], kp2[to], st2, inc_to, 0.1, 0.2, 0.2)
        if fr in kp1 and to in kp1:
            poly_1 = estimate_polygon(kp1[fr], kp1[to], st1, inc_to, 0.1, 0.2, 0.2)
        else:
            fr = fr.replace('R', 'L') if fr[0] == 'R' else fr.replace('L', 'R')
            to = to.replace('R', 'L') if to[0] == 'R' else to.replace('L', 'R')
            if fr in kp1 and to in kp1:
                poly_1 = estimate_polygon(kp1[fr], kp1[to], st1, inc_to, 0.1, 0.2, 0.2)
            else:
                return no_point_tr
        return estimate_transform('affine', dst=poly_1, src=poly_2).params

    for (fr, to), inc in zip((('Rhip', 'Rkne'), ('Lhip', 'Lkne'), ('Rkne', 'Rank'), ('Lkne', 'Lank'), ('Rsho', 'Relb'),
                              ('Lsho', 'Lelb'), ('Relb', 'Rwri'), ('Lelb', 'Lwri')), (0.1, 0.1, 0.3, 0.3, 0.1, 0.1, 0.3, 0.3)):
        to_transforms(estimate_join(fr, to, inc))
    return np.array(transforms).reshape((-1, 9))[..., :-1]
