// Device-side data path: the per-sample numpy / skimage preprocessing of the reference's
// datasets/PoseTransfer_Dataset.py:78-109,163-189 -- Gaussian pose heat-maps (utils/pose_utils.py:79-86 cords_to_map) and the
// ten body-part masks (utils/pose_transform.py:143-214 pose_masks, estimate_polygon, mask_from_kp_array) -- computed on the
// GPU from the raw key-points, so that a training batch crosses PCIe as images + a few hundred integers instead of
// 2P float heat-map planes and ten float64 mask planes per sample (14.6 MB / image at 256x256).
// The per-part affine transforms (pose_transform.py:216-289: ten 3x3 least-squares fits per sample) stay on the host
// (pose_transfer_b200/utils/pose_geometry.py): 80 floats per sample.
// (compiled with -fmad=false: the polygon vertices must round exactly like numpy's separate multiply / add)
#include "common.cuh"

namespace ptk {

constexpr int kMissing = -1;      // utils/pose_utils.py:42 MISSING_VALUE

// out[n, c0 + p, y, x] = exp(-((y - ky)^2 + (x - kx)^2) / (2 sigma^2)), zero plane for a missing key-point.
// The reference evaluates the exponent in float64 and stores float32 (result is a float32 array): same here.
__global__ void __launch_bounds__(256)
pose_heatmaps_kernel(const int* __restrict__ kp, int P, int H, int W, double two_sigma2, float* __restrict__ out, int C_total, int c0) {
  pdl_trigger();
  const int n = blockIdx.z, p = blockIdx.y;
  const int ky = kp[(n * P + p) * 2], kx = kp[(n * P + p) * 2 + 1];
  const bool missing = ky == kMissing || kx == kMissing;
  float* plane = out + ((int64_t)n * C_total + c0 + p) * H * W;
  const int HW = H * W;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += gridDim.x * blockDim.x) {
    float v = 0.f;
    if (!missing) {
      const int y = i / W, x = i - y * W;
      const long long d2 = (long long)(y - ky) * (y - ky) + (long long)(x - kx) * (x - kx);
      v = (float)exp(-((double)d2 / two_sigma2));      // numpy: -(int64 / int) -> float64 true division, exp, cast
    }
    plane[i] = v;
  }
}

struct PoseNames {          // key-point indices by name for the label set in use (-1: the name does not exist in the set)
  int st[4];                // Rhip, Rsho, Lhip, Lsho                        (compute_st_distance, pose_transform.py:121-124)
  int head[5];              // Leye, Reye, Lear, Rear, nose                  (pose_transform.py:152)
  int jfrom[8], jto[8];     // Rhip-Rkne, Lhip-Lkne, Rkne-Rank, Lkne-Lank, Rsho-Relb, Lsho-Lelb, Relb-Rwri, Lelb-Lwri
};

__device__ __forceinline__ bool kp_present(const int* kp, int idx) {
  return idx >= 0 && kp[idx * 2] != kMissing && kp[idx * 2 + 1] != kMissing;
}

// point-in-polygon of skimage.measure.grid_points_in_poly (W. R. Franklin's pnpoly as shipped by scikit-image 0.13/0.14,
// skimage/_shared/_geometry / measure/_pnpoly.pyx): vx = rows, vy = columns of the vertices, (x, y) = (row, column) tested.
__device__ __forceinline__ bool pnpoly4(const double* vx, const double* vy, double x, double y) {
  bool c = false;
  int j = 3;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if ((((vy[i] <= y) && (y < vy[j])) || ((vy[j] <= y) && (y < vy[i]))) &&
        (x < (vx[j] - vx[i]) * (y - vy[i]) / (vy[j] - vy[i]) + vx[i]))
      c = !c;
    j = i;
  }
  return c;
}

// masks[n, part, y, x] (float64, the reference's dtype).  grid (chunks, 10, N).
__global__ void __launch_bounds__(256)
pose_masks_kernel(const int* __restrict__ kps, int P, int H, int W, const __grid_constant__ PoseNames nm, double* __restrict__ masks) {
  pdl_trigger();
  const int n = blockIdx.z, part = blockIdx.y;
  const int* kp = kps + (int64_t)n * P * 2;
  double* plane = masks + ((int64_t)n * 10 + part) * H * W;
  __shared__ double s_v[8];      // polygon: rows[4], cols[4]
  __shared__ int s_rect[4];      // head rectangle: y0, y1, x0, x1
  __shared__ int s_kind;         // 0 empty, 1 all ones, 2 rectangle, 3 polygon
  if (threadIdx.x == 0) {
    int kind = 0;
    // st2 (pose_transform.py:121-124); key-points are stored (y, x) and used as (x, y) (array[i][::-1], :98,103)
    double st = 0.0;
    {
      const double rhx = kp[nm.st[0] * 2 + 1], rhy = kp[nm.st[0] * 2], rsx = kp[nm.st[1] * 2 + 1], rsy = kp[nm.st[1] * 2];
      const double lhx = kp[nm.st[2] * 2 + 1], lhy = kp[nm.st[2] * 2], lsx = kp[nm.st[3] * 2 + 1], lsy = kp[nm.st[3] * 2];
      const double d1 = (rhx - rsx) * (rhx - rsx) + (rhy - rsy) * (rhy - rsy);
      const double d2 = (lhx - lsx) * (lhx - lsx) + (lhy - lsy) * (lhy - lsy);
      st = sqrt((d1 + d2) / 2.0);
    }
    if (part == 0) {
      kind = 1;                                    // body mask: all ones (pose_transform.py:149-150)
    } else if (part == 1) {
      double sx = 0.0, sy = 0.0;
      int cnt = 0;
      for (int q = 0; q < 5; ++q)
        if (kp_present(kp, nm.head[q])) { sx += kp[nm.head[q] * 2 + 1]; sy += kp[nm.head[q] * 2]; ++cnt; }
      if (cnt > 0) {
        // center_of_mass.astype(int) truncates; border = int(0.40 * st2); clipped to the image (mask_from_kp_array :127-138)
        const int cx = (int)(sx / cnt), cy = (int)(sy / cnt);
        const int border = (int)(0.40 * st);
        int x0 = cx - border, x1 = cx + border, y0 = cy - border, y1 = cy + border;
        x0 = x0 < 0 ? 0 : x0; y0 = y0 < 0 ? 0 : y0;
        x1 = x1 > W ? W : x1; y1 = y1 > H ? H : y1;
        s_rect[0] = y0; s_rect[1] = y1; s_rect[2] = x0; s_rect[3] = x1;
        kind = 2;
      }
    } else {
      const int jq = part - 2;
      const int a = nm.jfrom[jq], b = nm.jto[jq];
      if (kp_present(kp, a) && kp_present(kp, b)) {
        const double inc_to = (jq == 2 || jq == 3 || jq == 6 || jq == 7) ? 0.5 : 0.1;     // pose_transform.py:172-182
        const double inc_from = 0.1, p_to = 0.2, p_from = 0.2;
        // estimate_polygon (pose_transform.py:186-210), points as (x, y); `to` uses the already extended `fr`
        double frx = kp[a * 2 + 1], fry = kp[a * 2], tox = kp[b * 2 + 1], toy = kp[b * 2];
        frx = frx + (frx - tox) * inc_from; fry = fry + (fry - toy) * inc_from;
        tox = tox + (tox - frx) * inc_to;   toy = toy + (toy - fry) * inc_to;
        double nx = -(fry - toy), ny = frx - tox;
        // np.linalg.norm = sqrt(dot(v, v)); the BLAS dot accumulates with a fused multiply-add: sqrt(fma(v1, v1, v0 * v0))
        const double norm = sqrt(fma(ny, ny, nx * nx));
        double vx[4], vy[4];       // polygon vertices (x, y)
        if (norm == 0.0) {
          vx[0] = frx + 1; vy[0] = fry + 1; vx[1] = frx - 1; vy[1] = fry - 1;
          vx[2] = tox - 1; vy[2] = toy - 1; vx[3] = tox + 1; vy[3] = toy + 1;
        } else {
          nx /= norm; ny /= norm;
          vx[0] = frx + st * p_from * nx; vy[0] = fry + st * p_from * ny;
          vx[1] = frx - st * p_from * nx; vy[1] = fry - st * p_from * ny;
          vx[2] = tox - st * p_to * nx;   vy[2] = toy - st * p_to * ny;
          vx[3] = tox + st * p_to * nx;   vy[3] = toy + st * p_to * ny;
        }
        // grid_points_in_poly(img_size, polygon[:, ::-1]): vertices as (row, column)
        for (int q = 0; q < 4; ++q) { s_v[q] = vy[q]; s_v[4 + q] = vx[q]; }
        kind = 3;
      }
    }
    s_kind = kind;
  }
  __syncthreads();
  const int kind = s_kind;
  const int HW = H * W;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += gridDim.x * blockDim.x) {
    double v = 0.0;
    if (kind == 1) {
      v = 1.0;
    } else if (kind == 2) {
      const int y = i / W, x = i - y * W;
      v = (y >= s_rect[0] && y < s_rect[1] && x >= s_rect[2] && x < s_rect[3]) ? 1.0 : 0.0;
    } else if (kind == 3) {
      const int y = i / W, x = i - y * W;
      v = pnpoly4(s_v, s_v + 4, (double)y, (double)x) ? 1.0 : 0.0;
    }
    plane[i] = v;
  }
}

}  // namespace ptk

using namespace ptk;

extern "C" int ptk_pose_heatmaps(const int* kp, int N, int P, int H, int W, float sigma, float* out, int C_total, int c0,
                                 void* stream) {
  PTK_REQUIRE(kp && out && N > 0 && P > 0 && P <= 65535 && N <= 65535 && H > 0 && W > 0 && sigma > 0.f, "pose_heatmaps: bad arguments");
  PTK_REQUIRE(c0 >= 0 && c0 + P <= C_total, "pose_heatmaps: channel slice outside the tensor");
  int chunks = (H * W + 255) / 256;
  if (chunks > 64) chunks = 64;
  pose_heatmaps_kernel<<<dim3((unsigned)chunks, (unsigned)P, (unsigned)N), 256, 0, (cudaStream_t)stream>>>(
      kp, P, H, W, 2.0 * (double)sigma * (double)sigma, out, C_total, c0);
  PTK_LAUNCH_CHECK("pose_heatmaps_kernel");
  return 0;
}

extern "C" int ptk_pose_masks(const int* kp, int N, int P, int H, int W, double* masks, void* stream) {
  PTK_REQUIRE(kp && masks && N > 0 && N <= 65535 && H > 0 && W > 0, "pose_masks: bad arguments");
  PTK_REQUIRE(P == 16 || P == 18, "pose_masks: the reference only labels 16 or 18 key-points (pose_transform.py:94-104)");
  PoseNames nm;
  if (P == 18) {
    // LABELS_PAF (utils/pose_utils.py:36-37)
    const PoseNames paf = {{8, 2, 11, 5}, {14, 15, 16, 17, 0}, {8, 11, 9, 12, 2, 5, 3, 6}, {9, 12, 10, 13, 3, 6, 4, 7}};
    nm = paf;
  } else {
    // LABELS (utils/pose_utils.py:27): spells 'Rknee' / 'Lknee', so the joints named 'Rkne' / 'Lkne' never exist and the
    // head candidates are absent: those masks are empty, exactly as in the reference
    const PoseNames shg = {{2, 12, 3, 13}, {-1, -1, -1, -1, -1}, {2, 3, -1, -1, 12, 13, 11, 14}, {-1, -1, 0, 5, 11, 14, 10, 15}};
    nm = shg;
  }
  int chunks = (H * W + 255) / 256;
  if (chunks > 32) chunks = 32;
  pose_masks_kernel<<<dim3((unsigned)chunks, 10u, (unsigned)N), 256, 0, (cudaStream_t)stream>>>(kp, P, H, W, nm, masks);
  PTK_LAUNCH_CHECK("pose_masks_kernel");
  return 0;
}
