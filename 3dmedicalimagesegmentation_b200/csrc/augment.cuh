// GPU-side crop sampling and augmentation: the per-iteration tail of the reference's training transforms
//   RandCropByPosNegLabeld(pos=1, neg=1, num_samples=4, image_key="image", image_threshold=0)   unetr_segmentation_3d.py:341-350
//   RandFlipd x3, RandRotate90d(max_k=3), RandShiftIntensityd(offsets=0.10)                       unetr_segmentation_3d.py:351-375
//   RandSpatialCropSamplesd(random_size=False)                                                    unetr_ranking_pretraining_3d.py:365-369
//   ConvertToMultiChannelBasedOnBratsClassesd                                                     unetr_segmentation_3d.py:65-93
// on a volume that is resident in HBM.  The random DRAWS stay on the host (numpy RandomState, the generator MONAI uses: a handful of
// numbers per crop, and the only way to reproduce its streams); everything that touches voxels runs here:
//   * foreground / background voxel sets are never materialised as index lists (MONAI: map_binary_to_indices builds two int64 arrays of
//     up to 67 M entries per 512x512x256 volume): per-block counts + one prefix scan, computed once per volume, let a kernel find
//     "the k-th foreground voxel in raster order" (= fg_indices[k]) by a binary search over the block prefix and a scan of one block;
//   * crop, the three flips, rot90 and the intensity shift are ONE gather: the composed signed axis permutation is applied to the
//     output coordinate, so every output voxel is read once and written once (coalesced writes), for all samples of a batch in one launch.
// All of it is HBM-bound: bytes = the crops written + the crops read (+ one pass over label / image per volume for the counts).
#pragma once
#include "common.cuh"

namespace b200 {

static constexpr int kAugBlock = 1024;      // voxels per counting block (256 threads x 4)

// fg = any_c(label[c] != 0); bg = !fg and (image ? any_c(image[c] > thr) : true)      (monai.transforms.utils.map_binary_to_indices)
__device__ __forceinline__ void aug_flags(const float* __restrict__ label, int Cl, const float* __restrict__ image, int Ci, float thr, long V, long v,
                                          bool& fg, bool& bg) {
  fg = false;
  for (int c = 0; c < Cl; ++c) fg = fg || (label[(long)c * V + v] != 0.f);
  bool img = image == nullptr;
  for (int c = 0; c < Ci && image; ++c) img = img || (image[(long)c * V + v] > thr);
  bg = !fg && img;
}

// counts[0][b] = foreground voxels of block b, counts[1][b] = background voxels
static __global__ void __launch_bounds__(256) aug_fgbg_counts_kernel(const float* __restrict__ label, int Cl, const float* __restrict__ image, int Ci,
                                                                     float thr, long V, int nblk, int* __restrict__ counts) {
  const long base = (long)blockIdx.x * kAugBlock;
  int nf = 0, nb = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long v = base + 4 * threadIdx.x + i;
    if (v < V) { bool fg, bg; aug_flags(label, Cl, image, Ci, thr, V, v, fg, bg); nf += fg; nb += bg; }
  }
  __shared__ int s[2][8];
  for (int o = 16; o > 0; o >>= 1) { nf += __shfl_xor_sync(0xffffffffu, nf, o); nb += __shfl_xor_sync(0xffffffffu, nb, o); }
  if ((threadIdx.x & 31) == 0) { s[0][threadIdx.x >> 5] = nf; s[1][threadIdx.x >> 5] = nb; }
  __syncthreads();
  if (threadIdx.x < 2) {
    int t = 0;
    for (int w = 0; w < 8; ++w) t += s[threadIdx.x][w];
    counts[(long)threadIdx.x * nblk + blockIdx.x] = t;
  }
}

// in-place exclusive scan of both rows (one block per row), totals[row] = number of set voxels
static __global__ void __launch_bounds__(1024) aug_prefix_kernel(int* __restrict__ counts, int nblk, long long* __restrict__ totals) {
  int* row = counts + (long)blockIdx.x * nblk;
  const int per = (nblk + 1023) / 1024;
  const int lo = threadIdx.x * per, hi = min(nblk, lo + per);
  long long sum = 0;
  for (int i = lo; i < hi; ++i) sum += row[i];
  __shared__ long long part[1024];
  part[threadIdx.x] = sum;
  __syncthreads();
  for (int o = 1; o < 1024; o <<= 1) {          // Hillis-Steele inclusive scan of the per-thread sums
    long long v = threadIdx.x >= o ? part[threadIdx.x - o] : 0;
    __syncthreads();
    part[threadIdx.x] += v;
    __syncthreads();
  }
  long long run = threadIdx.x ? part[threadIdx.x - 1] : 0;
  for (int i = lo; i < hi; ++i) { const int c = row[i]; row[i] = (int)run; run += c; }
  if (threadIdx.x == 1023) totals[blockIdx.x] = part[1023];
}

struct AugPick { int use_fg; long long k; };               // the k-th (0-based, raster order) voxel of the foreground / background set
struct AugPicks { AugPick p[16]; int n; };
struct AugDims { int D, H, W, r0, r1, r2; };

// monai.transforms.utils.correct_crop_centers (0.6.0) for one axis, then SpatialCrop's start = max(center - size // 2, 0)
__device__ __forceinline__ int aug_crop_start(int c, int size, int roi) {
  const int valid_start = roi / 2;
  int valid_end = (int)(unsigned short)(size + 1 - roi / 2.0);       // np.subtract(shape + 1, roi / 2).astype(np.uint16)
  if (valid_start == valid_end) valid_end += 1;
  if (c < valid_start) c = valid_start;
  if (c >= valid_end) c = valid_end - 1;
  return max(c - roi / 2, 0);
}

// one block per pick: locate the counting block by binary search over the exclusive prefix, then the voxel inside it
static __global__ void __launch_bounds__(256) aug_pick_kernel(const float* __restrict__ label, int Cl, const float* __restrict__ image, int Ci, float thr,
                                                              const int* __restrict__ prefix, int nblk, AugPicks picks, AugDims g, int* __restrict__ starts) {
  const AugPick pk = picks.p[blockIdx.x];
  const long V = (long)g.D * g.H * g.W;
  const int* row = prefix + (pk.use_fg ? 0 : (long)nblk);
  int lo = 0, hi = nblk - 1;                     // largest b with prefix[b] <= k
  while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if ((long long)row[mid] <= pk.k) lo = mid; else hi = mid - 1; }
  const int target = (int)(pk.k - row[lo]);      // rank inside the block
  const long base = (long)lo * kAugBlock;
  bool f[4]; int cnt = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long v = base + 4 * threadIdx.x + i;
    f[i] = false;
    if (v < V) { bool fg, bg; aug_flags(label, Cl, image, Ci, thr, V, v, fg, bg); f[i] = pk.use_fg ? fg : bg; }
    cnt += f[i];
  }
  // exclusive scan of cnt over the 256 threads
  __shared__ int wsum[8];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int inc = cnt;
  for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
  if (lane == 31) wsum[w] = inc;
  __syncthreads();
  int before = inc - cnt;
  for (int i = 0; i < w; ++i) before += wsum[i];
  if (target >= before && target < before + cnt) {
    int r = target - before;
    long v = -1;
#pragma unroll
    for (int i = 0; i < 4; ++i) if (f[i]) { if (r == 0 && v < 0) v = base + 4 * threadIdx.x + i; --r; }
    const int z = (int)(v % g.W), y = (int)((v / g.W) % g.H), x = (int)(v / ((long)g.W * g.H));
    starts[3 * blockIdx.x] = aug_crop_start(x, g.D, g.r0);
    starts[3 * blockIdx.x + 1] = aug_crop_start(y, g.H, g.r1);
    starts[3 * blockIdx.x + 2] = aug_crop_start(z, g.W, g.r2);
  }
}

// Composed per-sample transform of the crop: output coordinate o (in the augmented crop) reads crop coordinate
// q[perm[a]] = flip[a] ? ext[perm[a]] - 1 - o[a] : o[a]   for output axis a;  + shift on the image channels.
struct AugMap { int perm[3]; int flip[3]; float shift; };
struct AugMaps { AugMap m[16]; int n; };

// out_image [n][Ci][r0][r1][r2], out_label [n][Cl_out][r0][r1][r2]; brats != 0: the single-channel label map becomes the 4 multi-hot
// channels (background, TC = 2|3, WT = 1|2|3, ET = 3) of ConvertToMultiChannelBasedOnBratsClassesd
static __global__ void __launch_bounds__(256) aug_crop_kernel(const float* __restrict__ image, int Ci, const float* __restrict__ label, int Cl, AugDims g,
                                                              const int* __restrict__ starts, AugMaps maps, int brats,
                                                              float* __restrict__ out_image, float* __restrict__ out_label) {
  const int s = blockIdx.y;
  const AugMap mp = maps.m[s];
  const int s0 = starts[3 * s], s1 = starts[3 * s + 1], s2 = starts[3 * s + 2];
  const long per = (long)g.r0 * g.r1 * g.r2, V = (long)g.D * g.H * g.W;
  const int ext[3] = {g.r0, g.r1, g.r2};
  const int Cl_out = brats ? 4 : Cl;
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < per; e += (long)gridDim.x * blockDim.x) {
    int o[3]; long r = e;
    o[2] = (int)(r % g.r2); r /= g.r2; o[1] = (int)(r % g.r1); o[0] = (int)(r / g.r1);
    int q[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) q[mp.perm[a]] = mp.flip[a] ? ext[mp.perm[a]] - 1 - o[a] : o[a];
    const long src = ((long)(s0 + q[0]) * g.H + (s1 + q[1])) * g.W + (s2 + q[2]);
    if (out_image)
      for (int c = 0; c < Ci; ++c) out_image[((long)s * Ci + c) * per + e] = image[(long)c * V + src] + mp.shift;
    if (out_label) {
      if (brats) {
        const float l = label[src];
        float* ol = out_label + (long)s * 4 * per + e;
        ol[0] = l == 0.f ? 1.f : 0.f;
        ol[per] = (l == 2.f || l == 3.f) ? 1.f : 0.f;
        ol[2 * per] = (l == 1.f || l == 2.f || l == 3.f) ? 1.f : 0.f;
        ol[3 * per] = l == 3.f ? 1.f : 0.f;
      } else {
        for (int c = 0; c < Cl_out; ++c) out_label[((long)s * Cl_out + c) * per + e] = label[(long)c * V + src];
      }
    }
  }
}

}  // namespace b200
