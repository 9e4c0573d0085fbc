// View-weight masks of the texture trainer: which view may paint which pixel.
//
// Reference: /root/reference/src/training/trainer.py:155-249.  create_face_view_map (:155-216) turns the rasteriser's
// face_idx [V,1,H,W] (int64, < 0 = background) into rows (face, view, i, j) of the covered pixels, in (view, pixel)
// order; compare_face_normals_between_views (:218-249) takes z = face_normals[view, 2, face] per row, the per-face
// maximum over all rows (torch_scatter.scatter_max, :227 -- a third-party CUDA dependency that blocks the reference on
// this image) and clears the mask of every pixel whose own z is below that maximum.
//
// z depends on (view, face) only, so the segmented maximum over <= V*H*W rows collapses to: face f is VISIBLE in view
// v (some pixel shows it); maxz[f] = max over the visible views of normals[v,2,f]; mask[v,i,j] = !(z < maxz[f]).
// Three HBM-bound passes, no row table, no float atomics:
//   mark   : read face_idx (8 B/pixel), set visible[v,f] (plain byte stores; every writer stores 1)
//   reduce : per face, max over the visible views (V*F bytes + floats)
//   apply  : read face_idx again, write the boolean mask (8 + 1 B/pixel)
// The row table itself (create_face_view_map) is an order-preserving stream compaction: per-block counts, one
// single-block scan of the counts, fill.
#include "ctx_common.cuh"

namespace ctx {

constexpr int kVmThreads = 256;
constexpr int kVmPerThread = 4;                       // pixels per thread (int64 ids: two 16-byte loads)
constexpr int kVmBlockPixels = kVmThreads * kVmPerThread;

__global__ void __launch_bounds__(kVmThreads)
viewmask_mark_kernel(const long long* __restrict__ face_idx, int64_t n_pix_per_view, int64_t n_total, int F,
                     unsigned char* __restrict__ visible) {
  const int64_t base = ((int64_t)blockIdx.x * kVmThreads + threadIdx.x) * kVmPerThread;
  if (base >= n_total) return;
  long long f[kVmPerThread];
  if (base + kVmPerThread <= n_total) {
    const longlong2 a = __ldg(reinterpret_cast<const longlong2*>(face_idx + base));
    const longlong2 b = __ldg(reinterpret_cast<const longlong2*>(face_idx + base + 2));
    f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y;
  } else {
#pragma unroll
    for (int k = 0; k < kVmPerThread; ++k) f[k] = base + k < n_total ? face_idx[base + k] : -1;
  }
  long long last = -1; int64_t last_v = -1;
#pragma unroll
  for (int k = 0; k < kVmPerThread; ++k) {
    if (f[k] < 0 || f[k] >= F) continue;
    const int64_t v = (base + k) / n_pix_per_view;
    if (f[k] == last && v == last_v) continue;         // neighbouring pixels mostly show the same face
    visible[v * F + f[k]] = 1;
    last = f[k]; last_v = v;
  }
}

__global__ void viewmask_reduce_kernel(const float* __restrict__ normals, const unsigned char* __restrict__ visible,
                                       int V, int F, float* __restrict__ maxz) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= F) return;
  // scatter_max semantics: a plain running maximum over the rows of the face (NaN never wins a '>' comparison);
  // faces no view shows keep -inf (their entry is never read back)
  float m = -INFINITY;
  for (int v = 0; v < V; ++v)
    if (visible[(int64_t)v * F + f]) {
      const float z = normals[((int64_t)v * 3 + 2) * F + f];
      m = (z > m) ? z : m;
    }
  maxz[f] = m;
}

__global__ void __launch_bounds__(kVmThreads)
viewmask_apply_kernel(const long long* __restrict__ face_idx, const float* __restrict__ normals,
                      const float* __restrict__ maxz, int64_t n_pix_per_view, int64_t n_total, int F,
                      unsigned char* __restrict__ mask) {
  const int64_t base = ((int64_t)blockIdx.x * kVmThreads + threadIdx.x) * kVmPerThread;
  if (base >= n_total) return;
  unsigned char out[kVmPerThread];
#pragma unroll
  for (int k = 0; k < kVmPerThread; ++k) {
    out[k] = 1;                                        // background pixels keep the default True (:220)
    if (base + k < n_total) {
      const long long f = face_idx[base + k];
      if (f >= 0 && f < F) {
        const int64_t v = (base + k) / n_pix_per_view;
        const float z = __ldg(normals + (v * 3 + 2) * F + f);
        out[k] = (z < __ldg(maxz + f)) ? 0 : 1;        // ~(z < max), :233-247
      }
    }
  }
  if (base + kVmPerThread <= n_total && (reinterpret_cast<uintptr_t>(mask + base) & 3) == 0) {
    *reinterpret_cast<uchar4*>(mask + base) = make_uchar4(out[0], out[1], out[2], out[3]);
  } else {
#pragma unroll
    for (int k = 0; k < kVmPerThread; ++k) if (base + k < n_total) mask[base + k] = out[k];
  }
}

// ---- create_face_view_map: order-preserving compaction of the covered pixels into rows (face, view, i, j) ----
__global__ void __launch_bounds__(kVmThreads)
faceview_count_kernel(const long long* __restrict__ face_idx, int64_t n_total, long long* __restrict__ block_counts) {
  const int64_t base = ((int64_t)blockIdx.x * kVmThreads + threadIdx.x) * kVmPerThread;
  int c = 0;
#pragma unroll
  for (int k = 0; k < kVmPerThread; ++k) c += (base + k < n_total && face_idx[base + k] >= 0) ? 1 : 0;
  c = (int)warp_sum((float)c);                        // <= 128 per warp: exact in fp32
  __shared__ int red[kVmThreads / 32];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int w = 0; w < kVmThreads / 32; ++w) t += red[w];
    block_counts[blockIdx.x] = t;
  }
}

// exclusive scan of block_counts[0..n) in place by ONE block; total -> block_counts[n]
__global__ void __launch_bounds__(1024) faceview_scan_kernel(long long* __restrict__ counts, int64_t n) {
  __shared__ long long warp_tot[32];
  __shared__ long long carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int64_t c0 = 0; c0 < n; c0 += 1024) {
    const int64_t i = c0 + threadIdx.x;
    const long long v = i < n ? counts[i] : 0;
    long long incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const long long t = __shfl_up_sync(CTX_FULL_MASK, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) warp_tot[w] = incl;
    __syncthreads();
    long long before = carry_s;
    for (int k = 0; k < w; ++k) before += warp_tot[k];
    if (i < n) counts[i] = before + incl - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry_s = before + incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) counts[n] = carry_s;
}

__global__ void __launch_bounds__(kVmThreads)
faceview_fill_kernel(const long long* __restrict__ face_idx, int64_t n_pix_per_view, int64_t n_total, int W,
                     const long long* __restrict__ block_offsets, long long* __restrict__ rows) {
  const int64_t base = ((int64_t)blockIdx.x * kVmThreads + threadIdx.x) * kVmPerThread;
  long long f[kVmPerThread];
  int c = 0;
#pragma unroll
  for (int k = 0; k < kVmPerThread; ++k) {
    f[k] = base + k < n_total ? face_idx[base + k] : -1;
    c += f[k] >= 0 ? 1 : 0;
  }
  // exclusive prefix of c over the block (warp scan + per-warp totals)
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int incl = c;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(CTX_FULL_MASK, incl, o);
    if (lane >= o) incl += t;
  }
  __shared__ int wt[kVmThreads / 32];
  if (lane == 31) wt[w] = incl;
  __syncthreads();
  int before = 0;
  for (int k = 0; k < w; ++k) before += wt[k];
  long long pos = block_offsets[blockIdx.x] + before + incl - c;
#pragma unroll
  for (int k = 0; k < kVmPerThread; ++k) {
    if (f[k] >= 0) {
      const int64_t g = base + k, v = g / n_pix_per_view, p = g - v * n_pix_per_view;
      longlong2* r = reinterpret_cast<longlong2*>(rows + pos * 4);
      r[0] = make_longlong2(f[k], v);
      r[1] = make_longlong2(p / W, p % W);
      ++pos;
    }
  }
}

}  // namespace ctx

// face_normals [V,3,F] fp32, face_idx [V,H,W] int64 -> mask [V,H,W] bytes (1 = True).  visible [V*F] bytes and
// maxz [F] floats are caller-owned scratch (visible is zeroed here).
extern "C" int ctx_view_weight_masks(const float* face_normals, const int64_t* face_idx, int V, int F, int H, int W,
                                     unsigned char* visible, float* maxz, unsigned char* mask, void* stream) {
  if (V < 0 || F < 0 || H < 0 || W < 0) return CTX_ERR_BAD_ARG;
  const int64_t npv = (int64_t)H * W, total = npv * V;
  if (total == 0) return 0;
  if (!face_idx || !mask || (F > 0 && (!face_normals || !visible || !maxz))) return CTX_ERR_BAD_ARG;
  if (reinterpret_cast<uintptr_t>(face_idx) % 16 != 0) return CTX_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const int blocks = (int)ctx::ceil_div(total, ctx::kVmBlockPixels);
  if (F > 0) {
    cudaError_t e = cudaMemsetAsync(visible, 0, (size_t)V * F, st);
    if (e != cudaSuccess) return (int)e;
    ctx::viewmask_mark_kernel<<<blocks, ctx::kVmThreads, 0, st>>>((const long long*)face_idx, npv, total, F, visible);
    ctx::viewmask_reduce_kernel<<<(F + 255) / 256, 256, 0, st>>>(face_normals, visible, V, F, maxz);
  }
  ctx::viewmask_apply_kernel<<<blocks, ctx::kVmThreads, 0, st>>>((const long long*)face_idx, face_normals, maxz, npv,
                                                                total, F, mask);
  CTX_RETURN_LAST();
}

// Rows (face, view, i, j) of the covered pixels in (view, pixel) order.  block_counts: int64 scratch of
// ctx_face_view_map_blocks(V*H*W) + 1 entries; after pass 0 its last entry holds the number of rows N (read it back,
// allocate rows [N,4], run pass 1).
extern "C" int64_t ctx_face_view_map_blocks(int64_t n_pixels) { return ctx::ceil_div(n_pixels, ctx::kVmBlockPixels); }

extern "C" int ctx_face_view_map(const int64_t* face_idx, int V, int H, int W, int64_t* block_counts, int64_t* rows,
                                 int pass, void* stream) {
  if (V < 0 || H < 0 || W < 0 || !block_counts) return CTX_ERR_BAD_ARG;
  const int64_t npv = (int64_t)H * W, total = npv * V;
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t blocks = ctx::ceil_div(total, ctx::kVmBlockPixels);
  if (pass == 0) {
    if (total == 0) return (int)cudaMemsetAsync(block_counts, 0, sizeof(int64_t), st);
    if (!face_idx) return CTX_ERR_BAD_ARG;
    ctx::faceview_count_kernel<<<(int)blocks, ctx::kVmThreads, 0, st>>>((const long long*)face_idx, total,
                                                                        (long long*)block_counts);
    ctx::faceview_scan_kernel<<<1, 1024, 0, st>>>((long long*)block_counts, blocks);
  } else {
    if (total == 0) return 0;
    if (!face_idx || !rows) return CTX_ERR_BAD_ARG;
    ctx::faceview_fill_kernel<<<(int)blocks, ctx::kVmThreads, 0, st>>>((const long long*)face_idx, npv, total, W,
                                                                       (const long long*)block_counts, (long long*)rows);
  }
  CTX_RETURN_LAST();
}
