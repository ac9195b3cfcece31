// K1 (per-frame ECC preparation) and K6 (W1 sharpness weight map) kernels.
//
// Reference semantics reproduced here:
//   scaleImage (cv::pyrDown of the ECC image)          core/proc/image_registration/c_frame_registration.cc:230-250
//   c_ecch::set_current_image / set_reference_image    core/proc/image_registration/ecc2.cc:984-1120
//        (Gaussian sepFilter2D BORDER_REPLICATE, pyramid by cv::pyrDown(img, nextSize))
//   extract_channel(gray) = cv::cvtColor(BGR2GRAY)     core/proc/extract_channel.cc:607-697
//   compute_local_variance_map                         core/proc/sharpness_measure/c_local_variance_sharpness_measure.cc:193-247
// All kernels are batched over the frames of a batch through blockIdx.z.
#include "ssk_prep.cuh"

namespace ssk {

namespace {

template <int DEPTH>
__device__ __forceinline__ float load_gray(const Img &im, int y, int x) {
  if (im.cn == 1) return load_px<DEPTH>(im, y, x, 0);
  // cv::cvtColor(COLOR_BGR2GRAY) on float data: 0.114 B + 0.587 G + 0.299 R
  const float b = load_px<DEPTH>(im, y, x, 0), g = load_px<DEPTH>(im, y, x, 1), r = load_px<DEPTH>(im, y, x, 2);
  return fmaf(r, 0.299f, fmaf(g, 0.587f, b * 0.114f));
}

constexpr int PD_OW = 32, PD_OH = 8;
constexpr int PD_IW = 2 * PD_OW + 3, PD_IH = 2 * PD_OH + 3;

template <int DEPTH>
__global__ void __launch_bounds__(256) k_pyrdown(const PyrDownArgs a) {
  __shared__ float s_in[PD_IH][PD_IW + 1];
  __shared__ float s_h[PD_IH][PD_OW + 1];
  const int b = blockIdx.z;
  Img src = a.src;
  if (a.src_ptrs) src.data = a.src_ptrs[b];
  float *dst = a.dst_ptrs ? a.dst_ptrs[b] : a.dst;
  const int ox0 = blockIdx.x * PD_OW, oy0 = blockIdx.y * PD_OH;
  const int ix0 = 2 * ox0 - 2, iy0 = 2 * oy0 - 2;
  const int width0 = min((src.cols - 3) / 2 + 1, a.dst_cols);   // PyrDownInvoker: columns free of border handling
  const int nsimd_h = 4 * ((width0 - 1) / 4);
  for (int k = threadIdx.x; k < PD_IH * PD_IW; k += blockDim.x) {
    const int r = k / PD_IW, c = k - r * PD_IW;
    const int yy = border_idx(iy0 + r, src.rows, SSK_BORDER_REFLECT101);
    const int xx = border_idx(ix0 + c, src.cols, SSK_BORDER_REFLECT101);
    s_in[r][c] = load_gray<DEPTH>(src, yy, xx);
  }
  __syncthreads();
  // horizontal: row[x] = s[2x]*6 + (s[2x-1] + s[2x+1])*4 + s[2x-2] + s[2x+2]
  for (int k = threadIdx.x; k < PD_IH * PD_OW; k += blockDim.x) {
    const int r = k / PD_OW, x = k - r * PD_OW;
    const float *p = &s_in[r][2 * x];   // p[0] = s[2x-2]
    // cv::pyrDown, bit-exact against cv2 4.13 (oracle/cvmodel.py::pyrdown_f32): columns covered by the 4-lane
    // PyrDownVecH loop use s0*6 + ((s-1 + s1)*4 + (s-2 + s2)); the border column and the scalar tail use
    // ((s0*6 + (s-1 + s1)*4) + s-2) + s2
    const int gx = ox0 + x;
    const float a1 = __fmul_rn(__fadd_rn(p[1], p[3]), 4.f), c6 = __fmul_rn(p[2], 6.f);
    s_h[r][x] = (gx >= 1 && gx < 1 + nsimd_h) ? __fadd_rn(c6, __fadd_rn(a1, __fadd_rn(p[0], p[4])))
                                              : __fadd_rn(__fadd_rn(__fadd_rn(c6, a1), p[0]), p[4]);
  }
  __syncthreads();
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int ox = ox0 + tx, oy = oy0 + ty;
  if (ox < a.dst_cols && oy < a.dst_rows) {
    const int r = 2 * ty;
    // PyrDownVecV (4 lanes): ((r1 + r3) + r2)*4 + ((r0 + r4) + (r2 + r2)); scalar tail: ((r2*6 + (r1 + r3)*4) + r0) + r4
    const float c2 = s_h[r + 2][tx], a13 = __fadd_rn(s_h[r + 1][tx], s_h[r + 3][tx]);
    float v;
    if (ox < (a.dst_cols & ~3))
      v = __fadd_rn(__fmul_rn(__fadd_rn(a13, c2), 4.f), __fadd_rn(__fadd_rn(s_h[r][tx], s_h[r + 4][tx]), __fadd_rn(c2, c2)));
    else
      v = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(c2, 6.f), __fmul_rn(a13, 4.f)), s_h[r][tx]), s_h[r + 4][tx]);
    v = __fmul_rn(v, 1.0f / 256.0f);
    if (a.post_scale != 1.f) v = __fmul_rn(v, a.post_scale);
    dst[(int64_t)oy * a.dst_cols + ox] = v;
  }
}

template <int DEPTH>
__global__ void __launch_bounds__(256) k_to_gray(const Img im, const void *const *src_ptrs, float *dst, float *const *dst_ptrs) {
  const int b = blockIdx.z;
  Img src = im;
  if (src_ptrs) src.data = src_ptrs[b];
  float *d = dst_ptrs ? dst_ptrs[b] : dst;
  const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x < src.cols && y < src.rows) d[(int64_t)y * src.cols + x] = load_gray<DEPTH>(src, y, x);
}

constexpr int SF_W = 32, SF_H = 8, SF_R = (kMaxTaps - 1) / 2;

__global__ void __launch_bounds__(256) k_sepfilter(const SepFilterArgs a) {
  __shared__ float s_in[SF_H + 2 * SF_R][SF_W + 2 * SF_R + 1];
  __shared__ float s_h[SF_H + 2 * SF_R][SF_W + 1];
  const int b = blockIdx.z;
  const float *src = a.src_ptrs ? a.src_ptrs[b] : a.src;
  float *dst = a.dst_ptrs ? a.dst_ptrs[b] : a.dst;
  const int rx = a.kxn >> 1, ry = a.kyn >> 1;
  const int x0 = blockIdx.x * SF_W, y0 = blockIdx.y * SF_H;
  const int iw = SF_W + 2 * rx, ih = SF_H + 2 * ry;
  for (int k = threadIdx.x; k < ih * iw; k += blockDim.x) {
    const int r = k / iw, c = k - r * iw;
    const int yy = min(max(y0 - ry + r, 0), a.rows - 1), xx = min(max(x0 - rx + c, 0), a.cols - 1);
    s_in[r][c] = __ldg(src + (int64_t)yy * a.cols + xx);
  }
  __syncthreads();
  // row / column arithmetic follows OpenCV's filter engine (found bit-exact against cv2 4.13 for the kernels of this
  // path: 5-tap derivative, 3-tap smoothing, 7-tap Gaussian)
  const bool xsym = a.kx[0] == a.kx[a.kxn - 1];
  for (int k = threadIdx.x; k < ih * SF_W; k += blockDim.x) {
    const int r = k / SF_W, x = k - r * SF_W;
    const float *p = &s_in[r][x + rx];
    float acc;
    if (a.kxn <= 5) {
      // SymmRowSmallFilter: k0*x0 then fma over the (anti)symmetric pairs
      if (xsym) {
        acc = __fmul_rn(a.kx[rx], p[0]);
        for (int i = 1; i <= rx; ++i) acc = __fmaf_rn(__fadd_rn(p[i], p[-i]), a.kx[rx + i], acc);
      } else {
        acc = rx >= 1 ? __fmul_rn(__fsub_rn(p[1], p[-1]), a.kx[rx + 1]) : 0.f;
        for (int i = 2; i <= rx; ++i) acc = __fmaf_rn(__fsub_rn(p[i], p[-i]), a.kx[rx + i], acc);
      }
    } else {
      // RowFilter (RowVec_32f): taps in order, fma chain
      acc = __fmul_rn(p[-rx], a.kx[0]);
      for (int i = 1; i < a.kxn; ++i) acc = __fmaf_rn(p[i - rx], a.kx[i], acc);
    }
    s_h[r][x] = acc;
  }
  __syncthreads();
  const bool ysym = a.ky[0] == a.ky[a.kyn - 1];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int x = x0 + tx, y = y0 + ty;
  if (x < a.cols && y < a.rows) {
    const int r = ty + ry;
    // SymmColumnFilter (SymmColumnVec_32f): centre tap then fma over the (anti)symmetric pairs
    float acc;
    if (ysym) {
      acc = __fmul_rn(a.ky[ry], s_h[r][tx]);
      for (int i = 1; i <= ry; ++i) acc = __fmaf_rn(__fadd_rn(s_h[r + i][tx], s_h[r - i][tx]), a.ky[ry + i], acc);
    } else {
      acc = ry >= 1 ? __fmul_rn(__fsub_rn(s_h[r + 1][tx], s_h[r - 1][tx]), a.ky[ry + 1]) : 0.f;
      for (int i = 2; i <= ry; ++i) acc = __fmaf_rn(__fsub_rn(s_h[r + i][tx], s_h[r - i][tx]), a.ky[ry + i], acc);
    }
    dst[(int64_t)y * a.cols + x] = acc;
  }
}

// ---- W1 ------------------------------------------------------------------------------------------
constexpr int W1_W = 32, W1_H = 8, W1_RMAX = 4;

__global__ void __launch_bounds__(256) k_w1_grad(const W1Args a, int nblocks) {
  __shared__ float s_in[W1_H + 2 * W1_RMAX][W1_W + 2 * W1_RMAX + 1];
  __shared__ double s_red[2][8];
  const int b = blockIdx.z;
  const float *M = a.M_ptrs ? a.M_ptrs[b] : a.M;
  float *gmap = a.gmap_ptrs ? a.gmap_ptrs[b] : a.gmap;
  const int r = a.kradius;
  const int x0 = blockIdx.x * W1_W, y0 = blockIdx.y * W1_H;
  const int iw = W1_W + 2 * r, ih = W1_H + 2 * r;
  for (int k = threadIdx.x; k < ih * iw; k += blockDim.x) {
    const int rr = k / iw, c = k - rr * iw;
    const int yy = min(max(y0 - r + rr, 0), a.rows - 1), xx = min(max(x0 - r + c, 0), a.cols - 1);
    s_in[rr][c] = __ldg(M + (int64_t)yy * a.cols + xx);
  }
  __syncthreads();
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int x = x0 + tx, y = y0 + ty;
  double sg = 0.0, sg4 = 0.0;
  if (x < a.cols && y < a.rows) {
    float mx = -3.4e38f, mn = 3.4e38f;
    for (int dy = 0; dy <= 2 * r; ++dy)
      for (int dx = 0; dx <= 2 * r; ++dx) {
        const float v = s_in[ty + dy][tx + dx];
        mx = fmaxf(mx, v);
        mn = fminf(mn, v);
      }
    const float g = __fsub_rn(mx, mn);
    const float ms = (float)(a.depth_scale * a.depth_scale * a.depth_scale);
    gmap[(int64_t)y * a.cols + x] = __fmul_rn(__fmul_rn(__fmul_rn(g, g), g), ms);
    sg = (double)fabsf(g);
    sg4 = (double)__fmul_rn(__fmul_rn(__fmul_rn(g, g), g), g);
  }
  sg = warp_sum(sg);
  sg4 = warp_sum(sg4);
  if (tx == 0) { s_red[0][ty] = sg; s_red[1][ty] = sg4; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t0 = 0, t1 = 0;
    for (int i = 0; i < 8; ++i) { t0 += s_red[0][i]; t1 += s_red[1][i]; }
    const int blk = blockIdx.y * gridDim.x + blockIdx.x;
    a.partials[((int64_t)b * 2 + 0) * nblocks + blk] = t0;
    a.partials[((int64_t)b * 2 + 1) * nblocks + blk] = t1;
  }
}

__global__ void __launch_bounds__(256) k_w1_final(const W1Args a, int nblocks) {
  __shared__ double s0[256], s1[256];
  const int b = blockIdx.x;
  double t0 = 0, t1 = 0;
  for (int i = threadIdx.x; i < nblocks; i += 256) {
    t0 += a.partials[((int64_t)b * 2 + 0) * nblocks + i];
    t1 += a.partials[((int64_t)b * 2 + 1) * nblocks + i];
  }
  s0[threadIdx.x] = t0; s1[threadIdx.x] = t1;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) { s0[threadIdx.x] += s0[threadIdx.x + o]; s1[threadIdx.x] += s1[threadIdx.x + o]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const double W = s0[0];
    const double ds3 = a.depth_scale * a.depth_scale * a.depth_scale;
    // the reference accumulates g^4 in a float (atomic CAS loop); narrow the total once
    const double Q = W > 0 ? (ds3 / W) * (double)(float)s1[0] : 0.0;
    a.stats[b * 4 + 0] = W;
    a.stats[b * 4 + 1] = s1[0];
    a.stats[b * 4 + 2] = Q;
    a.stats[b * 4 + 3] = (double)(float)(0.05 * Q);
  }
}

// cv::resize(map + 0.05 Q, full size, INTER_LINEAR) (c_local_variance_sharpness_measure.cc:176-184, 239-243)
__global__ void __launch_bounds__(256) k_w1_upsample(const W1Args a) {
  const int b = blockIdx.z;
  const float *g = a.gmap_ptrs ? a.gmap_ptrs[b] : a.gmap;
  float *out = a.out_ptrs ? a.out_ptrs[b] : a.out;
  const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= a.full_cols || y >= a.full_rows) return;
  const float add = (float)a.stats[b * 4 + 3];
  if (a.full_cols == a.cols && a.full_rows == a.rows) {
    out[(int64_t)y * a.full_cols + x] = __fadd_rn(g[(int64_t)y * a.cols + x], add);
    return;
  }
  const double scx = (double)a.cols / a.full_cols, scy = (double)a.rows / a.full_rows;
  float fx = (float)((x + 0.5) * scx - 0.5), fy = (float)((y + 0.5) * scy - 0.5);
  int sx = (int)floorf(fx), sy = (int)floorf(fy);
  fx -= sx; fy -= sy;
  if (sx < 0) { fx = 0; sx = 0; }
  if (sx >= a.cols - 1) { fx = 0; sx = a.cols - 1; }
  if (sy < 0) { fy = 0; sy = 0; }
  if (sy >= a.rows - 1) { fy = 0; sy = a.rows - 1; }
  const int sx1 = min(sx + 1, a.cols - 1), sy1 = min(sy + 1, a.rows - 1);
  const float a0 = 1.f - fx, a1 = fx, b0 = 1.f - fy, b1 = fy;
  const float v00 = __fadd_rn(g[(int64_t)sy * a.cols + sx], add), v01 = __fadd_rn(g[(int64_t)sy * a.cols + sx1], add);
  const float v10 = __fadd_rn(g[(int64_t)sy1 * a.cols + sx], add), v11 = __fadd_rn(g[(int64_t)sy1 * a.cols + sx1], add);
  const float r0 = __fadd_rn(__fmul_rn(v00, a0), __fmul_rn(v01, a1));
  const float r1 = __fadd_rn(__fmul_rn(v10, a0), __fmul_rn(v11, a1));
  out[(int64_t)y * a.full_cols + x] = __fadd_rn(__fmul_rn(r0, b0), __fmul_rn(r1, b1));
}

__global__ void __launch_bounds__(256) k_erode5_u8(const uint8_t *src, int64_t sstep, uint8_t *dst, int64_t dstep, int rows,
                                                   int cols, int border_replicate) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= cols || y >= rows) return;
  uint8_t m = 255;
  for (int dy = -2; dy <= 2; ++dy)
    for (int dx = -2; dx <= 2; ++dx) {
      int yy = y + dy, xx = x + dx;
      if (border_replicate) { yy = min(max(yy, 0), rows - 1); xx = min(max(xx, 0), cols - 1); }
      else if ((unsigned)yy >= (unsigned)rows || (unsigned)xx >= (unsigned)cols) continue;
      m = min(m, src[(int64_t)yy * sstep + xx]);
    }
  dst[(int64_t)y * dstep + x] = m;
}

}  // namespace

int launch_pyrdown(const PyrDownArgs &a, cudaStream_t s) {
  SSK_REQUIRE(abs(a.dst_cols * 2 - a.src.cols) <= 2 && abs(a.dst_rows * 2 - a.src.rows) <= 2, "pyrDown: bad dstsize");
  dim3 grid(div_up(a.dst_cols, PD_OW), div_up(a.dst_rows, PD_OH), a.batch);
  if (a.src.depth == SSK_32F) k_pyrdown<SSK_32F><<<grid, 256, 0, s>>>(a);
  else if (a.src.depth == SSK_16U) k_pyrdown<SSK_16U><<<grid, 256, 0, s>>>(a);
  else if (a.src.depth == SSK_8U) k_pyrdown<SSK_8U><<<grid, 256, 0, s>>>(a);
  else { set_error("pyrDown: unsupported depth"); return SSK_ERR_INVALID; }
  SSK_LAUNCH_CHECK();
  return SSK_OK;
}

int launch_to_gray(const Img &src, const void *const *src_ptrs, float *dst, float *const *dst_ptrs, int batch, cudaStream_t s) {
  dim3 grid(div_up(src.cols, 32), div_up(src.rows, 8), batch);
  if (src.depth == SSK_32F) k_to_gray<SSK_32F><<<grid, 256, 0, s>>>(src, src_ptrs, dst, dst_ptrs);
  else if (src.depth == SSK_16U) k_to_gray<SSK_16U><<<grid, 256, 0, s>>>(src, src_ptrs, dst, dst_ptrs);
  else if (src.depth == SSK_8U) k_to_gray<SSK_8U><<<grid, 256, 0, s>>>(src, src_ptrs, dst, dst_ptrs);
  else { set_error("to_gray: unsupported depth"); return SSK_ERR_INVALID; }
  SSK_LAUNCH_CHECK();
  return SSK_OK;
}

int launch_sepfilter(const SepFilterArgs &a, cudaStream_t s) {
  SSK_REQUIRE((a.kxn & 1) && (a.kyn & 1) && a.kxn <= kMaxTaps && a.kyn <= kMaxTaps, "sepFilter2D: odd kernels up to 31 taps");
  dim3 grid(div_up(a.cols, SF_W), div_up(a.rows, SF_H), a.batch);
  k_sepfilter<<<grid, 256, 0, s>>>(a);
  SSK_LAUNCH_CHECK();
  return SSK_OK;
}

int w1_num_blocks(int rows, int cols) { return div_up(cols, W1_W) * div_up(rows, W1_H); }

int launch_w1(const W1Args &a, cudaStream_t s) {
  SSK_REQUIRE(a.kradius >= 1 && a.kradius <= W1_RMAX, "local variance map: kradius 1..4");
  const int nblocks = w1_num_blocks(a.rows, a.cols);
  dim3 grid(div_up(a.cols, W1_W), div_up(a.rows, W1_H), a.batch);
  k_w1_grad<<<grid, 256, 0, s>>>(a, nblocks);
  SSK_LAUNCH_CHECK();
  k_w1_final<<<a.batch, 256, 0, s>>>(a, nblocks);
  SSK_LAUNCH_CHECK();
  if (a.out || a.out_ptrs) {
    dim3 g2(div_up(a.full_cols, 32), div_up(a.full_rows, 8), a.batch);
    k_w1_upsample<<<g2, 256, 0, s>>>(a);
    SSK_LAUNCH_CHECK();
  }
  return SSK_OK;
}

int launch_erode5_u8(const uint8_t *src, int64_t sstep, uint8_t *dst, int64_t dstep, int rows, int cols,
                     int border_replicate, cudaStream_t s) {
  dim3 grid(div_up(cols, 32), div_up(rows, 8));
  k_erode5_u8<<<grid, 256, 0, s>>>(src, sstep, dst, dstep, rows, cols, border_replicate);
  SSK_LAUNCH_CHECK();
  return SSK_OK;
}

}  // namespace ssk
