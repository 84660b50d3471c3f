// MaxPool2d forward (with window-argmax byte index) and backward on NHWC activations (HBM-bound).
//
// Reference: _max_pool2d / _max_pool2d_backward, /root/reference/src/tortto/autograd/grad_nn.py:784-826.
//  * forward: padding behaves as -inf; the selected element is the FIRST maximum in row-major (r, s) window order
//    (nanargmax over kw, then over kh, :789-805).
//  * backward: `expanded[pos] = dy` (:820) is an assignment through a fancy index, so when two overlapping windows
//    chose the same input element the window that comes LAST in (n, p, q) raster order wins - gradients are not
//    summed.  Implemented here as a gather (each input element searches the windows that cover it, last one
//    first), so it is deterministic and needs neither atomics nor a zero-fill pass.
#include <math.h>

#include "common.cuh"

namespace ttb {

__global__ void __launch_bounds__(256)
maxpool_fwd_kernel(ttb_pool_desc d, const float* __restrict__ x, float* __restrict__ y, uint8_t* __restrict__ idx,
                   int64_t total, int cvec) {
  pdl_entry();
  // one thread per (n, p, q, channel group of `cvec` channels); cvec is 4 (float4) or 1
  const int cg = d.c / cvec;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
    int g = (int)(t % cg);
    int64_t pix = t / cg;
    int q = (int)(pix % d.q);
    int64_t t2 = pix / d.q;
    int p = (int)(t2 % d.p);
    int n = (int)(t2 / d.p);
    float best[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
    int bi[4] = {-1, -1, -1, -1};
    for (int r = 0; r < d.kh; ++r) {
      int h = p * d.stride_h - d.pad_h + r * d.dil_h;
      if (h < 0 || h >= d.h) continue;
      for (int s = 0; s < d.kw; ++s) {
        int w = q * d.stride_w - d.pad_w + s * d.dil_w;
        if (w < 0 || w >= d.w) continue;
        const float* px = x + (((int64_t)n * d.h + h) * d.w + w) * d.c + g * cvec;
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if (cvec == 4) {
          float4 f = ld_f4(px);
          v[0] = f.x; v[1] = f.y; v[2] = f.z; v[3] = f.w;
        } else {
          v[0] = px[0];
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          // strict '>' keeps the first maximum; a NaN ranks as -inf (nanargmax ignores NaNs)
          float key = (v[j] == v[j]) ? v[j] : -INFINITY;
          if (j < cvec && (bi[j] < 0 || key > best[j])) {
            best[j] = key;
            bi[j] = r * d.kw + s;
          }
        }
      }
    }
    int64_t o = pix * d.c + g * cvec;
    for (int j = 0; j < cvec; ++j) {
      y[o + j] = best[j];
      idx[o + j] = (uint8_t)(bi[j] < 0 ? 0 : bi[j]);
    }
  }
}

template <bool ACCUM>
__global__ void __launch_bounds__(256)
maxpool_bwd_kernel(ttb_pool_desc d, const float* __restrict__ dy, const uint8_t* __restrict__ idx,
                   float* __restrict__ dx, int64_t total) {
  pdl_entry();
  // one thread per input element (n, h, w, c); c fastest -> coalesced
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
    int c = (int)(t % d.c);
    int64_t pix = t / d.c;
    int w = (int)(pix % d.w);
    int64_t t2 = pix / d.w;
    int h = (int)(t2 % d.h);
    int n = (int)(t2 / d.h);
    float acc = 0.f;
    bool done = false;
    // r ascending <=> p descending, s ascending <=> q descending: the first hit is the last writer in raster order
    for (int r = 0; r < d.kh && !done; ++r) {
      int hp = h + d.pad_h - r * d.dil_h;
      if (hp < 0) break;
      if (hp % d.stride_h) continue;
      int p = hp / d.stride_h;
      if (p >= d.p) continue;
      for (int s = 0; s < d.kw; ++s) {
        int wq = w + d.pad_w - s * d.dil_w;
        if (wq < 0) break;
        if (wq % d.stride_w) continue;
        int q = wq / d.stride_w;
        if (q >= d.q) continue;
        int64_t o = (((int64_t)n * d.p + p) * d.q + q) * d.c + c;
        if (idx[o] == (uint8_t)(r * d.kw + s)) {
          if (ACCUM) {
            acc += dy[o];
          } else {
            acc = dy[o];
            done = true;
            break;
          }
        }
      }
    }
    dx[t] = acc;
  }
}

// Specialised forward for the two geometries the BASELINE networks use (3x3 / stride 2: ResNet stem; 2x2 / stride 2: UNet),
// dilation 1, 4 channels per thread: compile-time window and stride (no integer division in the window loop), one
// float4 load per tap, one float4 store of y and one 32-bit store of the 4 index bytes.
template <int KH, int KW, int SH, int SW>
__global__ void __launch_bounds__(256)
maxpool_fwd4_kernel(ttb_pool_desc d, const float* __restrict__ x, float* __restrict__ y, uint8_t* __restrict__ idx,
                    int64_t total4) {
  pdl_entry();
  const int cq = d.c / 4;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total4; t += stride) {
    const int c0 = (int)(t % cq) * 4;
    int64_t pix = t / cq;
    const int q = (int)(pix % d.q);
    int64_t t2 = pix / d.q;
    const int p = (int)(t2 % d.p);
    const int n = (int)(t2 / d.p);
    float best[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
    int bi[4] = {-1, -1, -1, -1};
#pragma unroll
    for (int r = 0; r < KH; ++r) {
      const int h = p * SH - d.pad_h + r;
      if (h < 0 || h >= d.h) continue;
#pragma unroll
      for (int s_ = 0; s_ < KW; ++s_) {
        const int w = q * SW - d.pad_w + s_;
        if (w < 0 || w >= d.w) continue;
        const float4 f = ld_f4_stream(x + (((int64_t)n * d.h + h) * d.w + w) * d.c + c0);
        const float v[4] = {f.x, f.y, f.z, f.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float key = (v[j] == v[j]) ? v[j] : -INFINITY;  // a NaN ranks as -inf (nanargmax ignores NaNs)
          if (bi[j] < 0 || key > best[j]) {                      // strict '>' keeps the first maximum
            best[j] = key;
            bi[j] = r * KW + s_;
          }
        }
      }
    }
    const int64_t o = pix * d.c + c0;
    st_f4(y + o, make_float4(best[0], best[1], best[2], best[3]));
    uint32_t packed = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) packed |= (uint32_t)(bi[j] < 0 ? 0 : bi[j]) << (8 * j);
    *reinterpret_cast<uint32_t*>(idx + o) = packed;
  }
}

// the same gather, 4 channels per thread: one 32-bit load brings the 4 index bytes of a window, dy is fetched as one float4
// only when at least one of the 4 channels selected this input element, dx is one float4 store
// KH/KW/SH/SW > 0: compile-time geometry (dilation 1) - the generic form spends its time in integer divisions by the
// run-time strides (measured 0.11 of the HBM roofline on the ResNet-50 stem pool); 0: run-time geometry.
template <bool ACCUM, int KH, int KW, int SH, int SW>
__global__ void __launch_bounds__(256)
maxpool_bwd4_kernel(ttb_pool_desc d, const float* __restrict__ dy, const uint8_t* __restrict__ idx, float* __restrict__ dx,
                    int64_t total4) {
  pdl_entry();
  if (KH > 0) {
    d.kh = KH; d.kw = KW; d.stride_h = SH; d.stride_w = SW; d.dil_h = 1; d.dil_w = 1;  // constants from here on
  }
  const int cq = d.c / 4;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total4; t += stride) {
    const int c0 = (int)(t % cq) * 4;
    int64_t pix = t / cq;
    const int w = (int)(pix % d.w);
    int64_t t2 = pix / d.w;
    const int h = (int)(t2 % d.h);
    const int n = (int)(t2 / d.h);
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    unsigned done = 0;  // bit j: channel j already has its (last-writer) value
    if constexpr (KH > 0) {
      // Compile-time geometry: an input element is covered by at most NR x NS windows (2 x 2 for 3x3 / stride 2).  Their
      // index words and gradients are ALL loaded first (independent loads; dy is re-read from L1 / L2, a quarter of dx's
      // size) and the last-writer selection runs on registers afterwards: the walk that loaded an index, tested it and
      // only then loaded the gradient was a chain of up to 8 dependent loads per thread (0.95 ms on the ResNet-50 stem pool
      // at batch 256, 0.18 of the HBM roofline).
      constexpr int NR = (KH + SH - 1) / SH, NS = (KW + SW - 1) / SW;
      const int r0 = (h + d.pad_h) % SH, s0 = (w + d.pad_w) % SW;
      uint32_t ib[NR * NS], code[NR * NS];
      float4 g[NR * NS];
#pragma unroll
      for (int i = 0; i < NR; ++i) {
        const int r = r0 + i * SH, hp = h + d.pad_h - r;
        const int pp = hp / SH;
        const bool okr = r < KH && hp >= 0 && pp < d.p;
#pragma unroll
        for (int j = 0; j < NS; ++j) {
          const int sx = s0 + j * SW, wq = w + d.pad_w - sx;
          const int qq = wq / SW;
          const bool ok = okr && sx < KW && wq >= 0 && qq < d.q;
          const int64_t o = ok ? (((int64_t)n * d.p + pp) * d.q + qq) * d.c + c0 : 0;
          ib[i * NS + j] = ok ? *reinterpret_cast<const uint32_t*>(idx + o) : 0xFFFFFFFFu;  // (0xFF is never a tap code)
          g[i * NS + j] = ok ? ld_f4_stream(dy + o) : make_float4(0.f, 0.f, 0.f, 0.f);
          code[i * NS + j] = (uint32_t)(r * KW + sx);
        }
      }
#pragma unroll
      for (int e = 0; e < NR * NS; ++e) {  // r ascending, then s ascending: the first hit is the last writer in raster order
        unsigned hit = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) hit |= (((ib[e] >> (8 * j)) & 0xFFu) == code[e]) ? (1u << j) : 0u;
        if (!ACCUM) hit &= ~done;
        const float gv[4] = {g[e].x, g[e].y, g[e].z, g[e].w};
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (hit & (1u << j)) acc[j] = ACCUM ? acc[j] + gv[j] : gv[j];
        done |= hit;
      }
      st_f4(dx + 4 * t, make_float4(acc[0], acc[1], acc[2], acc[3]));
      continue;
    }
    for (int r = 0; r < d.kh && done != 0xFu; ++r) {
      int hp = h + d.pad_h - r * d.dil_h;
      if (hp < 0) break;
      if (hp % d.stride_h) continue;
      int p = hp / d.stride_h;
      if (p >= d.p) continue;
      for (int s = 0; s < d.kw; ++s) {
        int wq = w + d.pad_w - s * d.dil_w;
        if (wq < 0) break;
        if (wq % d.stride_w) continue;
        int q = wq / d.stride_w;
        if (q >= d.q) continue;
        const int64_t o = (((int64_t)n * d.p + p) * d.q + q) * d.c + c0;
        const uint32_t ib = *reinterpret_cast<const uint32_t*>(idx + o);
        const uint32_t code = (uint32_t)(r * d.kw + s);
        unsigned hit = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) hit |= (((ib >> (8 * j)) & 0xFFu) == code) ? (1u << j) : 0u;
        if (!ACCUM) hit &= ~done;
        if (hit) {
          const float4 g = ld_f4_stream(dy + o);
          const float gv[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (hit & (1u << j)) acc[j] = ACCUM ? acc[j] + gv[j] : gv[j];
          if (!ACCUM) {
            done |= hit;
            if (done == 0xFu) break;
          }
        }
      }
    }
    st_f4(dx + 4 * t, make_float4(acc[0], acc[1], acc[2], acc[3]));
  }
}

// 3 x 3 / stride 2 (the ResNet stem pool): one thread = a 2 x 2 block of input pixels x 4 channels.  With u = h + pad, the
// rows u = 2a+1 (covered by window p = a through tap r = 1) and u = 2a+2 (p = a+1 through r = 0, p = a through r = 2) see only
// the windows p in {a, a+1}, and likewise for the columns: the block's four pixels read the SAME four windows, whose index
// words and gradients are loaded once (4 + 4 loads for 4 outputs; the per-pixel gather loads 9 + 9).  The candidates of a
// pixel are visited r ascending, then s ascending - the first hit is the last writer in raster order (reference
// grad_nn.py:820), exactly as maxpool_bwd4_kernel does.
template <bool ACCUM>
__global__ void __launch_bounds__(256)
maxpool_bwd_k3s2_kernel(ttb_pool_desc d, const float* __restrict__ dy, const uint8_t* __restrict__ idx, float* __restrict__ dx,
                        int a_lo, int na, int b_lo, int nb) {
  pdl_entry();
  const int cq = d.c / 4;
  const int64_t total = (int64_t)d.n * na * nb * cq;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
    const int c0 = (int)(t % cq) * 4;
    int64_t rest = t / cq;
    const int b = b_lo + (int)(rest % nb);
    rest /= nb;
    const int a = a_lo + (int)(rest % na);
    const int n = (int)(rest / na);
    // the four windows (p, q) in {a, a+1} x {b, b+1}: index word and gradient, 0xFFFFFFFF / 0 when the window does not exist
    uint32_t ib[2][2];
    float4 g[2][2];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int p = a + i, q = b + j;
        const bool ok = p >= 0 && p < d.p && q >= 0 && q < d.q;
        const int64_t o = ok ? (((int64_t)n * d.p + p) * d.q + q) * d.c + c0 : 0;
        ib[i][j] = ok ? *reinterpret_cast<const uint32_t*>(idx + o) : 0xFFFFFFFFu;
        g[i][j] = ok ? ld_f4_stream(dy + o) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    // pixel (hi, wi) of the block: hi = 0 -> u odd (tap r = 1 of window a), hi = 1 -> u even (r = 0 of a+1, then r = 2 of a)
#pragma unroll
    for (int hi = 0; hi < 2; ++hi) {
      const int h = 2 * a + 1 + hi - d.pad_h;
      if (h < 0 || h >= d.h) continue;
#pragma unroll
      for (int wi = 0; wi < 2; ++wi) {
        const int w = 2 * b + 1 + wi - d.pad_w;
        if (w < 0 || w >= d.w) continue;
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        unsigned done = 0;
        // candidate list in (r ascending, s ascending) order: (window row offset i, tap r), (window column offset j, tap s)
        const int nr = hi ? 2 : 1, ns = wi ? 2 : 1;
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          if (e >= nr) break;
          const int i = hi ? 1 - e : 0, r = hi ? 2 * e : 1;
#pragma unroll
          for (int f = 0; f < 2; ++f) {
            if (f >= ns) break;
            const int j = wi ? 1 - f : 0, sx = wi ? 2 * f : 1;
            const uint32_t code = (uint32_t)(r * 3 + sx), word = ib[i][j];
            unsigned hit = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) hit |= (((word >> (8 * k)) & 0xFFu) == code) ? (1u << k) : 0u;
            if (!ACCUM) hit &= ~done;
            const float gv[4] = {g[i][j].x, g[i][j].y, g[i][j].z, g[i][j].w};
#pragma unroll
            for (int k = 0; k < 4; ++k)
              if (hit & (1u << k)) acc[k] = ACCUM ? acc[k] + gv[k] : gv[k];
            done |= hit;
          }
        }
        st_f4(dx + ((((int64_t)n * d.h + h) * d.w + w) * d.c + c0), make_float4(acc[0], acc[1], acc[2], acc[3]));
      }
    }
  }
}

// 2 x 2 / stride 2 (UNet's pools): windows do not overlap, so the 2 x 2 block of input pixels u = h + pad in {2a, 2a+1} x
// {2b, 2b+1} belongs to window (a, b) alone - one index word and one gradient load serve its four pixels (tap r*2 + s).
// ACCUM and last-writer-wins coincide (one candidate per pixel).
__global__ void __launch_bounds__(256)
maxpool_bwd_k2s2_kernel(ttb_pool_desc d, const float* __restrict__ dy, const uint8_t* __restrict__ idx, float* __restrict__ dx,
                        int a_lo, int na, int b_lo, int nb) {
  pdl_entry();
  const int cq = d.c / 4;
  const int64_t total = (int64_t)d.n * na * nb * cq;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
    const int c0 = (int)(t % cq) * 4;
    int64_t rest = t / cq;
    const int b = b_lo + (int)(rest % nb);
    rest /= nb;
    const int a = a_lo + (int)(rest % na);
    const int n = (int)(rest / na);
    const bool ok = a >= 0 && a < d.p && b >= 0 && b < d.q;
    const int64_t o = ok ? (((int64_t)n * d.p + a) * d.q + b) * d.c + c0 : 0;
    const uint32_t word = ok ? *reinterpret_cast<const uint32_t*>(idx + o) : 0xFFFFFFFFu;
    const float4 g = ok ? ld_f4_stream(dy + o) : make_float4(0.f, 0.f, 0.f, 0.f);
    const float gv[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int h = 2 * a + r - d.pad_h;
      if (h < 0 || h >= d.h) continue;
#pragma unroll
      for (int sx = 0; sx < 2; ++sx) {
        const int w = 2 * b + sx - d.pad_w;
        if (w < 0 || w >= d.w) continue;
        const uint32_t code = (uint32_t)(r * 2 + sx);
        float acc[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[k] = (((word >> (8 * k)) & 0xFFu) == code) ? gv[k] : 0.f;
        st_f4(dx + ((((int64_t)n * d.h + h) * d.w + w) * d.c + c0), make_float4(acc[0], acc[1], acc[2], acc[3]));
      }
    }
  }
}

}  // namespace ttb

using namespace ttb;

extern "C" {

int ttb_maxpool2d_fwd(const ttb_pool_desc* d, const float* x, float* y, uint8_t* idx, void* stream) {
  TTB_REQUIRE(d != nullptr, "maxpool2d_fwd: null descriptor");
  TTB_REQUIRE(d->kh * d->kw <= 255, "maxpool2d_fwd: window of %dx%d does not fit the byte index", d->kh, d->kw);
  int cvec = (d->c % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0) ? 4 : 1;
  int64_t total = (int64_t)d->n * d->p * d->q * (d->c / cvec);
  if (total <= 0) return 0;
  if (cvec == 4 && d->dil_h == 1 && d->dil_w == 1 && (reinterpret_cast<uintptr_t>(y) & 15) == 0 &&
      (reinterpret_cast<uintptr_t>(idx) & 3) == 0) {
    const int g4 = elementwise_grid(total, 256);
    if (d->kh == 3 && d->kw == 3 && d->stride_h == 2 && d->stride_w == 2) {
      launch_k(maxpool_fwd4_kernel<3, 3, 2, 2>, g4, 256, 0, as_stream(stream), *d, x, y, idx, total);
      return check_launch("maxpool2d_fwd");
    }
    if (d->kh == 2 && d->kw == 2 && d->stride_h == 2 && d->stride_w == 2) {
      launch_k(maxpool_fwd4_kernel<2, 2, 2, 2>, g4, 256, 0, as_stream(stream), *d, x, y, idx, total);
      return check_launch("maxpool2d_fwd");
    }
  }
  int grid = elementwise_grid(total, 256);
  launch_k(maxpool_fwd_kernel, grid, 256, 0, as_stream(stream), *d, x, y, idx, total, cvec);
  return check_launch("maxpool2d_fwd");
}

int ttb_maxpool2d_bwd(const ttb_pool_desc* d, const float* dy, const uint8_t* idx, float* dx, int accumulate,
                      void* stream) {
  TTB_REQUIRE(d != nullptr, "maxpool2d_bwd: null descriptor");
  int64_t total = (int64_t)d->n * d->h * d->w * d->c;
  if (total <= 0) return 0;
  if (d->c % 4 == 0 && ((reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(dx)) & 15) == 0 &&
      (reinterpret_cast<uintptr_t>(idx) & 3) == 0) {
    int grid4 = elementwise_grid(total / 4, 256);
    cudaStream_t st = as_stream(stream);
    const bool unit_dil = d->dil_h == 1 && d->dil_w == 1;
#define TTB_POOL_BWD(KH, KW, SH, SW)                                                                            \
  do {                                                                                                          \
    if (accumulate) launch_k(maxpool_bwd4_kernel<true, KH, KW, SH, SW>, grid4, 256, 0, st, *d, dy, idx, dx, total / 4);  \
    else launch_k(maxpool_bwd4_kernel<false, KH, KW, SH, SW>, grid4, 256, 0, st, *d, dy, idx, dx, total / 4);   \
  } while (0)
    if (unit_dil && d->kh == 3 && d->kw == 3 && d->stride_h == 2 && d->stride_w == 2) {
      // blocks of 2 x 2 input pixels: u = h + pad in {2a+1, 2a+2}, a = floor((u - 1) / 2) over u in [pad, H - 1 + pad]
      auto fl = [](int v) { return v >= 0 ? v / 2 : -((-v + 1) / 2); };  // floor(v / 2)
      const int a_lo = fl(d->pad_h - 1), a_hi = fl(d->h + d->pad_h - 2);
      const int b_lo = fl(d->pad_w - 1), b_hi = fl(d->w + d->pad_w - 2);
      const int na = a_hi - a_lo + 1, nb = b_hi - b_lo + 1;
      const int grid = elementwise_grid((int64_t)d->n * na * nb * (d->c / 4), 256);
      if (accumulate) launch_k(maxpool_bwd_k3s2_kernel<true>, grid, 256, 0, st, *d, dy, idx, dx, a_lo, na, b_lo, nb);
      else launch_k(maxpool_bwd_k3s2_kernel<false>, grid, 256, 0, st, *d, dy, idx, dx, a_lo, na, b_lo, nb);
    } else if (unit_dil && d->kh == 2 && d->kw == 2 && d->stride_h == 2 && d->stride_w == 2) {
      auto fl = [](int v) { return v >= 0 ? v / 2 : -((-v + 1) / 2); };  // floor(v / 2)
      const int a_lo = fl(d->pad_h), a_hi = fl(d->h - 1 + d->pad_h), b_lo = fl(d->pad_w), b_hi = fl(d->w - 1 + d->pad_w);
      const int na = a_hi - a_lo + 1, nb = b_hi - b_lo + 1;
      const int grid = elementwise_grid((int64_t)d->n * na * nb * (d->c / 4), 256);
      launch_k(maxpool_bwd_k2s2_kernel, grid, 256, 0, st, *d, dy, idx, dx, a_lo, na, b_lo, nb);
    }
    else TTB_POOL_BWD(0, 0, 0, 0);
#undef TTB_POOL_BWD
    return check_launch("maxpool2d_bwd");
  }
  int grid = elementwise_grid(total, 256);
  if (accumulate) launch_k(maxpool_bwd_kernel<true>, grid, 256, 0, as_stream(stream), *d, dy, idx, dx, total);
  else launch_k(maxpool_bwd_kernel<false>, grid, 256, 0, as_stream(stream), *d, dy, idx, dx, total);
  return check_launch("maxpool2d_bwd");
}

}  // extern "C"
