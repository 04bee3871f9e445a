// Token-level data movement and the mask sampler (HBM / latency bound).
//
//   ub_patchify      fp32 clip -> bf16 im2col rows (the A operand of the stride==kernel Conv3d as a GEMM)
//                    clip.py:123-128,146 (teacher conv1), modeling_finetune.py:165-174 (student PatchEmbed.proj)
//   ub_mask_select   attention-guided visible-token selection, bit-exact
//                    run_stage1.py:379-387 (multinomial == top-N_vis of attn/q, q ~ Exp(1) supplied by the caller)
//                    utils.py:89-120 get_greedy_masks (k committee members take ranks i, i+k, ...; q == NULL)
//   ub_gather_rows   out[i,:] = in[idx[i],:]  — replaces the boolean-mask gathers `x[~mask]`
//                    (modeling_adaptation.py:153,319; run_stage1.py:393) without their nonzero() host syncs
//   ub_colsum_bf16   bias gradients: out[n] += sum_m dY[m,n]
#include "common.cuh"
#include "../../include/unite_b200.h"

namespace ub {

// ------------------------------------------------------------------------------------------------
// patchify: x fp32 [B,3,T,H,W] -> out bf16 [B*(T/tub)*(H/p)*(W/p), 3*tub*p*p], p == 16
// token order (t,h,w); feature order (c,kt,kh,kw) == flattened Conv3d weight [D,3,tub,16,16].
// One thread moves one (token, c, kt, kh) run of 16 contiguous pixels: 64 B read, 32 B written; consecutive
// threads take consecutive runs of the same token so the bf16 writes of a warp are one contiguous 1 KB span.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) patchify_kernel(const float* __restrict__ x, bf16* __restrict__ out, int B, int T,
                                                       int H, int W, int tub, long total_runs) {
  pdl_grid_sync();
  const int gh = H / 16, gw = W / 16, Tp = T / tub;
  const int runs_per_tok = 3 * tub * 16;
  for (long id = (long)blockIdx.x * blockDim.x + threadIdx.x; id < total_runs; id += (long)gridDim.x * blockDim.x) {
    const int run = (int)(id % runs_per_tok);
    const long tok = id / runs_per_tok;
    const int kh = run % 16, kt = (run / 16) % tub, c = run / (16 * tub);
    const int pw = (int)(tok % gw), ph = (int)((tok / gw) % gh), tp = (int)((tok / ((long)gw * gh)) % Tp);
    const int b = (int)(tok / ((long)gw * gh * Tp));
    const float* src = x + ((((long)b * 3 + c) * T + (tp * tub + kt)) * H + (ph * 16 + kh)) * W + pw * 16;
    const float4 f0 = *reinterpret_cast<const float4*>(src), f1 = *reinterpret_cast<const float4*>(src + 4),
                 f2 = *reinterpret_cast<const float4*>(src + 8), f3 = *reinterpret_cast<const float4*>(src + 12);
    uint4 o0, o1;
    o0.x = pack_bf16x2(f0.x, f0.y); o0.y = pack_bf16x2(f0.z, f0.w); o0.z = pack_bf16x2(f1.x, f1.y); o0.w = pack_bf16x2(f1.z, f1.w);
    o1.x = pack_bf16x2(f2.x, f2.y); o1.y = pack_bf16x2(f2.z, f2.w); o1.z = pack_bf16x2(f3.x, f3.y); o1.w = pack_bf16x2(f3.z, f3.w);
    bf16* dst = out + tok * (long)(runs_per_tok * 16) + run * 16;
    stg_v4(dst, o0);
    stg_v4(dst + 8, o1);
  }
}

// ------------------------------------------------------------------------------------------------
// patchify_u8: decoded frames uint8 [B,T,H,W,3] (decord / NVDEC layout) -> ImageNet-normalised bf16 im2col rows, i.e.
// ToTensor + tensor_normalize (kinetics_sparse.py:236-243, :434-451) + the THWC->CTHW permute + patchify in one pass:
//   v = (u / 255 - mean[c]) / std[c]   in fp32 with IEEE divides (bit-identical to the torch CPU pipeline), then bf16.
// One thread = one (token, kt, kh) run of 16 pixels: 48 contiguous bytes in, three 32-byte runs out (one per channel).
// The step then needs 38.5 MB of H2D per 32-clip batch instead of the 154 MB fp32 clip.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) patchify_u8_kernel(const uint8_t* __restrict__ x, bf16* __restrict__ out, float m0, float m1,
                                                          float m2, float s0, float s1, float s2, int B, int T, int H, int W, int tub,
                                                          long total_runs) {
  pdl_grid_sync();
  const int gh = H / 16, gw = W / 16, Tp = T / tub;
  const int runs_per_tok = tub * 16;          // (kt, kh)
  const int feat = 3 * tub * 256;
  const float mean[3] = {m0, m1, m2}, sd[3] = {s0, s1, s2};
  for (long id = (long)blockIdx.x * blockDim.x + threadIdx.x; id < total_runs; id += (long)gridDim.x * blockDim.x) {
    const int run = (int)(id % runs_per_tok);
    const long tok = id / runs_per_tok;
    const int kh = run % 16, kt = run / 16;
    const int pw = (int)(tok % gw), ph = (int)((tok / gw) % gh), tp = (int)((tok / ((long)gw * gh)) % Tp);
    const int b = (int)(tok / ((long)gw * gh * Tp));
    const uint8_t* src = x + ((((long)b * T + (tp * tub + kt)) * H + (ph * 16 + kh)) * W + pw * 16) * 3;
    uint32_t w[12];
    *reinterpret_cast<uint4*>(&w[0]) = ldg_nc_v4(src);
    *reinterpret_cast<uint4*>(&w[4]) = ldg_nc_v4(src + 16);
    *reinterpret_cast<uint4*>(&w[8]) = ldg_nc_v4(src + 32);
    const uint8_t* by = reinterpret_cast<const uint8_t*>(w);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      uint32_t o[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float a = __fdiv_rn(__fsub_rn(__fdiv_rn((float)by[(2 * i) * 3 + c], 255.0f), mean[c]), sd[c]);
        const float d = __fdiv_rn(__fsub_rn(__fdiv_rn((float)by[(2 * i + 1) * 3 + c], 255.0f), mean[c]), sd[c]);
        o[i] = pack_bf16x2(a, d);
      }
      bf16* dst = out + tok * (long)feat + (long)c * (tub * 256) + kt * 256 + kh * 16;
      stg_v4(dst, make_uint4(o[0], o[1], o[2], o[3]));
      stg_v4(dst + 8, make_uint4(o[4], o[5], o[6], o[7]));
    }
  }
}

// ------------------------------------------------------------------------------------------------
// mask selection.  One warp per frame (row of attn): score = attn / q (IEEE fp32 divide, same as torch),
// rank_i = #{j : s_j > s_i or (s_j == s_i and j < i)};  member = rank % k, visible for that member iff rank / k < n_vis.
// Outputs (all per member m):
//   mask     uint8 [k, frames*P]        1 = masked (viewed as bool [k, B, T*P])
//   vis_idx  int32 [k, B, T*n_vis]      ascending token index inside the clip (t*P + p)
//   tea_rows int32 [k, B, T*n_vis]      row of that token in the teacher stream: (b*T + t)*(P+1) + 1 + p   (optional)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) mask_select_kernel(const float* __restrict__ attn, const float* __restrict__ q,
                                                          uint8_t* __restrict__ mask, int* __restrict__ vis_idx,
                                                          int* __restrict__ tea_rows, int frames, int P, int T, int k,
                                                          int n_vis) {
  pdl_grid_sync();
  extern __shared__ float s_scores[];  // [warps][P] scores, then ranks as int
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int f = blockIdx.x * (blockDim.x >> 5) + warp;
  if (f >= frames) return;
  float* sc = s_scores + warp * 2 * P;
  int* rk = reinterpret_cast<int*>(sc + P);
  for (int i = lane; i < P; i += 32) {
    const float a = attn[(long)f * P + i];
    sc[i] = q ? a / q[(long)f * P + i] : a;
  }
  __syncwarp();
  for (int i = lane; i < P; i += 32) {
    const float si = sc[i];
    int r = 0;
    for (int j = 0; j < P; ++j) {
      const float sj = sc[j];
      r += (sj > si) || (sj == si && j < i);
    }
    rk[i] = r;
  }
  __syncwarp();
  const int b = f / T, t = f % T;
  const long total = (long)frames * P;
  for (int m = 0; m < k; ++m) {
    int count = 0;
    for (int i0 = 0; i0 < P; i0 += 32) {
      const int i = i0 + lane;
      bool vis = false;
      if (i < P) {
        const int r = rk[i];
        vis = (r % k == m) && (r / k < n_vis);
        mask[(long)m * total + (long)f * P + i] = vis ? 0 : 1;
      }
      const unsigned bal = __ballot_sync(0xffffffffu, vis);
      if (vis) {
        const int slot = count + __popc(bal & ((1u << lane) - 1u));
        const long o = ((long)m * (frames / T) + b) * ((long)T * n_vis) + (long)t * n_vis + slot;
        vis_idx[o] = t * P + i;
        if (tea_rows) tea_rows[o] = (b * T + t) * (P + 1) + 1 + i;
      }
      count += __popc(bal);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// gather rows: out[i, :] = in[(base(i) +) idx[i], :], rows of `row_bytes` (multiple of 16) bytes
//   rows_per_group > 0: idx is relative to a group base:  src = (i / rows_per_group) * group_stride + idx[i]
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gather_rows_kernel(const uint4* __restrict__ in, const int* __restrict__ idx,
                                                          uint4* __restrict__ out, long n_rows, int chunks_per_row,
                                                          int rows_per_group, long group_stride) {
  pdl_grid_sync();
  const long total = n_rows * chunks_per_row;
  for (long id = (long)blockIdx.x * blockDim.x + threadIdx.x; id < total; id += (long)gridDim.x * blockDim.x) {
    const long r = id / chunks_per_row;
    const int c = (int)(id % chunks_per_row);
    long src = idx[r];
    if (rows_per_group > 0) src += (r / rows_per_group) * group_stride;
    out[id] = in[src * chunks_per_row + c];
  }
}

// ------------------------------------------------------------------------------------------------
// column sums of a bf16 matrix (bias gradients), accumulated into fp32 with red.add.
// CTA = 256 threads = 8 row-groups x 32 lanes; a lane owns 8 columns (one 16-byte load per row), the CTA covers
// 256 columns x ROWS_PER_CTA rows with 4 independent loads in flight per thread, then reduces the 8 row-groups
// through smem and issues one red.add per column.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) colsum_bf16_kernel(const bf16* __restrict__ x, int64_t ld, float* __restrict__ out,
                                                          int M, int N, int CS_ROWS, int skip_lo, int skip_hi) {
  pdl_grid_sync();
  __shared__ float s_part[8][256];
  const int lane = threadIdx.x & 31, rg = threadIdx.x >> 5;
  const int col = blockIdx.x * 256 + lane * 8;
  const int r0 = blockIdx.y * CS_ROWS;
  const int r1 = min(M, r0 + CS_ROWS);
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  if (col < N) {
    for (int r = r0 + rg; r < r1; r += 32) {
      uint4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u)
        v[u] = (r + u * 8 < r1) ? ldg_nc_v4(x + (int64_t)(r + u * 8) * ld + col) : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float2 f;
        f = unpack_bf16x2(v[u].x); acc[0] += f.x; acc[1] += f.y;
        f = unpack_bf16x2(v[u].y); acc[2] += f.x; acc[3] += f.y;
        f = unpack_bf16x2(v[u].z); acc[4] += f.x; acc[5] += f.y;
        f = unpack_bf16x2(v[u].w); acc[6] += f.x; acc[7] += f.y;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) s_part[rg][lane * 8 + i] = acc[i];
  __syncthreads();
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c < N && !(c >= skip_lo && c < skip_hi)) {      // skipped columns (the structurally zero key bias) are left untouched
    float s = 0.f;
#pragma unroll
    for (int g = 0; g < 8; ++g) s += s_part[g][threadIdx.x];
    atomicAdd(out + c, s);
  }
}

// out_bf16 = bf16(x * row_scale[row / rows_per_scale])   (fp32 -> bf16 cast of the residual-stream gradient)
__global__ void __launch_bounds__(256) cast_scale_bf16_kernel(const float4* __restrict__ x, uint2* __restrict__ out,
                                                              const float* __restrict__ row_scale, int rows_per_scale,
                                                              long n4, int d4) {
  pdl_grid_sync();
  for (long id = (long)blockIdx.x * blockDim.x + threadIdx.x; id < n4; id += (long)gridDim.x * blockDim.x) {
    float4 v = x[id];
    if (row_scale) {
      const float s = __ldg(row_scale + (id / d4) / rows_per_scale);
      v.x *= s; v.y *= s; v.z *= s; v.w *= s;
    }
    uint2 o;
    o.x = pack_bf16x2(v.x, v.y);
    o.y = pack_bf16x2(v.z, v.w);
    out[id] = o;
  }
}

static int flat_grid(long n_items, int block) {
  long want = (n_items + block - 1) / block;
  const long cap = (long)sm_count() * 16;
  return (int)(want < cap ? (want < 1 ? 1 : want) : cap);
}

}  // namespace ub

using namespace ub;

extern "C" int ub_patchify(const float* x, void* out, int B, int T, int H, int W, int tubelet, void* stream) {
  UB_REQUIRE(x && out, "patchify: null pointer");
  UB_REQUIRE(B > 0 && T > 0 && tubelet > 0 && T % tubelet == 0, "patchify: bad temporal shape T=%d tubelet=%d", T, tubelet);
  UB_REQUIRE(H % 16 == 0 && W % 16 == 0 && H > 0 && W > 0, "patchify: H, W must be multiples of the 16-pixel patch (H=%d W=%d)", H, W);
  const long tokens = (long)B * (T / tubelet) * (H / 16) * (W / 16);
  const long runs = tokens * 3 * tubelet * 16;
  UB_LAUNCH(patchify_kernel, flat_grid(runs, 256), 256, 0, (cudaStream_t)stream, x, (bf16*)out, B, T, H, W, tubelet, runs);
  return check_launch("patchify_kernel");
}

extern "C" int ub_patchify_u8(const uint8_t* x, void* out, const float* mean3, const float* std3, int B, int T, int H, int W,
                              int tubelet, void* stream) {
  UB_REQUIRE(x && out && mean3 && std3, "patchify_u8: null pointer");
  UB_REQUIRE(B > 0 && T > 0 && tubelet > 0 && T % tubelet == 0 && H % 16 == 0 && W % 16 == 0, "patchify_u8: bad shape B=%d T=%d H=%d W=%d tub=%d", B,
             T, H, W, tubelet);
  UB_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0, "patchify_u8: input must be 16-byte aligned");
  UB_REQUIRE(std3[0] != 0.f && std3[1] != 0.f && std3[2] != 0.f, "patchify_u8: zero std");
  const long runs = (long)B * (T / tubelet) * (H / 16) * (W / 16) * tubelet * 16;
  UB_LAUNCH(patchify_u8_kernel, flat_grid(runs, 256), 256, 0, (cudaStream_t)stream, x, (bf16*)out, mean3[0], mean3[1], mean3[2], std3[0],
            std3[1], std3[2], B, T, H, W, tubelet, runs);
  return check_launch("patchify_u8_kernel");
}

extern "C" int ub_mask_select(const float* attn, const float* q, uint8_t* mask, int* vis_idx, int* tea_rows, int frames,
                              int P, int T, int k, int n_vis, void* stream) {
  UB_REQUIRE(attn && mask && vis_idx, "mask_select: null pointer");
  UB_REQUIRE(frames > 0 && P > 0 && T > 0 && frames % T == 0, "mask_select: frames=%d must be a multiple of T=%d", frames, T);
  UB_REQUIRE(k >= 1 && n_vis >= 1 && (long)k * n_vis <= P, "mask_select: k=%d x n_vis=%d exceeds P=%d", k, n_vis, P);
  const int warps = 4;
  const size_t smem = (size_t)warps * 2 * P * sizeof(float);
  UB_REQUIRE(smem <= 48 * 1024, "mask_select: P=%d too large", P);
  UB_LAUNCH(mask_select_kernel, (frames + warps - 1) / warps, warps * 32, smem, (cudaStream_t)stream, attn, q, mask, vis_idx, tea_rows,
                                                                                               frames, P, T, k, n_vis);
  return check_launch("mask_select_kernel");
}

extern "C" int ub_gather_rows(const void* in, const int* idx, void* out, int64_t n_rows, int64_t row_bytes,
                              int rows_per_group, int64_t group_stride_rows, void* stream) {
  UB_REQUIRE(in && idx && out, "gather_rows: null pointer");
  UB_REQUIRE(n_rows > 0 && row_bytes > 0 && row_bytes % 16 == 0, "gather_rows: row_bytes=%lld must be a positive multiple of 16",
             (long long)row_bytes);
  const int cpr = (int)(row_bytes / 16);
  UB_LAUNCH(gather_rows_kernel, flat_grid(n_rows * cpr, 256), 256, 0, (cudaStream_t)stream, (const uint4*)in, idx, (uint4*)out, n_rows,
                                                                                    cpr, rows_per_group, group_stride_rows);
  return check_launch("gather_rows_kernel");
}

extern "C" int ub_colsum_bf16(const void* x, int64_t ld, float* out, int M, int N, int skip_lo, int skip_hi, void* stream) {
  UB_REQUIRE(x && out && M > 0 && N > 0 && N % 8 == 0 && ld % 8 == 0, "colsum_bf16: N and ld must be multiples of 8 (M=%d N=%d)", M, N);
  UB_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0, "colsum_bf16: x must be 16-byte aligned");
  // rows per CTA (multiple of 32): small enough that the grid keeps ~6 CTAs per SM in flight — this is a latency-bound pass
  // over a few tens of MB, so bytes in flight per SM matter more than the number of atomics at the end
  const int gx = (N + 255) / 256;
  const int want_y = (6 * sm_count() + gx - 1) / gx;
  int rows = ((M + want_y - 1) / want_y + 31) / 32 * 32;
  rows = rows < 32 ? 32 : (rows > 128 ? 128 : rows);
  dim3 grid(gx, (M + rows - 1) / rows);
  UB_LAUNCH(colsum_bf16_kernel, grid, 256, 0, (cudaStream_t)stream, (const bf16*)x, ld, out, M, N, rows, skip_lo, skip_hi);
  return check_launch("colsum_bf16_kernel");
}

extern "C" int ub_cast_scale_bf16(const float* x, void* out, const float* row_scale, int rows_per_scale, int64_t rows, int D,
                                  void* stream) {
  UB_REQUIRE(x && out && rows > 0 && D > 0 && D % 4 == 0, "cast_scale_bf16: bad arguments");
  UB_REQUIRE(row_scale == nullptr || rows_per_scale > 0, "cast_scale_bf16: rows_per_scale must be > 0");
  const long n4 = rows * (D / 4);
  UB_LAUNCH(cast_scale_bf16_kernel, flat_grid(n4, 256), 256, 0, (cudaStream_t)stream, (const float4*)x, (uint2*)out, row_scale,
                                                                              rows_per_scale, n4, D / 4);
  return check_launch("cast_scale_bf16_kernel");
}
