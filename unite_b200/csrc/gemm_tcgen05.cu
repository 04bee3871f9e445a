// tcgen05 / TMEM / TMA GEMM for sm_100a:   C[M,N] = epilogue(A[M,K] * B[N,K]^T), bf16 in, fp32 accumulate.
//
// Replaces every dense contraction of the UNITE step (see include/unite_b200.h for the reference
// call sites).  Design (B200-first, nothing like the reference's cuBLAS calls):
//   * persistent CTAs (one per SM), static round-robin over (m-tile, n-tile, k-split) work items;
//   * warp 0  : TMA producer  (cp.async.bulk.tensor, 128B-swizzled boxes, STAGES-deep mbarrier ring);
//   * warp 1  : single-thread tcgen05.mma issuer, 128 x BN x 16 UMMA, accumulators in TMEM,
//               two accumulator stages (2*BN columns) so the epilogue of tile i overlaps the MMAs of tile i+1;
//   * warps 2-9: epilogue, tcgen05.ld 32x32b -> registers -> swizzled smem transpose -> fused bias / activation /
//               DropPath scale / residual with fully coalesced 128-bit global loads and stores
//               (or red.global.add.v4.f32 for split-K weight gradients).
//   * operands may be K-major (activations, weights [out,in]) or MN-major (the same row-major tensors
//     contracted over their ROW index: dgrad uses W as B^T, wgrad contracts over tokens) — no transposes
//     are ever materialised.
#include "common.cuh"
#include <cstdlib>
#include <cstring>
#include "../../include/unite_b200.h"

namespace ub {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int GEMM_THREADS = 320;  // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue

struct GemmParams {
  void* C;
  int64_t ldc;
  int M, N, K;
  int splits, kb_per_split;
  ub_gemm_epilogue ep;
};

// EPI: 0 = bias/activation only, 1 = + fp32 residual, 2 = DGELU (reads the bf16 pre-activation)
// NCTA: 1 = one CTA per 128 x BN tile; 2 = a CTA pair (cluster of 2, tcgen05 cta_group::2) per 256 x BN tile: each CTA
//       holds its own 128 rows of A and HALF of B, which cuts the smem fill + operand-read traffic per MMA by a third.
template <int BN, int STAGES, bool A_MN, bool B_MN, int EPI, int NCTA>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
  constexpr int BN_L = BN / NCTA;            // rows of B resident in this CTA
  constexpr int A_BYTES = BM * BK * 2;
  constexpr int B_BYTES = BN_L * BK * 2;
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr uint32_t TMEM_COLS = 2 * BN;  // 256 or 512: power of two
  constexpr uint32_t IDESC = umma_idesc_bf16(BM * NCTA, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
  const uint32_t cta_rank = NCTA == 2 ? cluster_ctarank() : 0u;
  const bool leader = cta_rank == 0;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  uint8_t* epi_s = smem + STAGES * STAGE_BYTES + 256;  // 8 warps x 4 KB epilogue staging (after the barriers)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 8 * NCTA);   // the leader's copy collects the epilogue warps of both CTAs
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    if (NCTA == 2) {
      tmem_alloc_cg2(tmem_slot, TMEM_COLS);
      tmem_relinquish_cg2();
    } else {
      tmem_alloc(tmem_slot, TMEM_COLS);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  if (NCTA == 2) cluster_sync(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int m_tiles = (p.M + BM * NCTA - 1) / (BM * NCTA);
  const int n_tiles = (p.N + BN - 1) / BN;
  const int total_kb = (p.K + BK - 1) / BK;
  const int total_work = m_tiles * n_tiles * p.splits;
  const int w_first = blockIdx.x / NCTA, w_step = gridDim.x / NCTA;   // both CTAs of a pair walk the same work items

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    int stage = 0;
    uint32_t phase = 0;
    for (int w = w_first; w < total_work; w += w_step) {
      const int ks = w % p.splits;
      const int tile = w / p.splits;
      const int m0 = (tile / n_tiles) * (BM * NCTA) + (int)cta_rank * BM;      // this CTA's rows of A
      const int n0 = (tile % n_tiles) * BN + (int)cta_rank * BN_L;             // this CTA's rows of B
      const int kb0 = ks * p.kb_per_split;
      const int kb1 = min(total_kb, kb0 + p.kb_per_split);
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&empty[stage], phase ^ 1);
        if (lane == 0) {
          uint8_t* sA = smem + stage * STAGE_BYTES;
          uint8_t* sB = sA + A_BYTES;
          if (NCTA == 2) {
            // both CTAs' loads complete on the LEADER's full barrier, which expects the bytes of the whole pair
            const uint32_t lbar = mapa_u32(smem_u32(&full[stage]), 0);
            if (leader) mbar_expect_tx(&full[stage], STAGE_BYTES * 2);
            if (A_MN) {
#pragma unroll
              for (int j = 0; j < BM / 64; ++j) tma_load_2d_cg2(&tmA, lbar, sA + j * 8192, m0 + j * 64, kb * BK);
            } else {
              tma_load_2d_cg2(&tmA, lbar, sA, kb * BK, m0);
            }
            if (B_MN) {
#pragma unroll
              for (int j = 0; j < BN_L / 64; ++j) tma_load_2d_cg2(&tmB, lbar, sB + j * 8192, n0 + j * 64, kb * BK);
            } else {
              tma_load_2d_cg2(&tmB, lbar, sB, kb * BK, n0);
            }
          } else {
            mbar_expect_tx(&full[stage], STAGE_BYTES);
            if (A_MN) {
#pragma unroll
              for (int j = 0; j < BM / 64; ++j) tma_load_2d(&tmA, &full[stage], sA + j * 8192, m0 + j * 64, kb * BK);
            } else {
              tma_load_2d(&tmA, &full[stage], sA, kb * BK, m0);
            }
            if (B_MN) {
#pragma unroll
              for (int j = 0; j < BN_L / 64; ++j) tma_load_2d(&tmB, &full[stage], sB + j * 8192, n0 + j * 64, kb * BK);
            } else {
              tma_load_2d(&tmB, &full[stage], sB, kb * BK, n0);
            }
          }
        }
        __syncwarp();
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1 && leader) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA of a pair only)
    int stage = 0;
    uint32_t phase = 0;
    int as = 0;
    uint32_t aphase = 0;
    for (int w = w_first; w < total_work; w += w_step) {
      const int ks = w % p.splits;
      const int kb0 = ks * p.kb_per_split;
      const int kb1 = min(total_kb, kb0 + p.kb_per_split);
      if (NCTA == 2) mbar_wait_cluster(&tempty[as], aphase ^ 1); else mbar_wait(&tempty[as], aphase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(as * BN);
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t sA = smem_u32(smem + stage * STAGE_BYTES);
          const uint32_t sB = sA + A_BYTES;
          const uint64_t adesc = A_MN ? umma_desc_mnmajor_sw128(sA, 8192) : umma_desc_kmajor_sw128(sA);
          const uint64_t bdesc = B_MN ? umma_desc_mnmajor_sw128(sB, 8192) : umma_desc_kmajor_sw128(sB);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // advance along K by 16 elements: 32 B inside the swizzle atom (K-major) or two 8-row groups (MN-major)
            const uint64_t ad = adesc + (uint64_t)(A_MN ? (k * 2048) >> 4 : (k * 32) >> 4);
            const uint64_t bd = bdesc + (uint64_t)(B_MN ? (k * 2048) >> 4 : (k * 32) >> 4);
            if (NCTA == 2) umma_bf16_cg2(d_tmem, ad, bd, IDESC, (kb > kb0 || k > 0) ? 1u : 0u);
            else umma_bf16(d_tmem, ad, bd, IDESC, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          // frees the smem slot (in both CTAs of a pair) once these MMAs have read it
          if (NCTA == 2) umma_commit_cg2(&empty[stage]); else umma_commit(&empty[stage]);
          if (kb == kb1 - 1) {
            if (NCTA == 2) umma_commit_cg2(&tfull[as]); else umma_commit(&tfull[as]);
          }
        }
        __syncwarp();
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
      as ^= 1;
      if (as == 0) aphase ^= 1;
    }
  } else if (warp >= 2) {
    // ------------------------------------------------------------------ epilogue (warps 2..9)
    // Two warps per TMEM sub-partition, each owning half of the tile's columns, 32 columns at a time:
    //   phase 1  tcgen05.ld gives every lane one ROW (32 fp32) -> written to this warp's private 4 KB smem
    //            staging tile with a 16-byte XOR swizzle (conflict-free);
    //   phase 2  the tile is read back so that 8 consecutive lanes hold one row's 128 contiguous bytes, and ALL
    //            global traffic of the epilogue (bias, residual, aux, output, red.add) is issued in that layout:
    //            a warp instruction touches 4 full 128 B lines instead of 32 partial ones.
    // Global operands are prefetched one chunk ahead so their latency hides behind TMEM/smem work.
    const int sp = warp & 3;            // TMEM sub-partition this warp may read
    const int half = (warp - 2) >> 2;   // which half of the BN columns
    constexpr int CHUNKS = BN / 64;     // 32-column chunks per warp
    uint8_t* stg = epi_s + (warp - 2) * 4096;
    const uint32_t stg_a = smem_u32(stg);
    const int sub_r = lane >> 3;        // phase-2: row within a group of 4
    const int c4 = lane & 7;            // phase-2: 16-byte chunk (4 fp32 columns) of the 32-column slab
    int as = 0;
    uint32_t aphase = 0;
    const ub_gemm_epilogue& ep = p.ep;
    constexpr bool has_res = EPI == 1;
    constexpr bool has_aux_in = EPI == 2;
    for (int w = w_first; w < total_work; w += w_step) {
      const int tile = w / p.splits;
      const int m0 = (tile / n_tiles) * (BM * NCTA) + (int)cta_rank * BM;
      const int n0 = (tile % n_tiles) * BN;
      const int row_base = m0 + sp * 32;
      const int cbase = n0 + half * (BN / 2);
      // per-tile operands in the phase-2 layout
      float rs[8];
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int row = row_base + it * 4 + sub_r;
        rs[it] = (ep.row_scale != nullptr && row < p.M) ? __ldg(ep.row_scale + row / ep.rows_per_scale) : 1.0f;
      }
      float4 rn[has_res ? 8 : 1];
      uint2 an[has_aux_in ? 8 : 1];
      float4 bn;
      auto prefetch = [&](int c) {
        const int col = cbase + c * 32 + c4 * 4;
        bn = (ep.bias != nullptr && col < p.N) ? __ldg(reinterpret_cast<const float4*>(ep.bias + col)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int row = row_base + it * 4 + sub_r;
          const bool ok = row < p.M && col < p.N;
          if constexpr (has_res)
            rn[it] = ok ? *reinterpret_cast<const float4*>(ep.residual + (int64_t)row * ep.ldr + col) : make_float4(0.f, 0.f, 0.f, 0.f);
          if constexpr (has_aux_in)
            an[it] = ok ? *reinterpret_cast<const uint2*>(reinterpret_cast<const bf16*>(ep.aux_in) + (int64_t)row * ep.ld_aux + col)
                        : make_uint2(0u, 0u);
        }
      };
      prefetch(0);
      mbar_wait(&tfull[as], aphase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(sp * 32) << 16) + (uint32_t)(as * BN + half * (BN / 2));
#pragma unroll 1
      for (int c = 0; c < CHUNKS; ++c) {
        const int col = cbase + c * 32 + c4 * 4;
        if (cbase + c * 32 >= p.N) break;
        uint32_t r[32];
        tmem_ld_32x32(t_row + (uint32_t)(c * 32), r);
        float4 rc[has_res ? 8 : 1];
        uint2 ac[has_aux_in ? 8 : 1];
        const float4 bc = bn;
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          if constexpr (has_res) rc[it] = rn[it];
          if constexpr (has_aux_in) ac[it] = an[it];
        }
        tmem_ld_wait();
        // phase 1: row-per-lane registers -> swizzled staging tile
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stg_a + lane * 128 + ((j ^ (lane & 7)) << 4)),
                       "r"(r[4 * j]), "r"(r[4 * j + 1]), "r"(r[4 * j + 2]), "r"(r[4 * j + 3])
                       : "memory");
        }
        __syncwarp();
        if (c + 1 < CHUNKS) prefetch(c + 1);
        // phase 2: coalesced epilogue
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int rr = it * 4 + sub_r;
          const int row = row_base + rr;
          float4 v;
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                       : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                       : "r"(stg_a + rr * 128 + ((c4 ^ (rr & 7)) << 4)));
          if (row >= p.M || col >= p.N) continue;
          v.x += bc.x; v.y += bc.y; v.z += bc.z; v.w += bc.w;
          if (ep.act == UB_ACT_QUICKGELU) {
            v.x = quick_gelu(v.x); v.y = quick_gelu(v.y); v.z = quick_gelu(v.z); v.w = quick_gelu(v.w);
          } else if (ep.act == UB_ACT_GELU) {
            if (ep.aux_out != nullptr) {
              uint2 pk;
              pk.x = pack_bf16x2(v.x, v.y); pk.y = pack_bf16x2(v.z, v.w);
              *reinterpret_cast<uint2*>(reinterpret_cast<bf16*>(ep.aux_out) + (int64_t)row * ep.ld_aux + col) = pk;
            }
            v.x = gelu_erf(v.x); v.y = gelu_erf(v.y); v.z = gelu_erf(v.z); v.w = gelu_erf(v.w);
          }
          if constexpr (has_aux_in) {
            const float2 a0 = unpack_bf16x2(ac[it].x), a1 = unpack_bf16x2(ac[it].y);
            v.x *= gelu_erf_grad(a0.x); v.y *= gelu_erf_grad(a0.y); v.z *= gelu_erf_grad(a1.x); v.w *= gelu_erf_grad(a1.y);
          }
          if (ep.row_scale != nullptr) { v.x *= rs[it]; v.y *= rs[it]; v.z *= rs[it]; v.w *= rs[it]; }
          if constexpr (has_res) { v.x += rc[it].x; v.y += rc[it].y; v.z += rc[it].z; v.w += rc[it].w; }
          if (ep.out_fp32) {
            float* out = reinterpret_cast<float*>(p.C) + (int64_t)row * p.ldc + col;
            if (ep.accumulate) {
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(out), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                           : "memory");
            } else {
              *reinterpret_cast<float4*>(out) = v;
            }
          } else {
            uint2 pk;
            pk.x = pack_bf16x2(v.x, v.y); pk.y = pack_bf16x2(v.z, v.w);
            *reinterpret_cast<uint2*>(reinterpret_cast<bf16*>(p.C) + (int64_t)row * p.ldc + col) = pk;
          }
        }
        __syncwarp();   // staging tile is rewritten by the next chunk
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (NCTA == 2) mbar_arrive_cluster(mapa_u32(smem_u32(&tempty[as]), 0)); else mbar_arrive(&tempty[as]);
      }
      as ^= 1;
      if (as == 0) aphase ^= 1;
    }
  }

  tc_fence_before();
  if (NCTA == 2) cluster_sync(); else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if (NCTA == 2) tmem_dealloc_cg2(tmem_base, TMEM_COLS); else tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
  }
  return fn;
}

// 2D bf16 row-major tensor [rows, cols] with leading dimension ld (elements); box = {box_cols, box_rows}.
int make_tmap_bf16_2d(CUtensorMap* out, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_cols,
                      int box_rows) {
  EncodeTiledFn enc = get_encode_fn();
  UB_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available from the driver");
  UB_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA base pointer must be 16-byte aligned");
  UB_REQUIRE((ld * 2) % 16 == 0, "TMA leading dimension must be a multiple of 8 bf16 elements (ld=%lld)",
             (long long)ld);
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  UB_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult %d (rows=%lld cols=%lld ld=%lld)", (int)r,
             (long long)rows, (long long)cols, (long long)ld);
  return 0;
}

template <int BN, int STAGES, bool A_MN, bool B_MN, int EPI, int NCTA>
static int launch_gemm_epi(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p, int grid,
                           cudaStream_t stream) {
  constexpr int SMEM = STAGES * (BM * BK * 2 + (BN / NCTA) * BK * 2) + 1024 + 256 + 8 * 4096;
  static_assert(SMEM <= 232448, "exceeds the 227 KB per-CTA shared memory limit");
  static bool configured = false;
  auto kern = gemm_kernel<BN, STAGES, A_MN, B_MN, EPI, NCTA>;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    UB_REQUIRE(e == cudaSuccess, "cudaFuncSetAttribute(gemm smem=%d): %s", SMEM, cudaGetErrorString(e));
    configured = true;
  }
  if (NCTA == 1) {
    kern<<<grid, GEMM_THREADS, SMEM, stream>>>(tmA, tmB, p);
  } else {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(GEMM_THREADS);
    cfg.dynamicSmemBytes = SMEM;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, tmA, tmB, p);
    UB_REQUIRE(e == cudaSuccess, "gemm_kernel (CTA-pair) launch: %s", cudaGetErrorString(e));
  }
  return check_launch("gemm_kernel");
}

template <int BN, int STAGES, bool A_MN, bool B_MN, int NCTA>
static int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p, int grid, cudaStream_t stream) {
  if (p.ep.residual != nullptr) return launch_gemm_epi<BN, STAGES, A_MN, B_MN, 1, NCTA>(tmA, tmB, p, grid, stream);
  if (p.ep.act == UB_ACT_DGELU) return launch_gemm_epi<BN, STAGES, A_MN, B_MN, 2, NCTA>(tmA, tmB, p, grid, stream);
  return launch_gemm_epi<BN, STAGES, A_MN, B_MN, 0, NCTA>(tmA, tmB, p, grid, stream);
}

}  // namespace ub

extern "C" int ub_gemm_bf16(const void* A, int64_t lda, int a_mn_major, const void* B, int64_t ldb, int b_mn_major,
                            void* C, int64_t ldc, int M, int N, int K, const ub_gemm_epilogue* ep_in, int split_k,
                            void* stream) {
  using namespace ub;
  UB_REQUIRE(M > 0 && N > 0 && K > 0, "gemm: empty problem M=%d N=%d K=%d", M, N, K);
  UB_REQUIRE(N % 8 == 0, "gemm: N must be a multiple of 8 (N=%d)", N);
  UB_REQUIRE(A && B && C, "gemm: null operand");
  ub_gemm_epilogue ep;
  if (ep_in) ep = *ep_in; else { memset(&ep, 0, sizeof(ep)); }
  if (split_k < 1) split_k = 1;
  UB_REQUIRE(split_k == 1 || (ep.accumulate && ep.out_fp32), "gemm: split_k>1 needs fp32 accumulate output");
  UB_REQUIRE(!ep.accumulate || ep.out_fp32, "gemm: accumulate needs fp32 output");
  UB_REQUIRE(ep.row_scale == nullptr || ep.rows_per_scale > 0, "gemm: rows_per_scale must be > 0");
  UB_REQUIRE((ldc * (ep.out_fp32 ? 4 : 2)) % 16 == 0, "gemm: ldc must keep rows 16-byte aligned");
  UB_REQUIRE(ep.act != UB_ACT_DGELU || ep.aux_in != nullptr, "gemm: DGELU needs aux_in");
  UB_REQUIRE(!(ep.act == UB_ACT_DGELU && ep.residual != nullptr), "gemm: DGELU with a residual is not supported");

  const int total_kb = (K + BK - 1) / BK;
  if (split_k > total_kb) split_k = total_kb;
  int kb_per_split = (total_kb + split_k - 1) / split_k;
  split_k = (total_kb + kb_per_split - 1) / kb_per_split;  // no empty splits

  // tile choice: 256-wide tiles unless that costs a whole extra wave; CTA pairs (256 x 256 per pair) for the wide tiles
  const int sms = sm_count();
  static int force_ncta = -1;
  if (force_ncta < 0) {
    const char* e = getenv("UB_GEMM_NCTA");
    force_ncta = e ? atoi(e) : 0;
  }
  auto cost = [&](int bm, int bn, int units) {
    const long tiles = (long)((M + bm - 1) / bm) * ((N + bn - 1) / bn) * split_k;
    const long waves = (tiles + units - 1) / units;
    return waves * (long)bm * bn / (bm / BM);   // time ~ per-SM tile area x waves
  };
  const int bn = (N <= 128 || cost(BM, 128, sms) < cost(BM, 256, sms)) ? 128 : 256;
  int ncta = 1;
  if (bn == 256 && M > BM) ncta = (cost(2 * BM, 256, sms / 2) <= cost(BM, 256, sms)) ? 2 : 1;
  if (force_ncta == 1) ncta = 1;
  if (force_ncta == 2 && bn == 256) ncta = 2;

  CUtensorMap tmA, tmB;
  if (a_mn_major) {
    if (make_tmap_bf16_2d(&tmA, A, K, M, lda, 64, 64)) return 1;
  } else {
    if (make_tmap_bf16_2d(&tmA, A, M, K, lda, BK, BM)) return 1;
  }
  if (b_mn_major) {
    if (make_tmap_bf16_2d(&tmB, B, K, N, ldb, 64, 64)) return 1;
  } else {
    if (make_tmap_bf16_2d(&tmB, B, N, K, ldb, BK, bn / ncta)) return 1;
  }

  GemmParams p;
  p.C = C; p.ldc = ldc; p.M = M; p.N = N; p.K = K;
  p.splits = split_k; p.kb_per_split = kb_per_split; p.ep = ep;
  const long total_work = (long)((M + BM * ncta - 1) / (BM * ncta)) * ((N + bn - 1) / bn) * split_k;
  const int units = sms / ncta;
  const int grid = (int)(total_work < units ? total_work : units) * ncta;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);

#define UB_GEMM_CASE(AMN, BMN)                                                              \
  if ((a_mn_major != 0) == AMN && (b_mn_major != 0) == BMN) {                                \
    if (bn == 256 && ncta == 2) return launch_gemm<256, 6, AMN, BMN, 2>(tmA, tmB, p, grid, st); \
    return bn == 256 ? launch_gemm<256, 4, AMN, BMN, 1>(tmA, tmB, p, grid, st)               \
                     : launch_gemm<128, 6, AMN, BMN, 1>(tmA, tmB, p, grid, st);              \
  }
  UB_GEMM_CASE(false, false)
  UB_GEMM_CASE(false, true)
  UB_GEMM_CASE(true, true)
#undef UB_GEMM_CASE
  set_error("gemm: unsupported operand-major combination a_mn=%d b_mn=%d", a_mn_major, b_mn_major);
  return 1;
}
