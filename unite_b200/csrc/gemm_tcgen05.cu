// tcgen05 / TMEM / TMA GEMM for sm_100a:   C[M,N] = epilogue(A[M,K] * B[N,K]^T), bf16 in, fp32 accumulate.
//
// Replaces every dense contraction of the UNITE step (see include/unite_b200.h for the reference
// call sites).  Design (B200-first, nothing like the reference's cuBLAS calls):
//   * persistent CTAs (one per SM), static round-robin over (m-tile, n-tile, k-split) work items;
//   * warp 0  : TMA producer  (cp.async.bulk.tensor, 128B-swizzled boxes, STAGES-deep mbarrier ring);
//   * warp 1  : single-thread tcgen05.mma issuer, 128 x BN x 16 UMMA, accumulators in TMEM,
//               two accumulator stages (2*BN columns) so the epilogue of tile i overlaps the MMAs of tile i+1;
//   * warps 2-5: epilogue, tcgen05.ld 32x32b -> registers -> fused bias / activation / DropPath scale /
//               residual -> 128-bit global stores (or fp32 red.add for split-K weight gradients).
//   * operands may be K-major (activations, weights [out,in]) or MN-major (the same row-major tensors
//     contracted over their ROW index: dgrad uses W as B^T, wgrad contracts over tokens) — no transposes
//     are ever materialised.
#include "common.cuh"
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

template <int BN, int STAGES, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
  constexpr int A_BYTES = BM * BK * 2;
  constexpr int B_BYTES = BN * BK * 2;
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr uint32_t TMEM_COLS = 2 * BN;  // 256 or 512: power of two
  constexpr uint32_t IDESC = umma_idesc_bf16(BM, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  float* bias_s = reinterpret_cast<float*>(tmem_slot + 4);  // [2][BN]

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
      mbar_init(&tempty[i], 8);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int m_tiles = (p.M + BM - 1) / BM;
  const int n_tiles = (p.N + BN - 1) / BN;
  const int total_kb = (p.K + BK - 1) / BK;
  const int total_work = m_tiles * n_tiles * p.splits;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    int stage = 0;
    uint32_t phase = 0;
    for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
      const int ks = w % p.splits;
      const int tile = w / p.splits;
      const int m0 = (tile / n_tiles) * BM;
      const int n0 = (tile % n_tiles) * BN;
      const int kb0 = ks * p.kb_per_split;
      const int kb1 = min(total_kb, kb0 + p.kb_per_split);
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&empty[stage], phase ^ 1);
        if (lane == 0) {
          uint8_t* sA = smem + stage * STAGE_BYTES;
          uint8_t* sB = sA + A_BYTES;
          mbar_expect_tx(&full[stage], STAGE_BYTES);
          if (A_MN) {
#pragma unroll
            for (int j = 0; j < BM / 64; ++j) tma_load_2d(&tmA, &full[stage], sA + j * 8192, m0 + j * 64, kb * BK);
          } else {
            tma_load_2d(&tmA, &full[stage], sA, kb * BK, m0);
          }
          if (B_MN) {
#pragma unroll
            for (int j = 0; j < BN / 64; ++j) tma_load_2d(&tmB, &full[stage], sB + j * 8192, n0 + j * 64, kb * BK);
          } else {
            tma_load_2d(&tmB, &full[stage], sB, kb * BK, n0);
          }
        }
        __syncwarp();
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    int stage = 0;
    uint32_t phase = 0;
    int as = 0;
    uint32_t aphase = 0;
    for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
      const int ks = w % p.splits;
      const int kb0 = ks * p.kb_per_split;
      const int kb1 = min(total_kb, kb0 + p.kb_per_split);
      mbar_wait(&tempty[as], aphase ^ 1);
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
            umma_bf16(d_tmem, ad, bd, IDESC, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty[stage]);  // frees the smem slot once these MMAs have read it
          if (kb == kb1 - 1) umma_commit(&tfull[as]);
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
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..9)
    // Two warps per TMEM sub-partition, each owning half of the tile's columns.  Global operands of the
    // epilogue are never waited on serially: the bias slice is staged in smem once per tile, residual / aux
    // rows are prefetched one 32-column chunk ahead of the TMEM load they are combined with.
    const int sp = warp & 3;            // TMEM sub-partition this warp may read
    const int half = (warp - 2) >> 2;   // which half of the BN columns
    const int etid = threadIdx.x - 64;  // 0..255
    constexpr int CHUNKS = BN / 64;     // 32-column chunks per warp
    int as = 0;
    uint32_t aphase = 0;
    const ub_gemm_epilogue& ep = p.ep;
    const bool has_res = ep.residual != nullptr;
    const bool has_aux_in = ep.act == UB_ACT_DGELU;
    for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
      const int tile = w / p.splits;
      const int m0 = (tile / n_tiles) * BM;
      const int n0 = (tile % n_tiles) * BN;
      const int row = m0 + sp * 32 + lane;
      const bool row_ok = row < p.M;
      float* bias_tile = bias_s + as * BN;
      if (ep.bias != nullptr) {
        if (etid < BN) bias_tile[etid] = (n0 + etid < p.N) ? __ldg(ep.bias + n0 + etid) : 0.0f;
        asm volatile("bar.sync 1, 256;" ::: "memory");
      }
      float rscale = 1.0f;
      if (ep.row_scale != nullptr && row_ok) rscale = __ldg(ep.row_scale + row / ep.rows_per_scale);
      const int cbase = n0 + half * (BN / 2);
      const float* res_row = has_res ? ep.residual + (int64_t)row * ep.ldr : nullptr;
      const bf16* aux_row = has_aux_in ? reinterpret_cast<const bf16*>(ep.aux_in) + (int64_t)row * ep.ld_aux : nullptr;
      float4 rn[8];
      uint4 an[4];
      auto prefetch = [&](int c) {
        const int col0 = cbase + c * 32;
        if (has_res) {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            rn[j] = (row_ok && col0 + j * 4 < p.N) ? *reinterpret_cast<const float4*>(res_row + col0 + j * 4)
                                                  : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        if (has_aux_in) {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            an[j] = (row_ok && col0 + j * 8 < p.N) ? ldg_nc_v4(aux_row + col0 + j * 8) : make_uint4(0, 0, 0, 0);
        }
      };
      prefetch(0);
      mbar_wait(&tfull[as], aphase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(sp * 32) << 16) + (uint32_t)(as * BN + half * (BN / 2));
#pragma unroll 1
      for (int c = 0; c < CHUNKS; ++c) {
        const int col0 = cbase + c * 32;
        if (col0 >= p.N) break;
        uint32_t r[32];
        tmem_ld_32x32(t_row + (uint32_t)(c * 32), r);
        float4 rc[8];
        uint4 ac[4];
#pragma unroll
        for (int j = 0; j < 8; ++j) rc[j] = rn[j];
#pragma unroll
        for (int j = 0; j < 4; ++j) ac[j] = an[j];
        if (c + 1 < CHUNKS) prefetch(c + 1);
        tmem_ld_wait();
        if (!row_ok) continue;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int col = col0 + g * 8;
          if (col >= p.N) break;
          float v[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[g * 8 + i]);
          if (ep.bias != nullptr) {
            const float4 b0 = *reinterpret_cast<const float4*>(bias_tile + (col - n0));
            const float4 b1 = *reinterpret_cast<const float4*>(bias_tile + (col - n0) + 4);
            v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
            v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
          }
          if (ep.act == UB_ACT_QUICKGELU) {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = quick_gelu(v[i]);
          } else if (ep.act == UB_ACT_GELU) {
            if (ep.aux_out != nullptr) {
              uint4 pk;
              pk.x = pack_bf16x2(v[0], v[1]); pk.y = pack_bf16x2(v[2], v[3]);
              pk.z = pack_bf16x2(v[4], v[5]); pk.w = pack_bf16x2(v[6], v[7]);
              stg_v4(reinterpret_cast<bf16*>(ep.aux_out) + (int64_t)row * ep.ld_aux + col, pk);
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = gelu_erf(v[i]);
          } else if (ep.act == UB_ACT_DGELU) {
            const uint4 pk = ac[g];
            const float2 a0 = unpack_bf16x2(pk.x), a1 = unpack_bf16x2(pk.y), a2 = unpack_bf16x2(pk.z),
                         a3 = unpack_bf16x2(pk.w);
            v[0] *= gelu_erf_grad(a0.x); v[1] *= gelu_erf_grad(a0.y);
            v[2] *= gelu_erf_grad(a1.x); v[3] *= gelu_erf_grad(a1.y);
            v[4] *= gelu_erf_grad(a2.x); v[5] *= gelu_erf_grad(a2.y);
            v[6] *= gelu_erf_grad(a3.x); v[7] *= gelu_erf_grad(a3.y);
          }
          if (ep.row_scale != nullptr) {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] *= rscale;
          }
          if (has_res) {
            const float4 r0 = rc[g * 2], r1 = rc[g * 2 + 1];
            v[0] += r0.x; v[1] += r0.y; v[2] += r0.z; v[3] += r0.w;
            v[4] += r1.x; v[5] += r1.y; v[6] += r1.z; v[7] += r1.w;
          }
          if (ep.out_fp32) {
            float* out = reinterpret_cast<float*>(p.C) + (int64_t)row * p.ldc + col;
            if (ep.accumulate) {
#pragma unroll
              for (int i = 0; i < 8; ++i) atomicAdd(out + i, v[i]);
            } else {
              *reinterpret_cast<float4*>(out) = make_float4(v[0], v[1], v[2], v[3]);
              *reinterpret_cast<float4*>(out + 4) = make_float4(v[4], v[5], v[6], v[7]);
            }
          } else {
            uint4 pk;
            pk.x = pack_bf16x2(v[0], v[1]); pk.y = pack_bf16x2(v[2], v[3]);
            pk.z = pack_bf16x2(v[4], v[5]); pk.w = pack_bf16x2(v[6], v[7]);
            stg_v4(reinterpret_cast<bf16*>(p.C) + (int64_t)row * p.ldc + col, pk);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[as]);
      as ^= 1;
      if (as == 0) aphase ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
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

template <int BN, int STAGES, bool A_MN, bool B_MN>
static int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p, int grid,
                       cudaStream_t stream) {
  constexpr int SMEM = STAGES * (BM * BK * 2 + BN * BK * 2) + 1024 + 256 + 2 * BN * 4;
  static bool configured = false;
  auto kern = gemm_kernel<BN, STAGES, A_MN, B_MN>;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    UB_REQUIRE(e == cudaSuccess, "cudaFuncSetAttribute(gemm smem=%d): %s", SMEM, cudaGetErrorString(e));
    configured = true;
  }
  kern<<<grid, GEMM_THREADS, SMEM, stream>>>(tmA, tmB, p);
  return check_launch("gemm_kernel");
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

  const int total_kb = (K + BK - 1) / BK;
  if (split_k > total_kb) split_k = total_kb;
  int kb_per_split = (total_kb + split_k - 1) / split_k;
  split_k = (total_kb + kb_per_split - 1) / kb_per_split;  // no empty splits

  // tile-N choice: fewer, larger tiles unless that costs a whole extra wave
  const int sms = sm_count();
  const int m_tiles = (M + BM - 1) / BM;
  auto cost = [&](int bn) {
    const long tiles = (long)m_tiles * ((N + bn - 1) / bn) * split_k;
    const long waves = (tiles + sms - 1) / sms;
    return waves * bn;
  };
  const int bn = (N <= 128 || cost(128) < cost(256)) ? 128 : 256;

  CUtensorMap tmA, tmB;
  if (a_mn_major) {
    if (make_tmap_bf16_2d(&tmA, A, K, M, lda, 64, 64)) return 1;
  } else {
    if (make_tmap_bf16_2d(&tmA, A, M, K, lda, BK, BM)) return 1;
  }
  if (b_mn_major) {
    if (make_tmap_bf16_2d(&tmB, B, K, N, ldb, 64, 64)) return 1;
  } else {
    if (make_tmap_bf16_2d(&tmB, B, N, K, ldb, BK, bn)) return 1;
  }

  GemmParams p;
  p.C = C; p.ldc = ldc; p.M = M; p.N = N; p.K = K;
  p.splits = split_k; p.kb_per_split = kb_per_split; p.ep = ep;
  const long total_work = (long)m_tiles * ((N + bn - 1) / bn) * split_k;
  const int grid = (int)(total_work < sms ? total_work : sms);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);

#define UB_GEMM_CASE(AMN, BMN)                                                              \
  if ((a_mn_major != 0) == AMN && (b_mn_major != 0) == BMN) {                                \
    return bn == 256 ? launch_gemm<256, 4, AMN, BMN>(tmA, tmB, p, grid, st)                  \
                     : launch_gemm<128, 6, AMN, BMN>(tmA, tmB, p, grid, st);                 \
  }
  UB_GEMM_CASE(false, false)
  UB_GEMM_CASE(false, true)
  UB_GEMM_CASE(true, true)
#undef UB_GEMM_CASE
  set_error("gemm: unsupported operand-major combination a_mn=%d b_mn=%d", a_mn_major, b_mn_major);
  return 1;
}
