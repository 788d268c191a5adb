// Tall-skinny projection C (rows x r) = A (rows x K) * V (K x r) on the tcgen05 tensor cores at fp32 accuracy
// ("3xTF32": A = A_hi + A_lo, V = V_hi + V_lo, C = A_hi V_hi + A_hi V_lo + A_lo V_hi, error ~2^-21 relative,
// i.e. the accuracy class of an fp32 FFMA product — a plain TF32 projection would put a 2^-11 relative error
// straight into the reconstruction).  This is the "C <- C V_r" step of the sweep (round.py:181, tensor.py:2081-2083).
//
//   * A row blocks (128 rows x 32 k) are staged by TMA (K-major, SWIZZLE_128B), 4-stage mbarrier ring; the matching
//     V_hi^T / V_lo^T chunks (r x 32 k) ride in the same stage;
//   * the tensor core truncates fp32 operands to TF32 itself, so A_hi is the raw tile; A_lo = A - trunc(A) is
//     produced by four "split" warps straight on the swizzled bytes (an elementwise map keeps the layout) into a
//     second buffer, fenced into the async proxy;
//   * one elected thread issues 3 x (32/8) tcgen05.mma.kind::tf32 (M=128, N=r_pad, K=8) per stage into one of
//     four TMEM accumulator slots; four epilogue warps drain finished tiles (tcgen05.ld) to global while the
//     next row block is already being multiplied (persistent CTAs, static round-robin over row blocks).
//
// Bound: HBM (reads A once: 4 B/element for 2*r flop) — the three MMAs per k-step cost 3*128*r*8 MACs per
// 128x8 elements, far below the tensor-pipe limit; the shared-memory pipe (TMA fill + split read/write + 3 MMA
// operand reads) is the secondary limit.
#pragma once
#include "gram_tc.cuh"

namespace tnb {

constexpr int PT_BM = 128, PT_KC = 32, PT_STAGES = 4, PT_ACC_SLOTS = 4, PT_THREADS = 320;
constexpr int PT_A_BYTES = PT_BM * PT_KC * 4;      // 16 KB
constexpr int PT_MAX_N = 64;                        // r padded to a multiple of 16, <= 64
constexpr int PT_B_BYTES = PT_MAX_N * PT_KC * 4;    // 8 KB slot per V part
constexpr int PT_STAGE_BYTES = 2 * PT_A_BYTES + 2 * PT_B_BYTES;  // A, A_lo, Vhi, Vlo
constexpr int PT_SMEM_BYTES = PT_STAGES * PT_STAGE_BYTES + 1024 + 512;

struct ProjTcParams {
  int64_t rows;
  int K;
  int r;       // real output columns
  int npad;    // r rounded up to a multiple of 16
  int64_t num_row_blocks;
  int nk;      // K chunks
  int slab;    // chunks per accumulator slab: the TMEM accumulator truncates on every add, so long K ranges are cut
               // into slabs of PT_SLAB_CHUNKS*32 columns whose partial tiles are summed in fp32 (RN) by the epilogue
  int nslabs;
  float* C;
};
constexpr int PT_SLAB_CHUNKS = 8;

// K-major operand, SWIZZLE_128B: 8-row groups 1024 B apart (SBO); LBO unused for swizzled K-major layouts.
__device__ __forceinline__ uint64_t make_k_major_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;                       // LBO (ignored)
  d |= (uint64_t)((1024u >> 4) & 0x3FFF) << 32; // SBO
  d |= (uint64_t)1 << 46;                       // version
  d |= (uint64_t)2 << 61;                       // SWIZZLE_128B
  return d;
}
__host__ __device__ inline uint32_t make_idesc_tf32_kk(int M, int N) {
  uint32_t d = 0;
  d |= 1u << 4;                    // c_format = F32
  d |= 2u << 7;                    // a_format = TF32
  d |= 2u << 10;                   // b_format = TF32
  d |= (uint32_t)(N >> 3) << 17;   // n_dim   (a_major = b_major = K-major = 0)
  d |= (uint32_t)(M >> 4) << 24;   // m_dim
  return d;
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__global__ void __launch_bounds__(PT_THREADS, 1)
project_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_vhi,
                  const __grid_constant__ CUtensorMap tmap_vlo, const ProjTcParams p) {
  extern __shared__ unsigned char pt_smem_raw[];
  const uint32_t raw_addr = smem_u32(pt_smem_raw);
  const uint32_t pad = (1024u - (raw_addr & 1023u)) & 1023u;
  unsigned char* stage_base = pt_smem_raw + pad;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(stage_base + PT_STAGES * PT_STAGE_BYTES);
  uint64_t* split_bar = full_bar + PT_STAGES;
  uint64_t* empty_bar = split_bar + PT_STAGES;
  uint64_t* acc_full = empty_bar + PT_STAGES;
  uint64_t* acc_empty = acc_full + PT_ACC_SLOTS;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(acc_empty + PT_ACC_SLOTS);

  const int warp_idx = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < PT_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&split_bar[s], 4);   // one arrival per split warp
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < PT_ACC_SLOTS; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], 4);   // one arrival per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp_idx == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)),
                 "r"((uint32_t)(PT_ACC_SLOTS * PT_MAX_N))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  const int64_t my_blocks = (p.num_row_blocks - blockIdx.x + gridDim.x - 1) / gridDim.x;  // may be 0
  const int64_t total_items = my_blocks * p.nk;

  if (warp_idx == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      const uint32_t tx_bytes = (uint32_t)PT_A_BYTES + 2u * (uint32_t)(p.npad * PT_KC * 4);
      int stage = 0;
      uint32_t phase = 0;
      for (int64_t item = 0; item < total_items; ++item) {
        const int64_t rb = blockIdx.x + (item / p.nk) * (int64_t)gridDim.x;
        const int kc = (int)(item % p.nk);
        mbar_wait(&empty_bar[stage], phase ^ 1u);
        unsigned char* sb = stage_base + stage * PT_STAGE_BYTES;
        mbar_expect_tx(&full_bar[stage], tx_bytes);
        tma_load_2d(sb, &tmap_a, &full_bar[stage], kc * PT_KC, (int)(rb * PT_BM));
        tma_load_2d(sb + 2 * PT_A_BYTES, &tmap_vhi, &full_bar[stage], kc * PT_KC, 0);
        tma_load_2d(sb + 2 * PT_A_BYTES + PT_B_BYTES, &tmap_vlo, &full_bar[stage], kc * PT_KC, 0);
        if (++stage == PT_STAGES) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp_idx == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      const uint32_t idesc = make_idesc_tf32_kk(PT_BM, p.npad);
      int stage = 0;
      uint32_t phase = 0;
      const int64_t tiles = my_blocks * p.nslabs;
      for (int64_t tile = 0; tile < tiles; ++tile) {
        const int slot = (int)(tile % PT_ACC_SLOTS);
        const uint32_t acc_phase = (uint32_t)((tile / PT_ACC_SLOTS) & 1);
        const int sl = (int)(tile % p.nslabs);
        const int kc_begin = sl * p.slab, kc_end = (kc_begin + p.slab < p.nk) ? kc_begin + p.slab : p.nk;
        mbar_wait(&acc_empty[slot], acc_phase ^ 1u);  // epilogue has drained this accumulator
        tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(slot * PT_MAX_N);
        for (int kc = kc_begin; kc < kc_end; ++kc) {
          const uint32_t sb = smem_u32(stage_base + stage * PT_STAGE_BYTES);
          mbar_wait(&full_bar[stage], phase);
          tcgen05_fence_after();
#pragma unroll
          for (int ks = 0; ks < PT_KC / 8; ++ks) {  // A_hi (raw, truncated by the tensor core) x V_hi, x V_lo
            const uint64_t a = make_k_major_desc(sb + ks * 32u);
            tcgen05_mma_tf32(tmem_d, a, make_k_major_desc(sb + 2 * PT_A_BYTES + ks * 32u), idesc, (kc > kc_begin || ks > 0) ? 1u : 0u);
            tcgen05_mma_tf32(tmem_d, a, make_k_major_desc(sb + 2 * PT_A_BYTES + PT_B_BYTES + ks * 32u), idesc, 1u);
          }
          mbar_wait(&split_bar[stage], phase);  // A_lo written and fenced
          tcgen05_fence_after();
#pragma unroll
          for (int ks = 0; ks < PT_KC / 8; ++ks)
            tcgen05_mma_tf32(tmem_d, make_k_major_desc(sb + PT_A_BYTES + ks * 32u),
                             make_k_major_desc(sb + 2 * PT_A_BYTES + ks * 32u), idesc, 1u);
          tcgen05_commit(&empty_bar[stage]);
          if (++stage == PT_STAGES) { stage = 0; phase ^= 1u; }
        }
        tcgen05_commit(&acc_full[slot]);
      }
    }
  } else if (warp_idx < 6) {
    // ================= split warps: A_lo = A - trunc_tf32(A), elementwise on the swizzled tile =================
    const int t = threadIdx.x - 64;  // 0..127
    int stage = 0;
    uint32_t phase = 0;
    for (int64_t item = 0; item < total_items; ++item) {
      mbar_wait(&full_bar[stage], phase);
      const float4* src = reinterpret_cast<const float4*>(stage_base + stage * PT_STAGE_BYTES);
      float4* dst = reinterpret_cast<float4*>(stage_base + stage * PT_STAGE_BYTES + PT_A_BYTES);
#pragma unroll
      for (int i = 0; i < PT_A_BYTES / 16 / 128; ++i) {
        const float4 v = src[t + i * 128];
        float4 o;
        o.x = v.x - __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u);
        o.y = v.y - __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u);
        o.z = v.z - __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u);
        o.w = v.w - __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u);
        dst[t + i * 128] = o;
      }
      fence_proxy_async_smem();  // generic-proxy writes -> visible to the tensor core's async-proxy reads
      __syncwarp();
      if (lane == 0) mbar_arrive(&split_bar[stage]);
      if (++stage == PT_STAGES) { stage = 0; phase ^= 1u; }
    }
  } else {
    // ================= epilogue warps =================
    const int lane_group = warp_idx & 3;
    const int row_in_tile = lane_group * 32 + lane;
    const int64_t tiles = my_blocks * p.nslabs;
    for (int64_t tile = 0; tile < tiles; ++tile) {
      const int slot = (int)(tile % PT_ACC_SLOTS);
      const uint32_t acc_phase = (uint32_t)((tile / PT_ACC_SLOTS) & 1);
      const int64_t blk = tile / p.nslabs;
      const bool add = (tile % p.nslabs) != 0;  // later slabs of a row block add to what this warp stored before
      const int64_t rb = blockIdx.x + blk * (int64_t)gridDim.x;
      const int64_t grow = rb * PT_BM + row_in_tile;
      mbar_wait(&acc_full[slot], acc_phase);
      tcgen05_fence_after();
      for (int c0 = 0; c0 < p.npad; c0 += 32) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(lane_group * 32) << 16) + (uint32_t)(slot * PT_MAX_N + c0), v);
        tmem_ld_wait();
        if (grow < p.rows) {
          float* out = p.C + grow * p.r + c0;
          if ((p.r & 3) == 0) {
#pragma unroll
            for (int q = 0; q < 8; ++q)
              if (c0 + 4 * q + 3 < p.r) {
                float4 o = make_float4(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]),
                                       __uint_as_float(v[4 * q + 2]), __uint_as_float(v[4 * q + 3]));
                float4* dst = reinterpret_cast<float4*>(out + 4 * q);
                if (add) {
                  const float4 old = *dst;
                  o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
                }
                *dst = o;
              }
          } else {
#pragma unroll
            for (int q = 0; q < 32; ++q)
              if (c0 + q < p.r) out[q] = add ? out[q] + __uint_as_float(v[q]) : __uint_as_float(v[q]);
          }
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[slot]);
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp_idx == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"((uint32_t)(PT_ACC_SLOTS * PT_MAX_N))
                 : "memory");
  }
}

// Vt_hi / Vt_lo (npad x K, row-major, K contiguous) from V (K x r): hi = tf32-truncated V, lo = V - hi.
__global__ void split_v_kernel(const float* __restrict__ V, int K, int r, int npad, float* __restrict__ Vhi,
                               float* __restrict__ Vlo) {
  const int total = npad * K;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int j = idx / K, k = idx % K;
    float v = 0.f;
    if (j < r) v = V[(size_t)k * r + j];
    const float hi = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
    Vhi[idx] = hi;
    Vlo[idx] = v - hi;
  }
}

inline bool project_tc_shape_ok(int64_t rows, int64_t K, int64_t r, const void* A, const void* C) {
  return r >= 1 && r <= PT_MAX_N && K % 4 == 0 && K >= 32 && K <= (1 << 24) && rows >= 128 &&
         rows < ((int64_t)1 << 31) - 256 && (reinterpret_cast<uintptr_t>(A) & 15u) == 0 &&
         (reinterpret_cast<uintptr_t>(C) & 15u) == 0;
}
inline size_t project_tc_workspace_bytes(int64_t K, int64_t r) {
  const int64_t npad = (r + 15) / 16 * 16;
  return 2 * align_up((size_t)npad * K * sizeof(float));
}

inline int encode_kmajor_f32(CUtensorMap* tmap, const float* ptr, int64_t rows, int64_t cols, int box_rows) {
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)cols * sizeof(float)};
  cuuint32_t box[2] = {(cuuint32_t)PT_KC, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult cr = get_encode_tiled()(tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), gdim, gstride, box,
                                   estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) return fail(TNB_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)cr);
  return TNB_OK;
}

// C (rows x r) = A (rows x K) V (K x r); ws: project_tc_workspace_bytes(K, r).
inline int project_tc_f32(const float* A, int64_t rows, int64_t K, const float* V, int r, float* C, void* ws, size_t ws_bytes,
                          cudaStream_t st) {
  if (!tc_path_available()) return fail(TNB_ERR_UNSUPPORTED, "project_tc: needs an sm_100 device");
  if (!project_tc_shape_ok(rows, K, r, A, C)) return fail(TNB_ERR_UNSUPPORTED, "project_tc: unsupported shape");
  if (ws_bytes < project_tc_workspace_bytes(K, r)) return fail(TNB_ERR_WORKSPACE, "project_tc: workspace too small");
  ProjTcParams p;
  p.rows = rows; p.K = (int)K; p.r = r; p.npad = (r + 15) / 16 * 16;
  p.num_row_blocks = (rows + PT_BM - 1) / PT_BM;
  p.nk = (int)((K + PT_KC - 1) / PT_KC);
  p.slab = PT_SLAB_CHUNKS;
  p.nslabs = (p.nk + p.slab - 1) / p.slab;
  p.C = C;
  float* Vhi = static_cast<float*>(ws);
  float* Vlo = reinterpret_cast<float*>(static_cast<char*>(ws) + align_up((size_t)p.npad * K * sizeof(float)));
  split_v_kernel<<<grid_for((int64_t)p.npad * K), 256, 0, st>>>(V, (int)K, r, p.npad, Vhi, Vlo);
  TNB_LAUNCH_CHECK();
  CUtensorMap ta, th, tl;
  TNB_TRY(encode_kmajor_f32(&ta, A, rows, K, PT_BM));
  TNB_TRY(encode_kmajor_f32(&th, Vhi, p.npad, K, p.npad));
  TNB_TRY(encode_kmajor_f32(&tl, Vlo, p.npad, K, p.npad));
  static bool attr_set = false;
  if (!attr_set) {
    TNB_CUDA(cudaFuncSetAttribute(project_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PT_SMEM_BYTES));
    attr_set = true;
  }
  const int sms = usable_sms();
  const int64_t grid = p.num_row_blocks < sms ? p.num_row_blocks : sms;
  project_tc_kernel<<<(unsigned)grid, PT_THREADS, PT_SMEM_BYTES, st>>>(ta, th, tl, p);
  TNB_LAUNCH_CHECK();
  return TNB_OK;
}

}  // namespace tnb
