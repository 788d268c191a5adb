// Tall-skinny projection C (rows x r) = A (rows x n) * V (n x r), fp32 in / fp32 accumulate (FFMA).
// This is the "C <- C V_r" step of the sweep (round.py:181 `M @ left`, tensor.py:2081-2083 absorb): it
// must keep fp32 accuracy (a TF32 projection would add a 2^-11 relative error straight into the
// reconstruction), and at r = 32 its FFMA work (2*rows*n*r) sits right at the HBM roofline of the
// A read, so the kernel streams A once with 128-bit loads, software-pipelined through registers into a
// transposed shared-memory tile, with V chunks broadcast from shared memory.
#pragma once
#include "common.cuh"

namespace tnb {

constexpr int PROJ_KC = 32;

template <int TX>
__global__ void __launch_bounds__(256) project_f32_kernel(const float* __restrict__ A, int64_t rows, int n,
                                                          const float* __restrict__ V, int r,
                                                          float* __restrict__ C) {
  constexpr int TY = 256 / TX, BR = TY * 4, RP = TX * 4, KC = PROJ_KC;
  constexpr int A4 = BR * (KC / 4) / 256;   // float4 loads of A per thread per chunk
  constexpr int VN = KC * RP / 256;         // V elements per thread per chunk
  __shared__ __align__(16) float As[2][KC][BR + 4];
  __shared__ __align__(16) float Vs[2][KC][RP];
  const int tid = threadIdx.x;
  const int tx = tid % TX, ty = tid / TX;
  float acc[4][4];
  float4 areg[A4];
  float vreg[VN];
  const int nchunks = (n + KC - 1) / KC;
  const int64_t nblocks = (rows + BR - 1) / BR;
  // persistent CTA: a flat stream of (row block, k chunk) work items, software-pipelined through registers
  // across row-block boundaries so that global loads are always one chunk ahead of the FFMAs
  const int64_t my_blocks = (nblocks - blockIdx.x + gridDim.x - 1) / gridDim.x;
  const int64_t total = my_blocks * nchunks;

  auto prefetch = [&](int64_t item) {
    const int64_t row0 = (blockIdx.x + (item / nchunks) * (int64_t)gridDim.x) * BR;
    const int k0 = (int)(item % nchunks) * KC;
#pragma unroll
    for (int i = 0; i < A4; ++i) {
      const int idx = tid + i * 256;
      const int lrow = idx >> 3, kg = idx & 7;
      const int64_t grow = row0 + lrow;
      const int gk = k0 + kg * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (grow < rows && gk < n) v = __ldg(reinterpret_cast<const float4*>(A + grow * n + gk));
      areg[i] = v;
    }
#pragma unroll
    for (int i = 0; i < VN; ++i) {
      const int idx = tid + i * 256;
      const int k = idx / RP, cc = idx % RP;
      float v = 0.f;
      if (k0 + k < n && cc < r) v = __ldg(V + (int64_t)(k0 + k) * r + cc);
      vreg[i] = v;
    }
  };
  auto stage = [&](int buf) {
#pragma unroll
    for (int i = 0; i < A4; ++i) {
      const int idx = tid + i * 256;
      const int lrow = idx >> 3, kg = idx & 7;
      // XOR-swizzle the row index with the k-group: the 32 lanes of a warp (8 k-groups x 4 rows) then hit 32
      // different banks, and groups of 4 rows stay contiguous for the 128-bit reads below
      const int srow = lrow ^ (kg << 2);
      As[buf][kg * 4 + 0][srow] = areg[i].x;
      As[buf][kg * 4 + 1][srow] = areg[i].y;
      As[buf][kg * 4 + 2][srow] = areg[i].z;
      As[buf][kg * 4 + 3][srow] = areg[i].w;
    }
#pragma unroll
    for (int i = 0; i < VN; ++i) {
      const int idx = tid + i * 256;
      Vs[buf][idx / RP][idx % RP] = vreg[i];
    }
  };

  if (total > 0) prefetch(0);
  for (int64_t item = 0; item < total; ++item) {
    const int buf = (int)(item & 1);
    const int c = (int)(item % nchunks);
    if (c == 0) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    }
    stage(buf);
    __syncthreads();
    if (item + 1 < total) prefetch(item + 1);
#pragma unroll
    for (int k = 0; k < KC; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&As[buf][k][(ty * 4) ^ ((k >> 2) << 2)]);
      const float4 b = *reinterpret_cast<const float4*>(&Vs[buf][k][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w};
      const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (c == nchunks - 1) {
      const int64_t row0 = (blockIdx.x + (item / nchunks) * (int64_t)gridDim.x) * BR;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int64_t grow = row0 + ty * 4 + i;
        if (grow >= rows) continue;
        float* out = C + grow * r + tx * 4;
        if ((r & 3) == 0 && tx * 4 + 3 < r) {
          __stcs(reinterpret_cast<float4*>(out), make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]));
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (tx * 4 + j < r) out[j] = acc[i][j];
        }
      }
    }
  }
}

inline bool project_f32_fast_ok(int64_t rows, int64_t n, int64_t r, const void* A, const void* C) {
  return r >= 1 && r <= 64 && n % 4 == 0 && n >= 4 && n < ((int64_t)1 << 31) && rows >= 1 &&
         (reinterpret_cast<uintptr_t>(A) & 15u) == 0 && (reinterpret_cast<uintptr_t>(C) & 15u) == 0;
}

inline int project_f32_fast(const float* A, int64_t rows, int64_t n, const float* V, int r, float* C, cudaStream_t st) {
  const int tx = r <= 32 ? 8 : 16;  // 128 x 32 or 64 x 64 output tile (static shared memory stays below 48 KB)
  const int br = (256 / tx) * 4;
  int64_t blocks = ceil_div<int64_t>(rows, br);
  const int sms = usable_sms();
  if (blocks > (int64_t)sms * 3) blocks = (int64_t)sms * 3;  // persistent: 3 resident CTAs per SM loop over the row blocks
  if (tx == 8)
    project_f32_kernel<8><<<(unsigned)blocks, 256, 0, st>>>(A, rows, (int)n, V, r, C);
  else
    project_f32_kernel<16><<<(unsigned)blocks, 256, 0, st>>>(A, rows, (int)n, V, r, C);
  TNB_LAUNCH_CHECK();
  return TNB_OK;
}

}  // namespace tnb
