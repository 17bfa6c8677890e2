// tcgen05 / TMEM / TMA GEMM and implicit-GEMM 3x3 convolution for sm_100a.
//
// One persistent, warp-specialised kernel:  warp 0 = TMA producer, warp 1 = tcgen05.mma issuer (one elected thread),
// warp 2 = TMEM allocator, warps 4..7 = epilogue.  Tiles are 128 (M) x BN (N) x 64 (K, one 128-byte swizzle row of
// fp16); accumulators are double-buffered in TMEM so the epilogue of tile i overlaps the MMAs of tile i+1.
// A 3x3 convolution is the same kernel with nine K segments: tap (dy,dx) loads the NHWC activation box shifted by
// (dy,dx) through a 4-D TMA descriptor and the out-of-bounds zero fill of TMA provides the padding; a channel concat
// is two descriptors walked back to back along K.
//
// Epilogue (v2): TMEM -> registers (thread = output row) -> fused bias / time-embedding row vector / residual / GEGLU ->
// fp16 -> 64-byte-swizzled shared-memory sub-tile (128 rows x 32 columns) -> TMA store.  The residual sub-tile arrives
// by TMA load into a second swizzled buffer.  HBM therefore only sees full-line bulk transfers (the v1 epilogue wrote
// 32-byte pieces per thread and was 4-8x off the HBM roofline for the small-K GEMMs).  Tiny-N outputs (conv_out, N=4)
// keep the direct-store path.
//
// Replaces (reference call sites): InflatedConv3d 3x3 / 1x1 (models/animatediff_models/resnet.py:9-18), every
// nn.Linear / 1x1 conv of Transformer3DModel (attention.py:65-93,174-204) and the motion module
// (motion_module.py:113-136,202-218), GEGLU (diffusers FeedForward).
#include "common.cuh"
#include "kernels.h"

namespace vs {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int A_STAGE_BYTES = BM * BK * 2;  // 16 KB
constexpr int GEMM_THREADS = 384;                   // 4 control warps + 2 epilogue warpgroups
constexpr int SMEM_LIMIT = 227 * 1024;
constexpr int EPI_COLS = 32;                        // output columns per epilogue sub-tile (64 B of fp16: SWIZZLE_64B)
constexpr int EPI_BUF_BYTES = BM * EPI_COLS * 2;    // 8 KB
constexpr int EPI_OUT_BUFS = 3, EPI_RES_BUFS = 2;   // per epilogue warpgroup
constexpr int EPI_GROUP_BYTES = (EPI_OUT_BUFS + EPI_RES_BUFS) * EPI_BUF_BYTES;
constexpr int EPI_BYTES = 2 * EPI_GROUP_BYTES;      // two warpgroups

struct GemmParams {
  CUtensorMap tmA, tmA2, tmB, tmC, tmR;
  int M, N;
  int num_kb;        // total K blocks
  int kb_per_tap;    // K blocks per tap (both concat sources)
  int kb_src1;       // of which from source 1
  int taps;
  int a_rank;        // 2 = plain rows, 4 = NHWC conv
  int nimg, H, W, TH, TW, TN, tiles_x, tiles_y;
  int m_tiles, n_tiles;
  const float* bias;
  const float* rowvec;
  int ldrv;
  int pix_per_batch;
  const __half* residual;
  int ldr;
  __half* out;
  int ldc;
  int mode;
  int tma_epi;       // 1 = smem-staged TMA-store epilogue
};

template <int BN>
struct Cfg {
  static constexpr int B_STAGE_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  static constexpr int STAGES_RAW = (SMEM_LIMIT - 2048 - EPI_BYTES) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
  static constexpr int TMEM_COLS = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
  static_assert(2 * BN <= 512, "two accumulator stages must fit TMEM");
  static_assert(B_STAGE_BYTES % 1024 == 0, "B stage must keep 1024-byte alignment for SWIZZLE_128B");
  static_assert(STAGES >= 3, "pipeline too shallow");
};

__device__ __forceinline__ uint32_t sw64_off(int row, int chunk) {   // byte offset inside a [128 x 64 B] SWIZZLE_64B tile
  return (uint32_t)(row * 64 + ((chunk ^ ((row >> 1) & 3)) << 4));
}

template <int BN>
__global__ void __launch_bounds__(GEMM_THREADS, 1) gemm_tc_kernel(const __grid_constant__ GemmParams p) {
  using C = Cfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t epi_base = smem_base + C::STAGES * C::STAGE_BYTES;       // out0 | out1 | res0 | res1
  const uint32_t bar_base = epi_base + EPI_BYTES;
  // barrier layout: full[S], empty[S], tmem_full[2], tmem_empty[2], res_full[2], then the TMEM base address word
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (C::STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * C::STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * C::STAGES + 2 + a); };
  auto rfull_bar = [&](int grp, int a) { return bar_base + 8u * (2 * C::STAGES + 4 + grp * 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * C::STAGES + 8);
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_tiles = p.m_tiles * p.n_tiles;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&p.tmA);
    prefetch_tmap(&p.tmB);
    if (p.kb_src1 < p.kb_per_tap) prefetch_tmap(&p.tmA2);
    if (p.tma_epi) {
      prefetch_tmap(&p.tmC);
      if (p.residual) prefetch_tmap(&p.tmR);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 256);
      mbar_init(rfull_bar(0, a), 1);
      mbar_init(rfull_bar(1, a), 1);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, C::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    if (lane == 0) {
      // ================================================================= TMA producer
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int m_tile = tile / p.n_tiles, n_tile = tile % p.n_tiles;
        int x0 = 0, y0 = 0, i0 = 0;
        if (p.a_rank == 4) {
          x0 = (m_tile % p.tiles_x) * p.TW;
          y0 = ((m_tile / p.tiles_x) % p.tiles_y) * p.TH;
          i0 = (m_tile / (p.tiles_x * p.tiles_y)) * p.TN;
        }
        for (int kb = 0; kb < p.num_kb; ++kb, ++it) {
          const int s = it % C::STAGES;
          const uint32_t ph = (it / C::STAGES) & 1;
          mbar_wait(empty_bar(s), ph ^ 1);
          mbar_expect_tx(full_bar(s), C::STAGE_BYTES);
          const uint32_t a_dst = smem_base + s * C::STAGE_BYTES;
          const uint32_t b_dst = a_dst + A_STAGE_BYTES;
          const int tap = kb / p.kb_per_tap;
          const int r = kb - tap * p.kb_per_tap;
          const CUtensorMap* tm = (r < p.kb_src1) ? &p.tmA : &p.tmA2;
          const int c = (r < p.kb_src1) ? r * BK : (r - p.kb_src1) * BK;
          if (p.a_rank == 4) {
            const int dy = (p.taps == 9) ? tap / 3 - 1 : 0;
            const int dx = (p.taps == 9) ? tap % 3 - 1 : 0;
            tma_load_4d(a_dst, tm, full_bar(s), c, x0 + dx, y0 + dy, i0);
          } else {
            tma_load_2d(a_dst, tm, full_bar(s), c, m_tile * BM);
          }
          tma_load_2d(b_dst, &p.tmB, full_bar(s), kb * BK, n_tile * BN);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ================================================================= MMA issuer
      constexpr uint32_t idesc = umma_idesc_f16(BM, BN);
      uint32_t it = 0, t = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++t) {
        const int acc = t & 1;
        const uint32_t aph = (t >> 1) & 1;
        mbar_wait(tempty_bar(acc), aph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < p.num_kb; ++kb, ++it) {
          const int s = it % C::STAGES;
          const uint32_t ph = (it / C::STAGES) & 1;
          mbar_wait(full_bar(s), ph);
          tc_fence_after();
          const uint32_t a_addr = smem_base + s * C::STAGE_BYTES;
          const uint32_t b_addr = a_addr + A_STAGE_BYTES;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t ad = umma_desc_sw128_kmajor(a_addr + k * 32);
            const uint64_t bd = umma_desc_sw128_kmajor(b_addr + k * 32);
            tc_mma_f16(d_tmem, ad, bd, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          tc_commit(empty_bar(s));   // frees the smem slot once these MMAs have read it
        }
        tc_commit(tfull_bar(acc));   // accumulator complete -> epilogue
      }
    }
  } else if (warp >= 4) {
    // =================================================================== epilogue
    const int q = warp & 3;                 // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;          // row of the tile handled by this thread
    const int grp = (warp - 4) >> 2;        // epilogue warpgroup: sub-tile s of a tile is handled by group s % 2
    const bool leader = (threadIdx.x & 127) == 0;
    const bool has_res = p.residual != nullptr;
    const uint32_t grp_base = epi_base + grp * EPI_GROUP_BYTES;   // out0 | out1 | out2 | res0 | res1
    uint32_t t = 0, g = 0;                  // tile counter, sub-tiles processed by this group (staging-buffer rotation)
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++t) {
      const int m_tile = tile / p.n_tiles, n_tile = tile % p.n_tiles;
      const int acc = t & 1;
      const uint32_t aph = (t >> 1) & 1;
      long long pix;
      bool valid;
      int x0 = 0, y0 = 0, i0 = 0;
      if (p.a_rank == 4) {
        x0 = (m_tile % p.tiles_x) * p.TW;
        y0 = ((m_tile / p.tiles_x) % p.tiles_y) * p.TH;
        i0 = (m_tile / (p.tiles_x * p.tiles_y)) * p.TN;
        const int per_img = p.TH * p.TW;
        const int ti = row / per_img, rem = row - ti * per_img;
        const int y = y0 + rem / p.TW, x = x0 + rem % p.TW, img = i0 + ti;
        valid = (img < p.nimg) && (y < p.H) && (x < p.W);
        pix = ((long long)img * p.H + y) * p.W + x;
      } else {
        pix = (long long)m_tile * BM + row;
        valid = pix < p.M;
      }
      const float* rv = (p.rowvec != nullptr && valid) ? p.rowvec + (pix / p.pix_per_batch) * p.ldrv : nullptr;
      const int n0 = n_tile * BN;

      mbar_wait(tfull_bar(acc), aph);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN;

      if (p.tma_epi) {
        // ------------------------------------------------------------ v2: staged, TMA-stored sub-tiles
        const bool geglu = p.mode == EPI_GEGLU;
        const int oc0 = geglu ? n_tile * (BN / 2) : n0;                    // first output column of this tile
        const int ocols = geglu ? BN / 2 : ((p.N - n0) < BN ? (p.N - n0) : BN);
        const int nsub = ocols / EPI_COLS;
        auto load_res = [&](int s, int buf) {
          mbar_expect_tx(rfull_bar(grp, buf), EPI_BUF_BYTES);
          const uint32_t dst = grp_base + (EPI_OUT_BUFS + buf) * EPI_BUF_BYTES;
          if (p.a_rank == 4) tma_load_4d(dst, &p.tmR, rfull_bar(grp, buf), oc0 + s * EPI_COLS, x0, y0, i0);
          else tma_load_2d(dst, &p.tmR, rfull_bar(grp, buf), oc0 + s * EPI_COLS, m_tile * BM);
        };
        if (leader && has_res) {
          if (grp < nsub) load_res(grp, g & 1);
          if (grp + 2 < nsub) load_res(grp + 2, (g + 1) & 1);
        }
#pragma unroll 1
        for (int s = grp; s < nsub; s += 2, ++g) {
          const int buf = g & 1;               // residual staging buffer
          float f[32];
          {
            uint32_t v[32];
            tmem_ld32(taddr + (geglu ? s * EPI_COLS : s * EPI_COLS), v);
            if (geglu) {
              uint32_t gt[32];
              tmem_ld32(taddr + BN / 2 + s * EPI_COLS, gt);
              tmem_ld_wait();
              const float* bv = p.bias + n0 + s * EPI_COLS;
              const float* bg = p.bias + n0 + BN / 2 + s * EPI_COLS;
#pragma unroll
              for (int j = 0; j < 32; ++j)
                f[j] = (__uint_as_float(v[j]) + __ldg(bv + j)) * gelu_erf_fast(__uint_as_float(gt[j]) + __ldg(bg + j));
            } else {
              tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
              const int n = n0 + s * EPI_COLS;
              if (p.bias) {
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                  const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + n + j));
                  f[j] += b.x; f[j + 1] += b.y; f[j + 2] += b.z; f[j + 3] += b.w;
                }
              }
              if (rv) {
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                  const float4 b = __ldg(reinterpret_cast<const float4*>(rv + n + j));
                  f[j] += b.x; f[j + 1] += b.y; f[j + 2] += b.z; f[j + 3] += b.w;
                }
              }
            }
          }
          if (has_res) {
            mbar_wait(rfull_bar(grp, buf), (g >> 1) & 1);
            const uint32_t rb = grp_base + (EPI_OUT_BUFS + buf) * EPI_BUF_BYTES;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              uint32_t r0, r1, r2, r3;
              asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
                           : "r"(rb + sw64_off(row, c)));
              const uint32_t rr[4] = {r0, r1, r2, r3};
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                const float2 h = __half22float2(*reinterpret_cast<const __half2*>(&rr[u]));
                f[c * 8 + 2 * u] += h.x;
                f[c * 8 + 2 * u + 1] += h.y;
              }
            }
          }
          uint32_t pk[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const __half2 h = __floats2half2_rn(f[2 * j], f[2 * j + 1]);
            pk[j] = *reinterpret_cast<const uint32_t*>(&h);
          }
          // Output staging rotates over three buffers: buffer g % 3 was last handed to a TMA store at sub-tile g - 3, and
          // the leader's wait_group.read(1) before the barrier of sub-tile g - 1 guaranteed that store had drained, so one
          // barrier per sub-tile suffices.
          const uint32_t ob = grp_base + (g % EPI_OUT_BUFS) * EPI_BUF_BYTES;
#pragma unroll
          for (int c = 0; c < 4; ++c)
            asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(ob + sw64_off(row, c)), "r"(pk[4 * c]),
                         "r"(pk[4 * c + 1]), "r"(pk[4 * c + 2]), "r"(pk[4 * c + 3]) : "memory");
          fence_proxy_async();
          if (leader) tma_store_wait_read<1>();
          named_bar_sync(1 + grp, 128);
          if (leader) {
            if (p.a_rank == 4) tma_store_4d(&p.tmC, ob, oc0 + s * EPI_COLS, x0, y0, i0);
            else tma_store_2d(&p.tmC, ob, oc0 + s * EPI_COLS, m_tile * BM);
            tma_store_commit();
            if (has_res && s + 4 < nsub) load_res(s + 4, buf);
          }
        }
      } else if (p.mode == EPI_LINEAR && grp == 0) {
        // ------------------------------------------------------------ v1: direct stores (tiny / unaligned N)
        __half* orow = p.out + pix * p.ldc;
        const __half* rrow = p.residual ? p.residual + pix * p.ldr : nullptr;
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 16) {
          if (n0 + c0 >= p.N) break;   // warp-uniform
          uint32_t v[16];
          tmem_ld16(taddr + c0, v);
          tmem_ld_wait();
          if (!valid) continue;
          const int n = n0 + c0;
          for (int j = 0; j < 16 && n + j < p.N; ++j) {
            float f = __uint_as_float(v[j]);
            if (p.bias) f += p.bias[n + j];
            if (rv) f += rv[n + j];
            if (rrow) f += __half2float(rrow[n + j]);
            orow[n + j] = __float2half_rn(f);
          }
        }
      }
      tc_fence_before();
      mbar_arrive(tempty_bar(acc));
    }
    if (leader && p.tma_epi) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

struct ConvTile { int tw, th, tn; };

ConvTile pick_conv_tile(int nimg, int H, int W) {
  ConvTile best{1, 1, 128};
  long long best_cost = -1;
  for (int tw = 1; tw <= 128; tw *= 2) {
    for (int th = 1; tw * th <= 128; th *= 2) {
      const int tn = 128 / (tw * th);
      if (tw > 256 || th > 256 || tn > 256) continue;
      const long long cost = (long long)((W + tw - 1) / tw) * tw * ((H + th - 1) / th) * th * ((nimg + tn - 1) / tn) * tn;
      if (best_cost < 0 || cost < best_cost || (cost == best_cost && tw > best.tw)) {
        best_cost = cost;
        best = ConvTile{tw, th, tn};
      }
    }
  }
  return best;
}

template <int BN>
int launch(cudaStream_t st, GemmParams& p) {
  using C = Cfg<BN>;
  static bool configured = false;
  if (!configured) {
    VS_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    configured = true;
  }
  const int total = p.m_tiles * p.n_tiles;
  int grid = total < num_sms() ? total : num_sms();
  gemm_tc_kernel<BN><<<grid, GEMM_THREADS, C::SMEM_BYTES, st>>>(p);
  VS_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// 2-D (plain rows) or 4-D (NHWC pixel box) descriptor over an output-side tensor with `cols` columns, row pitch `ld`.
int make_epi_tmap(CUtensorMap* tm, const __half* base, int cols, int ld, const GemmParams& p) {
  if (p.a_rank == 4) {
    const uint64_t dims[4] = {(uint64_t)cols, (uint64_t)p.W, (uint64_t)p.H, (uint64_t)p.nimg};
    const uint64_t str[3] = {(uint64_t)ld * 2, (uint64_t)ld * 2 * p.W, (uint64_t)ld * 2 * p.W * p.H};
    const uint32_t box[4] = {EPI_COLS, (uint32_t)p.TW, (uint32_t)p.TH, (uint32_t)p.TN};
    return make_tmap_f16(tm, base, 4, dims, str, box, 2);
  }
  const uint64_t dims[2] = {(uint64_t)cols, (uint64_t)p.M};
  const uint64_t str[1] = {(uint64_t)ld * 2};
  const uint32_t box[2] = {EPI_COLS, BM};
  return make_tmap_f16(tm, base, 2, dims, str, box, 2);
}

}  // namespace

int gemm_tc(cudaStream_t st, const GemmArgs& a) {
  VS_REQUIRE(a.A && a.Bw && a.out, "gemm_tc: null pointer");
  VS_REQUIRE(a.taps == 1 || a.taps == 9, "gemm_tc: taps must be 1 or 9");
  VS_REQUIRE(a.K1 % 8 == 0 && a.K2 % 8 == 0, "gemm_tc: K must be a multiple of 8 (TMA 16-byte strides)");
  const bool two = a.A2 != nullptr && a.K2 > 0;
  if (two || a.taps == 9) VS_REQUIRE(a.K1 % BK == 0 && a.K2 % BK == 0, "gemm_tc: concat/conv sources need C %% 64 == 0 (got %d,%d)", a.K1, a.K2);
  GemmParams p;
  memset(&p, 0, sizeof(p));
  const int Ktap = a.K1 + (two ? a.K2 : 0);
  const int Ktot = Ktap * a.taps;
  p.taps = a.taps;
  p.kb_src1 = (a.K1 + BK - 1) / BK;
  p.kb_per_tap = p.kb_src1 + (two ? a.K2 / BK : 0);
  p.num_kb = p.kb_per_tap * a.taps;
  p.M = a.M;
  p.N = a.N;
  p.bias = a.bias;
  p.rowvec = a.rowvec;
  p.ldrv = a.ldrv > 0 ? a.ldrv : a.N;
  p.pix_per_batch = a.pix_per_batch > 0 ? a.pix_per_batch : 1;
  p.residual = a.residual;
  p.ldr = a.ldr;
  p.out = a.out;
  p.ldc = a.ldc;
  p.mode = a.mode;

  int bn = a.force_bn;
  if (a.mode == EPI_GEGLU) {
    bn = 2 * kGegluGranule;
    VS_REQUIRE(a.N % bn == 0, "gemm_tc: GEGLU needs N %% %d == 0 (N=%d)", bn, a.N);
    VS_REQUIRE(a.bias != nullptr && a.residual == nullptr && a.rowvec == nullptr, "gemm_tc: GEGLU takes a bias only");
  } else if (bn == 0) {
    if (a.N <= 64) bn = 64;
    else if (a.N % 160 == 0) bn = 160;
    else bn = 128;
  }
  VS_REQUIRE(bn == 64 || bn == 128 || bn == 160, "gemm_tc: unsupported BLOCK_N %d", bn);
  p.n_tiles = (a.N + bn - 1) / bn;

  if (a.taps == 9) {
    VS_REQUIRE(a.nimg > 0 && a.H > 0 && a.W > 0 && a.M == a.nimg * a.H * a.W, "gemm_tc: bad conv geometry");
    p.a_rank = 4;
    const ConvTile t = pick_conv_tile(a.nimg, a.H, a.W);
    p.nimg = a.nimg; p.H = a.H; p.W = a.W; p.TW = t.tw; p.TH = t.th; p.TN = t.tn;
    p.tiles_x = (a.W + t.tw - 1) / t.tw;
    p.tiles_y = (a.H + t.th - 1) / t.th;
    p.m_tiles = p.tiles_x * p.tiles_y * ((a.nimg + t.tn - 1) / t.tn);
    const uint32_t box[4] = {BK, (uint32_t)t.tw, (uint32_t)t.th, (uint32_t)t.tn};
    {
      const uint64_t dims[4] = {(uint64_t)a.K1, (uint64_t)a.W, (uint64_t)a.H, (uint64_t)a.nimg};
      const uint64_t str[3] = {(uint64_t)a.lda1 * 2, (uint64_t)a.lda1 * 2 * a.W, (uint64_t)a.lda1 * 2 * a.W * a.H};
      if (make_tmap_f16(&p.tmA, a.A, 4, dims, str, box, 1)) return 3;
    }
    if (two) {
      const uint64_t dims[4] = {(uint64_t)a.K2, (uint64_t)a.W, (uint64_t)a.H, (uint64_t)a.nimg};
      const uint64_t str[3] = {(uint64_t)a.lda2 * 2, (uint64_t)a.lda2 * 2 * a.W, (uint64_t)a.lda2 * 2 * a.W * a.H};
      if (make_tmap_f16(&p.tmA2, a.A2, 4, dims, str, box, 1)) return 3;
    }
  } else {
    p.a_rank = 2;
    p.m_tiles = (a.M + BM - 1) / BM;
    const uint32_t box[2] = {BK, BM};
    {
      const uint64_t dims[2] = {(uint64_t)a.K1, (uint64_t)a.M};
      const uint64_t str[1] = {(uint64_t)a.lda1 * 2};
      if (make_tmap_f16(&p.tmA, a.A, 2, dims, str, box, 1)) return 3;
    }
    if (two) {
      const uint64_t dims[2] = {(uint64_t)a.K2, (uint64_t)a.M};
      const uint64_t str[1] = {(uint64_t)a.lda2 * 2};
      if (make_tmap_f16(&p.tmA2, a.A2, 2, dims, str, box, 1)) return 3;
    }
  }
  {
    const uint64_t dims[2] = {(uint64_t)Ktot, (uint64_t)a.N};
    const uint64_t str[1] = {(uint64_t)Ktot * 2};
    const uint32_t box[2] = {BK, (uint32_t)bn};
    if (make_tmap_f16(&p.tmB, a.Bw, 2, dims, str, box, 1)) return 3;
  }
  // staged TMA-store epilogue whenever the output geometry allows it (16-byte strides, whole 32-column sub-tiles)
  const int out_cols = (a.mode == EPI_GEGLU) ? a.N / 2 : a.N;
  p.tma_epi = (out_cols % EPI_COLS == 0) && (a.ldc % 8 == 0) && ((reinterpret_cast<uintptr_t>(a.out) & 15) == 0) &&
              (!a.residual || ((a.ldr % 8 == 0) && ((reinterpret_cast<uintptr_t>(a.residual) & 15) == 0)));
  if (p.tma_epi) {
    if (make_epi_tmap(&p.tmC, a.out, out_cols, a.ldc, p)) return 3;
    if (a.residual && make_epi_tmap(&p.tmR, a.residual, out_cols, a.ldr, p)) return 3;
  } else {
    VS_REQUIRE(a.mode == EPI_LINEAR, "gemm_tc: GEGLU output must be TMA-storable");
  }
  ProfScope prof(st, a.taps == 9 ? PC_CONV : PC_GEMM, 2.0 * a.M * (double)a.N * Ktot);
  switch (bn) {
    case 64: return launch<64>(st, p);
    case 128: return launch<128>(st, p);
    default: return launch<160>(st, p);
  }
}

}  // namespace vs
