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
constexpr int EPI_BYTES = 8 * 4096;                 // per epilogue warp: 2 KB residual + 2 KB output staging

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
  int tma_epi;       // 1 = smem-staged (warp-transposed, coalesced) epilogue; 0 = direct stores for tiny / unaligned N
  int b_resident;    // stationary mode: 0 off, 1 = weight panel [BN x K] resident, 2 = activation panel [128 x K] resident
  int cluster;       // 1, or 2 = CTA pairs sharing the weight tile through TMA multicast
  int stages;        // 0 = all, else limits the smem ring depth (pipeline-depth experiments)
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
  const uint32_t bfull_bar = bar_base + 8u * (2 * C::STAGES + 4);   // resident weight panel loaded
  const uint32_t bfree_bar = bar_base + 8u * (2 * C::STAGES + 5);   // resident weight panel no longer read
  const uint32_t tmem_slot = bar_base + 8u * (2 * C::STAGES + 6);
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_tiles = p.m_tiles * p.n_tiles;
  const int nst = (p.stages > 0 && p.stages < C::STAGES) ? p.stages : C::STAGES;   // ring depth (debug knob: "gemm_stages")
  // Tile order.  Default (stationary = 0): n fastest, tiles round-robin over CTAs.
  // Stationary modes (small K, one operand panel over the whole K extent fits in shared memory next to the ring of the
  // other operand); each CTA owns a contiguous range of tiles:
  //   1  weight panel [BN x K] resident, m fastest, only A is streamed;
  //   2  activation panel [128 x K] resident, n fastest, only B (L2-resident weights) is streamed: A is read from HBM
  //      exactly once (ncu on the K=320 QKV GEMM showed A being fetched 3x from DRAM with an L2 hit rate of 41%).
  //
  // Cluster mode (p.cluster == 2, stationary = 0): two CTAs of a cluster take the two M tiles of a tile *pair* with the
  // same N tile; each loads its own A tile and HALF of the weight tile, multicast into both CTAs' shared memory, so the
  // L2 -> SM traffic per MMA drops from 36 KB to 26 KB per k-block (ncu: the 128x160 tiles need ~31 TB/s of L2 feed at
  // full tensor rate; the conv kernel sat at 44% tensor-pipe utilisation with nothing else saturated).
  const int stat = p.b_resident;
  const bool wres = stat != 0;
  const bool cl2 = p.cluster == 2;
  const uint32_t crank = cl2 ? cluster_ctarank() : 0u;
  const int pair_tiles = ((p.m_tiles + 1) / 2) * p.n_tiles;
  const int t_begin = cl2 ? (int)(blockIdx.x >> 1) : wres ? (int)((long long)blockIdx.x * total_tiles / gridDim.x) : (int)blockIdx.x;
  const int t_end = cl2 ? pair_tiles : wres ? (int)((long long)(blockIdx.x + 1) * total_tiles / gridDim.x) : total_tiles;
  const int t_step = cl2 ? (int)(gridDim.x >> 1) : wres ? 1 : (int)gridDim.x;
  const uint32_t panel_bytes = stat == 1 ? (uint32_t)p.num_kb * C::B_STAGE_BYTES : stat == 2 ? (uint32_t)p.num_kb * A_STAGE_BYTES : 0u;
  const uint32_t ring_bytes = stat == 1 ? (uint32_t)A_STAGE_BYTES : stat == 2 ? (uint32_t)C::B_STAGE_BYTES : (uint32_t)C::STAGE_BYTES;
  auto a_stage = [&](int s, int kb) {
    return stat == 2 ? smem_base + kb * A_STAGE_BYTES : stat == 1 ? smem_base + panel_bytes + s * A_STAGE_BYTES : smem_base + s * C::STAGE_BYTES;
  };
  auto b_stage = [&](int s, int kb) {
    return stat == 1 ? smem_base + kb * C::B_STAGE_BYTES
                     : stat == 2 ? smem_base + panel_bytes + s * C::B_STAGE_BYTES : smem_base + s * C::STAGE_BYTES + A_STAGE_BYTES;
  };
  auto decode = [&](int tile, int& m_tile, int& n_tile) {
    if (stat == 1) { n_tile = tile / p.m_tiles; m_tile = tile - n_tile * p.m_tiles; }
    else { m_tile = tile / p.n_tiles; n_tile = tile - m_tile * p.n_tiles; }
    if (cl2) m_tile = 2 * m_tile + (int)crank;      // may be == m_tiles (odd count): an all-out-of-bounds dummy tile
  };
  auto panel_key = [&](int tile) { return stat == 1 ? tile / p.m_tiles : tile / p.n_tiles; };   // n-tile or m-tile id

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&p.tmA);
    prefetch_tmap(&p.tmB);
    if (p.kb_src1 < p.kb_per_tap) prefetch_tmap(&p.tmA2);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), cl2 ? 2 : 1);       // cluster mode: both CTAs' MMAs must have drained the stage
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 256);
    }
    mbar_init(bfull_bar, 1);
    mbar_init(bfree_bar, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, C::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  if (cl2) cluster_sync_all();                    // peers' barriers are initialised before any multicast / remote arrive
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  // The two single-issuer roles below run with the WHOLE warp (all values warp-uniform) and predicate only the issuing
  // instructions on one elected lane.  ncu showed that with the loops inside `if (lane == 0)` the MMA warp spent ~95% of
  // the kernel executing its own 129-instruction loop body (ELECT / VOTEU / R2UR.BROADCAST sequences re-materialising
  // the descriptors in uniform registers) and capped the tensor pipe at 44%; the loops also carry no division / modulo.
  if (warp == 0) {
    // =================================================================== TMA producer
    const int num_kb = p.num_kb, kb_per_tap = p.kb_per_tap, kb_src1 = p.kb_src1;
    const bool conv = p.a_rank == 4;
    const int tap0 = p.taps == 9 ? -1 : 0;
    uint32_t pr_s = 0, pr_ph = 0;              // ring stage / phase, carried across tiles
    for (int tile = t_begin; tile < t_end; tile += t_step) {
      int m_tile, n_tile;
      decode(tile, m_tile, n_tile);
      int x0 = 0, y0 = 0, i0 = 0;
      if (conv) {
        x0 = (m_tile % p.tiles_x) * p.TW;
        y0 = ((m_tile / p.tiles_x) % p.tiles_y) * p.TH;
        i0 = (m_tile / (p.tiles_x * p.tiles_y)) * p.TN;
      }
      const int n0 = n_tile * BN;
      const int bn_row = cl2 ? n0 + (int)crank * (BN / 2) : n0;
      const uint32_t b_half = cl2 ? crank * (C::B_STAGE_BYTES / 2) : 0u;
      const int m0 = m_tile * BM;
      int kcoord = 0;                          // K coordinate into the weight panel (kb * 64)
      int dy = tap0, dx = tap0;                // tap offsets, advanced like an odometer
      int r = 0;                               // k-block inside the tap
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(empty_bar(pr_s), pr_ph ^ 1);
        if (elect_one()) {
          const uint32_t fb = full_bar(pr_s);
          const uint32_t a_dst = smem_base + pr_s * C::STAGE_BYTES;
          mbar_expect_tx(fb, C::STAGE_BYTES);
          const bool first = r < kb_src1;
          const CUtensorMap* tm = first ? &p.tmA : &p.tmA2;
          const int c = (first ? r : r - kb_src1) * BK;
          if (conv) tma_load_4d(a_dst, tm, fb, c, x0 + dx, y0 + dy, i0);
          else tma_load_2d(a_dst, tm, fb, c, m0);
          if (cl2) tma_load_2d_mc(a_dst + A_STAGE_BYTES + b_half, &p.tmB, fb, kcoord, bn_row, (uint16_t)0x3);   // my half -> both CTAs
          else tma_load_2d(a_dst + A_STAGE_BYTES, &p.tmB, fb, kcoord, n0);
        }
        __syncwarp();
        kcoord += BK;
        if (++r == kb_per_tap) {               // next tap
          r = 0;
          if (++dx == 2) { dx = -1; ++dy; }
        }
        if (++pr_s == (uint32_t)nst) { pr_s = 0; pr_ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // =================================================================== MMA issuer
    constexpr uint32_t idesc = umma_idesc_f16(BM, BN);
    const int num_kb = p.num_kb;
    uint32_t mm_s = 0, mm_ph = 0, t = 0;
    for (int tile = t_begin; tile < t_end; tile += t_step, ++t) {
      const int acc = t & 1;
      const uint32_t aph = (t >> 1) & 1;
      mbar_wait(tempty_bar(acc), aph ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BN;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(full_bar(mm_s), mm_ph);
        tc_fence_after();
        if (elect_one()) {
          // one descriptor per operand per k-block; the K = 16 sub-steps advance the 16-byte-unit address field by 2
          const uint32_t a_addr = smem_base + mm_s * C::STAGE_BYTES;
          const uint64_t ad = umma_desc_sw128_kmajor(a_addr);
          const uint64_t bd = umma_desc_sw128_kmajor(a_addr + A_STAGE_BYTES);
          tc_mma_f16(d_tmem, ad, bd, idesc, kb != 0 ? 1u : 0u);
          tc_mma_f16(d_tmem, ad + 2, bd + 2, idesc, 1u);
          tc_mma_f16(d_tmem, ad + 4, bd + 4, idesc, 1u);
          tc_mma_f16(d_tmem, ad + 6, bd + 6, idesc, 1u);
          if (cl2) tc_commit_mc(empty_bar(mm_s), (uint16_t)0x3);   // the peer's producer also writes into this slot
          else tc_commit(empty_bar(mm_s));   // frees the smem slot once these MMAs have read it
          if (kb == num_kb - 1) tc_commit(tfull_bar(acc));   // accumulator complete -> epilogue
        }
        __syncwarp();
        if (++mm_s == (uint32_t)nst) { mm_s = 0; mm_ph ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // =================================================================== epilogue
    const int q = warp & 3;                 // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;          // row of the tile handled by this thread
    const int grp = (warp - 4) >> 2;        // epilogue warpgroup: sub-tile s of a tile is handled by group s % 2
    const bool has_res = p.residual != nullptr;
    uint32_t t = 0;                         // tile counter
    for (int tile = t_begin; tile < t_end; tile += t_step, ++t) {
      int m_tile, n_tile;
      decode(tile, m_tile, n_tile);
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
        // Warp-private staging (32 rows x 64 B for the residual, the same for the output): the accumulator layout is
        // thread = row, HBM wants lanes along columns, so each warp transposes its own 32 rows through shared memory with
        // nothing but __syncwarp.  (TMA stores were tried first: they queue behind the producer's prefetched TMA loads
        // in the same engine and their completion wait serialised the epilogue.)
        const uint32_t res_stage = epi_base + (warp - 4) * 4096, out_stage = res_stage + 2048;
        const int tr = lane >> 2, tch = lane & 3;        // transposed mapping: 8 rows x four 16-byte chunks per instruction
        long long tpix[4];
        bool tvalid[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          tpix[i] = __shfl_sync(0xffffffffu, pix, i * 8 + tr);
          tvalid[i] = __shfl_sync(0xffffffffu, (int)valid, i * 8 + tr) != 0;
        }
#pragma unroll 1
        for (int s = grp; s < nsub; s += 2) {
          const int col0 = oc0 + s * EPI_COLS;
          uint4 rres[4];
          if (has_res) {                       // coalesced residual loads, issued before the TMEM round trip
#pragma unroll
            for (int i = 0; i < 4; ++i)
              rres[i] = tvalid[i] ? __ldg(reinterpret_cast<const uint4*>(p.residual + tpix[i] * p.ldr + col0 + tch * 8))
                                  : make_uint4(0, 0, 0, 0);
          }
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
#pragma unroll
            for (int i = 0; i < 4; ++i)
              asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(res_stage + sw64_off(i * 8 + tr, tch)),
                           "r"(rres[i].x), "r"(rres[i].y), "r"(rres[i].z), "r"(rres[i].w) : "memory");
            __syncwarp();
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              uint32_t r0, r1, r2, r3;
              asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
                           : "r"(res_stage + sw64_off(lane, c)));
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
#pragma unroll
          for (int c = 0; c < 4; ++c)
            asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(out_stage + sw64_off(lane, c)), "r"(pk[4 * c]),
                         "r"(pk[4 * c + 1]), "r"(pk[4 * c + 2]), "r"(pk[4 * c + 3]) : "memory");
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 4; ++i) {       // 8 rows x 64 contiguous bytes per store instruction
            uint4 o;
            asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(o.x), "=r"(o.y), "=r"(o.z), "=r"(o.w)
                         : "r"(out_stage + sw64_off(i * 8 + tr, tch)));
            if (tvalid[i]) *reinterpret_cast<uint4*>(p.out + tpix[i] * p.ldc + col0 + tch * 8) = o;
          }
          __syncwarp();                        // staging is reused by this warp's next sub-tile
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
  }

  tc_fence_before();
  __syncthreads();
  if (cl2) cluster_sync_all();                    // nobody leaves while the peer may still signal or write into it
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
  if (p.cluster == 2) {
    const int pairs = ((p.m_tiles + 1) / 2) * p.n_tiles;
    const int max_clusters = num_sms() / 2;
    const int clusters = pairs < max_clusters ? pairs : max_clusters;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * clusters);
    cfg.blockDim = dim3(GEMM_THREADS);
    cfg.dynamicSmemBytes = C::SMEM_BYTES;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    VS_CHECK_CUDA(cudaLaunchKernelEx(&cfg, gemm_tc_kernel<BN>, p));
    return 0;
  }
  const int total = p.m_tiles * p.n_tiles;
  int grid = total < num_sms() ? total : num_sms();
  gemm_tc_kernel<BN><<<grid, GEMM_THREADS, C::SMEM_BYTES, st>>>(p);
  VS_CHECK_CUDA(cudaGetLastError());
  return 0;
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
  // stationary modes (small K) and cluster mode (everything else with >= 2 M tiles) are mutually exclusive
  // (The operand-stationary tile orders that were tried for K <= 320 -- weight panel or activation panel resident in
  // shared memory -- measured no gain on the B200 and were removed from the kernel)
  p.b_resident = 0;
  p.stages = get_option("gemm_stages");
  p.cluster = (p.b_resident == 0 && p.m_tiles >= 2 && get_option("gemm_cluster") != 0) ? 2 : 1;
  {
    const uint64_t dims[2] = {(uint64_t)Ktot, (uint64_t)a.N};
    const uint64_t str[1] = {(uint64_t)Ktot * 2};
    const uint32_t box[2] = {BK, (uint32_t)(p.cluster == 2 ? bn / 2 : bn)};   // cluster mode: each CTA fetches half
    if (make_tmap_f16(&p.tmB, a.Bw, 2, dims, str, box, 1)) return 3;
  }
  // staged TMA-store epilogue whenever the output geometry allows it (16-byte strides, whole 32-column sub-tiles)
  const int out_cols = (a.mode == EPI_GEGLU) ? a.N / 2 : a.N;
  p.tma_epi = (out_cols % EPI_COLS == 0) && (a.ldc % 8 == 0) && ((reinterpret_cast<uintptr_t>(a.out) & 15) == 0) &&
              (!a.residual || ((a.ldr % 8 == 0) && ((reinterpret_cast<uintptr_t>(a.residual) & 15) == 0)));
  if (!p.tma_epi) VS_REQUIRE(a.mode == EPI_LINEAR, "gemm_tc: GEGLU output needs 32-column aligned, 16-byte strided rows");
  // weight-stationary mode: plain GEMM, the whole K extent of the weight panel fits next to the A ring, enough M tiles
  ProfScope prof(st, a.taps == 9 ? PC_CONV : PC_GEMM, 2.0 * a.M * (double)a.N * Ktot, 1, a.M, a.N, Ktot);
  switch (bn) {
    case 64: return launch<64>(st, p);
    case 128: return launch<128>(st, p);
    default: return launch<160>(st, p);
  }
}

}  // namespace vs
