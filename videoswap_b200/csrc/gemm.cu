// tcgen05 / TMEM / TMA GEMM and implicit-GEMM 3x3 convolution for sm_100a.
//
// One persistent, warp-specialised kernel:  warp 0 = TMA producer, warp 1 = tcgen05.mma issuer, warp 2 = TMEM allocator,
// warps 4..11 = two epilogue warpgroups.  Tiles are 128 (M) x BN (N) x 64 (K, one 128-byte swizzle row of fp16);
// accumulators are double-buffered in TMEM so the epilogue of tile i overlaps the MMAs of tile i+1.
// A 3x3 convolution is the same kernel with nine K segments: tap (dy,dx) loads the NHWC activation box shifted by
// (dy,dx) through a 4-D TMA descriptor and the out-of-bounds zero fill of TMA provides the padding; a channel concat
// is two descriptors walked back to back along K.
//
// What bounds it (ncu):
//  * main loop: the L2 -> SM operand feed.  A 128 x 160 tile needs 36 KB per k-block (320 tensor cycles); every shape
//    measured -- GEMM, conv, any K -- plateaus where that feed is ~45-60 B/clk/SM, i.e. ~55% of the tensor pipe.  Wider
//    tiles need fewer bytes per flop, so BN is picked per shape by a wave-quantisation x bytes-per-k-block model
//    (BN = 256: 48 KB per 512 tensor cycles) and, for K >= 768, two CTAs of a cluster share one 256-row tile with
//    cta_group::2 MMAs (PAIR: each CTA loads half of the weight tile, 32 KB per 512 tensor cycles).
//  * K <= 640: the epilogue.  With two epilogue warps per scheduler it is bound by the serial instruction stream of each
//    warp, so the epilogue is specialised at compile time (GEGLU / residual / row vector), epilogue group g owns
//    accumulator stage g (whole tiles: the per-tile set-up is paid once), and its body is kept small.
//
// Epilogue: TMEM -> registers (thread = output row) -> [folded LayerNorm: rstd * acc - rstd * mean * u] + bias (+ warp-
// uniform time-embedding / positional row vector; both come from a per-warp shared-memory slice) / GEGLU -> fp16 -> warp-
// private swizzled shared-memory transpose -> (+ fp16 residual, loaded coalesced one sub-tile ahead) -> coalesced 16-byte
// global stores.  Tiny-N outputs (conv_out, N = 4) keep a direct-store path.
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
constexpr int EPI_WARP_BYTES = 4096;                // per epilogue warp: 2 KB transpose staging + 1 KB bias + 1 KB LN-fold slice
constexpr int EPI_BYTES = 8 * EPI_WARP_BYTES;
enum { EPI_F_GEGLU = 1, EPI_F_RES = 2, EPI_F_RV = 4, EPI_F_LN = 8, EPI_F_LNOUT = 16, EPI_F_RESST = 32 };   // compile-time epilogue features

struct GemmParams {
  CUtensorMap tmA, tmA2, tmB;
  int M, N;
  int num_kb;        // total K blocks
  int kb_per_tap;    // K blocks per tap (both concat sources)
  int kb_src1;       // of which from source 1
  int taps;
  int a_rank;        // 2 = plain rows, 4 = NHWC conv
  int nimg, H, W, TH, TW, TN, tiles_x, tiles_y;
  int tap_x0, tap_y0, tap_w;   // conv taps: dy in [tap_y0, tap_y0 + taps / tap_w), dx in [tap_x0, tap_x0 + tap_w)
  int osy, osx, ooy, oox, OH, OW;   // output pixel of input pixel (y, x): (y * osy + ooy, x * osx + oox) in an OH x OW image
  int tw_log, thw_log;   // log2(TW), log2(TH * TW): the conv tile dims are powers of two
  int m_tiles, n_tiles;
  const float* bias;
  const float* rowvec;
  int ldrv;
  int pix_per_batch;
  int rv_mod;        // row-vector index = (pix / pix_per_batch) % rv_mod when > 0 (per-frame vectors)
  const float2* ln_stats;   // folded LayerNorm: per-row (rstd, -mean * rstd)
  const float* ln_u;        // ... and per-column sum of the gamma-scaled weight row
  const float2* ln_parts;   // alternative to ln_stats: [ln_nparts][M] per-row partial (sum, sum of squares) written by the
  int ln_nparts;            // epilogue of the GEMM that PRODUCED this GEMM's A operand (ln_sums_out below)
  float ln_inv_c;           // 1 / (LayerNorm width)
  float2* ln_sums_out;      // EPI_F_LNOUT: [n_tiles][M] per-row (sum, sum of squares) of this tile's stored fp16 outputs
  const __half* residual;
  int ldr;
  __half* out;
  int ldc;
  int epi_prefetch;  // 1 = the accumulator columns of sub-tile s+1 are loaded from TMEM while sub-tile s is processed
  int res_stage;     // 1 = linear layers with a residual and K <= 320 keep the residual tile in shared memory (EPI_F_RESST)
  int res_prefetch;  // residual rows of this epilogue group's NEXT tile are prefetched into L2 while the current tile is
                     // drained: 1 = prefetch.global.L2 per 128 bytes, 2 = one cp.async.bulk.prefetch.L2 per row, 0 = off
  int staged;        // 1 = smem-transposed coalesced epilogue; 0 = direct stores for tiny / unaligned N
  int stages;        // 0 = all, else limits the smem ring depth (pipeline-depth experiments)
};

// RESST: the epilogue keeps the residual rows of a whole tile in shared memory (one private [32 rows x BN] slice per epilogue
// warp, filled with cp.async a whole tile ahead); it costs ring stages, so only tiles up to BN = 160 take it.
template <int BN, bool PAIR = false, bool RESST = false>
struct Cfg {
  static constexpr int B_STAGE_BYTES = (PAIR ? BN / 2 : BN) * BK * 2;   // a CTA pair splits the weight tile along N
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  // row stride of a residual slice: 64 B of padding when the row would be a multiple of 128 B, so that the two rows a
  // quarter-warp reads per LDS.128 phase fall into different bank halves
  static constexpr int RES_ROW_BYTES = BN * 2 + ((BN * 2) % 128 == 0 ? 64 : 0);
  static constexpr int RES_WARP_BYTES = RESST ? 32 * RES_ROW_BYTES : 0;
  static constexpr int RES_BYTES = 8 * RES_WARP_BYTES;
  static constexpr int STAGES_RAW = (SMEM_LIMIT - 2048 - EPI_BYTES - RES_BYTES) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
  static constexpr int TMEM_COLS = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BYTES + 1024 /*align slack*/ + 256 /*barriers*/ + RES_BYTES;
  static_assert(2 * BN <= 512, "two accumulator stages must fit TMEM");
  static_assert(B_STAGE_BYTES % 1024 == 0, "B stage must keep 1024-byte alignment for SWIZZLE_128B");
  static_assert(STAGES >= 3, "pipeline too shallow");
};

__host__ __device__ constexpr bool residual_staged(int BN, int EPI) {
  return (EPI & EPI_F_RESST) != 0 && (EPI & EPI_F_RES) != 0 && (EPI & EPI_F_GEGLU) == 0 && BN <= 160;
}

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint32_t sw64_off(int row, int chunk) {   // byte offset inside a [32 x 64 B] SWIZZLE_64B tile
  return (uint32_t)(row * 64 + ((chunk ^ ((row >> 1) & 3)) << 4));
}

// PAIR: two CTAs of a cluster (an SM pair) compute a 256 x BN tile with cta_group::2 MMAs: each loads its own 128 rows
// of A and HALF of the weight tile, so the L2 -> SM operand feed per flop -- the bound of the main loop -- drops by a
// third (BN = 256: 32 KB instead of 48 KB per 512 tensor cycles and CTA).
template <int BN, int EPI, bool PAIR>
__global__ void __launch_bounds__(GEMM_THREADS, 1) gemm_tc_kernel(const __grid_constant__ GemmParams p) {
  constexpr bool RESST = residual_staged(BN, EPI);
  using C = Cfg<BN, PAIR, RESST>;
  constexpr bool GEGLU = (EPI & EPI_F_GEGLU) != 0, HAS_RES = (EPI & EPI_F_RES) != 0, HAS_RV = (EPI & EPI_F_RV) != 0;
  // LN: the A operand is the RAW input of a LayerNorm whose affine map is folded into the weights:
  //   LN(x) W^T = rstd (x W'^T) - rstd mean u + c,  W' = W * gamma, u[n] = sum_k W'[n,k], c = beta W^T + bias (the `bias`)
  constexpr bool LN = (EPI & EPI_F_LN) != 0;
  // LNOUT: this GEMM's output feeds a LayerNorm (folded into the NEXT GEMM): the row statistics of what is stored -- the
  // fp16-rounded values, after the residual add -- are accumulated here, in the shadow of the store path, and written as
  // one (sum, sum of squares) pair per (column tile, row).  Deterministic (no atomics); replaces the stand-alone
  // ln_stats_kernel pass, which re-read every such tensor from HBM (4.5 GB per step).
  constexpr bool LNOUT = (EPI & EPI_F_LNOUT) != 0;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t epi_base = smem_base + C::STAGES * C::STAGE_BYTES;
  const uint32_t bar_base = epi_base + EPI_BYTES;
  // barrier layout: full[S], empty[S], tmem_full[2], tmem_empty[2], then the TMEM base address word
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (C::STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * C::STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * C::STAGES + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * C::STAGES + 4);
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t crank = PAIR ? cluster_ctarank() : 0u;        // 0 = leader (issues the MMAs)
  const int total_tiles = (PAIR ? (p.m_tiles + 1) / 2 : p.m_tiles) * p.n_tiles;   // PAIR: tiles of 256 rows
  const int nst = (p.stages > 0 && p.stages < C::STAGES) ? p.stages : C::STAGES;   // ring depth (debug knob: "gemm_stages")
  // Tile order: n fastest, tiles round-robin over the CTAs (CTAs running at the same time share the A row panel and a
  // few weight tiles in L2).  m-fastest and operand-stationary orders measured 3-40% slower.
  const int t_begin = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x, t_step = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&p.tmA);
    prefetch_tmap(&p.tmB);
    if (p.kb_src1 < p.kb_per_tap) prefetch_tmap(&p.tmA2);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), PAIR ? 256 : 128);   // one epilogue warpgroup per accumulator stage (of each CTA)
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    if (PAIR) tmem_alloc_pair(tmem_slot, C::TMEM_COLS);
    else tmem_alloc(tmem_slot, C::TMEM_COLS);
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();                  // the peer's barriers are initialised before any remote arrive / TMA credit
  tc_fence_after();
  pdl_trigger();                                 // the next kernel may start its own set-up
  pdl_wait();                                    // everything above overlapped the previous kernel's tail; its data is visible now
  const uint32_t tmem_base = *tmem_slot_ptr;

  // The two single-issuer roles below run with the WHOLE warp (all values warp-uniform) and predicate only the issuing
  // instructions on one elected lane.  ncu showed that with the loops inside `if (lane == 0)` the MMA warp spent ~95% of
  // the kernel executing its own 129-instruction loop body (ELECT / VOTEU / R2UR.BROADCAST sequences re-materialising
  // the descriptors in uniform registers) and capped the tensor pipe at 44%; the loops also carry no division / modulo.
  if (warp == 0) {
    // =================================================================== TMA producer
    const int num_kb = p.num_kb, kb_per_tap = p.kb_per_tap, kb_src1 = p.kb_src1;
    const bool conv = p.a_rank == 4;
    const int tap_x0 = p.tap_x0, tap_x1 = p.tap_x0 + p.tap_w;
    uint32_t pr_s = 0, pr_ph = 0;              // ring stage / phase, carried across tiles
    for (int tile = t_begin; tile < total_tiles; tile += t_step) {
      const int mt = tile / p.n_tiles, n_tile = tile - mt * p.n_tiles;
      const int m_tile = PAIR ? 2 * mt + (int)crank : mt;      // may be == m_tiles (odd count): an all-out-of-bounds tile
      int x0 = 0, y0 = 0, i0 = 0;
      if (conv) {
        x0 = (m_tile % p.tiles_x) * p.TW;
        y0 = ((m_tile / p.tiles_x) % p.tiles_y) * p.TH;
        i0 = (m_tile / (p.tiles_x * p.tiles_y)) * p.TN;
      }
      const int n0 = n_tile * BN + (PAIR ? (int)crank * (BN / 2) : 0);   // PAIR: this CTA's half of the weight tile
      const int m0 = m_tile * BM;
      int kcoord = 0;                          // K coordinate into the weight panel (kb * 64)
      int dy = p.tap_y0, dx = tap_x0;          // tap offsets, advanced like an odometer
      int r = 0;                               // k-block inside the tap
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(empty_bar(pr_s), pr_ph ^ 1);
        if (elect_one()) {
          const uint32_t fb = full_bar(pr_s);
          const uint32_t a_dst = smem_base + pr_s * C::STAGE_BYTES;
          const bool first = r < kb_src1;
          const CUtensorMap* tm = first ? &p.tmA : &p.tmA2;
          const int c = (first ? r : r - kb_src1) * BK;
          if (PAIR) {                          // both CTAs' bytes are credited to the LEADER's full barrier
            const uint32_t fbl = mapa_shared(fb, 0);
            if (crank == 0) mbar_expect_tx(fb, 2 * C::STAGE_BYTES);
            if (conv) tma_load_4d_pair(a_dst, tm, fbl, c, x0 + dx, y0 + dy, i0);
            else tma_load_2d_pair(a_dst, tm, fbl, c, m0);
            tma_load_2d_pair(a_dst + A_STAGE_BYTES, &p.tmB, fbl, kcoord, n0);
          } else {
            mbar_expect_tx(fb, C::STAGE_BYTES);
            if (conv) tma_load_4d(a_dst, tm, fb, c, x0 + dx, y0 + dy, i0);
            else tma_load_2d(a_dst, tm, fb, c, m0);
            tma_load_2d(a_dst + A_STAGE_BYTES, &p.tmB, fb, kcoord, n0);
          }
        }
        __syncwarp();
        kcoord += BK;
        if (++r == kb_per_tap) {               // next tap
          r = 0;
          if (++dx == tap_x1) { dx = tap_x0; ++dy; }
        }
        if (++pr_s == (uint32_t)nst) { pr_s = 0; pr_ph ^= 1; }
      }
    }
  } else if (warp == 1 && crank == 0) {
    // =================================================================== MMA issuer (PAIR: the leader CTA only)
    constexpr uint32_t idesc = umma_idesc_f16(PAIR ? 2 * BM : BM, BN);
    const int num_kb = p.num_kb;
    uint32_t mm_s = 0, mm_ph = 0, t = 0;
    for (int tile = t_begin; tile < total_tiles; tile += t_step, ++t) {
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
          if (PAIR) {
            tc_mma_f16_pair(d_tmem, ad, bd, idesc, kb != 0 ? 1u : 0u);
            tc_mma_f16_pair(d_tmem, ad + 2, bd + 2, idesc, 1u);
            tc_mma_f16_pair(d_tmem, ad + 4, bd + 4, idesc, 1u);
            tc_mma_f16_pair(d_tmem, ad + 6, bd + 6, idesc, 1u);
            tc_commit_pair(empty_bar(mm_s), (uint16_t)0x3);   // frees the slot in BOTH CTAs
            if (kb == num_kb - 1) tc_commit_pair(tfull_bar(acc), (uint16_t)0x3);
          } else {
            tc_mma_f16(d_tmem, ad, bd, idesc, kb != 0 ? 1u : 0u);
            tc_mma_f16(d_tmem, ad + 2, bd + 2, idesc, 1u);
            tc_mma_f16(d_tmem, ad + 4, bd + 4, idesc, 1u);
            tc_mma_f16(d_tmem, ad + 6, bd + 6, idesc, 1u);
            tc_commit(empty_bar(mm_s));        // frees the smem slot once these MMAs have read it
            if (kb == num_kb - 1) tc_commit(tfull_bar(acc));   // accumulator complete -> epilogue
          }
        }
        __syncwarp();
        if (++mm_s == (uint32_t)nst) { mm_s = 0; mm_ph ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // =================================================================== epilogue
    // Warpgroup g drains accumulator stage g, i.e. every second tile of this CTA.
    const int q = warp & 3;                 // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;          // row of the tile handled by this thread
    const int grp = (warp - 4) >> 2;
    const int acc = grp;
    const uint32_t out_stage = epi_base + (warp - 4) * EPI_WARP_BYTES, bias_stage = out_stage + 2048, u_stage = bias_stage + 1024;
    const int tr = lane >> 2, tch = lane & 3;   // transposed mapping: 8 rows x four 16-byte chunks per instruction
    const bool has_bias = p.bias != nullptr;
    const bool staged = p.staged != 0;
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN;
    uint32_t aph = 0;
    const uint32_t tempty_arrive = PAIR ? mapa_shared(tempty_bar(acc), 0) : tempty_bar(acc);   // PAIR: the leader's barrier
    // output pixel (row of the GEMM) this thread owns in `tile`, -1 = outside the problem
    auto tile_pixel = [&](int tile, int& n_tile) -> int {
      const int mt = tile / p.n_tiles;
      n_tile = tile - mt * p.n_tiles;
      const int m_tile = PAIR ? 2 * mt + (int)crank : mt;
      if (p.a_rank == 4) {
        const int x0 = (m_tile % p.tiles_x) * p.TW;
        const int y0 = ((m_tile / p.tiles_x) % p.tiles_y) * p.TH;
        const int i0 = (m_tile / (p.tiles_x * p.tiles_y)) * p.TN;
        const int ti = row >> p.thw_log, rem = row & ((1 << p.thw_log) - 1);
        const int y = y0 + (rem >> p.tw_log), x = x0 + (rem & (p.TW - 1)), img = i0 + ti;
        return ((m_tile < p.m_tiles) && (img < p.nimg) && (y < p.H) && (x < p.W))
                   ? (img * p.OH + y * p.osy + p.ooy) * p.OW + x * p.osx + p.oox : -1;
      }
      const int px = m_tile * BM + row;
      return px < p.M ? px : -1;
    };
    // The residual add made the K <= 640 epilogues DRAM-latency-bound: each 32-column sub-tile waited for residual rows
    // requested only one sub-tile earlier (ncu: the top stall of the epilogue warps was the long scoreboard on those loads,
    // DRAM at 46 %, tensor pipe at 22 %).  The rows of the group's NEXT tile are therefore pulled into L2 a whole tile ahead.
    auto prefetch_residual = [&](int tile) {
      if (tile >= total_tiles) return;
      int nt;
      const int px = tile_pixel(tile, nt);
      if (px < 0) return;
      const int n0p = nt * BN;
      const int bytes = (p.N - n0p < BN ? p.N - n0p : BN) * 2;
      const char* src = reinterpret_cast<const char*>(p.residual + (long long)px * p.ldr + n0p);
      if (p.res_prefetch == 2) {
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
      } else {
        for (int o = 0; o < bytes; o += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(src + o));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(src + bytes - 1));
      }
    };
    const bool res_pf = HAS_RES && !GEGLU && !RESST && staged && p.res_prefetch != 0;
    if (res_pf) prefetch_residual(t_begin + grp * t_step);
    // RESST -- whole-tile residual stage.  With the residual rows requested one 32-column sub-tile ahead (register double
    // buffer, below) every epilogue warp sat out one full memory latency per sub-tile: 5 x ~1.2 us + work = the 8.6 us a
    // BN = 160 tile took (64 us vs 42 us without the residual at M = 131072, N = K = 320; fewer instructions or an L2
    // prefetch did not move it; this stage brings it to 58 us -- it costs ring stages, see launch_linear for where it is used).  Now each lane owns the sixteen-byte
    // chunks it will add after the transpose -- rows i*8 + tr, chunk tch of every sub-tile -- in a private shared-memory
    // slice: it copies them in with cp.async, one commit group per sub-tile, and refills a sub-tile's chunks with the NEXT
    // tile's rows right after it has consumed them.  A chunk is written and read by the same lane only, so no barrier is
    // involved: program order covers write-after-read, cp.async.wait_group covers read-after-write.  Every tile commits
    // exactly MAXSUB groups (empty ones for missing sub-tiles or a missing next tile), so when sub-tile s is consumed the
    // groups younger than its own are always MAXSUB - 1: (MAXSUB - 1 - s) of this tile + s refills.
    constexpr int MAXSUB_T = (GEGLU ? BN / 2 : BN) / EPI_COLS;
    const uint32_t res_lane = bar_base + 256u + (uint32_t)((warp - 4) * C::RES_WARP_BYTES + tr * C::RES_ROW_BYTES + tch * 16);
    const __half* rnext[4] = {nullptr, nullptr, nullptr, nullptr};   // this lane's 4 residual rows of the group's next tile
    int nsub_next = 0;
    auto residual_rows = [&](int tile) {
      nsub_next = 0;
      if (tile < total_tiles) {               // warp-uniform
        int nt;
        const int px = tile_pixel(tile, nt);
        const int n0p = nt * BN;
        nsub_next = (p.N - n0p < BN ? p.N - n0p : BN) / EPI_COLS;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int tp = __shfl_sync(0xffffffffu, px, i * 8 + tr);
          rnext[i] = tp >= 0 ? p.residual + (long long)tp * p.ldr + n0p + tch * 8 : nullptr;
        }
      }
    };
    auto stage_residual = [&](int s) {       // this lane's chunks of sub-tile s of the next tile; always exactly one group
      if (s < nsub_next) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (rnext[i] != nullptr) cp_async16(res_lane + (uint32_t)(i * 8 * C::RES_ROW_BYTES + s * 64), rnext[i] + s * EPI_COLS);
      }
      cp_async_commit();
    };
    if (RESST && staged) {
      residual_rows(t_begin + grp * t_step);
#pragma unroll 1
      for (int s = 0; s < MAXSUB_T; ++s) stage_residual(s);
    }
    for (int tile = t_begin + grp * t_step; tile < total_tiles; tile += 2 * t_step, aph ^= 1) {
      int n_tile;
      const int pix = tile_pixel(tile, n_tile);
      if (res_pf) prefetch_residual(tile + 2 * t_step);
      if (RESST && staged) residual_rows(tile + 2 * t_step);
      const int n0 = n_tile * BN;

      if (staged) {
        constexpr int OC = GEGLU ? BN / 2 : BN;                            // output columns of a full tile
        const int oc0 = n_tile * OC;                                        // first output column of this tile
        const int nsub = (p.N - n0 < BN ? (GEGLU ? (p.N - n0) / 2 : p.N - n0) : OC) / EPI_COLS;
        const float* rv = nullptr;
        // Row vectors (time embedding per batch element, positional offsets per frame): the 32 rows of a warp nearly always
        // share one vector, which is then folded into the warp's bias slice; mixed warps read theirs per sub-tile.
        bool rv_uniform = false;
        const float* rv_warp = nullptr;
        if (HAS_RV) {
          int ri = -1;
          if (pix >= 0) {
            ri = pix / p.pix_per_batch;
            if (p.rv_mod > 0) ri %= p.rv_mod;
            rv = p.rowvec + (long long)ri * p.ldrv + n0;
          }
          const unsigned live = __ballot_sync(0xffffffffu, pix >= 0);
          if (live != 0u) {
            const int r0 = __shfl_sync(0xffffffffu, ri, __ffs(live) - 1);
            rv_uniform = __all_sync(0xffffffffu, pix < 0 || ri == r0);
            rv_warp = p.rowvec + (long long)r0 * p.ldrv;
          }
        }
        float la = 1.f, lb = 0.f;           // folded LayerNorm: row scale and row shift factor
        if (LN && pix >= 0) {
          if (p.ln_nparts > 0) {            // statistics from the producer GEMM's per-column-tile partial sums
            float S = 0.f, Q = 0.f;
            for (int k = 0; k < p.ln_nparts; ++k) {
              const float2 v = __ldg(p.ln_parts + (long long)k * p.M + pix);
              S += v.x; Q += v.y;
            }
            const float mean = S * p.ln_inv_c;
            la = rsqrtf(fmaxf(fmaf(-mean, mean, Q * p.ln_inv_c), 0.f) + 1e-5f);
            lb = -mean * la;
          } else {
            const float2 st2 = __ldg(p.ln_stats + pix);
            la = st2.x; lb = st2.y;
          }
        }
        float rs[4] = {0.f, 0.f, 0.f, 0.f}, rq[4] = {0.f, 0.f, 0.f, 0.f};   // LNOUT: sums of the 4 rows this lane stores
        int tpix[4];                        // pixels of the rows this lane stores after the transpose
#pragma unroll
        for (int i = 0; i < 4; ++i) tpix[i] = __shfl_sync(0xffffffffu, pix, i * 8 + tr);
        // Row pointers of the 4 rows this lane stores, computed once per tile: the sub-tile loop below is fully unrolled for
        // the linear epilogues, so every access becomes [pointer + immediate].  (Computed per access inside a rolled loop --
        // 64-bit multiplies, constant-bank reloads after every asm volatile, one branch per row, register copies for the
        // residual double buffer -- the sub-tile body was ~300 instructions per warp; the diet to ~210 bought 2 % of the GEMM
        // time, the residual latency above was the larger term: profiles/r02_ncu_proj.json, r02_epilogue_shapes_ab.json.)
        // Rows outside the problem point at row 0 (loads are harmless, stores are predicated).
        __half* orow[4];
        const __half* rrow[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const long long px = tpix[i] >= 0 ? tpix[i] : 0;
          orow[i] = p.out + px * p.ldc + oc0 + tch * 8;
          rrow[i] = (HAS_RES && !RESST) ? p.residual + px * p.ldr + oc0 + tch * 8 : nullptr;
        }
        // Residual rows are requested one sub-tile ahead of their use into the OTHER of two register sets (the sub-tile loop is
        // unrolled by two so the sets alternate without copies).
        uint4 res_a[4], res_b[4];
        auto load_res = [&](int s, uint4 (&dst)[4]) {
#pragma unroll
          for (int i = 0; i < 4; ++i) dst[i] = __ldg(reinterpret_cast<const uint4*>(rrow[i] + s * EPI_COLS));
        };
        if (HAS_RES && !RESST) load_res(0, res_a);
        // The tile's bias vector goes to this warp's own shared-memory slice while the accumulator is still being
        // computed; the sub-tiles then read it as broadcast LDS.  (Read with __ldg inside the sub-tile loop, the epilogue
        // warps of the K = 320 GEMMs spent 31% of their time waiting for those loads.)
        const bool stage_bias = has_bias || (HAS_RV && rv_uniform);
        if (stage_bias) {
          __syncwarp();                     // the previous tile's readers are done
#pragma unroll
          for (int c = lane; c < BN / 4; c += 32) {
            const int n = n0 + 4 * c;
            float4 b = (has_bias && n + 3 < p.N) ? __ldg(reinterpret_cast<const float4*>(p.bias + n)) : make_float4(0.f, 0.f, 0.f, 0.f);
            if (HAS_RV && rv_uniform && n + 3 < p.N) {
              const float4 r4 = __ldg(reinterpret_cast<const float4*>(rv_warp + n));
              b.x += r4.x; b.y += r4.y; b.z += r4.z; b.w += r4.w;
            }
            asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(bias_stage + 16 * c), "f"(b.x), "f"(b.y), "f"(b.z), "f"(b.w) : "memory");
            if (LN) {
              const float4 u4 = n + 3 < p.N ? __ldg(reinterpret_cast<const float4*>(p.ln_u + n)) : make_float4(0.f, 0.f, 0.f, 0.f);
              asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(u_stage + 16 * c), "f"(u4.x), "f"(u4.y), "f"(u4.z), "f"(u4.w) : "memory");
            }
          }
          __syncwarp();
        }
        mbar_wait(tfull_bar(acc), aph);
        tc_fence_after();
        // Linear epilogues prefetch the next 32 accumulator columns (tcgen05.ld is asynchronous until tcgen05.wait::ld), so the
        // TMEM round trip of sub-tile s+1 overlaps the bias / transpose / store work of sub-tile s.
        uint32_t vn[32];
        const bool prefetch = !GEGLU && p.epi_prefetch != 0;
        if (prefetch) tmem_ld32(taddr, vn);
        auto subtile = [&](const int s, uint4 (&rcur)[4], uint4 (&rnxt)[4]) {
          float f[32];
          if (prefetch) {
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(vn[j]);
            if (s + 1 < nsub) tmem_ld32(taddr + (s + 1) * EPI_COLS, vn);
          } else {
            uint32_t v[32];
            tmem_ld32(taddr + s * EPI_COLS, v);
            if (GEGLU) {
              uint32_t g[32];
              tmem_ld32(taddr + BN / 2 + s * EPI_COLS, g);
              tmem_ld_wait();
              const uint32_t bv = bias_stage + s * EPI_COLS * 4, bg = bv + (BN / 2) * 4;
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                float4 b = lds_f4(bv + j * 4);
                float4 c = lds_f4(bg + j * 4);
                if (LN) {                        // value / gate pre-activations of the folded LayerNorm
                  const float4 uv = lds_f4(bv + 1024 + j * 4), ug = lds_f4(bg + 1024 + j * 4);
                  b.x = fmaf(lb, uv.x, b.x); b.y = fmaf(lb, uv.y, b.y); b.z = fmaf(lb, uv.z, b.z); b.w = fmaf(lb, uv.w, b.w);
                  c.x = fmaf(lb, ug.x, c.x); c.y = fmaf(lb, ug.y, c.y); c.z = fmaf(lb, ug.z, c.z); c.w = fmaf(lb, ug.w, c.w);
#pragma unroll
                  for (int t = 0; t < 4; ++t) {
                    v[j + t] = __float_as_uint(__uint_as_float(v[j + t]) * la);
                    g[j + t] = __float_as_uint(__uint_as_float(g[j + t]) * la);
                  }
                }
                f[j] = (__uint_as_float(v[j]) + b.x) * gelu_sig(__uint_as_float(g[j]) + c.x);
                f[j + 1] = (__uint_as_float(v[j + 1]) + b.y) * gelu_sig(__uint_as_float(g[j + 1]) + c.y);
                f[j + 2] = (__uint_as_float(v[j + 2]) + b.z) * gelu_sig(__uint_as_float(g[j + 2]) + c.z);
                f[j + 3] = (__uint_as_float(v[j + 3]) + b.w) * gelu_sig(__uint_as_float(g[j + 3]) + c.w);
              }
            } else {
              tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
            }
          }
          if (s == nsub - 1) {              // accumulator fully read: hand the stage back to the MMA warp before the stores
            tc_fence_before();
            if (PAIR) mbar_arrive_cluster(tempty_arrive);
            else mbar_arrive(tempty_bar(acc));
          }
          if (HAS_RES && !RESST && s + 1 < nsub) load_res(s + 1, rnxt);
          if (!GEGLU) {
            if (LN) {                         // rstd * acc + (-mean rstd) * u + c in two FMAs per element
              const uint32_t up = u_stage + s * EPI_COLS * 4, bp = bias_stage + s * EPI_COLS * 4;
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                const float4 u4 = lds_f4(up + j * 4), b = lds_f4(bp + j * 4);
                f[j] = fmaf(la, f[j], fmaf(lb, u4.x, b.x)); f[j + 1] = fmaf(la, f[j + 1], fmaf(lb, u4.y, b.y));
                f[j + 2] = fmaf(la, f[j + 2], fmaf(lb, u4.z, b.z)); f[j + 3] = fmaf(la, f[j + 3], fmaf(lb, u4.w, b.w));
              }
            } else if (stage_bias) {
              const uint32_t bp = bias_stage + s * EPI_COLS * 4;
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                const float4 b = lds_f4(bp + j * 4);
                f[j] += b.x; f[j + 1] += b.y; f[j + 2] += b.z; f[j + 3] += b.w;
              }
            }
            if (HAS_RV && rv && !rv_uniform) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                const float4 b = __ldg(reinterpret_cast<const float4*>(rv + s * EPI_COLS + j));
                f[j] += b.x; f[j + 1] += b.y; f[j + 2] += b.z; f[j + 3] += b.w;
              }
            }
          }
          // thread = row -> lanes along columns: transpose this warp's 32 x 32 block through its own staging buffer.
          // TMA stores were tried first: they queue behind the producer's prefetched loads in the same engine and their
          // completion wait serialised the epilogue.
          __syncwarp();                       // the previous sub-tile's readers of the staging buffer are done
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            uint32_t pk[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const __half2 h = __floats2half2_rn(f[c * 8 + 2 * u], f[c * 8 + 2 * u + 1]);
              pk[u] = *reinterpret_cast<const uint32_t*>(&h);
            }
            asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(out_stage + sw64_off(lane, c)), "r"(pk[0]), "r"(pk[1]),
                         "r"(pk[2]), "r"(pk[3]) : "memory");
          }
          __syncwarp();
          if (RESST) cp_async_wait<MAXSUB_T - 1>();   // this lane's residual chunks of sub-tile s have landed
#pragma unroll
          for (int i = 0; i < 4; ++i) {       // 8 rows x 64 contiguous bytes per store instruction
            uint4 o;
            asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(o.x), "=r"(o.y), "=r"(o.z), "=r"(o.w)
                         : "r"(out_stage + sw64_off(i * 8 + tr, tch)));
            if (HAS_RES) {                    // fp16 add: the rounding order of the reference's `linear(x) + residual`
              __half2* oh = reinterpret_cast<__half2*>(&o);
              uint4 rr;
              if (RESST) {
                asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(rr.x), "=r"(rr.y), "=r"(rr.z), "=r"(rr.w)
                             : "r"(res_lane + (uint32_t)(i * 8 * C::RES_ROW_BYTES + s * 64)));
              } else {
                rr = rcur[i];
              }
              const __half2* rh = reinterpret_cast<const __half2*>(&rr);
#pragma unroll
              for (int u = 0; u < 4; ++u) oh[u] = __hadd2(oh[u], rh[u]);
            }
            if (tpix[i] >= 0) *reinterpret_cast<uint4*>(orow[i] + s * EPI_COLS) = o;
            if (LNOUT) {                      // off the store's critical path: 8 values of row i*8+tr
              const __half2* oh = reinterpret_cast<const __half2*>(&o);
              const float2 a = __half22float2(oh[0]), b = __half22float2(oh[1]), c = __half22float2(oh[2]), d = __half22float2(oh[3]);
              rs[i] += ((a.x + a.y) + (b.x + b.y)) + ((c.x + c.y) + (d.x + d.y));
              float q0 = fmaf(a.x, a.x, a.y * a.y), q1 = fmaf(b.x, b.x, b.y * b.y);
              float q2 = fmaf(c.x, c.x, c.y * c.y), q3 = fmaf(d.x, d.x, d.y * d.y);
              rq[i] += (q0 + q1) + (q2 + q3);
            }
          }
          if (RESST) stage_residual(s);       // the chunks just consumed are refilled with the next tile's rows
        };
#pragma unroll 1
        for (int s = 0; s < nsub; s += 2) {
          subtile(s, res_a, res_b);
          if (s + 1 < nsub) subtile(s + 1, res_b, res_a);
        }
        if (RESST) {
#pragma unroll 1
          for (int s = nsub; s < MAXSUB_T; ++s) stage_residual(s);   // keep the group count per tile constant
        }
        if (LNOUT) {                          // the 4 column chunks of a row sit in 4 adjacent lanes
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            rs[i] += __shfl_xor_sync(0xffffffffu, rs[i], 1); rq[i] += __shfl_xor_sync(0xffffffffu, rq[i], 1);
            rs[i] += __shfl_xor_sync(0xffffffffu, rs[i], 2); rq[i] += __shfl_xor_sync(0xffffffffu, rq[i], 2);
            if (tch == 0 && tpix[i] >= 0) p.ln_sums_out[(long long)n_tile * p.M + tpix[i]] = make_float2(rs[i], rq[i]);
          }
        }
      } else {
        // ------------------------------------------------------------ direct stores (tiny / unaligned N, linear only)
        mbar_wait(tfull_bar(acc), aph);
        tc_fence_after();
        const float* rv = (p.rowvec != nullptr && pix >= 0) ? p.rowvec + (long long)(pix / p.pix_per_batch) * p.ldrv : nullptr;
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 16) {
          if (n0 + c0 >= p.N) break;   // warp-uniform
          uint32_t v[16];
          tmem_ld16(taddr + c0, v);
          tmem_ld_wait();
          if (pix < 0) continue;
          __half* orow = p.out + (long long)pix * p.ldc;
          const __half* rrow = p.residual ? p.residual + (long long)pix * p.ldr : nullptr;
          const int n = n0 + c0;
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            if (n + j < p.N) {
              float f = __uint_as_float(v[j]);
              if (has_bias) f += p.bias[n + j];
              if (rv) f += rv[n + j];
              if (rrow) f += __half2float(rrow[n + j]);
              orow[n + j] = __float2half_rn(f);
            }
          }
        }
        tc_fence_before();
        if (PAIR) mbar_arrive_cluster(tempty_arrive);
        else mbar_arrive(tempty_bar(acc));
      }
    }
    if (RESST) cp_async_wait<0>();             // only empty groups can be left; nothing in flight when the CTA exits
  }

  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();                  // nobody leaves while its peer may still signal it or read its smem
  if (warp == 2) {
    tc_fence_after();
    if (PAIR) tmem_dealloc_pair(tmem_base, C::TMEM_COLS);
    else tmem_dealloc(tmem_base, C::TMEM_COLS);
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

template <int BN, int EPI, bool PAIR>
int launch(cudaStream_t st, const GemmParams& p) {
  using C = Cfg<BN, PAIR, residual_staged(BN, EPI)>;
  static bool configured = false;
  if (!configured) {
    VS_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN, EPI, PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    configured = true;
  }
  if (PAIR) {
    const int pairs = ((p.m_tiles + 1) / 2) * p.n_tiles;
    const int max_clusters = num_sms() / 2;
    return launch_pdl(gemm_tc_kernel<BN, EPI, PAIR>, dim3(2 * (pairs < max_clusters ? pairs : max_clusters)), dim3(GEMM_THREADS),
                      C::SMEM_BYTES, st, 2, p);
  }
  const int total = p.m_tiles * p.n_tiles;
  return launch_pdl(gemm_tc_kernel<BN, EPI, PAIR>, dim3(total < num_sms() ? total : num_sms()), dim3(GEMM_THREADS), C::SMEM_BYTES,
                    st, 1, p);
}

template <int BN, bool PAIR>
int launch_linear(cudaStream_t st, const GemmParams& p) {
  int epi = (p.residual ? EPI_F_RES : 0) | (p.rowvec ? EPI_F_RV : 0) | ((p.ln_stats || p.ln_parts) ? EPI_F_LN : 0) |
            (p.ln_sums_out ? EPI_F_LNOUT : 0);
  // Whole-tile residual stage in shared memory: pays only where the epilogue, not the main loop, sets the tile time AND a
  // 3-stage operand ring is deep enough -- measured: K = 320 (5 k-blocks) at M = 131072 -8 %; K = 640 +7 %, K = 1280 +20 %,
  // CTA pairs +15 % (profiles/r02_resst_shapes.json).  So: single-CTA linear layers with K <= 320 only.
  if constexpr (BN <= 160 && !PAIR) {
    if ((epi == EPI_F_RES || epi == (EPI_F_LNOUT | EPI_F_RES)) && p.res_stage != 0 && p.staged != 0 && p.a_rank != 4 && p.num_kb <= 5)
      epi |= EPI_F_RESST;
    switch (epi) {
      case EPI_F_RES | EPI_F_RESST: return launch<BN, EPI_F_RES | EPI_F_RESST, PAIR>(st, p);
      case EPI_F_LNOUT | EPI_F_RES | EPI_F_RESST: return launch<BN, EPI_F_LNOUT | EPI_F_RES | EPI_F_RESST, PAIR>(st, p);
      default: break;
    }
  }
  switch (epi) {
    case EPI_F_LNOUT: return launch<BN, EPI_F_LNOUT, PAIR>(st, p);
    case EPI_F_LNOUT | EPI_F_RES: return launch<BN, EPI_F_LNOUT | EPI_F_RES, PAIR>(st, p);
    case EPI_F_LN: return launch<BN, EPI_F_LN, PAIR>(st, p);
    case EPI_F_LN | EPI_F_RV: return launch<BN, EPI_F_LN | EPI_F_RV, PAIR>(st, p);
    case 0: return launch<BN, 0, PAIR>(st, p);
    case EPI_F_RES: return launch<BN, EPI_F_RES, PAIR>(st, p);
    case EPI_F_RV: return launch<BN, EPI_F_RV, PAIR>(st, p);
    case EPI_F_RES | EPI_F_RV: return launch<BN, EPI_F_RES | EPI_F_RV, PAIR>(st, p);
    default: set_error("gemm_tc: unsupported epilogue combination 0x%x", epi); return 2;
  }
}

// BLOCK_N by a two-term model: waves of the persistent grid x L2 -> SM bytes per k-block of one tile (the operand feed,
// not the tensor pipe, bounds the main loop; see the file header).  Ties go to the wider tile only for long K, where
// the main loop -- not the epilogue -- dominates.
int pick_bn(int m_tiles, int N, int num_kb) {
  if (N <= 64) return 64;
  int best = 128;
  long long best_cost = -1;
  const int cand[3] = {128, 160, 256};
  for (int i = 0; i < 3; ++i) {
    const int bn = cand[i];
    if (bn != 128 && N % bn != 0) continue;
    const long long tiles = (long long)m_tiles * ((N + bn - 1) / bn);
    const long long waves = (tiles + num_sms() - 1) / num_sms();
    const long long cost = waves * (16 + bn / 8);           // KB per k-block: 16 (A) + bn * 128 B (B)
    if (best_cost < 0 || cost < best_cost || (cost == best_cost && num_kb >= 40)) {
      best_cost = cost;
      best = bn;
    }
  }
  return best;
}

int ilog2(int v) { int l = 0; while ((1 << l) < v) ++l; return l; }

}  // namespace

// Column tiles the kernel will use for this problem (= number of LayerNorm partial-sum slices a producer writes).
int gemm_n_tiles(const GemmArgs& a) {
  if (a.taps != 1) return 0;
  int bn = a.force_bn;
  if (a.mode == EPI_GEGLU) bn = 2 * kGegluGranule;
  else if (bn == 0) bn = pick_bn((a.M + BM - 1) / BM, a.N, (a.K1 + BK - 1) / BK + a.K2 / BK);
  return (a.N + bn - 1) / bn;
}

int gemm_tc(cudaStream_t st, const GemmArgs& a) {
  VS_REQUIRE(a.A && a.Bw && a.out, "gemm_tc: null pointer");
  VS_REQUIRE(a.taps == 1 || a.taps == 9 || a.taps == 4, "gemm_tc: taps must be 1, 9 (3x3) or 4 (2x2 sub-pixel)");
  VS_REQUIRE(a.K1 % 8 == 0 && a.K2 % 8 == 0, "gemm_tc: K must be a multiple of 8 (TMA 16-byte strides)");
  VS_REQUIRE((long long)a.M * (a.ldc > a.ldr ? a.ldc : a.ldr) < (1LL << 40) && a.M < (1 << 30), "gemm_tc: M too large");
  const bool two = a.A2 != nullptr && a.K2 > 0;
  if (two || a.taps != 1) VS_REQUIRE(a.K1 % BK == 0 && a.K2 % BK == 0, "gemm_tc: concat/conv sources need C %% 64 == 0 (got %d,%d)", a.K1, a.K2);
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
  p.rv_mod = a.rv_mod;
  p.ln_stats = reinterpret_cast<const float2*>(a.ln_stats);
  p.ln_u = a.ln_u;
  p.ln_parts = reinterpret_cast<const float2*>(a.ln_parts);
  p.ln_nparts = a.ln_parts ? a.ln_nparts : 0;
  p.ln_inv_c = 1.f / (float)(a.K1 + a.K2);
  p.ln_sums_out = reinterpret_cast<float2*>(a.ln_sums_out);
  if (a.ln_parts) VS_REQUIRE(a.ln_stats == nullptr && a.ln_nparts >= 1 && a.ln_nparts <= 16 && a.taps == 1, "gemm_tc: bad LayerNorm partial-sum input");
  if (a.ln_stats || a.ln_parts) VS_REQUIRE(a.ln_u != nullptr && a.bias != nullptr && a.residual == nullptr, "gemm_tc: folded LayerNorm needs u and c vectors and no residual");
  if (a.ln_sums_out) VS_REQUIRE(a.taps == 1 && a.mode == EPI_LINEAR && a.rowvec == nullptr && !a.ln_stats && !a.ln_parts,
                                "gemm_tc: row statistics output is implemented for plain linear layers (+ bias / residual)");
  p.residual = a.residual;
  p.ldr = a.ldr;
  p.out = a.out;
  p.ldc = a.ldc;
  p.stages = get_option("gemm_stages");
  p.epi_prefetch = get_option("epi_prefetch");
  p.res_prefetch = get_option("res_prefetch");
  p.res_stage = get_option("res_stage");

  if (a.taps != 1) {
    VS_REQUIRE(a.nimg > 0 && a.H > 0 && a.W > 0 && a.M == a.nimg * a.H * a.W, "gemm_tc: bad conv geometry");
    p.a_rank = 4;
    // 3x3: taps (-1..1)^2.  2x2 (one output parity (py, px) of nearest-2x + 3x3, see pack_conv_subpixel): taps
    // {py - 1, py} x {px - 1, px}, output pixel (2 y + py, 2 x + px) of the 2H x 2W image.
    if (a.taps == 9) { p.tap_x0 = p.tap_y0 = -1; p.tap_w = 3; p.osy = p.osx = 1; p.ooy = p.oox = 0; p.OH = a.H; p.OW = a.W; }
    else {
      VS_REQUIRE((a.sub_py | 1) == 1 && (a.sub_px | 1) == 1, "gemm_tc: sub-pixel parity must be 0 or 1");
      p.tap_y0 = a.sub_py - 1; p.tap_x0 = a.sub_px - 1; p.tap_w = 2;
      p.osy = p.osx = 2; p.ooy = a.sub_py; p.oox = a.sub_px; p.OH = 2 * a.H; p.OW = 2 * a.W;
    }
    const ConvTile t = pick_conv_tile(a.nimg, a.H, a.W);
    p.nimg = a.nimg; p.H = a.H; p.W = a.W; p.TW = t.tw; p.TH = t.th; p.TN = t.tn;
    p.tw_log = ilog2(t.tw);
    p.thw_log = ilog2(t.tw * t.th);
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

  int bn = a.force_bn;
  if (a.mode == EPI_GEGLU) {
    bn = 2 * kGegluGranule;
    VS_REQUIRE(a.N % bn == 0, "gemm_tc: GEGLU needs N %% %d == 0 (N=%d)", bn, a.N);
    VS_REQUIRE(a.bias != nullptr && a.residual == nullptr && a.rowvec == nullptr, "gemm_tc: GEGLU takes a bias only");
  } else if (bn == 0) {
    bn = pick_bn(p.m_tiles, a.N, p.num_kb);
  }
  VS_REQUIRE(bn == 64 || bn == 128 || bn == 160 || bn == 256, "gemm_tc: unsupported BLOCK_N %d", bn);
  p.n_tiles = (a.N + bn - 1) / bn;
  // CTA pairs (cta_group::2) where the main loop dominates (K >= 768): measured -5..-13% there, but +5..+12% on the
  // epilogue-bound K <= 640 shapes (two CTAs in lockstep, remote barrier arrivals).  "gemm_pair": 0 = never (A/B
  // switch), 2 = whenever there are two row tiles.
  const int pair_opt = get_option("gemm_pair");
  const bool pair = pair_opt != 0 && p.m_tiles >= 2 && bn >= 128 && (p.num_kb >= 12 || pair_opt == 2);
  {
    const uint64_t dims[2] = {(uint64_t)Ktot, (uint64_t)a.N};
    const uint64_t str[1] = {(uint64_t)Ktot * 2};
    const uint32_t box[2] = {BK, (uint32_t)(pair ? bn / 2 : bn)};   // a CTA pair: each CTA fetches half of the weight tile
    if (make_tmap_f16(&p.tmB, a.Bw, 2, dims, str, box, 1)) return 3;
  }
  // staged (transposed, coalesced) epilogue whenever the output geometry allows it: 16-byte strides, whole 32-column
  // sub-tiles, float4-aligned bias / row vectors
  const int out_cols = (a.mode == EPI_GEGLU) ? a.N / 2 : a.N;
  p.staged = (out_cols % EPI_COLS == 0) && (a.ldc % 8 == 0) && ((reinterpret_cast<uintptr_t>(a.out) & 15) == 0) &&
             (!a.residual || ((a.ldr % 8 == 0) && ((reinterpret_cast<uintptr_t>(a.residual) & 15) == 0))) &&
             (!a.bias || (reinterpret_cast<uintptr_t>(a.bias) & 15) == 0) &&
             (!a.rowvec || ((reinterpret_cast<uintptr_t>(a.rowvec) & 15) == 0 && p.ldrv % 4 == 0));
  if (!p.staged) VS_REQUIRE(a.mode == EPI_LINEAR, "gemm_tc: GEGLU output needs 32-column aligned, 16-byte strided rows");
  if (a.ln_stats || a.ln_parts) VS_REQUIRE(p.staged && (reinterpret_cast<uintptr_t>(a.ln_u) & 15) == 0, "gemm_tc: folded LayerNorm needs the staged epilogue");
  if (a.ln_sums_out) VS_REQUIRE(p.staged, "gemm_tc: row statistics output needs the staged epilogue (N %% 32 == 0, aligned rows)");
  ProfScope prof(st, a.taps != 1 ? PC_CONV : PC_GEMM, 2.0 * a.M * (double)a.N * Ktot, 1, a.M, a.N, Ktot);
  if (a.mode == EPI_GEGLU) {
    if (p.ln_stats || p.ln_parts) return pair ? launch<256, EPI_F_GEGLU | EPI_F_LN, true>(st, p) : launch<256, EPI_F_GEGLU | EPI_F_LN, false>(st, p);
    return pair ? launch<256, EPI_F_GEGLU, true>(st, p) : launch<256, EPI_F_GEGLU, false>(st, p);
  }
  if (pair) {
    switch (bn) {
      case 128: return launch_linear<128, true>(st, p);
      case 256: return launch_linear<256, true>(st, p);
      default: return launch_linear<160, true>(st, p);
    }
  }
  switch (bn) {
    case 64: return launch_linear<64, false>(st, p);
    case 128: return launch_linear<128, false>(st, p);
    case 256: return launch_linear<256, false>(st, p);
    default: return launch_linear<160, false>(st, p);
  }
}

}  // namespace vs
