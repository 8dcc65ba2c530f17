// Column-sweep whole-network kernel of the bf16 tensor-core ResNet path (included by conv_tc.cu).
//
// Why a second formulation.  tools/umma_bench*.cu shows that a 128 x N x 16 tcgen05.mma with both
// operands in shared memory costs max(N/2, 32 + N/4) cycles: the 4 KB activation operand is re-read
// from shared memory (128 B/cycle) by every instruction, so the N = 48 MMAs of the position-major
// kernel (resnet_fused.cuh) run at 44 cycles instead of 24.  Here the three WIDTH taps of the 3x3
// convolution (/root/reference/model/resnet.py:20-26) are stacked on the N axis instead:
//
//   lanes (M = 128)  = 128 consecutive rows h of ONE map column w   ("strip")
//   one MMA          = X[rows + (dh-1) d, w, 16 ch]  x  [ W[dh][+1] | W[dh][0] | W[dh][-1] ]   (N = 3 CP = 144)
//   accumulators     = a ring of CP-column blocks in TMEM, one block per OUTPUT column; the MMA of input
//                      column w writes the three consecutive blocks of output columns w-d, w, w+d
//
// so input column w is read once and feeds all nine taps with 9 MMAs of 72 cycles (the tensor-pipe
// rate) instead of 27 MMAs of 44.  Columns are visited run by run (w = r, r+d, r+2d, ... for every
// residue r mod d), so consecutive steps always shift the accumulator window by exactly one block.
// A ring wrap or a fresh (not yet zeroed) block splits the MMA into N = 96 + 48 pieces.
//
// Activation layout ("column planar-8"): [slot][NP][W][H][8] bf16.  A TMA box {8 ch, 128 + 2d rows, 1
// column, NP planes} starting at row -d lands in shared memory as the K-major SWIZZLE_NONE canonical
// layout (row pitch 16 B, plane pitch = LBO); out-of-range rows are zero-filled by the TMA unit, which
// is exactly the reference's zero padding (resnet.py:22-24, padding = dilation).  The +-d row shift
// of a height tap is a descriptor start-address offset.
//
// There is no CTA-wide barrier between layers.  Dependencies are tracked per map column with
// mbarriers: the producer loads column w of layer l+1 as soon as the epilogue has stored column w of
// layer l (col_done), the epilogue of layer l+1 starts storing only after every MMA of layer l has
// retired (layer_done: the two layers ping-pong the same buffers), weights are double buffered.
#pragma once
#include <type_traits>

namespace kws {

// Warp roles: warps 0 .. 4*NKC-1 epilogue (warp % 4 = TMEM lane quarter), then one TMA producer warp and three
// MMA issuer warps.  The front-end warps have the HIGHEST warp ids on purpose: the SM's warp arbiter prefers
// higher ids, and the epilogue warps spend half their time polling barriers; with the issuers at the low ids
// every instruction of the (latency-critical) issue loops waited behind those polls.
constexpr int kSwFrontWarps = 4;
constexpr int kSwIssuers = 3;       // issuer m issues the MMAs of height tap dh = m
constexpr int kSwMaxStages = 12;
constexpr int kSwMaxRing = 32;
constexpr int kSwMaxW = 256;
constexpr int kSwWeightStep = 12;   // producer step of a layer at which the NEXT layer's weights are requested

__host__ __device__ constexpr int sw_epi_warps(int NKC) { return 4 * NKC; }
__host__ __device__ constexpr int sw_threads(int NKC) { return 32 * (kSwFrontWarps + sw_epi_warps(NKC)); }

// control block layout (bytes from the start of dynamic shared memory)
constexpr int kSwBarFull = 0;                                  // [kSwMaxStages]
constexpr int kSwBarEmpty = kSwBarFull + 8 * kSwMaxStages;     // [kSwMaxStages]
constexpr int kSwBarTfull = kSwBarEmpty + 8 * kSwMaxStages;    // [kSwMaxRing]
constexpr int kSwBarTempty = kSwBarTfull + 8 * kSwMaxRing;     // [kSwMaxRing]
constexpr int kSwBarWfull = kSwBarTempty + 8 * kSwMaxRing;     // [2]
constexpr int kSwBarLayer = kSwBarWfull + 16;                  // [2]  all MMAs of a layer retired
constexpr int kSwBarConv0 = kSwBarLayer + 16;                  // [1]  conv_0 output stored
constexpr int kSwBarCol = kSwBarConv0 + 8;                     // [2][kSwMaxW]  column stored by every epilogue warp
constexpr int kSwTmemSlot = kSwBarCol + 8 * 2 * kSwMaxW;       // u32
constexpr int kSwZero = kSwTmemSlot + 4;                       // u32, always 0 (see the MMA issuers)
constexpr int kSwPool = round_up(kSwZero + 4, 128);            // [64] f32 pooled sums
constexpr int kSwW0 = kSwPool + 256;                           // [64][12] f32 conv_0 weights
constexpr int kSwCtrlBytes = round_up(kSwW0 + 64 * 12 * 4, 1024);

struct SwParams {
  // Per-layer data is derived from kernel parameters only (constant bank => warp-uniform for the compiler, which
  // keeps the MMA issue loop on the uniform datapath): layer l (0-based) has dilation 2^(l/3) or 1
  // (resnet.py:21), reads Q and adds the skip tensor iff l is odd (resnet.py:51-53), and its packed weights /
  // epilogue constants sit at a fixed stride.
  const unsigned char* wpack0;  // layer 0 weights, [NKC][3 dh][2 K halves][3 blocks][CP][8] bf16; block k = width tap dw = 2 - k
  const unsigned char* kconst0; // layer 0 epilogue constants, [CP] f32 (see pad_bn_kernel)
  int64_t layer_stride;         // bytes between consecutive layers in both arrays
  const CUtensorMap* maps;      // [n_layers]  input tensor map of every layer (global memory, 64 B aligned)
  int use_dilation;
  const float* feat;            // [B][T][F]
  const float* conv0_w;         // [C][9]
  const float* last_scale;      // [CP]
  const float* out_w;           // [n_labels][C]
  const float* out_b;           // [n_labels]
  float* logits;                // [B][n_labels]
  __nv_bfloat16* P;             // [n_slots][NP][W][H][8]
  __nv_bfloat16* Q;
  int64_t B;
  int n_layers, C, n_labels, T, F, ph, pw, H, W;
  int n_strips;                 // ceil(H / 128)
  int smem_w_off[2], smem_ring_off, ring_slot_bytes, n_stages;
  int l2_policy;
  int bulk_rows;                // > 0 (single-strip maps): columns are staged with 1-D bulk copies of bulk_rows = H rows per plane
                                //   into a fixed [plane][128 + 2 dmax] slot whose pad rows stay zero; 0: TMA tensor boxes
  int dmax;                     // largest dilation of the network (bulk path: data row h sits at slot row dmax + h)
  int issue_style;              // 0: MMA operands in uniform registers, 1: ordinary registers + R2UR (experiments)
  long long* debug;             // optional [8] cycle counters of CTA 0's issuer
};

template <int NKC, bool DBG>
__global__ void __launch_bounds__(sw_threads(NKC), 1)
resnet_tc_sweep_kernel(const SwParams p) {
  constexpr int kEpiWarps = sw_epi_warps(NKC);
  constexpr int kEpiThreads = 32 * kEpiWarps;
  constexpr int CP = 16 * NKC;
  constexpr int NP = 2 * NKC;
  constexpr int NB = (512 / CP) < kSwMaxRing ? (512 / CP) : kSwMaxRing;   // accumulator ring, in blocks of CP columns
  constexpr int BLK_BYTES = CP * 16;                   // one [CP][8] weight block
  constexpr int W_LBO = 3 * BLK_BYTES;                 // distance between the two K halves of a weight slab
  constexpr int W_SLAB = 2 * W_LBO;                    // one (kc, dh) slab
  constexpr int W_BYTES = NKC * 3 * W_SLAB;
  extern __shared__ __align__(1024) unsigned char smem[];

  const uint32_t sbase = smem_u32(smem);
  auto full_bar = [&](int s) { return sbase + kSwBarFull + 8u * s; };
  auto empty_bar = [&](int s) { return sbase + kSwBarEmpty + 8u * s; };
  auto tfull_bar = [&](int a) { return sbase + kSwBarTfull + 8u * a; };
  auto tempty_bar = [&](int a) { return sbase + kSwBarTempty + 8u * a; };
  auto wfull_bar = [&](int i) { return sbase + kSwBarWfull + 8u * i; };
  auto layer_bar = [&](int i) { return sbase + kSwBarLayer + 8u * i; };
  const uint32_t conv0_bar = sbase + kSwBarConv0;
  auto col_bar = [&](int par, int w) { return sbase + kSwBarCol + 8u * (par * kSwMaxW + w); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kSwTmemSlot);
  float* s_pool = reinterpret_cast<float*>(smem + kSwPool);
  float* s_w0 = reinterpret_cast<float*>(smem + kSwW0);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const int n_layers = p.n_layers, H = p.H, W = p.W, n_strips = p.n_strips;
  const int64_t n_my = (p.B - blockIdx.x + gridDim.x - 1) / gridDim.x;
  const int64_t n_seq = n_my * n_layers;

  // ---- one-time setup
  if (threadIdx.x == 0) {
    for (int s = 0; s < p.n_stages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < NB; ++a) { mbar_init(tfull_bar(a), kSwIssuers); mbar_init(tempty_bar(a), kEpiWarps); }
    for (int i = 0; i < 2; ++i) { mbar_init(wfull_bar(i), 1); mbar_init(layer_bar(i), kSwIssuers); }
    mbar_init(conv0_bar, kEpiWarps);
    for (int par = 0; par < 2; ++par)
      for (int w = 0; w < W; ++w) mbar_init(col_bar(par, w), (uint32_t)(kEpiWarps * n_strips));
    fence_barrier_init();
  }
  if (threadIdx.x == 32) *reinterpret_cast<volatile uint32_t*>(smem + kSwZero) = 0u;
  if (warp == kEpiWarps + 1) tmem_alloc(smem_u32(tmem_slot), 512);
  for (int i = threadIdx.x; i < CP * 12; i += sw_threads(NKC)) {
    const int c = i / 12, k = i - c * 12;
    s_w0[i] = (k < 9 && c < p.C) ? p.conv0_w[c * 9 + k] : 0.f;
  }
  for (int i = threadIdx.x; i < CP; i += sw_threads(NKC)) s_pool[i] = 0.f;
  if (p.bulk_rows > 0) {
    // bulk-copy staging: the pad rows above and below the map are never written again and supply the zero padding
    uint4* ring = reinterpret_cast<uint4*>(smem + p.smem_ring_off);
    const int n16 = p.n_stages * (p.ring_slot_bytes >> 4);
    for (int i = threadIdx.x; i < n16; i += sw_threads(NKC)) ring[i] = make_uint4(0u, 0u, 0u, 0u);
    fence_async_smem();   // generic-proxy zeros -> visible to the tensor core's (async proxy) operand reads
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // Every MMA accumulates (three issuer warps feed the same accumulators, in no particular order), so the ring
  // starts zeroed and the epilogue re-zeroes each block right after reading it.
  if (warp < kEpiWarps) {
    const int q = warp & 3, j = warp >> 2;
    for (int a = 0; a < NB; ++a) tmem_st16_zero(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * CP + 16 * j));
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  const int64_t plane_stride = (int64_t)W * H;                          // 16-byte units
  const int64_t slot_base = (int64_t)blockIdx.x * NP * plane_stride;    // this CTA's utterance slot
  uint4* bufP = reinterpret_cast<uint4*>(p.P) + slot_base;
  uint4* bufQ = reinterpret_cast<uint4*>(p.Q) + slot_base;
  const bool use_pol = p.l2_policy != 0;
  const uint64_t pol_keep = l2_policy_evict_last();
  const uint64_t pol_stream = l2_policy_evict_first();

  auto layer_dil = [&](int l) { return p.use_dilation ? (1 << (l / 3)) : 1; };
  // rows per plane of a staged column = plane pitch in shared memory (kept a multiple of 8 rows = 128 B)
  auto box_rows_of = [&](int d) { return p.bulk_rows > 0 ? 128 + 2 * p.dmax : ((128 + 2 * d + 7) & ~7); };
  // slot row that the height tap dh = 0 of output row 0 reads
  auto row0_of = [&](int d) { return p.bulk_rows > 0 ? p.dmax - d : 0; };

  if (warp == kEpiWarps) {
    // ========================================= TMA producer =========================================
    // The whole warp walks the (warp-uniform) schedule; one elected lane issues the copies.
    const bool leader = elect_one();
    if (n_seq > 0) {
      int stage = 0;
      uint32_t sphase = 0;
      int64_t seq = 0, utt = 0;
      if (leader) {
        mbar_expect_tx(wfull_bar(0), W_BYTES);
        bulk_load(sbase + p.smem_w_off[0], p.wpack0, W_BYTES, wfull_bar(0));
      }
      for (int64_t b = blockIdx.x; b < p.B; b += gridDim.x, ++utt) {
        for (int l = 0; l < n_layers; ++l, ++seq) {
          const CUtensorMap* map = p.maps + l;
          const int d = layer_dil(l), box_rows = box_rows_of(d);
          const int n_runs = d < W ? d : W;
          const uint32_t tx = (uint32_t)(NP * box_rows * 16);
          const bool in_q = (l & 1) != 0;
          const uint64_t pol = in_q ? pol_stream : pol_keep;
          bool w_pending = seq + 1 < n_seq;
          auto request_weights = [&]() {
            // buffer (seq+1)&1 was read by the MMAs of layer seq-1: wait until they have retired
            if (seq >= 1) mbar_wait(layer_bar((int)((seq - 1) & 1)), (uint32_t)(((seq - 1) >> 1) & 1));
            const int nl = (l + 1 < n_layers) ? l + 1 : 0;
            const int nb = (int)((seq + 1) & 1);
            if (leader) {
              mbar_expect_tx(wfull_bar(nb), W_BYTES);
              bulk_load(sbase + p.smem_w_off[nb], p.wpack0 + nl * p.layer_stride, W_BYTES, wfull_bar(nb));
            }
            w_pending = false;
          };
          if (l == 0) mbar_wait(conv0_bar, (uint32_t)(utt & 1));   // conv_0 of this utterance is in P
          const int prev_par = (int)((seq - 1) & 1);
          const uint32_t prev_phase = (uint32_t)(((seq - 1) >> 1) & 1);
          int step = 0;
          for (int s = 0; s < n_strips; ++s)
            for (int r = 0; r < n_runs; ++r)
              for (int w = r; w < W; w += d, ++step) {
                if (w_pending && step == kSwWeightStep) request_weights();
                if (l > 0) mbar_wait(col_bar(prev_par, w), prev_phase);   // column w of the previous layer is stored
                mbar_wait(empty_bar(stage), sphase ^ 1);
                const uint32_t dst = sbase + p.smem_ring_off + (uint32_t)stage * p.ring_slot_bytes;
                if (p.bulk_rows > 0) {
                  // one contiguous H x 16 B run per 8-channel plane (a TMA box with 16-byte rows fetches a whole
                  // 32-byte sector per row: 2.6x the bytes, measured with ncu)
                  const uint32_t bytes = (uint32_t)p.bulk_rows * 16u;
                  if (leader) mbar_expect_tx(full_bar(stage), bytes * NP);
                  __syncwarp();
                  if (lane < NP) {
                    const uint4* src = (in_q ? bufQ : bufP) + (int64_t)lane * plane_stride + (int64_t)w * H;
                    bulk_load_hint(dst + (uint32_t)(lane * box_rows + p.dmax) * 16u, src, bytes, full_bar(stage), pol);
                  }
                } else if (leader) {
                  mbar_expect_tx(full_bar(stage), tx);
                  if (use_pol) tma_load_4d_hint(dst, map, full_bar(stage), 0, s * 128 - d, w, (int)blockIdx.x * NP, pol);
                  else tma_load_4d(dst, map, full_bar(stage), 0, s * 128 - d, w, (int)blockIdx.x * NP);
                }
                if (++stage == p.n_stages) { stage = 0; sphase ^= 1; }
              }
          if (w_pending) request_weights();
        }
      }
    }
    __syncwarp();
  } else if (warp > kEpiWarps) {
    // ========================================= MMA issuers (3 warps) =========================================
    // Issuer m owns every third STEP: it waits for the step's staged column and for the ring slots of the step's
    // window, issues the 3*NKC MMAs of the step (twice as many, half as wide, when the window wraps around the
    // ring) and commits; for the other two steps it only advances its counters.  While one issuer is inside its
    // burst the next one has already prepared its operands and queues right behind it.  All MMAs accumulate (the
    // ring blocks are zeroed by the epilogue), so the order in which the three warps' MMAs reach the tensor pipe
    // does not matter (tools/umma_bench2/4: concurrent accumulation from several warps into one accumulator is
    // exact and runs at the pipe rate).
    //
    // What limits this role is the INSTRUCTION COUNT of its single worker lane (ncu: ~5 cycles per instruction in
    // this branchy scalar code, the MMA instructions themselves are 3 % of its time), so the step is kept lean: ring
    // slots advance by increments (no divisions), the cycle accounting is compiled out unless DBG, barrier polls use
    // the hardware suspend hint, and the MMA burst is straight-line code without predicated-off instructions.
    //
    // An output block receives MMAs from the steps of its three neighbouring input columns, i.e. from all three
    // issuers: tfull counts three arrivals (each issuer commits on every block of its window; the run's first and
    // last block, which have only two contributing steps, get a plain arrival from the issuer of the edge step),
    // and every issuer waits for the ring slot of every block of its window to be drained, not only the fresh one,
    // because nothing orders the issuers among themselves.
    const int me = warp - kEpiWarps - 1;
    // Whole warp walks the schedule with warp-uniform values only (kernel parameters, counters, the TMEM base
    // broadcast by a shuffle), so the descriptors live in uniform registers and an MMA costs 2-3 instructions
    // instead of the 17-instruction ELECT / R2UR / branch sequence the compiler emits for per-thread operands;
    // one elected lane issues the MMAs, commits and arrivals.
    const bool leader = elect_one();
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    constexpr uint32_t nz = 0u;
    if (n_seq > 0) {
      int stage = 0;
      uint32_t sphase = 0;
      int owner = 0;            // issuer of the current step (global step counter mod 3)
      int sl = 0;               // ring slot of the output block of the CURRENT step's own column
      uint32_t pr = 0;          // use parity of that slot
      constexpr uint32_t desc_hi = (128u >> 4) | (1u << 14);   // SBO = 128 B, descriptor version 1
      constexpr uint32_t idesc0 = umma_idesc(128, 0);
      constexpr uint32_t idesc_blk = (uint32_t)(CP >> 3) << 17;   // one more block of CP columns
      constexpr uint32_t b_lbo = ((uint32_t)(W_LBO >> 4) & 0x3FFFu) << 16;
      constexpr uint32_t blk16 = BLK_BYTES >> 4;
      long long dbg_full = 0, dbg_tempty = 0, dbg_issue = 0, dbg_w = 0, dbg_utt = 0, dbg_other = 0, dbg_t = DBG ? clock64() : 0;
      const bool dbg = DBG && p.debug != nullptr && blockIdx.x == 0 && me == 0;
      auto stamp = [&](long long& bucket) {
        if constexpr (DBG) { if (dbg) { const long long t = clock64(); bucket += t - dbg_t; dbg_t = t; } }
      };
      int64_t seq = 0;
      for (int64_t b = blockIdx.x; b < p.B; b += gridDim.x) {
        for (int l = 0; l < n_layers; ++l, ++seq) {
          const int d = layer_dil(l), box_rows = box_rows_of(d);
          const int n_runs = d < W ? d : W;
          const int cur = (int)(seq & 1);
          const uint32_t plane16 = (uint32_t)box_rows;                       // plane pitch in 16-byte units
          const uint32_t a_lbo = (plane16 & 0x3FFFu) << 16;
          const uint32_t w16 = ((sbase + p.smem_w_off[cur]) >> 4) + nz;
          const uint32_t row0 = (uint32_t)row0_of(d);
          stamp(dbg_other);
          mbar_wait_lean(wfull_bar(cur), (uint32_t)((seq >> 1) & 1));
          stamp(dbg_w);
          for (int s = 0; s < n_strips; ++s)
            for (int r = 0; r < n_runs; ++r) {
              const int Lr = (W - r + d - 1) / d;   // output columns of this run
              for (int i = 0; i < Lr; ++i) {
                if (owner == me) {
                  // window = output blocks of columns i-1 (if any), i, i+1 (if any): n blocks, ring-contiguous from p0
                  const bool lo = i > 0, hi_ = i + 1 < Lr;
                  const int s_lo = sl > 0 ? sl - 1 : NB - 1, s_hi = sl + 1 < NB ? sl + 1 : 0;
                  const uint32_t par_lo = pr ^ (sl == 0 ? 1u : 0u), par_hi = pr ^ (sl + 1 == NB ? 1u : 0u);
                  const int p0 = lo ? s_lo : sl;
                  const int n = 1 + (lo ? 1 : 0) + (hi_ ? 1 : 0);
                  const int wrap_at = NB - p0 < n ? NB - p0 : n;              // blocks before the ring wraps
                  const int blk0 = lo ? 0 : 1;                                 // weight block of the first window block
                  stamp(dbg_other);
                  // the epilogue must have drained and re-zeroed the previous use of every slot of the window
                  if (lo) mbar_wait_lean(tempty_bar(s_lo), par_lo ^ 1u);
                  mbar_wait_lean(tempty_bar(sl), pr ^ 1u);
                  if (hi_) mbar_wait_lean(tempty_bar(s_hi), par_hi ^ 1u);
                  stamp(dbg_tempty);
                  mbar_wait_lean(full_bar(stage), sphase);
                  tc_fence_after();
                  if constexpr (DBG) { if (l == 0 && s == 0 && r == 0 && i < kSwIssuers) stamp(dbg_utt); else stamp(dbg_full); }
                  // All operands of the step are computed BEFORE the burst, and the burst is straight-line code
                  // without predicated-off MMAs (separate path for the wrapped window).
                  const uint32_t a16 = ((sbase + p.smem_ring_off + (uint32_t)stage * p.ring_slot_bytes) >> 4) + row0 + nz;
                  const uint32_t d1 = tmem_u + (uint32_t)(p0 * CP);
                  const uint32_t id1 = idesc0 + (uint32_t)wrap_at * idesc_blk;
                  uint32_t al[3 * NKC], bl[3 * NKC];
#pragma unroll
                  for (int kc = 0; kc < NKC; ++kc)
#pragma unroll
                    for (int dh = 0; dh < 3; ++dh) {
                      al[kc * 3 + dh] = ((a16 + (uint32_t)(2 * kc) * plane16 + (uint32_t)(dh * d)) & 0x3FFFu) | a_lbo;
                      bl[kc * 3 + dh] = ((w16 + (uint32_t)(((kc * 3 + dh) * W_SLAB) >> 4) + (uint32_t)blk0 * blk16) & 0x3FFFu) | b_lbo;
                    }
                  if (!leader) {
                    // (only the elected lane issues)
                  } else if (wrap_at == n) {
#pragma unroll
                    for (int k = 0; k < 3 * NKC; ++k) umma_f16_lohi<true>(d1, al[k], bl[k], desc_hi, id1);
                  } else {
                    // the window wraps around the ring: first wrap_at blocks at p0, the rest from slot 0
                    const uint32_t id2 = idesc0 + (uint32_t)(n - wrap_at) * idesc_blk;
                    const uint32_t bo2 = (uint32_t)wrap_at * blk16;
#pragma unroll
                    for (int k = 0; k < 3 * NKC; ++k) {
                      umma_f16_lohi<true>(d1, al[k], bl[k], desc_hi, id1);
                      umma_f16_lohi<true>(tmem_u, al[k], bl[k] + bo2, desc_hi, id2);   // (weights end far below 256 KB: no carry out of the address field)
                    }
                  }
                  if (leader) {
                  umma_commit(empty_bar(stage));   // stage reusable once these MMAs retire
                  // this issuer's share of every block of the window; the run's edge blocks have only two
                  // contributing steps, so the issuer of the edge step stands in for the missing third
                  if (lo) umma_commit(tfull_bar(s_lo));
                  umma_commit(tfull_bar(sl));
                  if (hi_) umma_commit(tfull_bar(s_hi));
                  if (!lo) mbar_arrive(tfull_bar(sl));
                  if (!hi_) mbar_arrive(tfull_bar(sl));
                  }
                  __syncwarp();
                  stamp(dbg_issue);
                }
                if (++owner == kSwIssuers) owner = 0;
                if (++stage == p.n_stages) { stage = 0; sphase ^= 1; }
                if (++sl == NB) { sl = 0; pr ^= 1u; }
              }
            }
          if (leader) umma_commit(layer_bar(cur));   // every MMA of this layer issued by this warp has retired
          __syncwarp();
        }
      }
      if constexpr (DBG) {
        if (dbg && leader) {
          p.debug[0] = dbg_w; p.debug[1] = dbg_tempty; p.debug[2] = dbg_full; p.debug[3] = dbg_issue;
          p.debug[4] = n_my; p.debug[5] = dbg_utt; p.debug[6] = dbg_other; p.debug[7] = 0;
        }
      }
    }
    __syncwarp();
  } else {
    // ========================================= epilogue (4*NKC warps) =========================================
    // warp e: TMEM lane quarter q = warp % 4 (rows q*32 .. q*32+31 of the strip), channel group j = e / 4
    // (16 channels = planes 2j, 2j+1).  One block = one output column of one strip.
    const int q = warp & 3;
    const int j = warp >> 2;
    const int et = threadIdx.x;
    int run_pos = 0;
    uint32_t run_par = 0;
    int64_t seq = 0;
    // cycle accounting of epilogue warp 0 of CTA 0 (HONK2_TC_DEBUG=1)
    const bool edbg = DBG && p.debug != nullptr && blockIdx.x == 0 && warp == 0;
    long long e_wait = 0, e_tmem = 0, e_pub = 0, e_math = 0, e_conv0 = 0, e_t = clock64();
    for (int64_t b = blockIdx.x; b < p.B; b += gridDim.x) {
      // ------------------------------ conv_0 -> P (resnet.py:40-44) ------------------------------
      {
        const float* src = p.feat + b * (int64_t)p.T * p.F;
        if (p.ph == 1 && p.pw == 1) {
          // item = (4 consecutive columns, one row); consecutive threads take consecutive rows so that a warp's
          // 16-byte stores to one (plane, column) are contiguous
          const int groups = (W + 3) >> 2;
          for (int item = et; item < H * groups; item += kEpiThreads) {
            const int gx = item / H, h = item - gx * H, w0 = gx * 4;
            float pch[3][6];
#pragma unroll
            for (int a = 0; a < 3; ++a)
#pragma unroll
              for (int e = 0; e < 6; ++e) {
                const int hh = h + a - 1, ww = w0 + e - 1;
                pch[a][e] = (hh >= 0 && hh < p.T && ww >= 0 && ww < p.F) ? __ldg(src + (int64_t)hh * p.F + ww) : 0.f;
              }
            for (int pl = 0; pl < NP; ++pl) {
              float a4[4][8];
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const float* wc9 = s_w0 + (pl * 8 + e) * 12;
                const float4 wa = *reinterpret_cast<const float4*>(wc9);
                const float4 wb = *reinterpret_cast<const float4*>(wc9 + 4);
                const float w8 = wc9[8];
#pragma unroll
                for (int px = 0; px < 4; ++px) {
                  float v = pch[0][px] * wa.x;
                  v = fmaf(pch[0][px + 1], wa.y, v); v = fmaf(pch[0][px + 2], wa.z, v);
                  v = fmaf(pch[1][px], wa.w, v); v = fmaf(pch[1][px + 1], wb.x, v); v = fmaf(pch[1][px + 2], wb.y, v);
                  v = fmaf(pch[2][px], wb.z, v); v = fmaf(pch[2][px + 1], wb.w, v); v = fmaf(pch[2][px + 2], w8, v);
                  a4[px][e] = fmaxf(v, 0.f);
                }
              }
#pragma unroll
              for (int px = 0; px < 4; ++px) {
                if (w0 + px < W) {
                  uint4 o;
                  __nv_bfloat162* ob = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
                  for (int e = 0; e < 4; ++e) ob[e] = __floats2bfloat162_rn(a4[px][2 * e], a4[px][2 * e + 1]);
                  uint4* dst = bufP + pl * plane_stride + (int64_t)(w0 + px) * H + h;
                  if (use_pol) st_hint(dst, o, pol_keep); else *dst = o;
                }
              }
            }
          }
        } else {
          const float inv = 1.f / (float)(p.ph * p.pw);
          for (int pix = et; pix < H * W; pix += kEpiThreads) {
            const int wo = pix / H, ho = pix - wo * H;
            for (int pl = 0; pl < NP; ++pl) {
              float a8[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) a8[e] = 0.f;
              for (int i = 0; i < p.ph; ++i)
                for (int jj = 0; jj < p.pw; ++jj) {
                  const int hc = ho * p.ph + i, wc = wo * p.pw + jj;   // centre of the 3x3 window
                  float xin[9];
#pragma unroll
                  for (int a = 0; a < 3; ++a)
#pragma unroll
                    for (int e = 0; e < 3; ++e) {
                      const int hh = hc + a - 1, ww = wc + e - 1;
                      xin[a * 3 + e] = (hh >= 0 && hh < p.T && ww >= 0 && ww < p.F) ? __ldg(src + (int64_t)hh * p.F + ww) : 0.f;
                    }
#pragma unroll
                  for (int e = 0; e < 8; ++e) {
                    const float* wc9 = s_w0 + (pl * 8 + e) * 12;
                    const float4 wa = *reinterpret_cast<const float4*>(wc9);
                    const float4 wb = *reinterpret_cast<const float4*>(wc9 + 4);
                    float v = xin[0] * wa.x;
                    v = fmaf(xin[1], wa.y, v); v = fmaf(xin[2], wa.z, v); v = fmaf(xin[3], wa.w, v);
                    v = fmaf(xin[4], wb.x, v); v = fmaf(xin[5], wb.y, v); v = fmaf(xin[6], wb.z, v);
                    v = fmaf(xin[7], wb.w, v); v = fmaf(xin[8], wc9[8], v);
                    a8[e] += fmaxf(v, 0.f);
                  }
                }
              uint4 o;
              __nv_bfloat162* ob = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
              for (int e = 0; e < 4; ++e) ob[e] = __floats2bfloat162_rn(a8[2 * e] * inv, a8[2 * e + 1] * inv);
              bufP[pl * plane_stride + pix] = o;
            }
          }
        }
        __threadfence();
        fence_async_all();   // generic-proxy global writes -> visible to the TMA (async proxy) reads of layer 1
        __syncwarp();
        if (lane == 0) mbar_arrive(conv0_bar);
        if (edbg) { const long long t = clock64(); e_conv0 += t - e_t; e_t = t; }
      }

      for (int l = 0; l < n_layers; ++l, ++seq) {
        const int d = layer_dil(l);
        const int n_runs = d < W ? d : W;
        const int cur = (int)(seq & 1);
        const float* kconst = reinterpret_cast<const float*>(p.kconst0 + l * p.layer_stride);
        float kc_reg[16];
#pragma unroll
        for (int c = 0; c < 16; ++c) kc_reg[c] = __ldg(kconst + 16 * j + c);
        // The layer body is instantiated per (skip, pooling) variant so that the registers of the skip prefetch
        // and of the pooled sums are not live together.
        auto layer_body = [&](auto skip_c, auto last_c) {
          constexpr bool HAS_SKIP = decltype(skip_c)::value;   // odd l: adds the skip tensor from P, stores to P in place
          constexpr bool LAST = decltype(last_c)::value;       // pooled instead of stored
          const uint4* skip_in = bufP + (int64_t)(2 * j) * plane_stride;
          uint4* y_out = (HAS_SKIP ? bufP : bufQ) + (int64_t)(2 * j) * plane_stride;
          const uint64_t pol_out = HAS_SKIP ? pol_keep : pol_stream;
          float psum[LAST ? 16 : 1];
#pragma unroll
          for (int c = 0; c < (LAST ? 16 : 1); ++c) psum[c] = 0.f;
          // this layer overwrites the buffer the previous layer READS through TMA: wait until all of its MMAs
          // (hence all of its loads) have retired before the first store
          bool guard = l > 0;
          int pending_w = -1;   // column whose stores still have to be published (one block behind)
          auto publish = [&](int wcol) {
            // generic-proxy global stores of this thread -> visible to the async proxy (the TMA loads of the next
            // layer, issued by this CTA's producer after it acquires the column barrier).  The all-space
            // fence.proxy.async compiles to MEMBAR.ALL.GPU and, issued by 12 warps per column, stalls the whole
            // SM; the .global form is a plain view fence.
            fence_async_global();
            __syncwarp();
            if (lane == 0) mbar_arrive(col_bar(cur, wcol));
          };
          // block iterator over (strip, run, output column of the run)
          struct It { int s, r, o, Lr, w; bool done; };
          auto it_begin = [&]() { It it; it.s = 0; it.r = 0; it.o = 0; it.Lr = (W + d - 1) / d; it.w = 0; it.done = false; return it; };
          auto it_next = [&](It it) {
            ++it.o; it.w += d;
            if (it.o == it.Lr) {
              it.o = 0;
              if (++it.r == n_runs) { it.r = 0; if (++it.s == n_strips) it.done = true; }
              it.w = it.r;
              it.Lr = (W - it.r + d - 1) / d;
            }
            return it;
          };
          auto load_skip = [&](uint4 (&pv)[2], const It& it) {
            if constexpr (HAS_SKIP) {
              const int row = it.s * 128 + q * 32 + lane;
              if (!it.done && row < H) {
                const int64_t off = (int64_t)it.w * H + row;
                pv[0] = use_pol ? ld_hint(skip_in + off, pol_keep) : skip_in[off];
                pv[1] = use_pol ? ld_hint(skip_in + off + plane_stride, pol_keep) : skip_in[off + plane_stride];
              }
            }
          };
          auto process = [&](const It& it, const uint4 (&pv)[2]) {
            const int row = it.s * 128 + q * 32 + lane;
            const bool valid = row < H;
            const int64_t off = (int64_t)it.w * H + row;
            const int t = run_pos + it.o, qd = t / NB, pos = t - qd * NB;
            mbar_wait_sleepy(tfull_bar(pos), run_par ^ (uint32_t)(qd & 1));
            tc_fence_after();
            if (edbg) { const long long t = clock64(); e_wait += t - e_t; e_t = t; }
            uint32_t v[16];
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(pos * CP + 16 * j);
            tmem_ld16(taddr, v);
            tmem_ld_wait();
            tmem_st16_zero(taddr);   // the next user of this ring slot accumulates from zero
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar(pos));   // accumulators are in registers, the slot is zero again: free
            if (edbg) { const long long t = clock64(); e_tmem += t - e_t; e_t = t; }
            if (pending_w >= 0) { publish(pending_w); pending_w = -1; }
            if (edbg) { const long long t = clock64(); e_pub += t - e_t; e_t = t; }
            if (guard) {
              mbar_wait_sleepy(layer_bar(cur ^ 1), (uint32_t)(((seq - 1) >> 1) & 1));
              guard = false;
            }
            if (valid) {
#pragma unroll
              for (int hf = 0; hf < 2; ++hf) {
                float x[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) x[e] = fmaxf(__uint_as_float(v[8 * hf + e]), 0.f) + kc_reg[8 * hf + e];
                if constexpr (HAS_SKIP) {
                  const __nv_bfloat162* pb = reinterpret_cast<const __nv_bfloat162*>(&pv[hf]);
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    const float2 f = __bfloat1622float2(pb[e]);
                    x[2 * e] += f.x;
                    x[2 * e + 1] += f.y;
                  }
                }
                if constexpr (LAST) {
#pragma unroll
                  for (int e = 0; e < 8; ++e) psum[8 * hf + e] += x[e];
                } else {
                  uint4 yo;
                  __nv_bfloat162* yb = reinterpret_cast<__nv_bfloat162*>(&yo);
#pragma unroll
                  for (int e = 0; e < 4; ++e) yb[e] = __floats2bfloat162_rn(x[2 * e], x[2 * e + 1]);
                  if (use_pol) st_hint(y_out + off + hf * plane_stride, yo, pol_out);
                  else y_out[off + hf * plane_stride] = yo;
                }
              }
            }
            if (edbg) { const long long t = clock64(); e_math += t - e_t; e_t = t; }
            pending_w = it.w;
            if (it.o == it.Lr - 1) {   // run finished: advance the ring bookkeeping
              const int t2 = run_pos + it.Lr, q2 = t2 / NB;
              run_pos = t2 - q2 * NB;
              run_par ^= (uint32_t)(q2 & 1);
            }
          };
          // The skip tensor is prefetched two blocks ahead (two register sets, the loop is unrolled by two) so that
          // its L2 latency hides behind a whole block period.
          uint4 pa[2], pb2[2];
          It ia = it_begin(), ib = it_next(ia);
          load_skip(pa, ia);
          load_skip(pb2, ib);
          while (!ia.done) {
            process(ia, pa);
            ia = it_next(ib);
            load_skip(pa, ia);
            if (ib.done) break;
            process(ib, pb2);
            ib = it_next(ia);
            load_skip(pb2, ib);
          }
          if (pending_w >= 0) publish(pending_w);
          if constexpr (LAST) {
            // fused global mean (resnet.py:57-58): warp-reduce the 32 rows, one shared atomic per channel
#pragma unroll
            for (int c = 0; c < 16; ++c) {
              float sum = psum[c];
#pragma unroll
              for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
              if (lane == c) atomicAdd(s_pool + 16 * j + c, sum);
            }
          }
        };
        const bool has_skip = (l & 1) != 0, last = l == n_layers - 1;
        if (has_skip) { if (last) layer_body(std::true_type{}, std::true_type{}); else layer_body(std::true_type{}, std::false_type{}); }
        else { if (last) layer_body(std::false_type{}, std::true_type{}); else layer_body(std::false_type{}, std::false_type{}); }
      }

      // ------------------------------ logits (resnet.py:59) ------------------------------
      asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
      for (int lb = et; lb < p.n_labels; lb += kEpiThreads) {
        const float inv = 1.f / (float)(H * W);
        float v = __ldg(p.out_b + lb);
        // s_pool holds sums of z = x - mean; the BatchNorm output mean is z_mean / sigma (resnet.py:55-58)
        for (int c = 0; c < p.C; ++c)
          v = fmaf(s_pool[c] * inv * __ldg(p.last_scale + c), __ldg(p.out_w + lb * p.C + c), v);
        p.logits[b * p.n_labels + lb] = v;
      }
      asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
      if (et < CP) s_pool[et] = 0.f;
      // the next writers of s_pool (last layer of the next utterance) are at most one accumulator ring
      // ahead of this thread, i.e. far behind this store
      if (edbg) { const long long t = clock64(); e_math += t - e_t; e_t = t; }
    }
    if (edbg && lane == 0) {
      p.debug[8] = e_wait; p.debug[9] = e_tmem; p.debug[10] = e_pub; p.debug[11] = e_math; p.debug[12] = e_conv0;
    }
  }

  // ---- teardown
  tc_fence_before();
  __syncthreads();
  if (warp == kEpiWarps + 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace kws
