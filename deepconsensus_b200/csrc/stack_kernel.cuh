// The whole encoder stack in ONE launch (included by kernels.cu after the helpers it uses).
//
// A CTA pair (tcgen05 cta_group::2, M = 256) owns two window-aligned 128-token tiles and walks them through ALL
// layers; the fp32 residual tile never leaves the SM: it lives in the 288 TMEM columns "Y" from the first layer to the
// last, and every sub-layer accumulates onto it in place (residual-in-accumulator):
//
//   per layer   P1  workers: Y (+= b2 of the previous FFN) -> [LayerNorm] -> bf16 operand tile sA
//               per head h: 3 UMMA blocks (N = 144: q, k, v of the head; K = 288) into ACC (TMEM 288..431),
//                           workers move each block as bf16 into padded shared-memory rows,
//                           banded softmax attention per 16-query block (mma.sync, two-pass) -> att_h (bf16, KC16)
//                           Y += att_h * Wo_h^T                     (encoder_stack.py:72-93, attention_layer.py:218)
//               P5  workers: Y (= x_mid) -> [LayerNorm] -> sA
//               16 hidden chunks: H = sA*W1c (TMEM 288..415) -> +b1, ReLU, bf16 -> shared -> Y += Hc*W2c
//   per tile    once: Y <- x (global), and after the last layer Y + b2 -> x (global) for the head kernel.
//
// Against the per-layer kernels this removes, per tile and layer, the fp32 residual round trip through HBM (294 KB),
// the attention-output and operand images (296 KB), and both tile hand-over phases in which the tensor pipe idled
// (measured 16 k + 11 k of 83 k cycles per tile in ffn_pair_kernel<true>).
//
// Weights stream through a ring of 13,824-byte slots in exactly the order they are consumed (a static program that
// producer, UMMA issuers and the peer's relay thread all walk): three dedicated slots, plus three more that live in the
// tail of the q/k/v staging area and are used only while that area holds the (smaller) hidden tiles of the FFN phase.
//
// Windows longer than one tile (128 < L <= 256, template kWide): the window-aligned layout uses Lw = 256, so a CTA pair
// owns exactly ONE window -- the leader its positions 0..127, the peer 128..255 (rows >= L are layout padding).  Every
// sub-layer except attention is row-wise and needs nothing else.  In attention the band reaches across the cut: the
// leader's last query block needs the peer's first 16 key / value rows and the peer's first block the leader's last
// 16.  That one warp per CTA loads the mma.sync B fragments of the foreign key tile straight from the partner's
// shared memory (ld.shared::cluster on mapa-translated addresses); `kv_peer` (one remote release-arrive per CTA and
// head, acquire-wait by the boundary warp) orders it after the partner's q/k/v staging, and the existing att_ready ->
// out-projection -> s_free chain keeps the partner from overwriting those rows before the reads are done.
//
// Cross-CTA protocol as in ffn_pair_kernel: the leader (cluster rank 0) issues every UMMA; tcgen05.commit multicasts
// completion to both CTAs; what the peer's warps produce is signalled with one default-semantics remote arrive per
// warp on the leader's barrier; the peer's relay thread forwards "my half of this stage has landed".
#pragma once

struct StackCfg {
  static constexpr int kABytes = (kDP / 8) * kTileM * 16;              // 73728: bf16 operand tile
  static constexpr int kStride = 152;                                  // padded q/k/v row (304 B, conflict-free)
  static constexpr int kMatBytes = kTileM * kStride * 2;               // 38912
  static constexpr int kStageBytes = 3 * kMatBytes;                    // 116736: q | k | v of one head
  static constexpr int kHBytes = (kFFChunk / 8) * kTileM * 16;         // 32768: one hidden chunk tile
  static constexpr int kSlotBytes = 13824;
  static constexpr int kSlotsA = 3, kSlotsF = 6;
  static constexpr int kQkvStepBytes = 2 * (kDHP / 2) * 16;            // 2304: one k-step of a q/k/v block (72 rows)
  static constexpr int kQkvStageK = 6, kQkvStages = 3;                 // 18 k-steps per block
  static constexpr int kQkvBlockBytes = (kDP / 16) * kQkvStepBytes;    // 41472
  static constexpr int kWoStepBytes = 2 * (kDP / 2) * 16;              // 4608: one k-step of Wo / W2 (144 rows)
  static constexpr int kWoStageK = 3, kWoStages = 3;                   // 9 k-steps per head
  static constexpr int kW1StepBytes = 2 * (kFFChunk / 2) * 16;         // 2048 (64 rows)
  static constexpr int kW1StageK = 6, kW1Stages = 3;
  static constexpr int kW2StageK = 2, kW2Stages = 4;
  static constexpr int kHalfChunkBytes = (kDP / 16) * kW1StepBytes + (kFFChunk / 16) * kWoStepBytes;   // 73728
  static constexpr int kOffA = 0;
  static constexpr int kOffS = kABytes;                                // staging: q|k|v, att_h, hidden tiles, LN stats
  static constexpr int kOffTail = kOffS + 2 * kHBytes;                 // ring slots 3..5 (FFN phase only)
  static constexpr int kOffRing = kOffS + kStageBytes;                 // ring slots 0..2
  static constexpr int kOffBars = kOffRing + kSlotsA * kSlotBytes;
  static constexpr int kSmemBytes = kOffBars + 256;
  static constexpr int kThreads = 384;   // warp 0 producer, 1 UMMA issuer A / relay, 2 UMMA issuer B, 3 prefetch, 4-11 workers
  static constexpr int kTmemY = 0, kTmemAcc = kDP, kTmemH = kDP, kTmemCols = 512;
};
static_assert(StackCfg::kSmemBytes <= 232448, "stack kernel shared memory budget");
static_assert(StackCfg::kOffTail + 3 * StackCfg::kSlotBytes <= StackCfg::kOffRing, "tail slots fit behind the hidden tiles");
static_assert(StackCfg::kQkvStageK * StackCfg::kQkvStepBytes == StackCfg::kSlotBytes, "qkv stage");
static_assert(StackCfg::kWoStageK * StackCfg::kWoStepBytes == StackCfg::kSlotBytes, "wo stage");
static_assert(StackCfg::kW1StageK * StackCfg::kW1StepBytes <= StackCfg::kSlotBytes, "w1 stage");
static_assert(StackCfg::kW2StageK * StackCfg::kWoStepBytes <= StackCfg::kSlotBytes, "w2 stage");
static_assert(StackCfg::kTmemAcc + kNC <= 512 && StackCfg::kTmemH + kFFChunk <= 512, "TMEM map");

__device__ __forceinline__ void ldmatrix_x4(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}

// Debug aid (-DDCB_WATCHDOG): every mbarrier wait gives up after ~50 M cycles, records (tag, parity) of EVERY waiting
// warp in g_ffn_trace[block*16 + warp] and lets the kernel run to completion (garbage results) so the host can read
// who was waiting on what (scripts/gpu_stack_check.py prints it).
#ifdef DCB_WATCHDOG
__device__ unsigned int g_stack_abort = 0;
template <bool kCluster>
__device__ __forceinline__ void stack_wait(uint64_t* bar, uint32_t parity, int tag) {
  const long long t0 = clock64();
  for (;;) {
    uint32_t ok;
    if (kCluster) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } else {
      ok = mbar_try_wait(bar, parity) ? 1u : 0u;
    }
    if (ok) return;
    if (clock64() - t0 > 50000000ll || *reinterpret_cast<volatile unsigned int*>(&g_stack_abort)) {
      g_stack_abort = 1;
      if (blockIdx.x < 256) {
        unsigned long long* tr = g_ffn_trace + blockIdx.x * 16 + (threadIdx.x >> 5);
        if (*tr == 0) *tr = 0x8000000000000000ull | ((unsigned long long)(clock64() - t0) << 24) | ((unsigned long long)parity << 16) | (unsigned long long)tag;
      }
      return;
    }
  }
}
#define SW(bar, par, tag) stack_wait<false>(bar, par, tag)
#define SWC(bar, par, tag) stack_wait<true>(bar, par, tag)
#else
#define SW(bar, par, tag) mbar_wait(bar, par)
#define SWC(bar, par, tag) mbar_wait_cluster(bar, par)
#endif

// remote (partner CTA) shared-memory loads for the wide-window attention halo
__device__ __forceinline__ uint32_t peer_smem_addr(const void* local, uint32_t peer_rank) {
  uint32_t raddr;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(smem_u32(local)), "r"(peer_rank));
  return raddr;
}
__device__ __forceinline__ uint32_t ld_peer_b32(uint32_t raddr) {
  uint32_t v;
  asm volatile("ld.shared::cluster.b32 %0, [%1];" : "=r"(v) : "r"(raddr) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t ld_peer_u16(uint32_t raddr) {
  uint16_t v;
  asm volatile("ld.shared::cluster.u16 %0, [%1];" : "=h"(v) : "r"(raddr) : "memory");
  return (uint32_t)v;
}
// release-arrive on the partner's barrier: everything this CTA wrote to its shared memory before (ordered by the
// preceding CTA barrier) is visible to a partner thread that acquire-waits on it
__device__ __forceinline__ void mbar_arrive_release_cluster(uint64_t* bar, uint32_t target_cta) {
  dcb_jitter();
  uint32_t raddr;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(smem_u32(bar)), "r"(target_cta));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(raddr) : "memory");
}

template <bool kWide>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(StackCfg::kThreads, 1)
stack_pair_kernel(float* __restrict__ xg, int ntiles, int L, int win, const __grid_constant__ StackParams P,
                  const __grid_constant__ HeadParams HP) {
  using C = StackCfg;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem + C::kOffA;
  uint8_t* sS = smem + C::kOffS;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kOffBars);
  uint64_t* full = bars;             // [6]
  uint64_t* empty = bars + 6;        // [6]
  uint64_t* a_ready = bars + 12;     // operand tile sA written by all workers of both CTAs (leader)
  uint64_t* acc_full = bars + 13;    // a q/k/v block landed in ACC (commit)
  uint64_t* acc_free = bars + 14;    // ... and has been read out by the workers (leader)
  uint64_t* att_ready = bars + 15;   // att_h written to the staging area (leader)
  uint64_t* s_free = bars + 16;      // out-proj of a head done: staging reusable / Y = x_mid after head 1 (commit)
  uint64_t* tail_free = bars + 17;   // attention part of the layer done: ring slots 3..5 usable (commit)
  uint64_t* h_full = bars + 18;
  uint64_t* h_free = bars + 19;
  uint64_t* hs_full = bars + 20;     // [2]
  uint64_t* hs_free = bars + 22;     // [2]
  uint64_t* y_full = bars + 24;      // FFN of the layer done (commit)
  uint64_t* kv_peer = bars + 25;     // kWide: the partner CTA has staged q/k/v of the current head (remote release-arrive)
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 26);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int npairs = (int)gridDim.x >> 1, pair = (int)blockIdx.x >> 1;
  const int tile_pairs = (ntiles + 1) >> 1;
  const int rounds = (tile_pairs + npairs - 1) / npairs;
  const int NL = P.num_layers;
  const int nchunks = P.ff / kFFChunk;
  auto tile_of = [&](int ti) { return ((ti * npairs + pair) << 1) + (int)rank; };
  auto slot_ptr = [&](int s) -> uint8_t* {
    return s < C::kSlotsA ? smem + C::kOffRing + s * C::kSlotBytes : smem + C::kOffTail + (s - C::kSlotsA) * C::kSlotBytes;
  };

  if (threadIdx.x == 0) {
    for (int i = 0; i < C::kSlotsF; ++i) { mbar_init(&full[i], leader ? 2 : 1); mbar_init(&empty[i], 1); }
    mbar_init(a_ready, 16);
    mbar_init(acc_full, 1);
    mbar_init(acc_free, 16);
    mbar_init(att_ready, 16);
    mbar_init(s_free, 1);
    mbar_init(tail_free, 1);
    mbar_init(h_full, 1);
    mbar_init(h_free, 16);
    mbar_init(&hs_full[0], 16);
    mbar_init(&hs_full[1], 16);
    mbar_init(&hs_free[0], 1);
    mbar_init(&hs_free[1], 1);
    mbar_init(y_full, 1);
    mbar_init(kv_peer, 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc_pair(tmem_holder, C::kTmemCols);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  if (warp < 4) {
    setmaxnreg_dec<56>();   // 128 x (168 - 56) registers released == 256 x (224 - 168) acquired by the workers
    if (warp == 0) {
      // ------------------------------------------------------------------ producer (both CTAs, own halves)
      if (lane == 0) {
        uint32_t par = 0;   // bit s: parity of the number of stages pushed into slot s
        auto push = [&](int s, const uint8_t* src, uint32_t bytes) {
          SW(&empty[s], ((par >> s) & 1) ^ 1, 701);
          mbar_arrive_expect_tx(&full[s], bytes);
          bulk_g2s(slot_ptr(s), src, bytes, &full[s]);
          par ^= 1u << s;
        };
        uint32_t li = 0;
        for (int ti = 0; ti < rounds; ++ti) {
          for (int n = 0; n < NL; ++n, ++li) {
            int sa = 0;
            // order of consumption: head 0 q,k,v | head 1 q (computed while the workers run the attention of head 0) |
            // Wo_0 | head 1 k,v | Wo_1
            auto push_blocks = [&](int h, int m_lo, int m_hi) {
              const uint8_t* wq = P.wq3[n] + (size_t)(h * 2 + rank) * 3 * C::kQkvBlockBytes;
              for (int m = m_lo; m < m_hi; ++m)
                for (int st = 0; st < C::kQkvStages; ++st) {
                  push(sa, wq + (size_t)m * C::kQkvBlockBytes + st * C::kSlotBytes, C::kSlotBytes);
                  sa = sa + 1 == C::kSlotsA ? 0 : sa + 1;
                }
            };
            auto push_wo = [&](int h) {
              const uint8_t* wo = P.wo2[n] + ((size_t)rank * kHeads + h) * C::kWoStages * C::kSlotBytes;
              for (int st = 0; st < C::kWoStages; ++st) {
                push(sa, wo + st * C::kSlotBytes, C::kSlotBytes);
                sa = sa + 1 == C::kSlotsA ? 0 : sa + 1;
              }
            };
#pragma unroll 1
            for (int step = 0; step < 5; ++step) {
              if (step == 2 || step == 4) push_wo(step == 4 ? 1 : 0);
              else push_blocks(step == 0 ? 0 : 1, step == 3 ? 1 : 0, step == 1 ? 1 : 3);
            }
            int sf = 0;
            bool tail_ok = false;
            auto pushf = [&](const uint8_t* src, uint32_t bytes) {
              if (sf >= C::kSlotsA && !tail_ok) { SW(tail_free, li & 1, 802); tail_ok = true; }
              push(sf, src, bytes);
              sf = sf + 1 == C::kSlotsF ? 0 : sf + 1;
            };
            auto half = [&](int c) { return P.wffn2[n] + ((size_t)c * 2 + rank) * C::kHalfChunkBytes; };
            auto push_w1 = [&](int c) {
              for (int s = 0; s < C::kW1Stages; ++s)
                pushf(half(c) + s * (C::kW1StageK * C::kW1StepBytes), C::kW1StageK * C::kW1StepBytes);
            };
            auto push_w2 = [&](int c) {
              const uint8_t* src = half(c) + (kDP / 16) * C::kW1StepBytes;
              for (int s = 0; s < C::kW2Stages; ++s)
                pushf(src + s * (C::kW2StageK * C::kWoStepBytes), C::kW2StageK * C::kWoStepBytes);
            };
            push_w1(0);
            for (int c = 0; c < nchunks; ++c) {
              if (c + 1 < nchunks) push_w1(c + 1);
              push_w2(c);
            }
          }
        }
      }
    } else if (warp == 1 || warp == 2) {
      if (leader || lane == 0) {   // leader: the whole warp walks the issue program (elected lane issues); peer: one relay thread
        if (leader) {
          // ---------------------------------------------------------------- UMMA issuers (leader only)
          // warp 1: q/k/v blocks, out-projections and every GEMM1; warp 2: every GEMM2 (the two share the tensor
          // pipe; causality between them runs through the workers: hs_full follows h_full follows GEMM1).
          constexpr uint32_t idesc_h = make_idesc_bf16(2 * kTileM, kFFChunk);
          constexpr uint32_t idesc_y = make_idesc_bf16(2 * kTileM, kNC);
          constexpr uint16_t kBoth = 3;
          const uint32_t s_addr = smem_u32(sS);
          // descriptors differ only in their 14-bit start-address field (bytes >> 4): build each family once and add
          const uint64_t adesc_a = make_kc16_desc(smem_u32(sA), kTileM * 16, 128);        // operand tile sA, + kstep * 256
          const uint64_t adesc_s = make_kc16_desc(s_addr, kTileM * 16, 128);              // att_h / hidden tiles in the staging area
          const uint64_t bdesc_72 = make_kc16_desc(0, (kDHP / 2) * 16, 128);              // + slot address >> 4
          const uint64_t bdesc_144 = make_kc16_desc(0, (kDP / 2) * 16, 128);
          const uint64_t bdesc_64 = make_kc16_desc(0, (kFFChunk / 2) * 16, 128);
          uint32_t cpar = 0;   // bit s: parity of the number of stages consumed from slot s (by either issuer)
          // Plain (CTA-scope) waits also on barriers the peer arrives on, as CUTLASS' ClusterBarrier::wait does: a
          // cluster-scope acquire makes ptxas put CCTL.IVALL (invalidate L1) into every poll, which evicted the
          // workers' bias / LayerNorm vectors continuously.
          auto use = [&](int s) -> uint32_t {
            SW(&full[s], (cpar >> s) & 1, 103);
            tc_fence_after();
            return smem_u32(slot_ptr(s)) >> 4;   // in descriptor address units
          };
          auto release = [&](int s) {
            umma_commit_pair_warp(&empty[s], kBoth);
            cpar ^= 1u << s;
          };
          uint32_t nn0 = 0;
          if (warp == 1) {
            uint32_t kblk = 0, katt = 0, kar = 0;
            long long t_ar1 = 0, t_accfree = 0, t_qkv = 0, t_attw = 0, t_oproj = 0, t_ar2 = 0, t_ffn = 0;
            const long long t_begin = clock64();
            for (int ti = 0; ti < rounds; ++ti) {
              for (int n = 0; n < NL; ++n) {
                TRACE_T0();
                SW(a_ready, kar & 1, 204); ++kar;
                TRACE_ADD(t_ar1);
                tc_fence_after();
                int sa = 0;
                auto qkv_blocks = [&](int count) {
                  for (int m = 0; m < count; ++m) {
                    SW(acc_free, (kblk & 1) ^ 1, 305); ++kblk;
                    TRACE_ADD(t_accfree);
                    tc_fence_after();
                    for (int st = 0; st < C::kQkvStages; ++st) {
                      const uint32_t sb = use(sa);
#pragma unroll
                      for (int kk = 0; kk < C::kQkvStageK; ++kk) {
                        const int kstep = st * C::kQkvStageK + kk;
                        umma_bf16_ss_pair_warp(tmem_base + C::kTmemAcc, adesc_a + kstep * 256,
                                               bdesc_72 + (sb + kk * (C::kQkvStepBytes >> 4)), idesc_y, kstep != 0);
                      }
                      release(sa);
                      sa = sa + 1 == C::kSlotsA ? 0 : sa + 1;
                    }
                    umma_commit_pair_warp(acc_full, kBoth);
                    TRACE_ADD(t_qkv);
                  }
                };
                auto out_proj = [&](int h) {
                  // Y += att_h * Wo_h^T
                  SW(att_ready, katt & 1, 406); ++katt;
                  TRACE_ADD(t_attw);
                  tc_fence_after();
                  for (int st = 0; st < C::kWoStages; ++st) {
                    const uint32_t sb = use(sa);
#pragma unroll
                    for (int kk = 0; kk < C::kWoStageK; ++kk) {
                      const int kstep = st * C::kWoStageK + kk;
#pragma unroll
                      for (int j = 0; j < 2; ++j)
                        umma_bf16_ss_pair_warp(tmem_base + C::kTmemY + j * kNC, adesc_s + h * (C::kMatBytes >> 4) + kstep * 256,
                                               bdesc_144 + (sb + kk * (C::kWoStepBytes >> 4) + j * (kNC / 2)), idesc_y, true);
                    }
                    release(sa);
                    sa = sa + 1 == C::kSlotsA ? 0 : sa + 1;
                  }
                  umma_commit_pair_warp(s_free, kBoth);
                  if (h == kHeads - 1) umma_commit_pair_warp(tail_free, kBoth);
                  TRACE_ADD(t_oproj);
                };
                // head 0: q,k,v | head 1: q (its block sits in ACC while the workers run the attention of head 0) | Wo_0 |
                // head 1: k,v | Wo_1 -- one call site each
#pragma unroll 1
                for (int step = 0; step < 5; ++step) {
                  if (step == 2 || step == 4) out_proj(step == 4 ? 1 : 0);
                  else qkv_blocks(step == 0 ? 3 : (step == 1 ? 1 : 2));
                }
                // ---- FFN: GEMM1 of every chunk
                SW(a_ready, kar & 1, 207); ++kar;
                    TRACE_ADD(t_ar2);
                tc_fence_after();
                int sf = 0;
                auto gemm1 = [&](uint32_t nn) {
                  SW(h_free, (nn & 1) ^ 1, 508);
                  tc_fence_after();
                  for (int s = 0; s < C::kW1Stages; ++s) {
                    const uint32_t sb = use(sf);
#pragma unroll
                    for (int kk = 0; kk < C::kW1StageK; ++kk) {
                      const int kstep = s * C::kW1StageK + kk;
                      umma_bf16_ss_pair_warp(tmem_base + C::kTmemH, adesc_a + kstep * 256,
                                             bdesc_64 + (sb + kk * (C::kW1StepBytes >> 4)), idesc_h, kstep != 0);
                    }
                    release(sf);
                    sf = sf + 1 == C::kSlotsF ? 0 : sf + 1;
                  }
                  umma_commit_pair_warp(h_full, kBoth);
                };
                auto skipf = [&](int count) {
                  for (int s = 0; s < count; ++s) { cpar ^= 1u << sf; sf = sf + 1 == C::kSlotsF ? 0 : sf + 1; }
                };
                gemm1(nn0);
                for (int c = 0; c < nchunks; ++c) {
                  if (c + 1 < nchunks) gemm1(nn0 + c + 1);
                  skipf(C::kW2Stages);
                }
                nn0 += nchunks;
                TRACE_ADD(t_ffn);
              }
            }
#ifdef DCB_TRACE
            if (blockIdx.x < 256) {
              unsigned long long* tr = g_ffn_trace + blockIdx.x * 16;
              tr[0] = clock64() - t_begin; tr[1] = t_ar1; tr[2] = t_accfree; tr[3] = t_qkv; tr[4] = t_attw; tr[5] = t_oproj;
              tr[6] = t_ar2; tr[7] = t_ffn;
            }
#endif
          } else {
            for (int ti = 0; ti < rounds; ++ti) {
              for (int n = 0; n < NL; ++n) {
                // the attention part uses each of slots 0..2 an even number of times: parities unchanged
                int sf = 0;
                auto skipf = [&](int count) {
                  for (int s = 0; s < count; ++s) { cpar ^= 1u << sf; sf = sf + 1 == C::kSlotsF ? 0 : sf + 1; }
                };
                skipf(C::kW1Stages);
                for (int c = 0; c < nchunks; ++c) {
                  if (c + 1 < nchunks) skipf(C::kW1Stages);
                  const uint32_t nn = nn0 + c, b = nn & 1;
                  SW(&hs_full[b], (nn >> 1) & 1, 609);
                  tc_fence_after();
                  const uint64_t adesc_h = adesc_s + b * (C::kHBytes >> 4);
                  for (int s = 0; s < C::kW2Stages; ++s) {
                    const uint32_t sb = use(sf);
#pragma unroll
                    for (int kk = 0; kk < C::kW2StageK; ++kk) {
                      const int kstep = s * C::kW2StageK + kk;
#pragma unroll
                      for (int j = 0; j < 2; ++j)
                        umma_bf16_ss_pair_warp(tmem_base + C::kTmemY + j * kNC, adesc_h + kstep * 256,
                                               bdesc_144 + (sb + kk * (C::kWoStepBytes >> 4) + j * (kNC / 2)), idesc_y, true);
                    }
                    release(sf);
                    sf = sf + 1 == C::kSlotsF ? 0 : sf + 1;
                  }
                  umma_commit_pair_warp(&hs_free[b], kBoth);
                }
                umma_commit_pair_warp(y_full, kBoth);
                nn0 += nchunks;
              }
            }
          }
        } else if (warp == 1) {
          // ---------------------------------------------------------------- relay (peer): forward "my half of
          // this stage has landed" to the leader's barrier, in consumption order
          uint32_t cpar = 0;
          auto relay = [&](int s) {
            SW(&full[s], (cpar >> s) & 1, 110);
            mbar_arrive_cluster(&full[s], 0);
            cpar ^= 1u << s;
          };
          for (int ti = 0; ti < rounds; ++ti)
            for (int n = 0; n < NL; ++n) {
              int sa = 0;
              for (int i = 0; i < kHeads * (3 * C::kQkvStages + C::kWoStages); ++i) { relay(sa); sa = sa + 1 == C::kSlotsA ? 0 : sa + 1; }
              int sf = 0;
              const int nf = C::kW1Stages * nchunks + C::kW2Stages * nchunks;
              for (int i = 0; i < nf; ++i) { relay(sf); sf = sf + 1 == C::kSlotsF ? 0 : sf + 1; }
            }
        }
      }
    } else {
      // warp 3: pull the next tile's residual image into L2 (it is read once per tile, by the workers' init pass)
      for (int ti = 0; ti + 1 < rounds; ++ti) {
        const int tile = min(tile_of(ti + 1), ntiles - 1);
        const uint8_t* base = reinterpret_cast<const uint8_t*>(xg + (size_t)tile * x_image_elems());
        for (int ln = lane; ln < (int)(x_image_elems() * 4 / 128); ln += 32)
          asm volatile("prefetch.global.L2 [%0];" ::"l"(base + (size_t)ln * 128));
        // pace: one tile ahead -- wait for the last layer of this round's tile to finish
        for (int n = 0; n < NL; ++n) SW(y_full, (uint32_t)(ti * NL + n) & 1, 911);
      }
    }
  } else {
    // -------------------------------------------------------------------- workers (8 warps, both CTAs)
    setmaxnreg_inc<224>();
    const int ew = warp - 4;
    const int q = warp & 3;
    const int r = q * 32 + lane;                  // token row (TMEM lane) of this thread
    const int halfsel = ew >> 2;                  // which half of the columns this thread handles in row passes
    const uint32_t tmem_row = tmem_base + ((uint32_t)(q * 32) << 16);
    const int g = lane >> 2, t = lane & 3;
    const int band = win;
    constexpr float kLog2e = 1.4426950408889634f;
    constexpr int kS = C::kStride;
    constexpr int kChunkElems = kTileM * 8;
    __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(sS);
    __nv_bfloat16* sK = sQ + kTileM * kS;
    __nv_bfloat16* sV = sK + kTileM * kS;
    float4* sStat = reinterpret_cast<float4*>(sS);   // [2][128] LayerNorm partial statistics
    const int cb0 = halfsel * 9;                  // this thread's 9 column blocks (of 16) of Y
    auto bar_red_or_workers = [&](bool pred) -> bool {   // OR of `pred` over the 256 worker threads (named barrier 1)
      uint32_t any;
      asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.b32 q, %1, 0;\n\tbar.red.or.pred p, 1, 256, q;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(any) : "r"((uint32_t)pred) : "memory");
      return any != 0;
    };
    auto arrive_leader = [&](uint64_t* bar) {     // one arrive per warp on the leader's barrier
      __syncwarp();
      if (lane == 0) {
        if (leader) mbar_arrive(bar); else mbar_arrive_cluster(bar, 0);
      }
    };

    // Row pass: Y (+= bias, written back) -> bf16 operand tile sA.  Two threads per row.
    // The loops are software-pipelined over two register buffers: the tcgen05.ld of column block cb + 1 is in flight
    // while block cb is processed (the serial ld -> wait -> math -> store form exposed one TMEM round trip per block:
    // measured 5.2 k cycles for 9 blocks).
    //
    // Pre-LayerNorm models: the normalisation is DEFERRED so that Y is read from TMEM once, not twice (common.h,
    // StackParams::deferred_ln).  The operand tile is bf16(x - shift) with shift = the row's exact mean at the PREVIOUS row
    // pass (for the first pass of a tile: summed while the tile is loaded); the weights carry gamma in rows 0..279 and the column
    // sums / beta^T W (+ b1) in the padding rows 280..287, against which this pass writes -(mean - shift) and 1 / rstd as
    // bf16 hi / lo pairs.  The accumulator then holds (LN(x) W + bw) / rstd and its reader multiplies by ln_rstd.
    // Centring on the previous mean keeps |x - shift| ~ |x - mean|, so the bf16 rounding error is that of the normalised
    // activations whatever the row's offset.
    float ln_shift = 0.f, ln_rstd = 1.f;
    auto row_pass = [&](const float* __restrict__ bias, const float bias_mean, const bool ln) {
      const uint32_t ycol = tmem_row + C::kTmemY + cb0 * 16;
      uint4* arow = reinterpret_cast<uint4*>(sA) + r;
      float s1 = 0.f, s2 = 0.f;
      const float* __restrict__ bias_now = bias;          // added (and stored back) by the first sweep only
      if (ln && bias) ln_shift += bias_mean;              // the row's mean moves by the mean of the bias that joins here
      auto emit = [&](uint32_t (&acc)[16], int cb) {
        if (bias_now) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias_now + (cb0 + cb) * 16) + i);
            acc[4 * i + 0] = __float_as_uint(__uint_as_float(acc[4 * i + 0]) + b4.x);
            acc[4 * i + 1] = __float_as_uint(__uint_as_float(acc[4 * i + 1]) + b4.y);
            acc[4 * i + 2] = __float_as_uint(__uint_as_float(acc[4 * i + 2]) + b4.z);
            acc[4 * i + 3] = __float_as_uint(__uint_as_float(acc[4 * i + 3]) + b4.w);
          }
          tmem_st16(ycol + cb * 16, acc);
        }
        float v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int col = (cb0 + cb) * 16 + i;
          v[i] = col < kD ? __uint_as_float(acc[i]) - ln_shift : 0.f;     // ln_shift == 0 for ReZero models
          if (ln) { s1 += v[i]; s2 = fmaf(v[i], v[i], s2); }
        }
        arow[(size_t)((cb0 + cb) * 2) * kTileM] = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]),
                                                             pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
        arow[(size_t)((cb0 + cb) * 2 + 1) * kTileM] = make_uint4(pack_bf16x2(v[8], v[9]), pack_bf16x2(v[10], v[11]),
                                                                 pack_bf16x2(v[12], v[13]), pack_bf16x2(v[14], v[15]));
      };
      float dmean = 0.f, sd = 1.f;
#pragma unroll 1
      for (int sweep = 0; sweep < 2; ++sweep) {
        s1 = 0.f; s2 = 0.f;
        uint32_t a[16], b[16];
        tmem_ld16(ycol, a);
#pragma unroll 1
        for (int cb = 0; cb < 8; cb += 2) {
          tmem_ld_wait();
          tmem_ld16(ycol + (cb + 1) * 16, b);
          emit(a, cb);
          tmem_ld_wait();
          tmem_ld16(ycol + (cb + 2) * 16, a);
          emit(b, cb + 1);
        }
        tmem_ld_wait();
        emit(a, 8);
        if (bias_now) { tmem_st_wait(); bias_now = nullptr; }
        if (!ln) break;
        // statistics of (x - shift) over the row's two halves; the staging area is idle until a_ready is signalled
        float4* slot = sStat + (sweep ? 4 : 0) * kTileM;
        slot[halfsel * kTileM + r] = make_float4(s1, s2, 0.f, 0.f);
        asm volatile("bar.sync 1, 256;" ::: "memory");
        const float4 o = slot[(1 - halfsel) * kTileM + r];
        dmean = (s1 + o.x) * (1.f / kD);                                  // mean - shift
        const float var = fmaxf((s2 + o.y) * (1.f / kD) - dmean * dmean, 0.f);
        sd = sqrtf(var + 1e-6f);
        if (sweep) break;
        // Guard: the tile was rounded around `shift`.  If a row's mean has moved by more than its standard deviation since
        // the previous pass, |x - shift| is no longer ~ |x - mean| and the bf16 rounding would cost precision: every
        // worker then sweeps once more, the rows concerned centred on their exact mean (CTA-wide vote; rare).
        const bool far = dmean * dmean > var;
        if (!bar_red_or_workers(far)) break;
        if (far) ln_shift += dmean;
      }
      if (ln) {
        ln_rstd = 1.f / sd;
        ln_shift += dmean;                                                // this pass's mean: the next pass's shift
      }
      if (halfsel) {
        // operand columns 280..287 against the weights' padding rows: -dmean and 1 / rstd, each as hi, hi, lo, lo
        // (ReZero: 0 and 1 -- the rows then only add b1 in the FFN)
        const float dh = __bfloat162float(__float2bfloat16(-dmean)), ih = __bfloat162float(__float2bfloat16(sd));
        arow[(size_t)(kD / 8) * kTileM] = make_uint4(pack_bf16x2(dh, dh), pack_bf16x2(-dmean - dh, -dmean - dh),
                                                     pack_bf16x2(ih, ih), pack_bf16x2(sd - ih, sd - ih));
      }
      tc_fence_before();
      fence_proxy_async_smem();
      arrive_leader(a_ready);
    };

    uint32_t k_acc = 0, k_sfree = 0, k_y = 0, nchunk = 0, k_kv = 0;
#ifdef DCB_TRACE
    // worker-side cycle trace (warp 4 of every CTA): slots 8..15 of g_ffn_trace
    long long w_rowpass = 0, w_stage = 0, w_qk = 0, w_soft = 0, w_pv = 0, w_store = 0, w_wait = 0, w_hid = 0;
#endif
    const int goff = kWide ? (int)rank * kTileM : 0;   // window position of this CTA's row 0
    for (int ti = 0; ti < rounds; ++ti) {
      const int tile_raw = tile_of(ti);
      const bool valid = tile_raw < ntiles;
      const int tile = min(tile_raw, ntiles - 1);
      float4* xrow = reinterpret_cast<float4*>(xg + (size_t)tile * x_image_elems()) + r;
      // ---- Y <- x (this thread's half row); pre-LayerNorm models: the row's mean on the way, as the first centring shift
      float xsum = 0.f;
#pragma unroll 3
      for (int cb = 0; cb < 9; ++cb) {
        uint32_t v[16];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 x4 = xrow[(size_t)((cb0 + cb) * 4 + i) * kTileM];     // padding columns of the image are zero
          xsum += (x4.x + x4.y) + (x4.z + x4.w);
          v[4 * i + 0] = __float_as_uint(x4.x); v[4 * i + 1] = __float_as_uint(x4.y);
          v[4 * i + 2] = __float_as_uint(x4.z); v[4 * i + 3] = __float_as_uint(x4.w);
        }
        tmem_st16(tmem_row + C::kTmemY + (cb0 + cb) * 16, v);
      }
      if (P.deferred_ln) {
        sStat[2 * kTileM + halfsel * kTileM + r] = make_float4(xsum, 0.f, 0.f, 0.f);
        asm volatile("bar.sync 1, 256;" ::: "memory");
        ln_shift = (xsum + sStat[2 * kTileM + (1 - halfsel) * kTileM + r].x) * (1.f / kD);
      }
      tmem_st_wait();

      for (int n = 0; n < NL; ++n) {
        const bool ln = P.deferred_ln != 0;           // pre-LayerNorm model (deferred normalisation, see row_pass)
        // ---- P1: operand tile of the attention sub-layer (+ b2 of the previous layer's FFN)
        if (n > 0) {
          SW(y_full, k_y & 1, 912); ++k_y;
          tc_fence_after();
        }
        { TRACE_T0(); row_pass(n > 0 ? P.b2[n - 1] : nullptr, n > 0 ? P.b2_mean[n - 1] : 0.f, ln); TRACE_ADD(w_rowpass); }

        for (int h = 0; h < kHeads; ++h) {
          TRACE_T0();
          // Staging areas rotate with the head: matrix m of head h lives in area (m + h) % 3, att_h over its q area.
          // Head 1's q and k therefore go over head 0's k and v (free once every worker of this CTA has finished the
          // attention of head 0) and only its v goes over att_0 (free once the out-projection of head 0 has read it): the
          // drain of the q block of head 1 overlaps the out-projection UMMAs instead of waiting for them.
          __nv_bfloat16* const qb = sQ + (size_t)((0 + h) % 3) * kTileM * kS;
          __nv_bfloat16* const kb = sQ + (size_t)((1 + h) % 3) * kTileM * kS;
          __nv_bfloat16* const vb = sQ + (size_t)((2 + h) % 3) * kTileM * kS;
          if (h > 0) {
            if (kWide) { SW(s_free, k_sfree & 1, 1013); ++k_sfree; }   // the partner's boundary warp may still read my k / v rows
            else asm volatile("bar.sync 1, 256;" ::: "memory");
          }
          TRACE_ADD(w_wait);
          // ---- q, k, v blocks: TMEM -> bf16 -> padded shared-memory rows
          for (int m = 0; m < 3; ++m) {
            SW(acc_full, k_acc & 1, 1114); ++k_acc;
            tc_fence_after();
            const int b0 = halfsel ? 5 : 0, nb = halfsel ? 4 : 5;
            uint32_t acc[5][16];
#pragma unroll
            for (int j = 0; j < 5; ++j)
              if (j < nb) tmem_ld16(tmem_row + C::kTmemAcc + (b0 + j) * 16, acc[j]);
            tmem_ld_wait();
            tc_fence_before();
            arrive_leader(acc_free);
            if (!kWide && h > 0 && m == 2) { SW(s_free, k_sfree & 1, 1013); ++k_sfree; }   // att_{h-1} consumed by its out-projection
            __nv_bfloat16* dst = sQ + (size_t)((m + h) % 3) * kTileM * kS + (size_t)r * kS;
#pragma unroll
            for (int j = 0; j < 5; ++j) {
              if (j < nb) {
                float v[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = ln_rstd * __uint_as_float(acc[j][i]);     // ln_rstd == 1 for ReZero
                *reinterpret_cast<uint4*>(dst + (2 * (b0 + j)) * 8) =
                    make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
                *reinterpret_cast<uint4*>(dst + (2 * (b0 + j) + 1) * 8) =
                    make_uint4(pack_bf16x2(v[8], v[9]), pack_bf16x2(v[10], v[11]), pack_bf16x2(v[12], v[13]), pack_bf16x2(v[14], v[15]));
              }
            }
          }
          asm volatile("bar.sync 1, 256;" ::: "memory");
          TRACE_ADD(w_stage);
          if (kWide && ew == 0 && lane == 0) mbar_arrive_release_cluster(kv_peer, rank ^ 1u);   // my q/k/v of this head are staged

          // ---- banded attention of query block `ew` (rows 16*ew .. +15), two-pass softmax (win <= 16)
          {
            const int i0 = ew * 16;
            const int r0 = i0 + g, r1 = r0 + 8;
            // A fragments of the 16 query rows: one ldmatrix.x4 per 16 dims (matrices: rows 0-7 | 8-15 x dims 0-7 | 8-15)
            const uint32_t qaddr = smem_u32(qb + (size_t)(i0 + (lane & 7) + ((lane >> 3) & 1) * 8) * kS + (lane >> 4) * 8);
            uint32_t qa[kDHP / 16][4];
#pragma unroll
            for (int ks = 0; ks < kDHP / 16; ++ks) ldmatrix_x4(qa[ks][0], qa[ks][1], qa[ks][2], qa[ks][3], qaddr + ks * 32);
            float o[kDHP / 8][4];
#pragma unroll
            for (int nt = 0; nt < kDHP / 8; ++nt) { o[nt][0] = o[nt][1] = o[nt][2] = o[nt][3] = 0.f; }
            float l0 = 0.f, l1 = 0.f;
            // key range in window coordinates (gq = window position of the block's first query); a tile whose local
            // row index falls outside [0, 128) lives in the partner CTA (kWide only: its rows 0..15 or 112..127)
            const int gq = goff + i0;
            int jlo = gq - band; if (jlo < 0) jlo = 0; jlo &= ~15;
            int jhi = gq + 15 + band + 1; if (jhi > L) jhi = L;
            const int nkt = (jhi - jlo + 15) >> 4;
            constexpr int kMaxKT = 3;
            const int lj0 = jlo - goff;                 // local row of key tile 0 (multiple of 16; -16 possible when kWide)
            auto tile_remote = [&](int kt) { return kWide && (lj0 + kt * 16 < 0 || lj0 + kt * 16 >= kTileM); };
            auto peer_row = [&](int kt) { return lj0 + kt * 16 < 0 ? lj0 + kt * 16 + kTileM : lj0 + kt * 16 - kTileM; };
            if (kWide) {
              bool any_remote = false;
#pragma unroll
              for (int kt = 0; kt < kMaxKT; ++kt) any_remote |= (kt < nkt) && tile_remote(kt);
              if (any_remote) SWC(kv_peer, k_kv & 1, 1418);   // the partner's k / v rows of this head are in place
            }
            float sc[kMaxKT][2][4];
#pragma unroll
            for (int kt = 0; kt < kMaxKT; ++kt)
#pragma unroll
              for (int nt = 0; nt < 2; ++nt) sc[kt][nt][0] = sc[kt][nt][1] = sc[kt][nt][2] = sc[kt][nt][3] = 0.f;
#pragma unroll
            for (int kt = 0; kt < kMaxKT; ++kt) {
              if (kt < nkt) {
                if (tile_remote(kt)) {
                  // the partner's key rows: B fragments by 32-bit remote loads (b0 = K[key g][dims 2t, 2t+1], b1 = dims + 8),
                  // all issued before the first use
                  const uint32_t kp = peer_smem_addr(kb + (size_t)(peer_row(kt) + g) * kS + 2 * t, rank ^ 1u);
                  uint32_t kf[kDHP / 16][4];
#pragma unroll
                  for (int ks = 0; ks < kDHP / 16; ++ks) {
                    kf[ks][0] = ld_peer_b32(kp + ks * 32);
                    kf[ks][1] = ld_peer_b32(kp + ks * 32 + 16);
                    kf[ks][2] = ld_peer_b32(kp + 8 * kS * 2 + ks * 32);
                    kf[ks][3] = ld_peer_b32(kp + 8 * kS * 2 + ks * 32 + 16);
                  }
#pragma unroll
                  for (int ks = 0; ks < kDHP / 16; ++ks) {
                    mma_bf16_16816(sc[kt][0], qa[ks], kf[ks][0], kf[ks][1]);
                    mma_bf16_16816(sc[kt][1], qa[ks], kf[ks][2], kf[ks][3]);
                  }
                } else {
                  // B fragments of 16 keys: one ldmatrix.x4 per 16 dims (keys 0-7 x dims 0-7 | 8-15, keys 8-15 x dims 0-7 | 8-15)
                  const uint32_t kaddr = smem_u32(kb + (size_t)(lj0 + kt * 16 + (lane & 7) + (lane >> 4) * 8) * kS + ((lane >> 3) & 1) * 8);
#pragma unroll
                  for (int ks = 0; ks < kDHP / 16; ++ks) {
                    uint32_t a0, a1, c0, c1;
                    ldmatrix_x4(a0, a1, c0, c1, kaddr + ks * 32);
                    mma_bf16_16816(sc[kt][0], qa[ks], a0, a1);
                    mma_bf16_16816(sc[kt][1], qa[ks], c0, c1);
                  }
                }
              }
            }
            TRACE_ADD(w_qk);
            float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
            for (int kt = 0; kt < kMaxKT; ++kt)
#pragma unroll
              for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const int i = goff + ((e < 2) ? r0 : r1);
                  const int j = jlo + kt * 16 + nt * 8 + 2 * t + (e & 1);
                  const int dlt = i - j;
                  const bool ok = (kt < nkt) && (j < L) && (dlt <= band) && (dlt >= -band);
                  const float v = ok ? sc[kt][nt][e] : -INFINITY;
                  sc[kt][nt][e] = v;
                  if (e < 2) mx0 = fmaxf(mx0, v); else mx1 = fmaxf(mx1, v);
                }
            mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
            mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
            mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
            mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
            const float base0 = mx0 == -INFINITY ? 0.f : mx0 * kLog2e, base1 = mx1 == -INFINITY ? 0.f : mx1 * kLog2e;
            uint32_t pa[kMaxKT][4];
#pragma unroll
            for (int kt = 0; kt < kMaxKT; ++kt) {
              float p[2][4];
#pragma unroll
              for (int nt = 0; nt < 2; ++nt) {
                p[nt][0] = exp2f(fmaf(sc[kt][nt][0], kLog2e, -base0));
                p[nt][1] = exp2f(fmaf(sc[kt][nt][1], kLog2e, -base0));
                p[nt][2] = exp2f(fmaf(sc[kt][nt][2], kLog2e, -base1));
                p[nt][3] = exp2f(fmaf(sc[kt][nt][3], kLog2e, -base1));
                l0 += p[nt][0] + p[nt][1];
                l1 += p[nt][2] + p[nt][3];
              }
              pa[kt][0] = pack_bf16x2(p[0][0], p[0][1]);
              pa[kt][1] = pack_bf16x2(p[0][2], p[0][3]);
              pa[kt][2] = pack_bf16x2(p[1][0], p[1][1]);
              pa[kt][3] = pack_bf16x2(p[1][2], p[1][3]);
            }
            TRACE_ADD(w_soft);
#pragma unroll
            for (int kt = 0; kt < kMaxKT; ++kt) {
              if (kt < nkt) {
                if (tile_remote(kt)) {
                  // the partner's value rows: b0 = {V[key 2t][dim g], V[key 2t+1][dim g]}, b1 = keys + 8 -- 16-bit remote
                  // loads, six dim blocks (24 loads) in flight at a time
                  const uint32_t vp = peer_smem_addr(vb + (size_t)(peer_row(kt) + 2 * t) * kS + g, rank ^ 1u);
#pragma unroll
                  for (int nb = 0; nb < kDHP / 8; nb += 6) {
                    uint32_t vf[6][4];
#pragma unroll
                    for (int u = 0; u < 6; ++u) {
                      const uint32_t a = vp + (nb + u) * 16;
                      vf[u][0] = ld_peer_u16(a);
                      vf[u][1] = ld_peer_u16(a + kS * 2);
                      vf[u][2] = ld_peer_u16(a + 8 * kS * 2);
                      vf[u][3] = ld_peer_u16(a + 9 * kS * 2);
                    }
#pragma unroll
                    for (int u = 0; u < 6; ++u)
                      mma_bf16_16816(o[nb + u], pa[kt], vf[u][0] | (vf[u][1] << 16), vf[u][2] | (vf[u][3] << 16));
                  }
                } else {
                  const int vrow = lj0 + kt * 16 + (lane & 15);
                  const uint32_t vbase = smem_u32(vb + (size_t)vrow * kS);
#pragma unroll
                  for (int nt = 0; nt < kDHP / 8; ++nt) {
                    uint32_t b0, b1;
                    ldmatrix_x2_trans(b0, b1, vbase + nt * 16);
                    mma_bf16_16816(o[nt], pa[kt], b0, b1);
                  }
                }
              }
            }
            TRACE_ADD(w_pv);
            l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
            l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
            l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
            l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
            // rows beyond the window (layout padding) contribute nothing to the out-projection
            const float inv0 = goff + r0 < L ? 1.f / l0 : 0.f, inv1 = goff + r1 < L ? 1.f / l1 : 0.f;
            // att_h as a KC16 operand tile [18 chunks][128 rows][8] over the (consumed) q area: every warp must hold its q
            // fragments before the first store (they were loaded at the start of the phase, so nobody waits here)
            asm volatile("bar.sync 2, 256;" ::: "memory");
            __nv_bfloat16* obase = qb + 2 * t;
#pragma unroll
            for (int nt = 0; nt < kDHP / 8; ++nt) {
              *reinterpret_cast<uint32_t*>(obase + (size_t)nt * kChunkElems + r0 * 8) = pack_bf16x2(o[nt][0] * inv0, o[nt][1] * inv0);
              *reinterpret_cast<uint32_t*>(obase + (size_t)nt * kChunkElems + r1 * 8) = pack_bf16x2(o[nt][2] * inv1, o[nt][3] * inv1);
            }
          }
          fence_proxy_async_smem();
          arrive_leader(att_ready);
          TRACE_ADD(w_store);
          if (kWide) ++k_kv;
        }
        { TRACE_T0(); SW(s_free, k_sfree & 1, 1015); ++k_sfree; TRACE_ADD(w_wait); }   // out-projection of the last head done: Y = x_mid
        tc_fence_after();

        // ---- P5: operand tile of the FFN
        { TRACE_T0(); row_pass(nullptr, 0.f, ln); TRACE_ADD(w_rowpass); }

        // ---- hidden-chunk epilogue: H (+b1, ReLU) -> bf16 -> shared memory operand of GEMM2
        for (int c = 0; c < nchunks; ++c, ++nchunk) {
          const uint32_t b = nchunk & 1;
          SW(h_full, nchunk & 1, 1216);
          tc_fence_after();
          SW(&hs_free[b], ((nchunk >> 1) & 1) ^ 1, 1317);
          TRACE_T0();
          uint4* hrow = reinterpret_cast<uint4*>(sS + b * C::kHBytes) + r;
          uint32_t acc[kFFChunk / 32][16];
#pragma unroll
          for (int cc = 0; cc < kFFChunk / 32; ++cc)
            tmem_ld16(tmem_row + C::kTmemH + (halfsel * (kFFChunk / 32) + cc) * 16, acc[cc]);
          tmem_ld_wait();
          tc_fence_before();
          arrive_leader(h_free);
#pragma unroll
          for (int cc = 0; cc < kFFChunk / 32; ++cc) {
            const int cb = halfsel * (kFFChunk / 32) + cc;
            float v[16];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              // b1 sits in the padding rows of W1 (StackParams): H = relu(rstd * acc), nothing to load
              v[4 * i + 0] = fmaxf(ln_rstd * __uint_as_float(acc[cc][4 * i + 0]), 0.f);
              v[4 * i + 1] = fmaxf(ln_rstd * __uint_as_float(acc[cc][4 * i + 1]), 0.f);
              v[4 * i + 2] = fmaxf(ln_rstd * __uint_as_float(acc[cc][4 * i + 2]), 0.f);
              v[4 * i + 3] = fmaxf(ln_rstd * __uint_as_float(acc[cc][4 * i + 3]), 0.f);
            }
            hrow[(size_t)(cb * 2) * kTileM] = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]),
                                                         pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
            hrow[(size_t)(cb * 2 + 1) * kTileM] = make_uint4(pack_bf16x2(v[8], v[9]), pack_bf16x2(v[10], v[11]),
                                                             pack_bf16x2(v[12], v[13]), pack_bf16x2(v[14], v[15]));
          }
          fence_proxy_async_smem();
          arrive_leader(&hs_full[b]);
          TRACE_ADD(w_hid);
        }
      }

      // ---- after the last layer
      float4 gwreg[3];                                   // fused head: this thread's share of the gamma * Wfc | b2 table,
      if (HP.bases != nullptr) {                         // fetched while the last GEMM2 drains
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          const int idx = (int)threadIdx.x - 128 + i * 256;
          gwreg[i] = idx < kD * 2 ? __ldg(reinterpret_cast<const float4*>(HP.gw8) + idx) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
      SW(y_full, k_y & 1, 917); ++k_y;
      tc_fence_after();
      const float* __restrict__ b2 = P.b2[NL - 1];
      if (HP.bases == nullptr) {
        // x <- Y + b2 (global), for a separate head kernel
#pragma unroll 3
        for (int cb = 0; cb < 9; ++cb) {
          uint32_t acc[16];
          tmem_ld16(tmem_row + C::kTmemY + (cb0 + cb) * 16, acc);
          tmem_ld_wait();
          if (valid) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int col = (cb0 + cb) * 16 + 4 * i;
              float4 o4;
              o4.x = col + 0 < kD ? __uint_as_float(acc[4 * i + 0]) + __ldg(b2 + col + 0) : 0.f;
              o4.y = col + 1 < kD ? __uint_as_float(acc[4 * i + 1]) + __ldg(b2 + col + 1) : 0.f;
              o4.z = col + 2 < kD ? __uint_as_float(acc[4 * i + 2]) + __ldg(b2 + col + 2) : 0.f;
              o4.w = col + 3 < kD ? __uint_as_float(acc[4 * i + 3]) + __ldg(b2 + col + 3) : 0.f;
              xrow[(size_t)((cb0 + cb) * 4 + i) * kTileM] = o4;
            }
          }
        }
      } else {
        // ---- fused head (encoder_stack.py:197, networks.py:342,238, quick_inference.py:377-414): final LayerNorm of
        // v = Y + b2, the five fc1 logits, then head_finish -- in ONE pass over the row, as head_kernel does it:
        //   logits_j = rstd * (sum_k (v_k - mean) g_k W_kj) + B_j + bfc_j,  v_k - shift = y_k + b2_k  (y = Y - shift), so with
        //   T_j = sum y_k gW_kj,  S1 = sum y,  S2 = sum y^2,  BB = sum y_k b2_k  (the row's two halves added up)
        //   m = (S1 + SB) / n,  var = (S2 + 2 BB + SBB) / n - m^2,  logits_j = rstd * (T_j + H_j - m A_j) + B_j
        // (A, B, H, SB, SBB: per-model constants in HP.ab; gW | b2 per column in HP.gw8, staged in the idle staging area).
        float* sShift = reinterpret_cast<float*>(sS);                       // [128]
        float4* sGW = reinterpret_cast<float4*>(sS + 8192);                 // [280][2]
        float* sPart = reinterpret_cast<float*>(sS + 8192 + kD * 32);       // [128][8] + [128] partial sums of the upper half
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          const int idx = (int)threadIdx.x - 128 + i * 256;
          if (idx < kD * 2) sGW[idx] = gwreg[i];
        }
        const uint32_t ycol = tmem_row + C::kTmemY + cb0 * 16;
        uint32_t a[16], b[16];
        tmem_ld16(ycol, a);
        tmem_ld_wait();
        if (!halfsel) sShift[r] = __uint_as_float(a[0]);
        asm volatile("bar.sync 1, 256;" ::: "memory");     // shift of the row published, table staged
        const float shift = sShift[r];
        float s1 = 0.f, s2 = 0.f, bb = 0.f, t[kVocab];
#pragma unroll
        for (int j = 0; j < kVocab; ++j) t[j] = 0.f;
        auto accum = [&](const uint32_t (&acc)[16], int cb) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int col = (cb0 + cb) * 16 + i;
            if (col < kD) {
              const float y = __uint_as_float(acc[i]) - shift;
              const float4 w0 = sGW[col * 2], w1 = sGW[col * 2 + 1];          // 16-byte broadcast reads
              s1 += y;
              s2 = fmaf(y, y, s2);
              t[0] = fmaf(y, w0.x, t[0]); t[1] = fmaf(y, w0.y, t[1]); t[2] = fmaf(y, w0.z, t[2]);
              t[3] = fmaf(y, w0.w, t[3]); t[4] = fmaf(y, w1.x, t[4]);
              bb = fmaf(y, w1.y, bb);
            }
          }
        };
#pragma unroll 1
        for (int cb = 0; cb < 8; cb += 2) {
          tmem_ld_wait();
          tmem_ld16(ycol + (cb + 1) * 16, b);
          accum(a, cb);
          tmem_ld_wait();
          tmem_ld16(ycol + (cb + 2) * 16, a);
          accum(b, cb + 1);
        }
        tmem_ld_wait();
        accum(a, 8);
        if (halfsel) {
          *reinterpret_cast<float4*>(sPart + r * 8) = make_float4(t[0], t[1], t[2], t[3]);
          *reinterpret_cast<float4*>(sPart + r * 8 + 4) = make_float4(t[4], s1, s2, bb);
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (!halfsel && valid && goff + r < L) {
          const float4 p0 = *reinterpret_cast<const float4*>(sPart + r * 8), p1 = *reinterpret_cast<const float4*>(sPart + r * 8 + 4);
          t[0] += p0.x; t[1] += p0.y; t[2] += p0.z; t[3] += p0.w; t[4] += p1.x;
          s1 += p1.y; s2 += p1.z; bb += p1.w;
          const float m = (s1 + __ldg(HP.ab + 24)) * (1.f / kD);
          const float var = fmaxf((s2 + 2.f * bb + __ldg(HP.ab + 25)) * (1.f / kD) - m * m, 0.f);
          const float rstd = rsqrtf(var + 1e-6f);
          float lg[kVocab];
#pragma unroll
          for (int j = 0; j < kVocab; ++j)
            lg[j] = rstd * (t[j] + __ldg(HP.ab + 16 + j) - m * __ldg(HP.ab + j)) + __ldg(HP.ab + 8 + j);
          // window-aligned layout: tile == window (kWide: tile pair == window), row == position
          head_finish(HP, lg, kWide ? (size_t)(tile >> 1) * L + goff + r : (size_t)tile * L + r);
        }
        asm volatile("bar.sync 2, 256;" ::: "memory");     // staging area free for the next tile
      }
      tc_fence_before();
    }
#ifdef DCB_TRACE
    if (warp == 4 && lane == 0 && blockIdx.x < 256) {
      unsigned long long* tr = g_ffn_trace + blockIdx.x * 16 + 8;
      tr[0] = w_rowpass; tr[1] = w_stage; tr[2] = w_qk; tr[3] = w_soft; tr[4] = w_pv; tr[5] = w_store; tr[6] = w_wait; tr[7] = w_hid;
    }
#endif
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, C::kTmemCols);
  }
}
