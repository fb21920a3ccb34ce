// pileup_ws.cu -- K1w: the tile-owner pileup of pileup_tiled.cu as a warp-specialised pipeline.
//
// Same arithmetic as K1f (bit-sliced positional popcount over TMA-staged reads, N recovered from the
// coverage identity, exclusive tile ownership) -- what changes is WHO does the per-tile bookkeeping.
// In K1f all 8 warps of a CTA stage a tile together (index load, bulk copy, per-read metadata,
// difference array, two CTA barriers, an mbarrier wait) and only then count it; ncu shows 62 % of the
// kernel's time in those phases at low issue rates, with only two CTAs per SM to overlap them.
// Here one persistent CTA per SM runs
//     4 PRODUCER warps: for every (tile, sub-chunk) item -- wait for a free stage, one bulk copy of the
//                       reads' bases (TMA, completes on the stage's `full` mbarrier), per-read
//                       metadata, coverage prefix sums, header; arrive on `full`;
//     8 CONSUMER warps: wait `full`, count their 64-slot window against the item's reads (the K1f inner
//                       loop, unchanged), flush at the end of a tile, arrive on the stage's `empty`;
// over a two-stage ring in ~200 KB of shared memory.  Consumers never touch global memory except for
// the table stores, never hit a CTA-wide barrier, and start a tile the moment its item is complete;
// the producers run one item ahead.
#include "kdl_common.cuh"

namespace kdl {

// Configurations.  WsCfg1 = K1w as validated on the GPU: one CTA per SM, full-tile stages.  WsCfg2 = K1w2
// (experimental, KDL_K1F=ws2, emulator-checked only): two CTAs per SM, i.e. 16 consumer warps per SM where the
// K1w experiment showed 8 to be the limit, with half-size stages; the register file is re-balanced between the
// roles with setmaxnreg (producers 40, consumers 96: 128*40 + 256*96 = 384*77 <= 384*80 of the launch bound).
struct WsCfg1 {
    static constexpr int kConsumers = 8;    // consumer warps (one 64-slot window each)
    static constexpr int kProducers = 4;    // producer warps
    static constexpr int kStages = 2;
    static constexpr int kRmax = F_RMAX;    // reads per item
    static constexpr int kCapW = 17920;     // words of packed bases per item (70 KB)
    static constexpr int kMinBlocks = 1;
    static constexpr bool kRebalance = false;
    static constexpr bool kLean = false;
};
struct WsCfg2 {
    static constexpr int kConsumers = 8;
    static constexpr int kProducers = 4;
    static constexpr int kStages = 2;
    static constexpr int kRmax = 512;
    static constexpr int kCapW = 9216;      // 36 KB: ~485 reads of 150 bases (19 words each)
    static constexpr int kMinBlocks = 2;
    static constexpr bool kRebalance = true;
    static constexpr bool kLean = true;     // the kLean trims of pileup_tiled.cu in both roles
};
constexpr int W_PRODUCER_REGS = 40, W_CONSUMER_REGS = 96;  // WsCfg2 only
constexpr int W_CONSUMERS = WsCfg1::kConsumers;
constexpr int W_PRODUCERS = WsCfg1::kProducers;
constexpr int W_THREADS = 32 * (W_CONSUMERS + W_PRODUCERS);  // the same for both configurations

enum : int { ITEM_FIRST = 1, ITEM_LAST = 2, ITEM_EMPTY = 4, ITEM_END = 8 };

template <class C>
struct WsStageT {
    uint32_t seq[C::kCapW];
    int4 meta[C::kRmax + 40 + (C::kRmax + 40) / 8];  // as in FastSmem: entry of read i at i + i/8
    int gs[C::kRmax + 32];
    int cov[KDL_TILE];                            // simple reads of THIS item covering each slot
    int diff[KDL_TILE + 32];                      // producers only
    long long tile_slot;
    int n_sub;
    int flags;
};

template <class C>
struct WsSmemT {
    WsStageT<C> st[C::kStages];
    int raw[3][C::kRmax];           // producers only: l_seq / ref_start / seq_off of the NEXT tile's first reads
    uint64_t full[C::kStages];      // producers -> consumers: 4 warp arrivals + the bulk copy's bytes
    uint64_t empty[C::kStages];     // consumers -> producers: 8 warp arrivals
};
using WsSmem = WsSmemT<WsCfg1>;
static_assert(WsCfg2::kConsumers == W_CONSUMERS && WsCfg2::kProducers == W_PRODUCERS,
              "the kernel body uses W_CONSUMERS / W_PRODUCERS for both configurations");
static_assert(sizeof(WsSmemT<WsCfg2>) <= 113 * 1024, "K1w2 must fit two CTAs per SM");
static_assert(offsetof(WsStageT<WsCfg2>, diff) % 16 == 0 && sizeof(WsStageT<WsCfg2>) % 16 == 0,
              "kLean reads the difference array with 128-bit loads");

#ifndef KDL_HOST_EMU
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx_only(uint64_t* bar, uint32_t bytes) {  // no arrival
    asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void producer_sync() {  // the 128 producer threads only
    asm volatile("bar.sync 1, %0;" ::"n"(32 * W_PRODUCERS) : "memory");
}
template <int kRegs> __device__ __forceinline__ void reg_dealloc() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegs)); }
template <int kRegs> __device__ __forceinline__ void reg_alloc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegs)); }
#endif  // KDL_HOST_EMU

template <bool kFresh, class C = WsCfg1>
__global__ void __launch_bounds__(W_THREADS, C::kMinBlocks)
pileup_ws_kernel(kdl_batch b, int32_t* __restrict__ counts, long long n_slots,
                 const uint32_t* __restrict__ tile_index, long long tile_lo, long long n_tiles) {
    KDL_DYNAMIC_SMEM(smem_raw);
    using Smem = WsSmemT<C>;
    using Stage = WsStageT<C>;
    constexpr int W_STAGES = C::kStages, W_RMAX = C::kRmax, W_CAPW = C::kCapW;
    Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int maxlen = b.max_simple_len;

    if (tid == 0) {
        for (int s = 0; s < W_STAGES; ++s) {
            mbar_init(&sm.full[s], W_PRODUCERS);
            mbar_init(&sm.empty[s], W_CONSUMERS);
        }
    }
    for (int s = 0; s < W_STAGES; ++s)
        for (int k = tid; k < KDL_TILE + 32; k += W_THREADS) sm.st[s].diff[k] = 0;
    __syncthreads();

    if (warp >= W_CONSUMERS) {
        // =========================== PRODUCERS ====================================================
        if constexpr (C::kRebalance) reg_dealloc<W_PRODUCER_REGS>();
        const int pw = warp - W_CONSUMERS;          // 0..3
        const int ptid = tid - 32 * W_CONSUMERS;    // 0..127
        long long item = 0;
        auto acquire_stage = [&](long long it) -> Stage& {
            const int s = (int)(it % W_STAGES);
            const uint32_t round = (uint32_t)(it / W_STAGES);
            if (round > 0) mbar_wait(&sm.empty[s], (round - 1) & 1u);  // consumers released its last use
            return sm.st[s];
        };
        auto publish = [&](long long it) {  // this warp's part of the item is written
            const int s = (int)(it % W_STAGES);
            __syncwarp();
            if (lane == 0) mbar_arrive(&sm.full[s]);
        };

        // The producers run a software pipeline of their own: the index entry of the next tile and the
        // three metadata words of its first reads are in flight (cp.async into sm.raw, each thread
        // fetching exactly the elements it will consume) while the current item is being prepared.
        constexpr int PT = 32 * W_PRODUCERS;
        constexpr int PER = W_RMAX / PT;  // reads per producer thread and item
        auto load_index = [&](long long t, uint4& ix, uint2& ic) {
            if (t < tile_lo + n_tiles) {
                ix = __ldg(reinterpret_cast<const uint4*>(tile_index + F_IDX * t));
                ic = __ldg(reinterpret_cast<const uint2*>(tile_index + F_IDX * t + 4));
            } else {
                ix = make_uint4(0, 0, 0, 0);
                ic = make_uint2(0, 0);
            }
        };
        auto prefetch_raw = [&](const uint4& ix) {
            const long long plo = ix.x, phi = ix.y;
            const int cnt = (int)(phi - plo < W_RMAX ? phi - plo : W_RMAX);
            for (int i = ptid; i < cnt; i += PT) {
                cp_async4(&sm.raw[0][i], b.l_seq + plo + i);
                cp_async4(&sm.raw[1][i], b.ref_start + plo + i);
                cp_async4(&sm.raw[2][i], b.seq_off + plo + i);
            }
        };
        uint4 ix, nix;
        uint2 ic, nic;
        load_index(tile_lo + blockIdx.x, ix, ic);
        long long cs = 0, ncs = 0;  // kLean: slot of the (next) tile's first contig, loaded a tile ahead
        if constexpr (C::kLean) cs = b.contig_slot[ic.x];
        prefetch_raw(ix);

        for (long long tile = tile_lo + blockIdx.x; tile < tile_lo + n_tiles; tile += gridDim.x) {
            load_index(tile + gridDim.x, nix, nic);  // consumed at the end of this iteration
            if constexpr (C::kLean) ncs = b.contig_slot[nic.x];
            const long long lo = ix.x, hi = ix.y;
            const long long tile_slot = tile * KDL_TILE;
            if (lo >= hi) {
                if (kFresh) {  // consumers must store zeros: a header-only item
                    Stage& st = acquire_stage(item);
                    if (ptid == 0) { st.tile_slot = tile_slot; st.n_sub = 0; st.flags = ITEM_FIRST | ITEM_LAST | ITEM_EMPTY; }
                    publish(item);
                    ++item;
                }
                prefetch_raw(nix);  // nothing was in flight for an empty tile
                ix = nix;
                ic = nic;
                cs = ncs;
                continue;
            }
            const bool one_contig = ic.x == ic.y;
            long long slot_base;
            if constexpr (C::kLean) slot_base = one_contig ? cs - tile_slot : 0;
            else slot_base = one_contig ? b.contig_slot[ic.x] - tile_slot : 0;
            long long c0 = lo;
            bool first = true;
            bool raw_pending = true;
            while (c0 < hi) {
                long long c1 = c0 + W_RMAX < hi ? c0 + W_RMAX : hi;
                const long long wa = c0 == lo ? (long long)ix.z : (long long)(b.seq_off[c0] & ~3u);
                long long wend = c1 == hi ? (long long)ix.w : (long long)b.seq_off[c1];
                bool skip = false;
                while (wend - wa > W_CAPW) {
                    if (c1 - c0 == 1) { skip = true; break; }  // one read too long to stage: never simple
                    if constexpr (C::kRebalance) {  // small stages: cut where the capacity ends, not in half
                        const long long n = c1 - c0;
                        long long n2 = n * W_CAPW / (wend - wa);
                        n2 = n2 >= n ? n - 1 : (n2 < 1 ? 1 : n2);
                        c1 = c0 + n2;
                    } else {
                        c1 = c0 + (c1 - c0) / 2;
                    }
                    wend = (long long)b.seq_off[c1];
                }
                const bool last = c1 >= hi;
                const int n_sub = skip ? 0 : (int)(c1 - c0);
                // this thread's reads of the item: from the prefetched words (first sub-chunk) or directly
                int l[PER], rs[PER];
                uint32_t so[PER];
                if (c0 == lo && raw_pending) {
                    cp_async_wait_all();
#pragma unroll
                    for (int k = 0; k < PER; ++k) {
                        const int i = ptid + k * PT;
                        const int ii = i < n_sub ? i : ptid;
                        l[k] = sm.raw[0][ii];
                        rs[k] = sm.raw[1][ii];
                        so[k] = (uint32_t)sm.raw[2][ii];
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < PER; ++k) {
                        const int i = ptid + k * PT;
                        const long long r = c0 + (i < n_sub ? i : 0);
                        l[k] = b.l_seq[r];
                        rs[k] = b.ref_start[r];
                        so[k] = b.seq_off[r];
                    }
                }
                if (raw_pending) {  // own elements are in registers: refill them for the next tile
                    prefetch_raw(nix);
                    raw_pending = false;
                }
                Stage& st = acquire_stage(item);
                const uint32_t seq_base = smem_u32(st.seq);
                if (!skip) {
                    const long long n_words = wend - wa;
                    const long long avail = b.seq4_words - wa;
                    const long long want = (n_words + 3) & ~3ll;
                    const long long bulk_words = want <= avail ? want : (avail & ~3ll);
                    const uint32_t tx = (uint32_t)(bulk_words * 4);
                    if (ptid == 0 && bulk_words) {  // announce the bytes, then let the TMA engine copy them
                        uint64_t* fb = &sm.full[(int)(item % W_STAGES)];
                        mbar_expect_tx_only(fb, tx);
                        bulk_g2s(st.seq, b.seq4 + wa, tx, fb);
                    }
                    if (bulk_words < n_words && ptid < 4) {
                        const long long w = bulk_words + ptid;
                        st.seq[w] = w < avail ? b.seq4[wa + w] : 0u;
                    }
                }
#pragma unroll
                for (int k = 0; k < PER; ++k) {
                    const int i = ptid + k * PT;
                    if (i < n_sub) {
                        long long g;
                        if (one_contig) {
                            g = slot_base + rs[k];
                        } else {
                            const int c = find_contig(b.contig_read_off, b.n_contigs, c0 + i);
                            g = b.contig_slot[c] + rs[k] - tile_slot;
                        }
                        if constexpr (!C::kLean) g = g < -0x10000000ll ? -0x10000000ll : (g > 0x10000000ll ? 0x10000000ll : g);
                        const int gs = (int)g;
                        int nb = 0;
                        if (l[k] > 0) {
                            nb = ((l[k] + 7) >> 3) << 2;
                            const int cs = gs < 0 ? 0 : gs, ce = gs + l[k] > KDL_TILE ? KDL_TILE : gs + l[k];
                            if (cs < ce) {
                                atomicAdd(st.diff + cs, 1);
                                atomicAdd(st.diff + ce, -1);
                            }
                        }
                        st.gs[i] = gs;
                        st.meta[i + (i >> 3)] = make_int4(((gs + 7) >> 3) << 2,
                                                          (int)(seq_base + (uint32_t)(((long long)so[k] - wa) << 2)), nb,
                                                          ((-gs) & 7) << 2);
                    }
                }
                if (ptid < 40) {  // sentinels behind the last read
                    const int i = n_sub + ptid;
                    if (ptid < 32) st.gs[i] = 0x10000000;
                    st.meta[i + (i >> 3)] = make_int4(0x10000000, (int)seq_base, 0, 0);
                }
                if (ptid == 0) {
                    st.tile_slot = tile_slot;
                    st.n_sub = n_sub;
                    st.flags = (first ? ITEM_FIRST : 0) | (last ? ITEM_LAST : 0);
                }
                producer_sync();  // difference array complete
                {   // coverage: producer warp pw scans slots [128 pw, 128 pw + 128)
                    const int w0 = (KDL_TILE / W_PRODUCERS) * pw;
                    int pre = 0;
                    if constexpr (C::kLean) {  // w0 is a multiple of 128, diff is 16-byte aligned
                        for (int k = 4 * lane; k < w0; k += 128) {
                            const int4 v4 = *reinterpret_cast<const int4*>(st.diff + k);
                            pre += (v4.x + v4.y) + (v4.z + v4.w);
                        }
                    } else {
                        for (int k = lane; k < w0; k += 32) pre += st.diff[k];
                    }
#pragma unroll
                    for (int d = 16; d; d >>= 1) pre += __shfl_xor_sync(0xffffffffu, pre, d);
                    constexpr int E = KDL_TILE / W_PRODUCERS / 32;  // entries per lane
                    int v[E];
                    int run = 0;
#pragma unroll
                    for (int k = 0; k < E; ++k) { v[k] = st.diff[w0 + E * lane + k]; run += v[k]; }
                    int incl = run;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        const int o = __shfl_up_sync(0xffffffffu, incl, d);
                        if (lane >= d) incl += o;
                    }
                    int acc = pre + incl - run;
#pragma unroll
                    for (int k = 0; k < E; ++k) { acc += v[k]; st.cov[w0 + E * lane + k] = acc; }
                }
                producer_sync();  // everybody has read diff: clean it for the stage's next use
                if constexpr (C::kLean) {  // 136 x 128-bit stores instead of 544 scalar ones
                    for (int k = ptid; k < (KDL_TILE + 32) / 4; k += PT)
                        reinterpret_cast<int4*>(st.diff)[k] = make_int4(0, 0, 0, 0);
                } else {
                    for (int k = ptid; k < KDL_TILE + 32; k += PT) st.diff[k] = 0;
                }
                publish(item);
                ++item;
                first = false;
                c0 = c1;
            }
            ix = nix;
            ic = nic;
            cs = ncs;
        }
        {   // END item
            Stage& st = acquire_stage(item);
            if (ptid == 0) { st.tile_slot = 0; st.n_sub = 0; st.flags = ITEM_END; }
            publish(item);
        }
        return;
    }

    // =============================== CONSUMERS ===================================================
    if constexpr (C::kRebalance) reg_alloc<W_CONSUMER_REGS>();
    const int quarter = lane >> 3;
    const int wlo = warp * F_WIN;
    const int p8b = (wlo >> 1) + 4 * (lane & 7);
    Planes acc;
    acc.clear();
    int rawacc[8], covacc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { rawacc[k] = 0; covacc[k] = 0; }
    int blocks_since_flush = 0;
    uint32_t pend8 = 0;  // kLean: weight-8 carry of an odd block, waiting for its partner
    bool stored = false;

    for (long long item = 0;; ++item) {
        const int s = (int)(item % W_STAGES);
        mbar_wait(&sm.full[s], (uint32_t)((item / W_STAGES) & 1));
        Stage& st = sm.st[s];
        const int flags = st.flags;
        const int n_sub = st.n_sub;
        const long long tile_slot = st.tile_slot;
        if (flags & ITEM_END) break;
        if (flags & ITEM_FIRST) {
            stored = false;
            blocks_since_flush = 0;  // (already 0 after the previous tile's final flush)
        }
        if (n_sub > 0) {
            {
                const int4 ca = *reinterpret_cast<const int4*>(st.cov + wlo + 8 * (lane & 7));
                const int4 cb = *reinterpret_cast<const int4*>(st.cov + wlo + 8 * (lane & 7) + 4);
                covacc[0] += ca.x; covacc[1] += ca.y; covacc[2] += ca.z; covacc[3] += ca.w;
                covacc[4] += cb.x; covacc[5] += cb.y; covacc[6] += cb.z; covacc[7] += cb.w;
            }
            int a, e;
            if constexpr (C::kLean) {
                lower_bound_warp2(st.gs, n_sub, wlo - maxlen + 1, wlo + F_WIN, lane, a, e);
            } else {
                a = lower_bound_warp(st.gs, n_sub, wlo - maxlen + 1, lane);
                e = lower_bound_warp(st.gs, n_sub, wlo + F_WIN, lane);
            }
            for (int base = a & ~7; base < e; base += 32) {
                uint32_t x[8];
                int4 mt[8];
                const int i0 = base + 8 * quarter;
                const int4* mp = st.meta + i0 + (i0 >> 3);
#pragma unroll
                for (int u = 0; u < 8; ++u) mt[u] = mp[u];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const uint32_t jb = (uint32_t)(p8b - mt[u].x);
                    const uint32_t addr = (uint32_t)mt[u].y + jb;
                    uint32_t hw, lw;
#ifndef KDL_HOST_EMU
                    asm("{\n"
                        ".reg .pred p, q;\n"
                        "setp.lt.u32 p, %2, %3;\n"
                        "setp.lt.u32 q, %4, %3;\n"
                        "mov.u32 %0, 0;\n"
                        "mov.u32 %1, 0;\n"
                        "@p ld.shared.u32 %0, [%5];\n"
                        "@q ld.shared.u32 %1, [%5+4];\n"
                        "}\n"
                        : "=&r"(hw), "=&r"(lw)
                        : "r"(jb), "r"((uint32_t)mt[u].z), "r"(jb + 4u), "r"(addr));
#else
                    hw = jb < (uint32_t)mt[u].z ? lds_u32(addr) : 0u;
                    lw = jb + 4u < (uint32_t)mt[u].z ? lds_u32(addr + 4u) : 0u;
#endif
                    x[u] = __funnelshift_l(lw, hw, (uint32_t)mt[u].w);
                }
                if constexpr (C::kLean) {
                    const uint32_t e8 = acc.add8_carry(x);
                    if (blocks_since_flush & 1) {  // second block of a pair: eights + eights -> sixteens, one ripple
                        uint32_t c16;
                        csa(c16, acc.p[3], acc.p[3], pend8, e8);
                        acc.template ripple<4>(c16);
                    } else {
                        pend8 = e8;
                    }
                } else {
                    acc.add8(x);
                }
                if (++blocks_since_flush == F_FLUSH_BLOCKS) {
                    if constexpr (C::kLean) acc.template ripple<3>(pend8);  // F_FLUSH_BLOCKS is odd: one carry is pending
                    if (kFresh && !stored)
                        flush_window<true, false>(acc, rawacc, covacc, counts, n_slots, tile_slot + wlo, lane);
                    else
                        flush_window<false, false>(acc, rawacc, covacc, counts, n_slots, tile_slot + wlo, lane);
                    stored = true;
                    blocks_since_flush = 0;
                }
            }
        }
        // this warp is done reading the stage: hand it back before the (global-memory) flush
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.empty[s]);
        if (flags & ITEM_LAST) {
            if constexpr (C::kLean) {
                if (blocks_since_flush & 1) acc.template ripple<3>(pend8);
            }
            if (kFresh && !stored) flush_window<true, true>(acc, rawacc, covacc, counts, n_slots, tile_slot + wlo, lane);
            else flush_window<false, true>(acc, rawacc, covacc, counts, n_slots, tile_slot + wlo, lane);
            blocks_since_flush = 0;
#pragma unroll
            for (int k = 0; k < 8; ++k) covacc[k] = 0;
        }
    }
}

}  // namespace kdl
