// pileup_wide.cu -- K1x: the tile-owner pileup with WIDE lanes (experimental, KDL_K1F=wide; not yet run on a GPU).
//
// K1f is bound by instruction issue; this variant does the same arithmetic with fewer
// instructions per base.  A lane owns 16 consecutive slots instead of 8 -- two words X0, X1 cut from THREE
// words of the read by two funnel shifts, so the per-read address / predicate work serves twice the bases --
// and a group of only 4 lanes owns the 64-slot window, so the 8 groups of a warp walk 8 different reads and a
// 150-base read keeps ~95 % of the lanes busy (K1f: quarters of 8 lanes, ~80 %).  Two plane sets per lane; the
// 8 streams are summed bit-sliced in three shuffle butterflies; at a flush group g transposes word g / 4,
// bit g % 4, i.e. every lane still writes 8 consecutive slots of one column.  Staging, metadata, coverage and
// the window search are K1f's.  The arithmetic is modelled and checked on the CPU (tests/k1f_model.py,
// pileup_model_wide) and this source runs under the host emulator (tests/emu/, KDL_HOST_EMU).
#include "kdl_common.cuh"

namespace kdl {

// sum of the same planes held by the 8 groups (lanes l ^ 4, l ^ 8, l ^ 16): F_P planes in, F_P + 3 out
__device__ __forceinline__ void octet_sum(const uint32_t (&in)[F_P], uint32_t (&out)[F_P + 3]) {
    uint32_t a[F_P + 3];
#pragma unroll
    for (int k = 0; k < F_P; ++k) a[k] = in[k];
    a[F_P] = a[F_P + 1] = a[F_P + 2] = 0;
#pragma unroll
    for (int stage = 0; stage < 3; ++stage) {
        const int width = F_P + stage;  // planes that can be non-zero before this stage
        uint32_t carry = 0;
#pragma unroll
        for (int k = 0; k < F_P + 3; ++k) {
            if (k < width) {
                const uint32_t o = __shfl_xor_sync(0xffffffffu, a[k], 4 << stage);
                uint32_t c2, s;
                csa(c2, s, a[k], o, carry);
                a[k] = s;
                carry = c2;
            } else if (k == width) {
                a[k] = carry;
            }
        }
    }
#pragma unroll
    for (int k = 0; k < F_P + 3; ++k) out[k] = a[k];
}

// counters of nibble bit `bit` from F_P + 3 (<= 12) planes: out[b] for slot b of the word (nibble 7 - b)
__device__ __forceinline__ void extract8_wide(const uint32_t (&pl)[F_P + 3], int bit, int (&out)[8]) {
    uint32_t v[3] = {0u, 0u, 0u};
#pragma unroll
    for (int k = 0; k < F_P + 3; ++k) v[k >> 2] |= ((pl[k] >> bit) & 0x11111111u) << (k & 3);
#pragma unroll
    for (int b = 0; b < 8; ++b) {
        const int sh = 4 * (7 - b);
        out[b] = (int)(((v[0] >> sh) & 0xFu) | (((v[1] >> sh) & 0xFu) << 4) | (((v[2] >> sh) & 0xFu) << 8));
    }
}

// Flush of a window: group g = lane / 4 transposes word g / 4 (slots 0-7 or 8-15 of the lane), column g % 4,
// and writes those 8 consecutive slots; groups 0 and 4 (bit 0) also write column N.  N from the coverage
// identity exactly as in K1f's flush_window.
template <bool kStore, bool kFinal>
__device__ __forceinline__ void flush_window_wide(Planes& acc0, Planes& acc1, int (&rawacc)[8], const int (&covacc)[8],
                                                  int32_t* __restrict__ counts, long long n_slots, long long slot0,
                                                  int lane) {
    const int grp = lane >> 2, j4 = lane & 3;
    const int word = grp >> 2, bit = grp & 3;
    const long long s = slot0 + 16 * j4 + 8 * word;
    int4* dst = reinterpret_cast<int4*>(counts + (long long)bit * n_slots + s);
    int4 v0 = make_int4(0, 0, 0, 0), v1 = v0;
    if (!kStore) { v0 = dst[0]; v1 = dst[1]; }
    uint32_t m0[F_P + 3], m1[F_P + 3];
    octet_sum(acc0.p, m0);
    octet_sum(acc1.p, m1);
    acc0.clear();
    acc1.clear();
    int cv[8], tot[8];
    if (word == 0) extract8_wide(m0, bit, cv);
    else extract8_wide(m1, bit, cv);
#pragma unroll
    for (int k = 0; k < 8; ++k) {  // A+C+G+T raw of each slot: the four groups that share this word
        int t = cv[k];
        t += __shfl_xor_sync(0xffffffffu, t, 4);
        t += __shfl_xor_sync(0xffffffffu, t, 8);
        tot[k] = t + rawacc[k];
    }
    int nn[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        if (kFinal) {
            nn[k] = (tot[k] - covacc[k]) / 3;
            rawacc[k] = 0;
        } else {
            nn[k] = 0;
            rawacc[k] = tot[k];
        }
    }
    v0.x += cv[0] - nn[0]; v0.y += cv[1] - nn[1]; v0.z += cv[2] - nn[2]; v0.w += cv[3] - nn[3];
    v1.x += cv[4] - nn[4]; v1.y += cv[5] - nn[5]; v1.z += cv[6] - nn[6]; v1.w += cv[7] - nn[7];
    dst[0] = v0;
    dst[1] = v1;
    if (bit == 0 && (kFinal || kStore)) {
        int4* dn = reinterpret_cast<int4*>(counts + (long long)KDL_W_N * n_slots + s);
        int4 n0 = make_int4(0, 0, 0, 0), n1 = n0;
        if (!kStore) { n0 = dn[0]; n1 = dn[1]; }
        n0.x += nn[0]; n0.y += nn[1]; n0.z += nn[2]; n0.w += nn[3];
        n1.x += nn[4]; n1.y += nn[5]; n1.z += nn[6]; n1.w += nn[7];
        dn[0] = n0;
        dn[1] = n1;
    }
}

struct WideSmem {
    uint32_t seq[F_CAPW];
    // per staged read:
    //   .x  byte offset, relative to the tile, of the first 8-slot group the read can serve:
    //       4 * ceil(start / 8)            (lane byte offset - this = word of the read, in bytes)
    //   .y  shared-memory address (u32) of the read's first word
    //   .z  bytes of packed bases (0 = not a simple read: adds nothing)
    //   .w  funnel-shift amount 4 * ((-start) & 7)
    int4 meta[F_RMAX + 72 + (F_RMAX + 72) / 8];  // entry of read i at i + i/8; 72 sentinels (a block is 64 reads)
    int gs[F_RMAX + 32];          // start slot relative to the tile (all reads: the array stays sorted)
    int diff[2][KDL_TILE + 32];   // +1 at read start, -1 at read end (double-buffered per sub-chunk)
    int cov[KDL_TILE];            // prefix sums of diff: simple reads covering each slot
    int raw[3][F_RMAX];           // l_seq / ref_start / seq_off of the NEXT tile's first reads (cp.async)
    uint64_t bar;                 // mbarrier the bulk copy of seq[] completes on
};

// kFresh: columns 0..4 hold stale data; the first flush of every window stores instead of adding,
// and windows / tiles without reads are stored as zeros.
template <bool kFresh>
__global__ void __launch_bounds__(F_THREADS, 2)
pileup_wide_kernel(kdl_batch b, int32_t* __restrict__ counts, long long n_slots,
                    const uint32_t* __restrict__ tile_index, long long tile_lo, long long n_tiles) {
    KDL_DYNAMIC_SMEM(smem_raw);
    WideSmem& sm = *reinterpret_cast<WideSmem*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int grp = lane >> 2;  // 8 groups of 4 lanes: 8 read streams per warp
    const int j4 = lane & 3;    // the lane's 16 slots inside the 64-slot window: 16 j4 .. 16 j4 + 15
    const int maxlen = b.max_simple_len;
    const uint32_t seq_base = smem_u32(sm.seq);
    uint32_t bar_parity = 0;
    int dbuf = 0;
    if (tid == 0) mbar_init(&sm.bar, 1);
    for (int k = tid; k < 2 * (KDL_TILE + 32); k += F_THREADS) (&sm.diff[0][0])[k] = 0;
    __syncthreads();

    // Software pipeline across this CTA's tiles: while tile t is being counted, the per-read
    // metadata of tile t + gridDim.x's first sub-chunk streams into sm.raw (cp.async, each thread
    // fetches exactly the elements it will later consume, so no barrier is needed for them).
    auto prefetch_raw = [&](long long t) {
        if (t >= tile_lo + n_tiles) return;
        const uint2 nx = __ldg(reinterpret_cast<const uint2*>(tile_index + F_IDX * t));
        const long long plo = nx.x, phi = nx.y;
        const int cnt = (int)(phi - plo < F_RMAX ? phi - plo : F_RMAX);
        for (int i = tid; i < cnt; i += F_THREADS) {
            cp_async4(&sm.raw[0][i], b.l_seq + plo + i);
            cp_async4(&sm.raw[1][i], b.ref_start + plo + i);
            cp_async4(&sm.raw[2][i], b.seq_off + plo + i);
        }
    };
    prefetch_raw(tile_lo + blockIdx.x);

    for (long long tile = tile_lo + blockIdx.x; tile < tile_lo + n_tiles; tile += gridDim.x) {
        const uint4 ix = __ldg(reinterpret_cast<const uint4*>(tile_index + F_IDX * tile));
        const uint2 ic = __ldg(reinterpret_cast<const uint2*>(tile_index + F_IDX * tile + 4));
        const long long lo = ix.x, hi = ix.y;
        const long long tile_slot = tile * KDL_TILE;
        bool raw_pending = true;  // sm.raw holds this tile's first reads; the next prefetch is still to issue
        if (lo >= hi) {  // uniform for the CTA: no read reaches this tile
            if (kFresh) {  // 5 columns x 512 slots of zeros, 128-bit stores
                for (int v = tid; v < 5 * (KDL_TILE / 4); v += F_THREADS) {
                    const int col = v / (KDL_TILE / 4), off = v % (KDL_TILE / 4);
                    reinterpret_cast<int4*>(counts + (long long)col * n_slots + tile_slot)[off] = make_int4(0, 0, 0, 0);
                }
            }
            prefetch_raw(tile + gridDim.x);  // nothing was prefetched for an empty tile: raw is free
            continue;
        }
        const bool one_contig = ic.x == ic.y;
        const long long slot_base = one_contig ? b.contig_slot[ic.x] - tile_slot : 0;
        const int wlo = warp * F_WIN;
        const int p8b = (wlo >> 1) + 8 * j4;  // byte offset, in a read aligned to the tile, of the lane's first word
        const int oslot = 16 * j4 + 8 * (grp >> 2);  // the 8 slots this lane transposes and writes at a flush
        Planes acc0, acc1;  // slots 0-7 and 8-15 of the lane
        acc0.clear();
        acc1.clear();
        int rawacc[8], covacc[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) { rawacc[k] = 0; covacc[k] = 0; }
        int blocks_since_flush = 0;
        bool stored = false;  // kFresh: has this window been written yet?

        long long c0 = lo;
        while (c0 < hi) {
            // ---- sub-chunk [c0, c1): at most F_RMAX reads and F_CAPW words; every thread derives
            // the same bounds from the same (broadcast) loads -- no elected thread, no extra barrier
            long long c1 = c0 + F_RMAX < hi ? c0 + F_RMAX : hi;
            const long long wa = c0 == lo ? (long long)ix.z : (long long)(b.seq_off[c0] & ~3u);
            long long wend = c1 == hi ? (long long)ix.w : (long long)b.seq_off[c1];
            bool skip = false;
            while (wend - wa > F_CAPW) {
                if (c1 - c0 == 1) { skip = true; break; }  // one read too long to stage: never simple
                c1 = c0 + (c1 - c0) / 2;
                wend = (long long)b.seq_off[c1];
            }
            if (skip) { c0 = c1; continue; }
            const int n_sub = (int)(c1 - c0);
            __syncthreads();  // previous sub-chunk (or tile) fully consumed
            // bases: ONE bulk copy of [wa, wend) by the TMA engine, completing on sm.bar; the
            // threads meanwhile fetch the per-read metadata.  16-byte granules; the (at most one)
            // partial granule at the very end of the array is copied by hand.
            {
                const long long n_words = wend - wa;
                const long long avail = b.seq4_words - wa;
                const long long want = (n_words + 3) & ~3ll;
                const long long bulk_words = want <= avail ? want : (avail & ~3ll);
                if (tid == 0) {
                    mbar_expect_tx(&sm.bar, (uint32_t)(bulk_words * 4));
                    if (bulk_words) bulk_g2s(sm.seq, b.seq4 + wa, (uint32_t)(bulk_words * 4), &sm.bar);
                }
                if (bulk_words < n_words && tid < 4) {
                    const long long w = bulk_words + tid;
                    sm.seq[w] = w < avail ? b.seq4[wa + w] : 0u;
                }
            }
            int* diff = sm.diff[dbuf];
            {   // metadata: all loads of this thread's (up to 4) reads first, then the stores
                int l[F_RMAX / F_THREADS], rs[F_RMAX / F_THREADS];
                uint32_t so[F_RMAX / F_THREADS];
                if (c0 == lo) {  // first sub-chunk: already in shared memory (prefetched during the last tile)
                    cp_async_wait_all();
#pragma unroll
                    for (int k = 0; k < F_RMAX / F_THREADS; ++k) {
                        const int i = tid + k * F_THREADS;
                        const int ii = i < n_sub ? i : tid;
                        l[k] = sm.raw[0][ii];
                        rs[k] = sm.raw[1][ii];
                        so[k] = (uint32_t)sm.raw[2][ii];
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < F_RMAX / F_THREADS; ++k) {
                        const int i = tid + k * F_THREADS;
                        const long long r = c0 + (i < n_sub ? i : 0);
                        l[k] = b.l_seq[r];
                        rs[k] = b.ref_start[r];
                        so[k] = b.seq_off[r];
                    }
                }
#pragma unroll
                for (int k = 0; k < F_RMAX / F_THREADS; ++k) {
                    const int i = tid + k * F_THREADS;
                    if (i < n_sub) {
                        long long g;
                        if (one_contig) {
                            g = slot_base + rs[k];
                        } else {
                            const int c = find_contig(b.contig_read_off, b.n_contigs, c0 + i);
                            g = b.contig_slot[c] + rs[k] - tile_slot;
                        }
                        g = g < -0x10000000ll ? -0x10000000ll : (g > 0x10000000ll ? 0x10000000ll : g);
                        const int gs = (int)g;
                        int nb = 0;
                        if (l[k] > 0) {  // simple read (bit 31 clear)
                            nb = ((l[k] + 7) >> 3) << 2;
                            const int cs = gs < 0 ? 0 : gs, ce = gs + l[k] > KDL_TILE ? KDL_TILE : gs + l[k];
                            if (cs < ce) {
                                atomicAdd(diff + cs, 1);
                                atomicAdd(diff + ce, -1);
                            }
                        }
                        sm.gs[i] = gs;
                        sm.meta[i + (i >> 3)] = make_int4(((gs + 7) >> 3) << 2,
                                               (int)(seq_base + (uint32_t)(((long long)so[k] - wa) << 2)), nb,
                                               ((-gs) & 7) << 2);
                    }
                }
                if (tid < 72) {  // sentinels: never overlap anything
                    const int i = n_sub + tid;
                    if (tid < 32) sm.gs[i] = 0x10000000;
                    sm.meta[i + (i >> 3)] = make_int4(0x10000000, (int)seq_base, 0, 0);
                }
                if (raw_pending) {  // this thread's raw elements are consumed: refill them for the next tile
                    prefetch_raw(tile + gridDim.x);
                    raw_pending = false;
                }
                int* other = sm.diff[dbuf ^ 1];  // clean the buffer the NEXT sub-chunk will use
                for (int k = tid; k < KDL_TILE + 32; k += F_THREADS) other[k] = 0;
            }
            __syncthreads();                 // metadata + difference array complete
            mbar_wait(&sm.bar, bar_parity);  // bases landed
            bar_parity ^= 1u;
            dbuf ^= 1;

            // ---- coverage of this warp's 64 slots: prefix sum of the difference array ------------
            {
                int pre = 0;
                for (int k = lane; k < wlo; k += 32) pre += diff[k];
#pragma unroll
                for (int d = 16; d; d >>= 1) pre += __shfl_xor_sync(0xffffffffu, pre, d);
                const int d0 = diff[wlo + 2 * lane], d1 = diff[wlo + 2 * lane + 1];
                int run = d0 + d1;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const int o = __shfl_up_sync(0xffffffffu, run, d);
                    if (lane >= d) run += o;
                }
                const int before = pre + run - d0 - d1;
                sm.cov[wlo + 2 * lane] = before + d0;
                sm.cov[wlo + 2 * lane + 1] = before + d0 + d1;
                __syncwarp();
                const int4 ca = *reinterpret_cast<const int4*>(sm.cov + wlo + oslot);
                const int4 cb = *reinterpret_cast<const int4*>(sm.cov + wlo + oslot + 4);
                covacc[0] += ca.x; covacc[1] += ca.y; covacc[2] += ca.z; covacc[3] += ca.w;
                covacc[4] += cb.x; covacc[5] += cb.y; covacc[6] += cb.z; covacc[7] += cb.w;
            }

            // ---- this warp's window against the sub-chunk: reads with start in (wlo - maxlen, wlo + 64)
            // two lower bounds over the sorted starts, each in two 32-wide probe rounds (n_sub <= 1024)
            const int a = lower_bound_warp(sm.gs, n_sub, wlo - maxlen + 1, lane);
            const int e = lower_bound_warp(sm.gs, n_sub, wlo + F_WIN, lane);

            for (int base = a & ~7; base < e; base += 64) {
                // 8 reads per lane and block: group g takes the 8 consecutive reads base + 8g .. + 7; a lane
                // needs THREE words of a read for its 16 slots (two funnel shifts).  As in K1f there is no
                // bounds logic: words outside [0, n_words) are predicated off and read as zero.
                uint32_t x0[8], x1[8];
                int4 mt[8];
                const int i0 = base + 8 * grp;
                const int4* mp = sm.meta + i0 + (i0 >> 3);
#pragma unroll
                for (int u = 0; u < 8; ++u) mt[u] = mp[u];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const uint32_t jb = (uint32_t)(p8b - mt[u].x);  // byte offset of the read's first needed word
                    const uint32_t addr = (uint32_t)mt[u].y + jb;
                    uint32_t w0, w1, w2;
#ifndef KDL_HOST_EMU
                    asm("{\n"
                        ".reg .pred p, q, r;\n"
                        "setp.lt.u32 p, %3, %4;\n"
                        "setp.lt.u32 q, %5, %4;\n"
                        "setp.lt.u32 r, %6, %4;\n"
                        "mov.u32 %0, 0;\n"
                        "mov.u32 %1, 0;\n"
                        "mov.u32 %2, 0;\n"
                        "@p ld.shared.u32 %0, [%7];\n"
                        "@q ld.shared.u32 %1, [%7+4];\n"
                        "@r ld.shared.u32 %2, [%7+8];\n"
                        "}\n"
                        : "=&r"(w0), "=&r"(w1), "=&r"(w2)
                        : "r"(jb), "r"((uint32_t)mt[u].z), "r"(jb + 4u), "r"(jb + 8u), "r"(addr));
#else
                    w0 = jb < (uint32_t)mt[u].z ? lds_u32(addr) : 0u;
                    w1 = jb + 4u < (uint32_t)mt[u].z ? lds_u32(addr + 4u) : 0u;
                    w2 = jb + 8u < (uint32_t)mt[u].z ? lds_u32(addr + 8u) : 0u;
#endif
                    x0[u] = __funnelshift_l(w1, w0, (uint32_t)mt[u].w);
                    x1[u] = __funnelshift_l(w2, w1, (uint32_t)mt[u].w);
                }
                acc0.add8(x0);
                acc1.add8(x1);
                if (++blocks_since_flush == F_FLUSH_BLOCKS) {
                    if (kFresh && !stored)
                        flush_window_wide<true, false>(acc0, acc1, rawacc, covacc, counts, n_slots, tile_slot + wlo, lane);
                    else
                        flush_window_wide<false, false>(acc0, acc1, rawacc, covacc, counts, n_slots, tile_slot + wlo, lane);
                    stored = true;
                    blocks_since_flush = 0;
                }
            }
            c0 = c1;
        }
        if (kFresh && !stored)
            flush_window_wide<true, true>(acc0, acc1, rawacc, covacc, counts, n_slots, tile_slot + wlo, lane);
        else
            flush_window_wide<false, true>(acc0, acc1, rawacc, covacc, counts, n_slots, tile_slot + wlo, lane);
    }
}

}  // namespace kdl
