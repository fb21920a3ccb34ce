// pileup_tiled.cu -- K0 tile index + K1f: owner-computes pileup of simple (nM) reads.
//
// What it computes is kindel/kindel.py:49-54 for reads whose CIGAR is one match op: every base
// adds 1 to weights[position][base].  How: as a positional population count.
//
//   * The slot space is cut into tiles of KDL_TILE = 512 slots; one CTA owns a tile, each of its 8
//     warps owns a 64-slot window, each lane owns 8 consecutive slots.  Because the flattened reads
//     are coordinate-sorted, the reads that can touch a tile are ONE contiguous index range, found
//     once per tile boundary by K0 (a binary search per boundary; this is the BAM linear index,
//     built on the device).  Ownership is exclusive, so the kernel needs no atomics at all.
//   * A BAM base is a one-hot nibble (A=1 C=2 G=4 T=8, N=15).  With 8 bases per 32-bit word, first
//     base in the top nibble, the 8 bases a read puts on a lane's 8 slots are ONE funnel shift of
//     two staged words; out-of-range words read as zero and add nothing.
//   * Counting is bit-sliced: the 32 bits of that word are 32 independent 1-bit inputs (8 slots x
//     A,C,G,T) added into vertical counters with a Harley-Seal carry-save tree -- 7 full adders
//     (14 LOP3) per 8 reads, instead of 8 x 150 read-modify-writes.  N (all four bits set) is not
//     counted at all: it is recovered per slot from A+C+G+T = coverage + 3N (see flush_window).
//   * The 4 quarter-warps walk 4 different reads at once (a 150-base read covers ~6 of 8 lanes of a
//     64-slot window, so quarter-warps keep ~80 % of lanes busy where a full warp would keep 40 %);
//     their planes are summed bit-sliced by two shuffle butterflies, transposed to integers once
//     per window and added to the table with 128-bit loads/stores.
//   * The read bytes of a tile are one contiguous range of seq4 (reads are laid out in read
//     order): staged into shared memory with 128-bit loads, each byte fetched ~1.3x (tile halo),
//     the second fetch normally an L2 hit because neighbouring tiles run concurrently.
//
// Preconditions (checked on the host side of the ABI, include/kindel_b200.h): reads_sorted,
// seq_off non-decreasing, simple reads clean (A,C,G,T,N only, trailing nibbles zero) and no longer
// than KDL_FAST_MAXLEN.  Anything else is a complex read and belongs to K1g.
#include <stddef.h>

#include "kdl_common.cuh"

namespace kdl {

constexpr int F_THREADS = 256;
constexpr int F_WIN = 64;              // slots per warp window
constexpr int F_RMAX = 1024;           // reads per staged sub-chunk
constexpr int F_CAPW = 17920;          // seq words per staged sub-chunk (70 KB): 2 CTAs x ~111 KB per SM
constexpr int F_P = 8;                 // bit planes per stream: up to 255 reads between flushes
constexpr int F_FLUSH_BLOCKS = 31;     // 31 blocks x 8 reads = 248 <= 255
static_assert(F_FLUSH_BLOCKS % 2 == 1, "kLean folds exactly one pending carry at a mid-window flush");

// ---- K0: per tile, what K1f needs to start without dependent global loads: 8 x int32
//   [0] lo, [1] hi   index range of reads whose first base lies in (tile_lo - maxlen, tile_hi)
//   [2] wa, [3] wend word range of their packed bases (wa rounded down to a 16-byte boundary)
//   [4] c_lo, [5] c_hi contigs of read lo and of read hi - 1
// global slot of a read's first base = contig_slot[c] + ref_start; reads are sorted by it.
constexpr int F_IDX = 8;  // int32 per tile in the index

// First read index whose global start slot is >= g, for TWO keys at once (a tile's lower and upper
// bound).  Warp-cooperative: the contig is found by every lane (few contigs), the read by a 32-ary
// search -- each round the 32 lanes probe 32 evenly spaced elements of the remaining range in ONE
// memory round trip (5 rounds for 10^7 reads, not 24); the two searches advance in lockstep so
// their round trips overlap.
struct Search {
    long long a, e, p;  // invariant: reads before a are < p, reads from e on are >= p (or e = end)
    bool live;
};

__device__ __forceinline__ Search search_begin(const kdl_batch& b, long long g) {
    Search s;
    s.live = false;
    s.a = s.e = 0;
    s.p = 0;
    if (b.n_contigs == 0) return s;
    int lo = 0, hi = b.n_contigs;  // first contig with slot + len + 1 > g
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (b.contig_slot[mid] + b.contig_len[mid] + 1 > g) hi = mid; else lo = mid + 1;
    }
    if (lo >= b.n_contigs) { s.a = s.e = b.n_reads; return s; }
    s.p = g - b.contig_slot[lo];  // position inside contig `lo` (may be < 0)
    s.a = b.contig_read_off[lo];
    s.e = b.contig_read_off[lo + 1];
    s.live = s.a < s.e;
    return s;
}

__device__ __forceinline__ void search_probe(const kdl_batch& b, const Search& s, int lane, long long& step, bool& ge) {
    step = (s.e - s.a + 31) >> 5;
    const long long idx = s.a + (long long)(lane + 1) * step - 1;  // last element of the lane's bucket
    ge = (s.live && idx < s.e) ? ((long long)b.ref_start[idx] >= s.p) : true;
}

__device__ __forceinline__ void search_narrow(Search& s, long long step, bool ge) {
    const unsigned m = __ballot_sync(0xffffffffu, ge);
    if (!s.live) return;
    if (m == 0u) { s.a = s.e; s.live = false; return; }  // even the very last element is < p
    const int k = __ffs(m) - 1;                          // first bucket whose last element is >= p
    const long long na = s.a + (long long)k * step;
    long long ne = s.a + (long long)(k + 1) * step - 1;  // that element is >= p: the answer is <= ne
    if (ne > s.e) ne = s.e;
    s.a = na;
    s.e = ne < na ? na : ne;
    if (step == 1) s.a = s.e;
    s.live = s.a < s.e;
}

__global__ void __launch_bounds__(256)
tile_index_kernel(kdl_batch b, long long tile_lo, long long n_tiles, uint32_t* __restrict__ index) {
    const int lane = threadIdx.x & 31;
    const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;  // one warp per tile
    if (w >= n_tiles) return;
    const long long t = tile_lo + w;
    const long long g0 = t * KDL_TILE;
    Search s1 = search_begin(b, g0 - b.max_simple_len + 1), s2 = search_begin(b, g0 + KDL_TILE);
    while (s1.live || s2.live) {  // warp-uniform
        long long st1, st2;
        bool ge1, ge2;
        search_probe(b, s1, lane, st1, ge1);  // both probes are issued before either is consumed
        search_probe(b, s2, lane, st2, ge2);
        search_narrow(s1, st1, ge1);
        search_narrow(s2, st2, ge2);
    }
    const long long lo = s1.a, hi = s2.a;
    if (lane) return;
    uint32_t* e = index + F_IDX * t;
    e[0] = (uint32_t)lo;
    e[1] = (uint32_t)hi;
    e[2] = lo < b.n_reads ? (b.seq_off[lo] & ~3u) : 0u;
    e[3] = hi < b.n_reads ? b.seq_off[hi] : (uint32_t)b.seq4_words;
    e[4] = lo < hi ? (uint32_t)find_contig(b.contig_read_off, b.n_contigs, lo) : 0u;
    e[5] = lo < hi ? (uint32_t)find_contig(b.contig_read_off, b.n_contigs, hi - 1) : 0u;
    e[6] = 0u;
    e[7] = 0u;
}

// (KDL_HOST_EMU: tests/emu/ compiles this file for the host and supplies functional stand-ins for the PTX
// helpers below; the device build never defines it.)
#ifndef KDL_HOST_EMU
// ---- 1-D bulk copy global -> shared (TMA engine, SASS UBLKCP) completing on an mbarrier --------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// 4-byte asynchronous global -> shared copy (LDGSTS): no register staging, completes in background
__device__ __forceinline__ void cp_async4(void* dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    }
}
#endif  // KDL_HOST_EMU

// first index i in [0, n) with g[i] >= key (n if none); g sorted, n <= 1024, whole warp calls
__device__ __forceinline__ int lower_bound_warp(const int* g, int n, int key, int lane) {
    const int step = (n + 31) >> 5;  // <= 32
    if (step == 0) return 0;
    const int i1 = (lane + 1) * step - 1;
    const unsigned m1 = __ballot_sync(0xffffffffu, i1 < n ? g[i1] >= key : true);
    if (m1 == 0u) return n;       // every probed element (incl. the last one) is < key
    const int k = __ffs(m1) - 1;  // first bucket whose last element is >= key
    const int i2 = k * step + lane;
    const unsigned m2 = __ballot_sync(0xffffffffu, (lane < step && i2 < n) ? g[i2] >= key : true);
    const int r = k * step + __ffs(m2) - 1;
    return r < n ? r : n;
}

// the same for TWO keys at once (key_a <= key_e): one probe load serves both searches and the two dependent
// chains (load -> ballot -> load -> ballot) overlap
__device__ __forceinline__ void lower_bound_warp2(const int* g, int n, int key_a, int key_e, int lane, int& ra, int& re) {
    const int step = (n + 31) >> 5;  // <= 32
    if (step == 0) { ra = re = 0; return; }
    const int i1 = (lane + 1) * step - 1;
    const int v1 = i1 < n ? g[i1] : 0x7fffffff;  // past the end: counts as >= key
    const unsigned ma = __ballot_sync(0xffffffffu, v1 >= key_a);
    const unsigned me = __ballot_sync(0xffffffffu, v1 >= key_e);
    const int ka = __ffs(ma) - 1, ke = __ffs(me) - 1;  // first bucket whose last element is >= key; -1: none
    const int ia = ka * step + lane, ie = ke * step + lane;
    const int va = (ka >= 0 && lane < step && ia < n) ? g[ia] : 0x7fffffff;
    const int ve = (ke >= 0 && lane < step && ie < n) ? g[ie] : 0x7fffffff;
    const unsigned m2a = __ballot_sync(0xffffffffu, va >= key_a);
    const unsigned m2e = __ballot_sync(0xffffffffu, ve >= key_e);
    const int qa = ka * step + __ffs(m2a) - 1, qe = ke * step + __ffs(m2e) - 1;
    ra = (ka < 0 || qa > n) ? n : qa;
    re = (ke < 0 || qe > n) ? n : qe;
}

// ---- bit-sliced counters ------------------------------------------------------------------------
__device__ __forceinline__ void csa(uint32_t& carry, uint32_t& sum, uint32_t a, uint32_t b, uint32_t c) {
    const uint32_t s = a ^ b ^ c;
    carry = (a & b) | (c & (a | b));
    sum = s;
}

struct Planes {
    uint32_t p[F_P];  // p[k] = bit k of 32 vertical counters
    __device__ __forceinline__ void clear() {
#pragma unroll
        for (int k = 0; k < F_P; ++k) p[k] = 0;
    }
    // Harley-Seal: 8 one-bit-per-counter inputs -> 7 full adders + a ripple from the 8s plane up
    __device__ __forceinline__ void add8(const uint32_t (&x)[8]) {
        uint32_t ta, tb, tc, td, fa, fb, e;
        csa(ta, p[0], p[0], x[0], x[1]);
        csa(tb, p[0], p[0], x[2], x[3]);
        csa(fa, p[1], p[1], ta, tb);
        csa(tc, p[0], p[0], x[4], x[5]);
        csa(td, p[0], p[0], x[6], x[7]);
        csa(fb, p[1], p[1], tc, td);
        csa(e, p[2], p[2], fa, fb);
#pragma unroll
        for (int k = 3; k < F_P; ++k) {
            const uint32_t t = p[k] & e;
            p[k] ^= e;
            e = t;
        }
    }
    // kLean: the same 7 full adders, but the carry out of the 4s plane (weight 8) is handed back instead of being
    // rippled up; the caller pairs two of them with one more full adder, so the ripple runs once per 16 reads
    __device__ __forceinline__ uint32_t add8_carry(const uint32_t (&x)[8]) {
        uint32_t ta, tb, tc, td, fa, fb, e;
        csa(ta, p[0], p[0], x[0], x[1]);
        csa(tb, p[0], p[0], x[2], x[3]);
        csa(fa, p[1], p[1], ta, tb);
        csa(tc, p[0], p[0], x[4], x[5]);
        csa(td, p[0], p[0], x[6], x[7]);
        csa(fb, p[1], p[1], tc, td);
        csa(e, p[2], p[2], fa, fb);
        return e;
    }
    template <int K0>  // add one plane of weight 2^K0
    __device__ __forceinline__ void ripple(uint32_t e) {
#pragma unroll
        for (int k = K0; k < F_P; ++k) {
            const uint32_t t = p[k] & e;
            p[k] ^= e;
            e = t;
        }
    }
};

// sum of the same planes held by the 4 quarter-warps (lanes l, l^8, l^16, l^24): bit-sliced ripple
// adders over two butterfly stages; F_P planes in, F_P + 2 planes out, identical in all 4 lanes.
__device__ __forceinline__ void quarter_sum(const uint32_t (&in)[F_P], uint32_t (&out)[F_P + 2]) {
    uint32_t a[F_P + 2];
#pragma unroll
    for (int k = 0; k < F_P; ++k) a[k] = in[k];
    a[F_P] = 0;
    a[F_P + 1] = 0;
#pragma unroll
    for (int stage = 0; stage < 2; ++stage) {
        const int width = F_P + stage;  // planes that can be non-zero before this stage
        uint32_t carry = 0;
#pragma unroll
        for (int k = 0; k < F_P + 2; ++k) {
            if (k < width) {
                const uint32_t o = __shfl_xor_sync(0xffffffffu, a[k], 8 << stage);
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
    for (int k = 0; k < F_P + 2; ++k) out[k] = a[k];
}

// counters of bit `bit` (0..3) of every nibble: 8 integers, out[b] for the lane's slot b
// (slot b sits in nibble 7-b).  Planes: F_P + 2 <= 12 bits per counter.
__device__ __forceinline__ void extract8(const uint32_t (&pl)[F_P + 2], int bit, int (&out)[8]) {
    uint32_t v[3] = {0u, 0u, 0u};  // 4 planes per packed word: nibble j of v[g] = bits 4g..4g+3 of counter j
#pragma unroll
    for (int k = 0; k < F_P + 2; ++k) v[k >> 2] |= ((pl[k] >> bit) & 0x11111111u) << (k & 3);
#pragma unroll
    for (int b = 0; b < 8; ++b) {
        const int sh = 4 * (7 - b);
        out[b] = (int)(((v[0] >> sh) & 0xFu) | (((v[1] >> sh) & 0xFu) << 4) | (((v[2] >> sh) & 0xFu) << 8));
    }
}

// Add the window's counters to the table.  Quarter q owns column q (A,C,G,T); quarter 0 also adds
// column 4 (N).  Each lane holds 8 consecutive slots -> two 128-bit read-modify-writes per column;
// no other thread of the grid touches these slots during this kernel.
// ---- flush: planes -> integers -> table ----------------------------------------------------------
// N is not counted: an N nibble (15) adds 1 to all four of A,C,G,T, so for every slot
//     A_raw + C_raw + G_raw + T_raw = cov + 3 N        (cov = simple reads covering the slot)
// and cov comes from a +1/-1 difference array over read starts/ends (two shared-memory atomics per
// staged read, one prefix sum per window).  The correction is applied by the FINAL flush of a window
// (kFinal); earlier flushes -- only needed when more than 248 reads per stream pile up on one
// window -- add raw counts and remember the raw total in `rawacc`.
// Quarter q owns column q (A,C,G,T); quarter 0 also writes column 4 (N).  Each lane holds 8
// consecutive slots: two 128-bit stores (kStore) or read-modify-writes per column; no other thread
// of the grid touches these slots during this kernel.
template <bool kStore, bool kFinal>
__device__ __forceinline__ void flush_window(Planes& acc, int (&rawacc)[8], const int (&covacc)[8],
                                             int32_t* __restrict__ counts, long long n_slots, long long slot0,
                                             int lane) {
    const int q = lane >> 3;
    const long long s = slot0 + 8 * (lane & 7);
    int4* dst = reinterpret_cast<int4*>(counts + (long long)q * n_slots + s);
    int4 v0 = make_int4(0, 0, 0, 0), v1 = v0;
    if (!kStore) { v0 = dst[0]; v1 = dst[1]; }  // issued first: latency hides behind the transposition
    uint32_t m[F_P + 2];
    quarter_sum(acc.p, m);
    acc.clear();
    int cv[8], tot[8];
    extract8(m, q, cv);
#pragma unroll
    for (int k = 0; k < 8; ++k) {  // A+C+G+T raw of each slot: sum of the four quarters' columns
        int t = cv[k];
        t += __shfl_xor_sync(0xffffffffu, t, 8);
        t += __shfl_xor_sync(0xffffffffu, t, 16);
        tot[k] = t + rawacc[k];
    }
    int nn[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        if (kFinal) {
            nn[k] = (tot[k] - covacc[k]) / 3;  // exact by construction
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
    if (q == 0 && (kFinal || kStore)) {
        int4* dn = reinterpret_cast<int4*>(counts + (long long)KDL_W_N * n_slots + s);
        int4 n0 = make_int4(0, 0, 0, 0), n1 = n0;
        if (!kStore) { n0 = dn[0]; n1 = dn[1]; }
        n0.x += nn[0]; n0.y += nn[1]; n0.z += nn[2]; n0.w += nn[3];
        n1.x += nn[4]; n1.y += nn[5]; n1.z += nn[6]; n1.w += nn[7];
        dn[0] = n0;
        dn[1] = n1;
    }
}

struct FastSmem {
    uint32_t seq[F_CAPW];
    // per staged read (32 sentinels follow the last one):
    //   .x  byte offset, relative to the tile, of the first 8-slot group the read can serve:
    //       4 * ceil(start / 8)            (lane byte offset - this = word of the read, in bytes)
    //   .y  shared-memory address (u32) of the read's first word
    //   .z  bytes of packed bases (0 = not a simple read: adds nothing)
    //   .w  funnel-shift amount 4 * ((-start) & 7)
    int4 meta[F_RMAX + 40 + (F_RMAX + 40) / 8];  // entry of read i at i + i/8 (one pad per 8: see main loop)
    int gs[F_RMAX + 32];          // start slot relative to the tile (all reads: the array stays sorted)
    int diff[2][KDL_TILE + 32];   // +1 at read start, -1 at read end (double-buffered per sub-chunk)
    int cov[KDL_TILE];            // prefix sums of diff: simple reads covering each slot
    int raw[3][F_RMAX];           // l_seq / ref_start / seq_off of the NEXT tile's first reads (cp.async)
    uint64_t bar;                 // mbarrier the bulk copy of seq[] completes on
};

static_assert(offsetof(FastSmem, diff) % 16 == 0 && (sizeof(int) * (KDL_TILE + 32)) % 16 == 0,
              "the difference arrays are read with 128-bit loads");

// kFresh: columns 0..4 hold stale data; the first flush of every window stores instead of adding,
// and windows / tiles without reads are stored as zeros.
//
// kLean (experimental, KDL_K1F=lean; not yet measured): the same kernel with two of the latency-bound phases
// that the ncu stall samples charge most (profiles/r01_k1f_final_regions.txt) shortened -- the tile-index entry
// and the contig slot of the NEXT tile are loaded one tile ahead into registers, so neither the tile setup nor
// the metadata prefetch waits on a global load, and the prefix of the difference array is summed with 128-bit
// shared loads.
template <bool kFresh, bool kLean = false>
__global__ void __launch_bounds__(F_THREADS, 2)
pileup_tiled_kernel(kdl_batch b, int32_t* __restrict__ counts, long long n_slots,
                    const uint32_t* __restrict__ tile_index, long long tile_lo, long long n_tiles) {
    KDL_DYNAMIC_SMEM(smem_raw);
    FastSmem& sm = *reinterpret_cast<FastSmem*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int quarter = lane >> 3;
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
    // kLean: index entry (and first contig's slot) of the tile this CTA handles NEXT, held in registers
    uint4 nix = make_uint4(0u, 0u, 0u, 0u);
    uint2 nic = make_uint2(0u, 0u);
    long long ncs = 0;
    auto load_next_index = [&](long long t) {
        if (t < tile_lo + n_tiles) {
            nix = __ldg(reinterpret_cast<const uint4*>(tile_index + F_IDX * t));
            nic = __ldg(reinterpret_cast<const uint2*>(tile_index + F_IDX * t + 4));
            ncs = b.contig_slot[nic.x];
        } else {
            nix = make_uint4(0u, 0u, 0u, 0u);  // lo == hi: nothing to prefetch
        }
    };
    auto prefetch_raw_lean = [&]() {  // the reads of the tile described by nix
        const long long plo = nix.x, phi = nix.y;
        const int cnt = (int)(phi - plo < F_RMAX ? phi - plo : F_RMAX);
        for (int i = tid; i < cnt; i += F_THREADS) {
            cp_async4(&sm.raw[0][i], b.l_seq + plo + i);
            cp_async4(&sm.raw[1][i], b.ref_start + plo + i);
            cp_async4(&sm.raw[2][i], b.seq_off + plo + i);
        }
    };
    if constexpr (kLean) {
        load_next_index(tile_lo + blockIdx.x);
        prefetch_raw_lean();
    } else {
        prefetch_raw(tile_lo + blockIdx.x);
    }

    for (long long tile = tile_lo + blockIdx.x; tile < tile_lo + n_tiles; tile += gridDim.x) {
        uint4 ix;
        uint2 ic;
        long long first_contig_slot = 0;
        if constexpr (kLean) {
            ix = nix;
            ic = nic;
            first_contig_slot = ncs;
            load_next_index(tile + gridDim.x);  // consumed a whole tile later
        } else {
            ix = __ldg(reinterpret_cast<const uint4*>(tile_index + F_IDX * tile));
            ic = __ldg(reinterpret_cast<const uint2*>(tile_index + F_IDX * tile + 4));
        }
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
            // nothing was prefetched for an empty tile: raw is free
            if constexpr (kLean) prefetch_raw_lean(); else prefetch_raw(tile + gridDim.x);
            continue;
        }
        const bool one_contig = ic.x == ic.y;
        long long slot_base;
        if constexpr (kLean) slot_base = one_contig ? first_contig_slot - tile_slot : 0;
        else slot_base = one_contig ? b.contig_slot[ic.x] - tile_slot : 0;
        const int wlo = warp * F_WIN;
        const int p8b = (wlo >> 1) + 4 * (lane & 7);  // 4 * (lane's first slot / 8): byte offset of its word
        Planes acc;
        acc.clear();
        int rawacc[8], covacc[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) { rawacc[k] = 0; covacc[k] = 0; }
        int blocks_since_flush = 0;
        uint32_t pend8 = 0;   // kLean: weight-8 carry of an odd block, waiting for its partner
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
                        // (the reads of [lo, hi) start inside (tile - maxlen, tile + 512) by construction of the
                        // index; the clamp only guards the int conversion against a corrupt index)
                        if constexpr (!kLean) g = g < -0x10000000ll ? -0x10000000ll : (g > 0x10000000ll ? 0x10000000ll : g);
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
                if (tid < 40) {  // sentinels: never overlap anything
                    const int i = n_sub + tid;
                    if (tid < 32) sm.gs[i] = 0x10000000;
                    sm.meta[i + (i >> 3)] = make_int4(0x10000000, (int)seq_base, 0, 0);
                }
                if (raw_pending) {  // this thread's raw elements are consumed: refill them for the next tile
                    if constexpr (kLean) prefetch_raw_lean(); else prefetch_raw(tile + gridDim.x);
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
                if constexpr (kLean) {  // wlo is a multiple of 64 and diff is 16-byte aligned
                    for (int k = 4 * lane; k < wlo; k += 128) {
                        const int4 v = *reinterpret_cast<const int4*>(diff + k);
                        pre += (v.x + v.y) + (v.z + v.w);
                    }
                } else {
                    for (int k = lane; k < wlo; k += 32) pre += diff[k];
                }
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
                const int4 ca = *reinterpret_cast<const int4*>(sm.cov + wlo + 8 * (lane & 7));
                const int4 cb = *reinterpret_cast<const int4*>(sm.cov + wlo + 8 * (lane & 7) + 4);
                covacc[0] += ca.x; covacc[1] += ca.y; covacc[2] += ca.z; covacc[3] += ca.w;
                covacc[4] += cb.x; covacc[5] += cb.y; covacc[6] += cb.z; covacc[7] += cb.w;
            }

            // ---- this warp's window against the sub-chunk: reads with start in (wlo - maxlen, wlo + 64)
            // two lower bounds over the sorted starts, each in two 32-wide probe rounds (n_sub <= 1024)
            int a, e;
            if constexpr (kLean) {
                lower_bound_warp2(sm.gs, n_sub, wlo - maxlen + 1, wlo + F_WIN, lane, a, e);
            } else {
                a = lower_bound_warp(sm.gs, n_sub, wlo - maxlen + 1, lane);
                e = lower_bound_warp(sm.gs, n_sub, wlo + F_WIN, lane);
            }

            for (int base = a & ~7; base < e; base += 32) {
                // 8 reads per lane and block: quarter q takes the 8 CONSECUTIVE reads base + 8q .. + 7.
                // With 19-word reads the four quarters then hit four disjoint groups of 8 banks
                // (8 reads = 152 words = 24 mod 32), and the metadata entries -- padded by one per 8
                // reads, so the quarters' entries are 144 B = 4 banks apart -- do not collide either.
                // No bounds logic: a read that does not reach the lane's 8 slots (the up to 7 reads
                // before a, reads [e, ...) right of the window, the sentinels behind the last read)
                // fails both range tests below and contributes zero.
                uint32_t x[8];
                int4 mt[8];
                const int i0 = base + 8 * quarter;
                const int4* mp = sm.meta + i0 + (i0 >> 3);
#pragma unroll
                for (int u = 0; u < 8; ++u) mt[u] = mp[u];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const uint32_t jb = (uint32_t)(p8b - mt[u].x);  // byte offset of the read's word
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
                if constexpr (kLean) {
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
                    if constexpr (kLean) acc.template ripple<3>(pend8);  // F_FLUSH_BLOCKS is odd: one carry is pending
                    if (kFresh && !stored)
                        flush_window<true, false>(acc, rawacc, covacc, counts, n_slots, tile_slot + wlo, lane);
                    else
                        flush_window<false, false>(acc, rawacc, covacc, counts, n_slots, tile_slot + wlo, lane);
                    stored = true;
                    blocks_since_flush = 0;
                }
            }
            c0 = c1;
        }
        if constexpr (kLean) {
            if (blocks_since_flush & 1) acc.template ripple<3>(pend8);
        }
        if (kFresh && !stored) flush_window<true, true>(acc, rawacc, covacc, counts, n_slots, tile_slot + wlo, lane);
        else flush_window<false, true>(acc, rawacc, covacc, counts, n_slots, tile_slot + wlo, lane);
    }
}

// zero columns [col_lo, col_hi) of slots [slot_lo, slot_hi) (multiples of 4), 128-bit stores
__global__ void __launch_bounds__(256)
zero_cols_kernel(int32_t* __restrict__ counts, long long n_slots, int col_lo, int col_hi, long long slot_lo,
                 long long slot_hi) {
    const long long per_col = (slot_hi - slot_lo) >> 2;
    const long long total = per_col * (col_hi - col_lo);
    for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < total;
         v += (long long)gridDim.x * blockDim.x) {
        const long long col = col_lo + v / per_col, off = v % per_col;
        reinterpret_cast<int4*>(counts + col * n_slots + slot_lo)[off] = make_int4(0, 0, 0, 0);
    }
}

}  // namespace kdl
