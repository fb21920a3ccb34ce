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
//     (14 LOP3) per 8 reads, instead of 8 x 150 read-modify-writes.  N (all four bits set) is
//     counted in a second set of planes and subtracted from A,C,G,T at the end.
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
#include "kdl_common.cuh"

namespace kdl {

constexpr int F_THREADS = 256;
constexpr int F_WIN = 64;              // slots per warp window
constexpr int F_RMAX = 1024;           // reads per staged sub-chunk
constexpr int F_CAPW = 20480;          // seq words per staged sub-chunk (80 KB)
constexpr int F_P = 8;                 // bit planes per stream: up to 255 reads between flushes
constexpr int F_FLUSH_BLOCKS = 31;     // 31 blocks x 8 reads = 248 <= 255

// ---- K0: per tile, the index range of reads whose first base lies in (tile_lo - maxlen, tile_hi)
// global slot of a read's first base = contig_slot[c] + ref_start; reads are sorted by it.
__device__ __forceinline__ long long first_read_at_or_after(const kdl_batch& b, long long g) {
    // first read index whose global start slot is >= g
    if (b.n_contigs == 0) return 0;
    // contig whose slot range contains g (or the first contig after g)
    int lo = 0, hi = b.n_contigs;  // first contig with slot + len + 1 > g
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (b.contig_slot[mid] + b.contig_len[mid] + 1 > g) hi = mid; else lo = mid + 1;
    }
    if (lo >= b.n_contigs) return b.n_reads;
    const long long p = g - b.contig_slot[lo];  // position inside contig `lo` (may be < 0)
    long long a = b.contig_read_off[lo], e = b.contig_read_off[lo + 1];
    while (a < e) {
        const long long mid = (a + e) >> 1;
        if ((long long)b.ref_start[mid] >= p) e = mid; else a = mid + 1;
    }
    return a;
}

__global__ void __launch_bounds__(256)
tile_index_kernel(kdl_batch b, long long n_tiles, uint32_t* __restrict__ index) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tiles) return;
    const long long g0 = t * KDL_TILE;
    index[2 * t] = (uint32_t)first_read_at_or_after(b, g0 - b.max_simple_len + 1);
    index[2 * t + 1] = (uint32_t)first_read_at_or_after(b, g0 + KDL_TILE);
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
__device__ __forceinline__ void flush_window(Planes& acc, Planes& accn, int32_t* __restrict__ counts,
                                             long long n_slots, long long slot0, int lane) {
    uint32_t m[F_P + 2], n[F_P + 2];
    quarter_sum(acc.p, m);
    quarter_sum(accn.p, n);
    acc.clear();
    accn.clear();
    const int q = lane >> 3;
    int cn[8], cv[8];
    extract8(n, 0, cn);
    extract8(m, q, cv);
    const long long s = slot0 + 8 * (lane & 7);
    int4* dst = reinterpret_cast<int4*>(counts + (long long)q * n_slots + s);
    int4 v0 = dst[0], v1 = dst[1];
    v0.x += cv[0] - cn[0]; v0.y += cv[1] - cn[1]; v0.z += cv[2] - cn[2]; v0.w += cv[3] - cn[3];
    v1.x += cv[4] - cn[4]; v1.y += cv[5] - cn[5]; v1.z += cv[6] - cn[6]; v1.w += cv[7] - cn[7];
    dst[0] = v0;
    dst[1] = v1;
    if (q == 0) {
        int4* dn = reinterpret_cast<int4*>(counts + (long long)KDL_W_N * n_slots + s);
        int4 n0 = dn[0], n1 = dn[1];
        n0.x += cn[0]; n0.y += cn[1]; n0.z += cn[2]; n0.w += cn[3];
        n1.x += cn[4]; n1.y += cn[5]; n1.z += cn[6]; n1.w += cn[7];
        dn[0] = n0;
        dn[1] = n1;
    }
}

struct FastSmem {
    uint32_t seq[F_CAPW];
    int gs[F_RMAX];        // start slot of the read relative to the tile's first slot
    uint32_t ww[F_RMAX];   // word offset in seq[] (low 16 bits) | n_words << 16 (0 = not a simple read)
    int c1;                // end of the current sub-chunk (broadcast)
    int skip;              // 1 = sub-chunk does not fit: its (necessarily complex) reads are skipped
};

__global__ void __launch_bounds__(F_THREADS, 2)
pileup_tiled_kernel(kdl_batch b, int32_t* __restrict__ counts, long long n_slots,
                    const uint32_t* __restrict__ tile_index, long long n_tiles) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    FastSmem& sm = *reinterpret_cast<FastSmem*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int quarter = lane >> 3;

    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long lo = tile_index[2 * tile], hi = tile_index[2 * tile + 1];
        if (lo >= hi) continue;  // uniform for the CTA
        const long long tile_slot = tile * KDL_TILE;
        const int p0 = warp * F_WIN + 8 * (lane & 7);  // lane's first slot, tile-relative
        Planes acc, accn;
        acc.clear();
        accn.clear();
        int blocks_since_flush = 0;

        long long c0 = lo;
        while (c0 < hi) {
            // ---- sub-chunk [c0, c1): at most F_RMAX reads and F_CAPW words ------------------------
            __syncthreads();  // previous sub-chunk fully consumed
            if (tid == 0) {
                long long c1 = c0 + F_RMAX < hi ? c0 + F_RMAX : hi;
                const long long wa = (long long)(b.seq_off[c0] & ~3u);
                int skip = 0;
                for (;;) {
                    const long long wend = c1 < b.n_reads ? (long long)b.seq_off[c1] : b.seq4_words;
                    if (wend - wa <= F_CAPW) break;
                    if (c1 - c0 == 1) { skip = 1; break; }
                    c1 = c0 + (c1 - c0) / 2;
                }
                sm.c1 = (int)(c1 - c0);
                sm.skip = skip;
            }
            __syncthreads();
            const long long c1 = c0 + sm.c1;
            const int n_sub = sm.c1;
            if (sm.skip) { c0 = c1; continue; }  // a read too long to stage is never a simple read
            const long long wa = (long long)(b.seq_off[c0] & ~3u);
            const long long wend = c1 < b.n_reads ? (long long)b.seq_off[c1] : b.seq4_words;
            // metadata
            for (int i = tid; i < n_sub; i += F_THREADS) {
                const long long r = c0 + i;
                const int l = b.l_seq[r];
                // every read keeps its place in the start-slot order (the flatten step guarantees it is
                // non-decreasing over ALL reads); complex reads get n_words = 0 and add nothing
                const int c = find_contig(b.contig_read_off, b.n_contigs, r);
                long long g = b.contig_slot[c] + b.ref_start[r] - tile_slot;
                g = g < -0x20000000ll ? -0x20000000ll : (g > 0x20000000ll ? 0x20000000ll : g);
                uint32_t ww = 0;
                if (l > 0)  // simple read (bit 31 clear)
                    ww = (uint32_t)((long long)b.seq_off[r] - wa) | ((uint32_t)((l + 7) >> 3) << 16);
                sm.gs[i] = (int)g;
                sm.ww[i] = ww;
            }
            // bases: 128-bit copies of [wa, wend)
            {
                const long long n_vec = (wend - wa + 3) >> 2;
                const uint4* src = reinterpret_cast<const uint4*>(b.seq4 + wa);
                uint4* dst = reinterpret_cast<uint4*>(sm.seq);
                for (long long v = tid; v < n_vec; v += F_THREADS) {
                    if (wa + 4 * v + 4 <= b.seq4_words) {
                        dst[v] = __ldg(src + v);
                    } else {  // last, partial vector of the whole array
                        uint32_t t4[4] = {0u, 0u, 0u, 0u};
                        for (int k = 0; k < 4; ++k)
                            if (wa + 4 * v + k < b.seq4_words) t4[k] = b.seq4[wa + 4 * v + k];
                        dst[v] = make_uint4(t4[0], t4[1], t4[2], t4[3]);
                    }
                }
            }
            __syncthreads();

            // ---- this warp's window against the sub-chunk --------------------------------------
            const int wlo = warp * F_WIN;
            int a, e;  // reads of the sub-chunk that can reach [wlo, wlo + 64): gs in (wlo - maxlen, wlo + 64)
            {
                int l0 = 0, h0 = n_sub;
                while (l0 < h0) {
                    const int mid = (l0 + h0) >> 1;
                    if (sm.gs[mid] + b.max_simple_len > wlo) h0 = mid; else l0 = mid + 1;
                }
                a = l0;
                int l1 = a, h1 = n_sub;
                while (l1 < h1) {
                    const int mid = (l1 + h1) >> 1;
                    if (sm.gs[mid] >= wlo + F_WIN) h1 = mid; else l1 = mid + 1;
                }
                e = l1;
            }

            for (int base = a; base < e; base += 32) {
                uint32_t x[8], xn[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int i = base + 4 * u + quarter;
                    uint32_t w = 0;
                    if (i < e) {
                        const uint32_t ww = sm.ww[i];
                        const int o = p0 - sm.gs[i];
                        const int j = o >> 3;
                        const int nw = (int)(ww >> 16);
                        const uint32_t* s = sm.seq + (ww & 0xFFFFu);
                        const uint32_t hw = ((unsigned)j < (unsigned)nw) ? s[j] : 0u;
                        const uint32_t lw = ((unsigned)(j + 1) < (unsigned)nw) ? s[j + 1] : 0u;
                        w = __funnelshift_l(lw, hw, o << 2);
                    }
                    x[u] = w;
                    xn[u] = w & (w >> 1) & 0x11111111u;  // nibble 15 (N): bits 0 and 1 both set
                }
                acc.add8(x);
                accn.add8(xn);
                if (++blocks_since_flush == F_FLUSH_BLOCKS) {
                    flush_window(acc, accn, counts, n_slots, tile_slot + wlo, lane);
                    blocks_since_flush = 0;
                }
            }
            c0 = c1;
        }
        if (blocks_since_flush) flush_window(acc, accn, counts, n_slots, tile_slot + warp * F_WIN, lane);
    }
}

}  // namespace kdl
