// kdl_common.cuh -- shared device helpers for the kindel_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/kindel_b200.h"

// The CTA's dynamic shared memory.  (KDL_HOST_EMU: tests/emu/ compiles the kernels for the host, where the
// array is an ordinary global and `__shared__` variables are statics; the device build never defines it.)
#ifndef KDL_HOST_EMU
#define KDL_DYNAMIC_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#else
#define KDL_DYNAMIC_SMEM(name) extern unsigned char name[]
#endif

namespace kdl {

// BAM nibble ("=ACMGRSVTWYHKDBN") -> weight column 0..4 (A,C,G,T,N) or -1.  Only the five keys
// of the reference's per-position dicts exist (kindel/kindel.py:29); anything else is a KeyError
// when it is used by an M or S op (kindel.py:52,72,79).  Packed 4 bits per entry: 0xF = invalid.
__device__ __forceinline__ int nib2col(int nib) {
    // nib:            0 1 2 3 4 5 6 7 8 9 a b c d e f
    // col (F = bad):  F 0 1 F 2 F F F 3 F F F F F F 4
    const unsigned long long lut = 0x4FFFFFF3FFF2F10FULL;
    int v = (int)((lut >> (nib * 4)) & 0xF);
    return v == 0xF ? -1 : v;
}

// 8 bases per 32-bit word, first base in the most significant nibble
__device__ __forceinline__ int nibble_at(const uint32_t* __restrict__ seq, long long q) {
    return (int)((seq[q >> 3] >> (28 - 4 * (int)(q & 7))) & 0xFu);
}

// SEQ length of a complex read's l_seq word: tile-eligible reads keep it in bits 0..15 (bits 16..22 hold their
// M-op count), KDL_HARD reads in bits 0..29
__device__ __forceinline__ long long complex_len(uint32_t lraw) {
    return (lraw & KDL_HARD) ? (long long)(lraw & 0x3fffffffu) : (long long)(lraw & KDL_LEN_MASK);
}

// Python list indexing (list length n): negative indices wrap once, otherwise IndexError (-1).
__device__ __forceinline__ long long pyindex(long long i, long long n) {
    if (i < 0) i += n;
    return (i < 0 || i >= n) ? -1 : i;
}

// contig c with contig_read_off[c] <= r < contig_read_off[c+1] (empty contigs skipped naturally)
__device__ __forceinline__ int find_contig(const int64_t* __restrict__ off, int n_contigs, long long r) {
    int lo = 0, hi = n_contigs;  // first index in (lo, hi] with off[idx] > r
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (off[mid + 1] > r) hi = mid; else lo = mid + 1;
    }
    return lo;
}

// consensus() over the five base counts (kindel/kindel.py:369-381): first maximum in dict order
// A,T,G,C,N; all-zero -> N with no tie; tie = another key holds the same non-zero maximum.
// Returns the emitted code (tie -> 4 = N) and the consensus key's count through *freq.
__device__ __forceinline__ int base_vote(int a, int c, int g, int t, int n, int* freq, int* raw_base) {
    int best = a, code = 0;
    if (t > best) { best = t; code = 3; }
    if (g > best) { best = g; code = 2; }
    if (c > best) { best = c; code = 1; }
    if (n > best) { best = n; code = 4; }
    const int ties = (a == best) + (c == best) + (g == best) + (t == best) + (n == best);
    if (best == 0) code = 4;  // ("N", 0): sum == 0 (counts are non-negative)
    *freq = best;
    *raw_base = code;
    return (best != 0 && ties > 1) ? 4 : code;
}

// One slot of consensus_sequence (kindel/kindel.py:402-424) in integer arithmetic:
//   del > 0.5*depth            <=> 2*del > depth
//   depth < min_depth          <=> depth < ceil(min_depth)
//   ins > min(0.5*d, 0.5*dn)   <=> 2*ins > min(d, dn)
__device__ __forceinline__ unsigned vote_slot(int a, int c, int g, int t, int n, int del, int ins,
                                              long long depth_next, long long min_depth_ceil) {
    const long long depth = (long long)a + c + g + t;  // N excluded (kindel.py:404)
    if (2ll * del > depth) return (1u << 4) | 4u;
    if (depth < min_depth_ceil) return (2u << 4) | 4u;
    const long long thr = depth < depth_next ? depth : depth_next;
    const unsigned change = (2ll * ins > thr) ? 3u : 0u;
    int freq, raw;
    const int code = base_vote(a, c, g, t, n, &freq, &raw);
    return (change << 4) | (unsigned)code;
}

}  // namespace kdl
