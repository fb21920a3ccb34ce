/* kindel_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C, single-threaded CPU restatement of the hot path of bede/kindel v1.2.1, used as the
 * checker for the CUDA engine (tests/, __graft_entry__.smoke(), bench.py's cpu_baseline leg).
 * The product path never links or calls this file.
 *
 * Parity pin: this restatement is checked (tests/test_oracle_pin.py, oracle/make_golden.py)
 *   (a) against the UNMODIFIED reference imported from /root/reference in the build container
 *       (oracle/refload.py) on every fixture BAM/SAM of the reference's test-suite and on synthetic
 *       edge-case reads, and
 *   (b) on the GPU box, where /root/reference does not exist, against the golden vectors committed
 *       under tests/golden/ that (a)'s script produced from the reference itself.
 *
 * It walks the same flattened buffers the engine consumes (include/kindel_b200.h), one read after
 * the other in reference iteration order, with Python's list-index semantics (negative indices
 * wrap once; anything else out of range is IndexError) so that the reference's edge behaviour
 *  falls out of the same arithmetic rather than from special cases.
 *
 *   oracle_pileup  <- parse_records            reference kindel/kindel.py:40-81
 *   oracle_vote    <- consensus_sequence       reference kindel/kindel.py:402-424
 *                     consensus                reference kindel/kindel.py:369-381
 *   oracle_derive  <- parse_records post-pass  reference kindel/kindel.py:83-96, build_report :450
 */
#include <stdint.h>
#include <stddef.h>
#include <string.h>

#define NCOL 19
enum { W_A = 0, W_N = 4, C_DEL = 5, C_INS = 6, C_CLIP_STARTS = 7, C_CLIP_ENDS = 8, CSW_A = 9, CEW_A = 14 };
enum { ERR_INDEX = 10, ERR_KEY = 11 };

typedef struct {
    int64_t n_reads, n_ops, seq4_words;
    const int32_t* ref_start;
    const uint32_t* seq_off;
    const int32_t* l_seq;
    const uint32_t* cig_off;
    const uint32_t* cigar;
    const uint32_t* seq4;
    int32_t n_contigs, reads_sorted, max_simple_len, reserved0;
    const int64_t* contig_read_off;
    const int32_t* contig_len;
    const int64_t* contig_slot;
    int64_t n_complex;
    const uint32_t* complex_idx;
    const uint32_t* evt_off;
} batch_t;

typedef struct {
    int32_t status, reserved;
    int64_t read;
    int32_t nibble, op_index;
} diag_t;

/* BAM nibble -> column offset A,C,G,T,N = 0..4; -1 for anything the reference's five-key dicts
 * (kindel.py:29) do not hold, which is a KeyError at kindel.py:52/72/79. */
static const int8_t NIB2COL[16] = {-1, 0, 1, -1, 2, -1, -1, -1, 3, -1, -1, -1, -1, -1, -1, 4};

/* 8 bases per 32-bit word, first base in the most significant nibble (include/kindel_b200.h) */
static inline int nibble_at(const uint32_t* s, int64_t q) {
    return (int)((s[q >> 3] >> (28 - 4 * (q & 7))) & 0xF);
}

/* Python list indexing: list of length n, index i.  Returns the wrapped index or -1 (IndexError). */
static inline int64_t pyindex(int64_t i, int64_t n) {
    if (i < 0) i += n;
    return (i < 0 || i >= n) ? -1 : i;
}

#define FAIL(code, nib)                   \
    do {                                  \
        diag->status = (code);            \
        diag->read = r;                   \
        diag->nibble = (nib);             \
        diag->op_index = (int32_t)i;      \
        return (code);                    \
    } while (0)

/* kindel.py:40-81.  counts must be zeroed by the caller (or hold a partial sum to add to).
 * ins_events rows: (slot, read, q_off, len) in iteration order; n_events_out receives the count. */
int oracle_pileup(const batch_t* b, int32_t* counts, int64_t n_slots, int32_t* ins_events,
                  int64_t* n_events_out, diag_t* diag) {
    int64_t n_evt = 0;
    memset(diag, 0, sizeof(*diag));
    for (int32_t c = 0; c < b->n_contigs; ++c) {
        const int64_t L = b->contig_len[c];
        const int64_t base = b->contig_slot[c];
        int32_t* col[NCOL];
        for (int k = 0; k < NCOL; ++k) col[k] = counts + (int64_t)k * n_slots + base;
        for (int64_t r = b->contig_read_off[c]; r < b->contig_read_off[c + 1]; ++r) {
            const int64_t lseq = (int64_t)(b->l_seq[r] & 0x7fffffff);
            const uint32_t* seq = b->seq4 + (size_t)b->seq_off[r];
            const uint32_t c0 = b->cig_off[r], c1 = b->cig_off[r + 1];
            int64_t r_pos = b->ref_start[r]; /* kindel.py:42 (already POS-1) */
            int64_t q_pos = 0;               /* kindel.py:41 */
            for (uint32_t i = 0; i < c1 - c0; ++i) { /* kindel.py:47 */
                const uint32_t cg = b->cigar[c0 + i];
                const int64_t len = cg >> 4;
                const int op = cg & 0xF;
                if (op == 0 || op == 7 || op == 8) { /* M = X  kindel.py:49-54 */
                    for (int64_t k = 0; k < len; ++k) {
                        if (q_pos >= lseq) FAIL(ERR_INDEX, 0); /* record.seq[q_pos] */
                        int nib = nibble_at(seq, q_pos);
                        int64_t idx = pyindex(r_pos, L); /* weights[r_pos] */
                        if (idx < 0) FAIL(ERR_INDEX, 0);
                        int cc = NIB2COL[nib];
                        if (cc < 0) FAIL(ERR_KEY, nib);
                        col[W_A + cc][idx] += 1;
                        r_pos += 1;
                        q_pos += 1;
                    }
                } else if (op == 1) { /* I  kindel.py:55-58 */
                    int64_t idx = pyindex(r_pos, L + 1); /* insertions has ref_len+1 slots, :38 */
                    if (idx < 0) FAIL(ERR_INDEX, 0);
                    col[C_INS][idx] += 1;
                    if (ins_events) {
                        int32_t* e = ins_events + 4 * n_evt;
                        e[0] = (int32_t)(base + idx);
                        e[1] = (int32_t)r;
                        e[2] = (int32_t)q_pos;
                        e[3] = (int32_t)len;
                    }
                    n_evt += 1;
                    q_pos += len;
                } else if (op == 2) { /* D  kindel.py:59-62 */
                    for (int64_t k = 0; k < len; ++k) {
                        int64_t idx = pyindex(r_pos + k, L + 1);
                        if (idx < 0) FAIL(ERR_INDEX, 0);
                        col[C_DEL][idx] += 1;
                    }
                    r_pos += len;
                } else if (op == 4) { /* S  kindel.py:63-81 */
                    if (i == 0) {     /* left clip, kindel.py:64-73 */
                        int64_t idx = pyindex(r_pos, L + 1);
                        if (idx < 0) FAIL(ERR_INDEX, 0);
                        col[C_CLIP_ENDS][idx] += 1;
                        for (int64_t g = 0; g < len; ++g) {
                            if (g >= lseq) FAIL(ERR_INDEX, 0); /* record.seq[gap_i] */
                            int nib = nibble_at(seq, g);
                            int64_t rel = r_pos - len + g;
                            if (rel >= 0) {
                                if (rel >= L) FAIL(ERR_INDEX, 0);
                                int cc = NIB2COL[nib];
                                if (cc < 0) FAIL(ERR_KEY, nib);
                                col[CEW_A + cc][rel] += 1;
                            }
                        }
                        q_pos += len;
                    } else { /* right clip (any S that is not op #0), kindel.py:74-81 */
                        int64_t idx = pyindex(r_pos - 1, L + 1);
                        if (idx < 0) FAIL(ERR_INDEX, 0);
                        col[C_CLIP_STARTS][idx] += 1;
                        for (int64_t k = 0; k < len; ++k) {
                            if (q_pos >= lseq) FAIL(ERR_INDEX, 0); /* evaluated before the guard */
                            int nib = nibble_at(seq, q_pos);
                            if (r_pos < L) {
                                int64_t w = pyindex(r_pos, L);
                                if (w < 0) FAIL(ERR_INDEX, 0);
                                int cc = NIB2COL[nib];
                                if (cc < 0) FAIL(ERR_KEY, nib);
                                col[CSW_A + cc][w] += 1;
                                r_pos += 1;
                                q_pos += 1;
                            }
                        }
                    }
                }
                /* N, H, P and anything else: no-op, cursors do not move (falls through :49-63) */
            }
        }
    }
    if (n_events_out) *n_events_out = n_evt;
    return 0;
}

/* consensus() on a five-key base dict, kindel.py:369-381: first maximum in dict order A,T,G,C,N,
 * ("N", 0) when all zero, tie = another key holds the same non-zero count. */
static inline void base_consensus(const int32_t w[5] /*A,C,G,T,N*/, int* base, int* tie, int32_t* freq) {
    static const int order[5] = {0, 3, 2, 1, 4}; /* A,T,G,C,N as column indices */
    int64_t sum = (int64_t)w[0] + w[1] + w[2] + w[3] + w[4];
    if (sum == 0) {
        *base = 4;
        *tie = 0;
        *freq = 0;
        return;
    }
    int best = order[0];
    for (int k = 1; k < 5; ++k)
        if (w[order[k]] > w[best]) best = order[k];
    int t = 0;
    for (int k = 0; k < 5; ++k)
        if (k != best && w[k] == w[best]) t = 1;
    *base = best;
    *tie = (w[best] != 0) && t;
    *freq = w[best];
}

/* consensus_sequence per-position decision, kindel.py:402-424, for every slot s.
 * calls[s]: bits 0-2 emitted base (tie -> N), bits 4-5 change (0 none, 1 D, 2 N, 3 I). */
void oracle_vote(const int32_t* counts, int64_t n_slots, int64_t min_depth_ceil, uint8_t* calls) {
    for (int64_t s = 0; s < n_slots; ++s) {
        int32_t w[5];
        for (int k = 0; k < 5; ++k) w[k] = counts[(int64_t)k * n_slots + s];
        int64_t ins = counts[(int64_t)C_INS * n_slots + s];
        int64_t del = counts[(int64_t)C_DEL * n_slots + s];
        int64_t depth = (int64_t)w[0] + w[1] + w[2] + w[3]; /* ACGT only, :404 */
        int64_t depth_next = 0;                              /* :405-410 */
        if (s + 1 < n_slots)
            for (int k = 0; k < 4; ++k) depth_next += counts[(int64_t)k * n_slots + s + 1];
        uint8_t out;
        if (2 * del > depth) { /* del_freq > aligned_depth * 0.5, :413 */
            out = (1 << 4) | 4;
        } else if (depth < min_depth_ceil) { /* :415 */
            out = (2 << 4) | 4;
        } else {
            int64_t thr = depth < depth_next ? depth : depth_next; /* min(...)*0.5, :412 */
            int change = (2 * ins > thr) ? 3 : 0;                   /* :419 */
            int base, tie;
            int32_t freq;
            base_consensus(w, &base, &tie, &freq); /* :423 */
            out = (uint8_t)((change << 4) | (tie ? 4 : base));
        }
        calls[s] = out;
    }
}

/* out[5][n_slots]: consensus_depth, clip_start_depth, clip_end_depth, clip_depth (kindel.py:83-96),
 * acgt_depth (kindel.py:450). */
void oracle_derive(const int32_t* counts, int64_t n_slots, int32_t* out) {
    for (int64_t s = 0; s < n_slots; ++s) {
        int32_t w[5];
        for (int k = 0; k < 5; ++k) w[k] = counts[(int64_t)k * n_slots + s];
        int base, tie;
        int32_t freq;
        base_consensus(w, &base, &tie, &freq);
        /* aligned_depth - discordant_depth == count of the consensus key (0 when all zero) */
        out[0 * n_slots + s] = freq;
        int32_t csd = 0, ced = 0;
        for (int k = 0; k < 4; ++k) {
            csd += counts[(int64_t)(CSW_A + k) * n_slots + s];
            ced += counts[(int64_t)(CEW_A + k) * n_slots + s];
        }
        out[1 * n_slots + s] = csd;
        out[2 * n_slots + s] = ced;
        out[3 * n_slots + s] = csd + ced;
        out[4 * n_slots + s] = w[0] + w[1] + w[2] + w[3];
    }
}

int oracle_abi_version(void) { return 1; }
