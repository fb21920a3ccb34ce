/* kindel_b200.h -- C ABI of the B200-native pileup/consensus engine (libkindel_b200.so).
 *
 * The reference (bede/kindel v1.2.1) is pure Python and has no FFI of its own; its seam for this
 * path is three Python callables:
 *     parse_records(ref_id, ref_len, records)          reference kindel/kindel.py:21-128
 *     consensus(weight)                                reference kindel/kindel.py:369-381
 *     consensus_sequence(weights, insertions, ...)     reference kindel/kindel.py:384-430
 * The entry points below are what a ctypes/cffi binding inside the reference's
 * `kindel/kindel.py` would call in place of those loops (the stub is shown in INTEGRATION.md;
 * the shipped host side that does exactly that is kindel_b200/kindel.py).
 *
 * Conventions
 *   - plain C: pointers + sizes, no torch / C++ types.  `stream` is a cudaStream_t passed as void*.
 *   - every function returns a kdl_status (0 = ok).  Nothing here allocates device memory except
 *     the kdl_ctx_* host-buffer path, which owns a growable workspace inside its context.
 *   - "device" entry points take DEVICE pointers and only enqueue work on `stream` (asynchronous;
 *     the caller synchronises).  "host" entry points (kdl_ctx_*) take HOST pointers, do the
 *     host->device copies, the kernels and the device->host copies themselves and return when the
 *     results are in the caller's buffers.
 *   - re-entrant per stream, no global mutable state.
 *
 * Data layout (all little-endian)
 *   reads, in the reference's iteration order (grouped by contig in first-seen order, file order
 *   inside a contig; records failing `mapped and len(seq) > 1`, kindel.py:43-46, already dropped):
 *     ref_start[n]   int32   0-based reference cursor at walk start (= SAM POS - 1; -1 if POS == 0)
 *     seq_off[n]     uint32  offset of the read's block in `seq4`, in 4-byte words (non-decreasing)
 *     l_seq[n]       int32   bits 0..29: SEQ length.  bit 31 (KDL_COMPLEX) clear = "simple" read: exactly
 *                            one M/=/X op whose length equals the SEQ length, fully inside the contig
 *                            (ref_start >= 0, ref_start + len <= L), every base one of A,C,G,T,N (nibbles
 *                            1,2,4,8,15), at most KDL_FAST_MAXLEN bases.  Its CIGAR is implied and never
 *                            travels.  bit 31 set = complex read: its CIGAR follows its bases in `seq4`
 *                            (below).  bit 30 (KDL_HARD) set as well = the general, atomic kernel K1g must
 *                            walk it (it may wrap a Python negative index, raise like the reference, or is
 *                            too long for a tile: SURVEY.md A-7..A-10); bit 30 clear = "tile-eligible": every
 *                            slot it touches lies in [1, L - 1] of its contig, its query span fits its SEQ,
 *                            it has at most KDL_TILE_MAXOPS ops and reaches at most KDL_TILE_MAXREACH slots
 *                            to either side of its start, and all its bases are A,C,G,T,N -- K1 and K1e
 *                            walk it without bounds or error logic.  A tile-eligible read keeps its SEQ
 *                            length (<= KDL_FAST_MAXLEN) in bits 0..15 and its number of M/=/X ops in bits
 *                            16..22, so a tile can be sized before its bases are staged; a hard read keeps
 *                            its length in bits 0..29.
 *                            The flatten step classifies (kindel_b200/bamio.py).
 *     seq4[n_words]  uint32  one block per read, in read order.  Bases: BAM nibble codes
 *                            "=ACMGRSVTWYHKDBN", 8 per 32-bit word, FIRST base in the MOST significant
 *                            nibble (base k of a read sits at bits [28-4*(k%8), 32-4*(k%8)) of word k/8),
 *                            unused trailing nibbles of the last word zero.  A COMPLEX read's block goes on
 *                            with  [n_ops] [evt_off] [n_ops x CIGAR word]  -- BAM encoding len << 4 | op,
 *                            op index into "MIDNSHP=X"; evt_off = number of I ops of all reads before this
 *                            one (row of its first insertion event).  A tile's reads are one contiguous
 *                            byte range of this array, CIGARs included: one bulk copy stages them.
 *   contigs:
 *     contig_read_off[n_contigs+1] int64  reads of contig c are [off[c], off[c+1])
 *     contig_len[n_contigs]        int32  reference length L_c
 *     contig_slot[n_contigs]       int64  first table slot of contig c; it owns L_c + 1 slots
 *   count table: counts[KDL_NCOL][n_slots] int32, column-major (one contiguous array per column):
 *     0-4  weights A,C,G,T,N          (kindel.py:29,49-54)
 *     5    deletions                  (kindel.py:39,59-62)
 *     6    insertion events (total)   (kindel.py:38,55-58; the string-keyed dict is rebuilt from
 *                                      the event list below)
 *     7    clip_starts   8 clip_ends  (kindel.py:36-37,66,75)
 *     9-13 clip_start_weights A,C,G,T,N   14-18 clip_end_weights A,C,G,T,N (kindel.py:30-35,67-81)
 *   insertion events: ins_events[n_events][4] int32 = (slot, read, q_off, len), event k of the j-th
 *     listed read stored at row evt_off[j] + k  (deterministic, reference iteration order).
 */
#ifndef KINDEL_B200_H
#define KINDEL_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KDL_ABI_VERSION 2
#define KDL_NCOL 19
#define KDL_NVOTE_COL 7 /* columns 0..6 are all the vote needs */
#define KDL_COMPLEX 0x80000000u
#define KDL_HARD 0x40000000u
#define KDL_LEN_MASK 0x0000ffffu   /* SEQ length of a complex read lives in bits 0..15 (longer reads are KDL_HARD
                                      and keep their length in bits 0..29 with the op count field zero) */
#define KDL_NM_SHIFT 16            /* bits 16..22: M/=/X op count of a complex read */
#define KDL_NM_MASK 0x7fu
#define KDL_TILE 512          /* slots per tile of the owner-computes pileup; n_slots % KDL_TILE == 0 */
#define KDL_FAST_MAXLEN 8192  /* longest read the flatten step may mark simple */
#define KDL_TILE_MAXOPS 64    /* most CIGAR ops of a tile-eligible complex read */
#define KDL_TILE_MAXREACH 1024 /* furthest slot, relative to its start, a tile-eligible complex read touches */

enum kdl_col {
    KDL_W_A = 0, KDL_W_C, KDL_W_G, KDL_W_T, KDL_W_N,
    KDL_DEL = 5, KDL_INS = 6, KDL_CLIP_STARTS = 7, KDL_CLIP_ENDS = 8,
    KDL_CSW_A = 9, KDL_CEW_A = 14
};

typedef enum kdl_status {
    KDL_OK = 0,
    KDL_ERR_INVALID_ARG = 1,
    KDL_ERR_CUDA = 2,
    KDL_ERR_NO_DEVICE = 3,
    /* data errors, mirroring the exceptions the reference raises: */
    KDL_ERR_INDEX = 10, /* IndexError: walk ran off the contig / off SEQ (kindel.py:51,52,61) */
    KDL_ERR_KEY = 11    /* KeyError: base outside A,C,G,T,N used in an M or S op (kindel.py:52,72,79) */
} kdl_status;

/* Flattened read batch.  Pointers are device pointers for the device entry points and host
 * pointers for the kdl_ctx_* entry points. */
typedef struct kdl_batch {
    int64_t n_reads;
    int64_t seq4_words;  /* 32-bit words in seq4 */
    const int32_t* ref_start;
    const uint32_t* seq_off;
    const int32_t* l_seq;
    const uint32_t* seq4;
    int32_t n_contigs;
    int32_t reads_sorted;   /* 1 = contig_slot + ref_start and seq_off are non-decreasing over all reads */
    int32_t max_simple_len; /* longest simple read (bases); 0 if there is none */
    int32_t reach_right;    /* max over simple and tile-eligible reads of (last touched slot - start + 1) */
    int32_t reach_left;     /* max over tile-eligible reads of (start - first touched slot) */
    int32_t reserved0;
    const int64_t* contig_read_off;
    const int32_t* contig_len;
    const int64_t* contig_slot;
    /* complex reads: how many there are (tile-eligible + hard; columns 5..18 are only ever written when
     * this is non-zero) and the ascending indices of the KDL_HARD ones, which K1g walks */
    int64_t n_complex;
    int64_t n_hard;
    const uint32_t* complex_idx; /* [n_complex] ascending indices of ALL complex reads; may be NULL when n_complex == 0 */
    const uint32_t* hard_idx;    /* [n_hard] ascending indices of the KDL_HARD ones; may be NULL when n_hard == 0 */
    /* scratch for the tile index kdl_pileup builds (K0): uint32[8 * n_slots / KDL_TILE], device
     * memory owned by the caller.  NULL, or reads_sorted == 0, selects the order-independent
     * atomic kernels instead of the tile-owner kernel. */
    uint32_t* tile_index;
} kdl_batch;

/* Error report written by the pileup (device memory, 4 x int32, zero it before the call):
 *   [0] != 0 : some read hit a data error; call kdl_diagnose for the exact first one. */
typedef struct kdl_diag {
    int32_t status;    /* KDL_OK, KDL_ERR_INDEX or KDL_ERR_KEY */
    int32_t reserved;
    int64_t read;      /* index of the first offending read in iteration order */
    int32_t nibble;    /* for KDL_ERR_KEY: BAM nibble code of the offending base */
    int32_t op_index;  /* CIGAR op at which the walk failed */
} kdl_diag;

int kdl_abi_version(void);
const char* kdl_status_string(int status);
/* number of CUDA kernels this library has launched in this process (bench.py: gpu_launches) */
int64_t kdl_launch_count(void);

/* K1 -- pileup.  Replaces the loop at kindel/kindel.py:40-81.
 * Adds every read's contribution to `counts` (caller zeroes it first, so several batches -- or
 * several read shards -- can accumulate into one table) and writes the insertion event rows.
 * err_flag: device int32[4], caller-zeroed; [0] becomes non-zero if any read raised.
 * Coordinate-sorted batches (reads_sorted, tile_index scratch given) take K0 (tile index) + the tile-owner
 * kernel K1 (simple reads and the M/=/X bases of tile-eligible complex reads: no atomics), K1e (the sparse
 * insertion / deletion / clip updates of those complex reads, once per read) and K1g (KDL_HARD reads, atomics);
 * anything else the order-independent atomic kernels K1s + K1g. */
int kdl_pileup(const kdl_batch* batch, int32_t* counts, int64_t n_slots, int32_t* ins_events,
               int32_t* err_flag, void* stream);

/* Same as kdl_pileup, restricted to the slot range [slot_lo, slot_hi) (multiples of KDL_TILE) that
 * contains everything the batch can touch (a shard's footprint), with control over zeroing:
 *   KDL_PILEUP_FRESH_WEIGHTS  columns 0..4 of the range hold stale data: the tile-owner kernel
 *                             OVERWRITES them (plain stores, no prior memset, no read-modify-write)
 *   KDL_PILEUP_ZERO_REST      columns 5..18 of the range are zeroed first (needed only when an
 *                             earlier pileup with complex reads dirtied them; with FRESH_WEIGHTS the tile-owner
 *                             kernel does it window by window in its flush, no separate pass) */
#define KDL_PILEUP_FRESH_WEIGHTS 1
#define KDL_PILEUP_ZERO_REST 2
int kdl_pileup_range(const kdl_batch* batch, int32_t* counts, int64_t n_slots, int64_t slot_lo,
                     int64_t slot_hi, int32_t flags, int32_t* ins_events, int32_t* err_flag, void* stream);

/* Exact first error in reference iteration order (only needed when err_flag[0] != 0).
 * `diag_dev`: device kdl_diag, written asynchronously. */
int kdl_diagnose(const kdl_batch* batch, kdl_diag* diag_dev, void* stream);

/* K2 -- per-position vote.  Replaces kindel/kindel.py:402-424 + 369-381 for every slot.
 * calls[s] : bits 0-2 = emitted base (0..4 = A,C,G,T,N; a tie emits N), bits 4-5 = change code
 *            (0 none, 1 'D' -> nothing emitted, 2 'N' -> 'N' emitted, 3 'I' -> insertion string
 *            precedes the base).  min_depth_ceil = ceil(min_depth). */
int kdl_vote(const int32_t* counts, int64_t n_slots, int64_t min_depth_ceil, uint8_t* calls,
             void* stream);

/* Derived per-position columns of the `alignment` tuple (kindel/kindel.py:83-96) + the ACGT depth
 * used by build_report (kindel.py:450).  out[5][n_slots] int32: consensus_depth, clip_start_depth,
 * clip_end_depth, clip_depth, acgt_depth. */
int kdl_derive(const int32_t* counts, int64_t n_slots, int32_t* out, void* stream);

/* K4 -- the per-position predicates of the --realign path (reference kindel/kindel.py:182-185,202,243-246,256) for
 * slots [slot_lo, slot_hi):  flags[s] bit 0 = clip-dominant for right-clipped reads: clip_start_depth /
 * (sum(weights) + deletions + 1) > 0.5; bit 1 = their clip consensus extends through s: clip_start_depth >
 * (sum(weights) + deletions) * clip_decay_threshold; bits 2, 3 = the same for left-clipped reads (clip_end_*).
 * bases[s]: low nibble = consensus()[0] of clip_start_weights[s] (0..4 = A,C,G,T,N), high nibble = of
 * clip_end_weights[s].  Masking of the contig ends, pairing and the LCS merge stay with the caller. */
int kdl_cdr_flags(const int32_t* counts, int64_t n_slots, int64_t slot_lo, int64_t slot_hi,
                  double clip_decay_threshold, uint8_t* flags, uint8_t* bases, void* stream);

/* K5 -- the consensus text of every contig from the call bytes (reference kindel/kindel.py:413-424): nothing for a
 * 'D' call, the base letter (N for an 'N' call or a tie) otherwise, preceded by the insertion string for an 'I'
 * call.  The strings of the 'I' slots come from the caller (ins_slot ascending, bytes ins_bytes[ins_off[k] ..
 * ins_off[k+1]) as they are to be printed: the modal inserted string in lower case, or "N" for a tie).
 * offsets: device uint32[n_slots + 1], out: offsets[s] = where slot s's text starts in `out`, offsets[n_slots] = total;
 * contig c's sequence is out[offsets[contig_slot[c]] .. offsets[contig_slot[c] + contig_len[c]]).
 * block_sums: device scratch of kdl_assemble_scratch_words(n_slots) uint32.  out: device bytes, at least
 * (number of positions + total insertion bytes).  All pointers are device pointers. */
int64_t kdl_assemble_scratch_words(int64_t n_slots);
int kdl_assemble(const uint8_t* calls, int64_t n_slots, const int64_t* contig_slot, const int32_t* contig_len,
                 int32_t n_contigs, const int64_t* ins_slot, const uint32_t* ins_off, const uint8_t* ins_bytes,
                 int64_t n_ins, uint32_t* block_sums, uint32_t* offsets, uint8_t* out, void* stream);

/* Fused cross-GPU count reduction + vote: sums the 7 vote columns of `n_peers`
 * tables that live on this and on peer GPUs (peer pointers mapped with CUDA IPC / P2P), votes on
 * slots [slot_lo, slot_hi) and writes calls for that range; optionally stores the reduced
 * columns into reduced[7][n_slots] (may be NULL). */
int kdl_vote_peers(const int32_t* const* peer_counts, int32_t n_peers, int64_t n_slots,
                   int64_t slot_lo, int64_t slot_hi, int64_t min_depth_ceil, uint8_t* calls,
                   int32_t* reduced, void* stream);
/* Same, with each table's footprint: peer p only holds non-zero counts in slots
 * [foot_lo[p], foot_hi[p]) (host int64 arrays, multiples of 4), so a read-sharded, coordinate-sorted
 * job pulls only the halo of its neighbours over NVLink instead of every table. */
int kdl_vote_peers_sparse(const int32_t* const* peer_counts, const int64_t* foot_lo, const int64_t* foot_hi,
                          int32_t n_peers, int64_t n_slots, int64_t slot_lo, int64_t slot_hi,
                          int64_t min_depth_ceil, uint8_t* calls, int32_t* reduced, void* stream);

/* ---- fully fused exchange (no NCCL on the data path) ------------------------------------------
 * Every rank owns one IPC block: [count table 19 x n_slots int32][calls n_slots bytes][flags].
 * Per step (epoch e = 1, 2, ...), on every rank, in stream order:
 *   kdl_pileup(...)                       its shard into its own table
 *   kdl_exchange_signal(x, e)             optional early "my table is complete" -> ready[p][rank] = e in
 *                                         every peer p (kdl_exchange_vote publishes it too, first thing)
 *   kdl_exchange_vote(x, ..., e)          K2x: waits for ready[rank][*] >= e, sums the 7 vote columns
 *                                         of its slot slice [slice_lo[rank], slice_hi[rank]) over the
 *                                         (footprint-clipped) peer tables through NVLink, votes,
 *                                         stores the call bytes of the slice locally; the last CTA
 *                                         out publishes done[p][rank] = e to every peer p
 *   kdl_exchange_wait(x, e)               K2g: per peer p, waits for done[rank][p] >= e and pulls p's
 *                                         call slice over NVLink into the local call buffer; when it
 *                                         ends every slice is here and nobody still reads this
 *                                         rank's table
 * All pointers of rank p (tables[p], calls[p], ready[p], done[p]) are this process's mappings of
 * rank p's block (own block: the local pointer). */
typedef struct kdl_exchange {
    int32_t n_ranks, rank;
    const int32_t* tables[16];
    uint8_t* calls[16];
    int32_t* ready[16]; /* int32[16] per rank */
    int32_t* done[16];  /* int32[16] per rank */
    int64_t foot_lo[16], foot_hi[16];   /* table p is zero outside [foot_lo[p], foot_hi[p]) */
    int64_t slice_lo[16], slice_hi[16]; /* slots rank p votes on (multiples of 4, a partition) */
    int32_t* counter;   /* local device int32, zero-initialised */
} kdl_exchange;

int kdl_exchange_signal(const kdl_exchange* x, int32_t epoch, void* stream);
int kdl_exchange_vote(const kdl_exchange* x, int64_t n_slots, int64_t min_depth_ceil, int32_t epoch,
                      void* stream);
int kdl_exchange_wait(const kdl_exchange* x, int32_t epoch, void* stream);

/* Count tables that peer GPUs (other processes of the same node) can map: plain cudaMalloc
 * memory exported / opened with CUDA IPC.  kdl_table_alloc zero-fills. */
int kdl_table_alloc(int64_t bytes, void** dev_ptr);
int kdl_table_free(void* dev_ptr);
int kdl_ipc_export(void* dev_ptr, uint8_t handle[64]);
int kdl_ipc_open(const uint8_t handle[64], void** dev_ptr);
int kdl_ipc_close(void* dev_ptr);

/* ---- host-buffer path (what a cgo/JNI/ctypes caller without its own CUDA runtime uses) ---- */
typedef struct kdl_ctx kdl_ctx;

int kdl_ctx_create(int device, kdl_ctx** out);
void kdl_ctx_destroy(kdl_ctx* ctx);

/* batch holds HOST pointers.  batch->seq_off may be NULL: the packed bases are then taken to be DENSE (read i
 * starts at the sum of ceil(l_seq[j] / 8) words over j < i, as every flattener here lays them out) and the
 * offsets are computed on the device instead of being copied (4 bytes per read less over PCIe).
 * Outputs (host, caller-allocated, any may be NULL to skip):
 *   calls_out[n_slots] uint8, counts_out[KDL_NCOL][n_slots] int32,
 *   ins_events_out[n_events][4] int32.  diag_out is always filled.
 * Returns KDL_OK, or KDL_ERR_INDEX / KDL_ERR_KEY with diag_out describing the first offender. */
int kdl_ctx_consensus(kdl_ctx* ctx, const kdl_batch* batch, int64_t n_slots, int64_t n_events,
                      int64_t min_depth_ceil, uint8_t* calls_out, int32_t* counts_out,
                      int32_t* ins_events_out, kdl_diag* diag_out);

/* device time (ms) of the last kdl_ctx_consensus call, H2D / kernels / D2H, from CUDA events */
int kdl_ctx_last_timing(kdl_ctx* ctx, float* h2d_ms, float* kernel_ms, float* d2h_ms);

/* ---- host-side BAM decode (no GPU involved; kindel_b200/csrc/bam_host.cpp) ----
 * Replaces the simplesam -> `samtools view` text round trip of kindel/kindel.py:136-145 for .bam input: BGZF blocks
 * inflated by zlib in C++ threads, records filtered (kindel.py:43-46), classified and written straight into the
 * layout above -- into caller-owned buffers, which may be pinned memory.
 *   kdl_bam_open     read + inflate + parse the header (text, reference dictionary).  Takes BGZF / gzip / plain BAM
 *                    and SAM text (plain or gzip): text lines are turned into BAM records in threads by a strict
 *                    parser that gives up (error) on anything unusual -- the caller then uses its own text reader
 *   kdl_bam_prepare  ref_len[n_ref] = contig lengths to classify against (the @SQ LN values the reference uses;
 *                    NULL = the binary dictionary's).  info[16] out: 0 records, 1 kept reads, 2 contigs seen,
 *                    3 CIGAR ops of kept reads, 4 words of seq4, 5 complex reads, 6 hard reads, 7 aligned bases,
 *                    9 reach_right, 10 reach_left, 11 longest simple read
 *   kdl_bam_contigs  order[n_seen] = reference ids in first-seen order (kindel.py:143-151), read_off[n_seen + 1]
 *   kdl_bam_fill     ref_start / seq_off / l_seq / seq_len [kept], cig_off [kept + 1], cigar [ops] (may be NULL),
 *                    seq4 [words], complex_idx [complex], hard_idx [hard] (may be NULL); contig_slot[n_seen] = the slot
 *                    layout; fills info[8] = insertion events, info[12] = reads_sorted */
typedef struct kdl_bam kdl_bam;
int kdl_bam_open(const char* path, int threads, kdl_bam** out);
void kdl_bam_close(kdl_bam* h);
const char* kdl_bam_header_text(const kdl_bam* h, int64_t* len);
int32_t kdl_bam_n_ref(const kdl_bam* h);
const char* kdl_bam_ref_name(const kdl_bam* h, int32_t ref_id);
int32_t kdl_bam_ref_len(const kdl_bam* h, int32_t ref_id);
int kdl_bam_prepare(kdl_bam* h, const int32_t* ref_len, int threads, int64_t* info);
int kdl_bam_contigs(const kdl_bam* h, int32_t* order, int64_t* read_off);
int kdl_bam_fill(kdl_bam* h, int threads, const int64_t* contig_slot, int32_t* ref_start, uint32_t* seq_off,
                 int32_t* l_seq, int32_t* seq_len, uint32_t* cig_off, uint32_t* cigar, uint32_t* seq4,
                 uint32_t* complex_idx, uint32_t* hard_idx, int64_t* info);

#ifdef __cplusplus
}
#endif
#endif /* KINDEL_B200_H */
