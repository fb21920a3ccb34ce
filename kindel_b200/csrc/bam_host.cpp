// bam_host.cpp -- host-side gather of BAM alignment records into the engine's flattened layout.
//
// This replaces, for .bam input, the record materialisation the reference does through
// simplesam -> `samtools view` text (reference kindel/kindel.py:136-145): BAM's on-disk encodings
// (CIGAR as len<<4|op uint32, SEQ as 4-bit nibbles) are already the device layout described in
// include/kindel_b200.h, so a record is flattened with two memcpy's and no text round trip.
// The inflated BAM byte stream comes from Python (zlib in a thread pool, kindel_b200/bamio.py).
//
// Two passes, because the output is grouped by contig in first-seen order (kindel.py:143-151):
//   kdl_bam_count : walk the block_size chain once; per contig: records seen (any flag), records
//                   kept (mapped and l_seq > 1, kindel.py:43-46), CIGAR ops and packed-SEQ words
//                   of kept records, and the first-seen rank of the contig.
//   kdl_bam_fill  : walk again and append every kept record at its contig's running cursor.
#include <cstdint>
#include <cstring>

#include "../../include/kindel_b200.h"

namespace {

struct RecView {
    int32_t ref_id, pos, l_seq;
    uint32_t n_cigar, flag, l_read_name;
    const uint8_t* cigar;
    const uint8_t* seq;
};

inline int32_t rd_i32(const uint8_t* p) {
    int32_t v;
    std::memcpy(&v, p, 4);
    return v;
}
inline uint16_t rd_u16(const uint8_t* p) {
    uint16_t v;
    std::memcpy(&v, p, 2);
    return v;
}

// Returns bytes consumed (4 + block_size) or 0 when the record is truncated / malformed.
inline int64_t parse_record(const uint8_t* p, int64_t avail, RecView* r) {
    if (avail < 36) return 0;
    const int32_t block_size = rd_i32(p);
    if (block_size < 32 || (int64_t)block_size + 4 > avail) return 0;
    const uint8_t* q = p + 4;
    r->ref_id = rd_i32(q);
    r->pos = rd_i32(q + 4);
    r->l_read_name = q[8];
    r->n_cigar = rd_u16(q + 12);
    r->flag = rd_u16(q + 14);
    r->l_seq = rd_i32(q + 16);
    const int64_t need = 32 + (int64_t)r->l_read_name + 4ll * r->n_cigar + ((int64_t)r->l_seq + 1) / 2;
    if (r->l_seq < 0 || need > block_size) return 0;
    r->cigar = q + 32 + r->l_read_name;
    r->seq = r->cigar + 4ll * r->n_cigar;
    return 4 + (int64_t)block_size;
}

inline bool kept(const RecView& r) { return !(r.flag & 0x4u) && r.l_seq > 1; }

}  // namespace

extern "C" {

// per_contig[n_ref][4] int64: seen, kept, ops, seq_words.  first_seen[n_ref] int32: rank or -1.
// totals[4]: records, kept, contigs seen, bytes consumed.
int kdl_bam_count(const uint8_t* bam, int64_t n_bytes, int64_t first_record, int32_t n_ref,
                  int64_t* per_contig, int32_t* first_seen, int64_t* totals) {
    if (!bam || !per_contig || !first_seen || !totals || first_record < 0 || first_record > n_bytes)
        return KDL_ERR_INVALID_ARG;
    std::memset(per_contig, 0, sizeof(int64_t) * 4 * (size_t)n_ref);
    for (int32_t c = 0; c < n_ref; ++c) first_seen[c] = -1;
    int64_t off = first_record, n_rec = 0, n_kept = 0;
    int32_t rank = 0;
    RecView r;
    while (off < n_bytes) {
        const int64_t used = parse_record(bam + off, n_bytes - off, &r);
        if (!used) return KDL_ERR_INVALID_ARG;
        off += used;
        ++n_rec;
        if (r.ref_id < 0) continue;  // rname '*' is dropped wholesale (kindel.py:147-148)
        if (r.ref_id >= n_ref) return KDL_ERR_INVALID_ARG;
        int64_t* pc = per_contig + 4ll * r.ref_id;
        if (first_seen[r.ref_id] < 0) first_seen[r.ref_id] = rank++;
        pc[0] += 1;
        if (kept(r)) {
            pc[1] += 1;
            pc[2] += r.n_cigar;
            pc[3] += ((int64_t)r.l_seq + 7) / 8;  // nibbles -> 4-byte words
            ++n_kept;
        }
    }
    totals[0] = n_rec;
    totals[1] = n_kept;
    totals[2] = rank;
    totals[3] = off;
    return KDL_OK;
}

// cursors[n_ref][3] int64: next read index, next op index, next seq word for each contig
// (pre-set by the caller from the prefix sums of kdl_bam_count's output; advanced in place).
// Outputs are sized for all kept records: ref_start/seq_off/l_seq/cig_start [n_kept],
// cigar [total ops], seq4 [total words] uint32.
int kdl_bam_fill(const uint8_t* bam, int64_t n_bytes, int64_t first_record, int32_t n_ref,
                 int64_t* cursors, int32_t* ref_start, uint32_t* seq_off, int32_t* l_seq,
                 uint32_t* cig_start, uint32_t* cigar, uint32_t* seq4, uint8_t* exotic) {
    if (!bam || !cursors) return KDL_ERR_INVALID_ARG;
    int64_t off = first_record;
    RecView r;
    while (off < n_bytes) {
        const int64_t used = parse_record(bam + off, n_bytes - off, &r);
        if (!used) return KDL_ERR_INVALID_ARG;
        off += used;
        if (r.ref_id < 0 || r.ref_id >= n_ref || !kept(r)) continue;
        int64_t* cur = cursors + 3ll * r.ref_id;
        const int64_t i = cur[0]++;
        const int64_t o = cur[1];
        const int64_t w = cur[2];
        cur[1] += r.n_cigar;
        cur[2] += ((int64_t)r.l_seq + 7) / 8;
        ref_start[i] = r.pos;  // BAM pos is 0-based == SAM POS - 1 (kindel.py:42)
        seq_off[i] = (uint32_t)w;
        l_seq[i] = r.l_seq;
        cig_start[i] = (uint32_t)o;
        std::memcpy(cigar + o, r.cigar, 4ull * r.n_cigar);
        // BAM packs two bases per byte, first base in the high nibble; the engine wants 8 bases
        // per 32-bit word with the first base in the most significant nibble: a byte-swapped copy.
        const int64_t n_bytes_seq = ((int64_t)r.l_seq + 1) / 2;
        const int64_t n_words_seq = ((int64_t)r.l_seq + 7) / 8;
        uint32_t bad = 0;
        for (int64_t k = 0; k < n_words_seq; ++k) {
            uint8_t b[4] = {0, 0, 0, 0};
            const int64_t left = n_bytes_seq - 4 * k;
            std::memcpy(b, r.seq + 4 * k, (size_t)(left < 4 ? left : 4));
            uint32_t v = ((uint32_t)b[0] << 24) | ((uint32_t)b[1] << 16) | ((uint32_t)b[2] << 8) | b[3];
            uint32_t chk = v;  // nibbles that must each be one of 1,2,4,8 (A,C,G,T) or 15 (N)
            if (k == n_words_seq - 1 && (r.l_seq & 7)) {
                const uint32_t pad = 0xFFFFFFFFu >> (4 * (r.l_seq & 7));
                v &= ~pad;
                chk = v | (pad & 0x11111111u);  // padding counts as fine
            }
            seq4[w + k] = v;
            const uint32_t h = chk | (chk >> 1), pair = chk & (chk >> 1);
            const uint32_t two_plus = (pair | (pair >> 2) | (h & (h >> 2))) & 0x11111111u;
            const uint32_t all4 = pair & (pair >> 2) & 0x11111111u;
            const uint32_t zero = ~(h | (h >> 2)) & 0x11111111u;
            bad |= (two_plus & ~all4) | zero;
        }
        if (exotic) exotic[i] = bad ? 1 : 0;
    }
    return KDL_OK;
}

}  // extern "C"
