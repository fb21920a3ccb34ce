// host_ctx.inl -- kdl_ctx_*: the host-buffer entry points (included at the end of api.cu).
//
// This is the call a host program without its own CUDA runtime makes (the cgo / JNI / ctypes
// binding of INTEGRATION.md): flattened reads in HOST memory in, calls (and optionally the count
// table and insertion events) in HOST memory out.  The context owns one stream and a grow-only
// device workspace, so repeated calls do not allocate.  Timing of the last call (H2D, kernels,
// D2H) is taken with CUDA events on the context's stream.

struct kdl_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    float ms[3] = {0.f, 0.f, 0.f};
    int64_t table_slots = -1;  // layout of the count table left by the previous call (-1: none)
    bool table_dirty_rest = false;
    struct Buf {
        void* p = nullptr;
        size_t cap = 0;
    };
    enum { B_REF_START, B_SEQ_OFF, B_L_SEQ, B_SEQ4, B_CREAD_OFF, B_CLEN, B_CSLOT, B_CX_IDX, B_HARD_IDX, B_COUNTS, B_EVENTS,
           B_CALLS, B_FLAG, B_DIAG, B_TILE_IDX, B_N };
    Buf buf[B_N];

    int ensure(int which, size_t bytes) {
        Buf& b = buf[which];
        if (bytes == 0) bytes = 16;
        if (b.cap >= bytes) return KDL_OK;
        if (b.p) cudaFree(b.p);
        b.p = nullptr;
        b.cap = 0;
        if (which == B_COUNTS) table_slots = -1;  // a new buffer holds garbage
        size_t want = bytes + bytes / 8;  // slack so slowly growing batches do not reallocate
        if (cudaMalloc(&b.p, want) != cudaSuccess) {
            if (cudaMalloc(&b.p, bytes) != cudaSuccess) return KDL_ERR_CUDA;
            want = bytes;
        }
        b.cap = want;
        return KDL_OK;
    }
};

extern "C" {

int kdl_ctx_create(int device, kdl_ctx** out) {
    if (!out) return KDL_ERR_INVALID_ARG;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) return KDL_ERR_NO_DEVICE;
    if (device < 0 || device >= n) return KDL_ERR_INVALID_ARG;
    if (cudaSetDevice(device) != cudaSuccess) return KDL_ERR_CUDA;
    kdl_ctx* c = new (std::nothrow) kdl_ctx();
    if (!c) return KDL_ERR_CUDA;
    c->device = device;
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete c;
        return KDL_ERR_CUDA;
    }
    for (auto& e : c->ev)
        if (cudaEventCreate(&e) != cudaSuccess) {
            kdl_ctx_destroy(c);
            return KDL_ERR_CUDA;
        }
    *out = c;
    return KDL_OK;
}

void kdl_ctx_destroy(kdl_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    for (auto& b : c->buf)
        if (b.p) cudaFree(b.p);
    for (auto& e : c->ev)
        if (e) cudaEventDestroy(e);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

// the body of kdl_ctx_consensus after the first enqueue: every exit goes back through the caller, which
// synchronises the stream (the caller's host buffers may be the source / target of copies still in flight)
static int ctx_consensus_enqueued(kdl_ctx* c, const kdl_batch* hb, int64_t n_slots, int64_t n_events,
                                  int64_t min_depth_ceil, uint8_t* calls_out, int32_t* counts_out,
                                  int32_t* ins_events_out, kdl_diag* diag_out, int32_t* flag_host) {
    const size_t n = (size_t)hb->n_reads, nc = (size_t)hb->n_contigs, nh = (size_t)hb->n_hard, nx = (size_t)hb->n_complex;
    struct Copy { int which; const void* src; size_t bytes; };
    const Copy copies[] = {
        {kdl_ctx::B_REF_START, hb->ref_start, n * 4},
        {kdl_ctx::B_SEQ_OFF, hb->seq_off, n * 4},
        {kdl_ctx::B_L_SEQ, hb->l_seq, n * 4},
        {kdl_ctx::B_SEQ4, hb->seq4, (size_t)hb->seq4_words * 4},
        {kdl_ctx::B_CREAD_OFF, hb->contig_read_off, (nc + 1) * 8},
        {kdl_ctx::B_CLEN, hb->contig_len, nc * 4},
        {kdl_ctx::B_CSLOT, hb->contig_slot, nc * 8},
        {kdl_ctx::B_CX_IDX, hb->complex_idx, nx * 4},
        {kdl_ctx::B_HARD_IDX, hb->hard_idx, nh * 4},
    };
    cudaStream_t st = c->stream;
    int rc;
    if (cudaEventRecord(c->ev[0], st) != cudaSuccess) return KDL_ERR_CUDA;
    for (const Copy& cp : copies)
        if (cp.bytes && cp.src &&
            cudaMemcpyAsync(c->buf[cp.which].p, cp.src, cp.bytes, cudaMemcpyHostToDevice, st) != cudaSuccess)
            return KDL_ERR_CUDA;
    if (cudaEventRecord(c->ev[1], st) != cudaSuccess) return KDL_ERR_CUDA;

    kdl_batch db = *hb;
    db.ref_start = (const int32_t*)c->buf[kdl_ctx::B_REF_START].p;
    db.seq_off = (const uint32_t*)c->buf[kdl_ctx::B_SEQ_OFF].p;
    db.l_seq = (const int32_t*)c->buf[kdl_ctx::B_L_SEQ].p;
    db.seq4 = (const uint32_t*)c->buf[kdl_ctx::B_SEQ4].p;
    db.contig_read_off = (const int64_t*)c->buf[kdl_ctx::B_CREAD_OFF].p;
    db.contig_len = (const int32_t*)c->buf[kdl_ctx::B_CLEN].p;
    db.contig_slot = (const int64_t*)c->buf[kdl_ctx::B_CSLOT].p;
    db.complex_idx = nx ? (const uint32_t*)c->buf[kdl_ctx::B_CX_IDX].p : nullptr;
    db.hard_idx = nh ? (const uint32_t*)c->buf[kdl_ctx::B_HARD_IDX].p : nullptr;
    db.tile_index = (n_slots % KDL_TILE) == 0 ? (uint32_t*)c->buf[kdl_ctx::B_TILE_IDX].p : nullptr;

    int32_t* d_counts = (int32_t*)c->buf[kdl_ctx::B_COUNTS].p;
    int32_t* d_events = (int32_t*)c->buf[kdl_ctx::B_EVENTS].p;
    uint8_t* d_calls = (uint8_t*)c->buf[kdl_ctx::B_CALLS].p;
    int32_t* d_flag = (int32_t*)c->buf[kdl_ctx::B_FLAG].p;
    if (cudaMemsetAsync(d_flag, 0, 16, st) != cudaSuccess) return KDL_ERR_CUDA;
    // the table is reused from call to call: memset only when its layout changed (or the buffer is
    // new); otherwise the kernels overwrite the weight columns and zero the rest only if dirty
    int32_t pflags = 0;
    if (c->table_slots != n_slots) {
        if (cudaMemsetAsync(d_counts, 0, (size_t)n_slots * KDL_NCOL * 4, st) != cudaSuccess) return KDL_ERR_CUDA;
    } else {
        pflags = KDL_PILEUP_FRESH_WEIGHTS | (c->table_dirty_rest ? KDL_PILEUP_ZERO_REST : 0);
    }
    c->table_slots = -1;  // until this call has gone through
    if ((rc = kdl_pileup_range(&db, d_counts, n_slots, 0, n_slots, pflags, n_events ? d_events : nullptr, d_flag,
                               st)) != KDL_OK)
        return rc;
    if ((rc = kdl_vote(d_counts, n_slots, min_depth_ceil, d_calls, st)) != KDL_OK) return rc;
    if (cudaEventRecord(c->ev[2], st) != cudaSuccess) return KDL_ERR_CUDA;

    if (cudaMemcpyAsync(flag_host, d_flag, 16, cudaMemcpyDeviceToHost, st) != cudaSuccess) return KDL_ERR_CUDA;
    if (calls_out && cudaMemcpyAsync(calls_out, d_calls, (size_t)n_slots, cudaMemcpyDeviceToHost, st) != cudaSuccess)
        return KDL_ERR_CUDA;
    if (counts_out && cudaMemcpyAsync(counts_out, d_counts, (size_t)n_slots * KDL_NCOL * 4,
                                      cudaMemcpyDeviceToHost, st) != cudaSuccess)
        return KDL_ERR_CUDA;
    if (ins_events_out && n_events &&
        cudaMemcpyAsync(ins_events_out, d_events, (size_t)n_events * 16, cudaMemcpyDeviceToHost, st) != cudaSuccess)
        return KDL_ERR_CUDA;
    if (cudaEventRecord(c->ev[3], st) != cudaSuccess) return KDL_ERR_CUDA;
    if (cudaStreamSynchronize(st) != cudaSuccess) return KDL_ERR_CUDA;
    c->table_slots = n_slots;
    c->table_dirty_rest = hb->n_complex > 0;
    for (int k = 0; k < 3; ++k)
        if (cudaEventElapsedTime(&c->ms[k], c->ev[k], c->ev[k + 1]) != cudaSuccess) c->ms[k] = 0.f;

    if (flag_host[0]) {  // some read raised: find the first one in reference iteration order
        kdl_diag* d_diag = (kdl_diag*)c->buf[kdl_ctx::B_DIAG].p;
        if ((rc = kdl_diagnose(&db, d_diag, st)) != KDL_OK) return rc;
        if (cudaMemcpyAsync(diag_out, d_diag, sizeof(kdl_diag), cudaMemcpyDeviceToHost, st) != cudaSuccess ||
            cudaStreamSynchronize(st) != cudaSuccess)
            return KDL_ERR_CUDA;
        return diag_out->status ? diag_out->status : KDL_ERR_CUDA;
    }
    return KDL_OK;
}

int kdl_ctx_consensus(kdl_ctx* c, const kdl_batch* hb, int64_t n_slots, int64_t n_events,
                      int64_t min_depth_ceil, uint8_t* calls_out, int32_t* counts_out,
                      int32_t* ins_events_out, kdl_diag* diag_out) {
    if (!c || !diag_out || !hb) return KDL_ERR_INVALID_ARG;
    int rc = validate_batch(hb);
    if (rc != KDL_OK) return rc;
    if (n_slots <= 0 || (n_slots & 3) || n_events < 0) return KDL_ERR_INVALID_ARG;
    if (cudaSetDevice(c->device) != cudaSuccess) return KDL_ERR_CUDA;
    std::memset(diag_out, 0, sizeof(*diag_out));
    diag_out->read = -1;
    const size_t n = (size_t)hb->n_reads, nc = (size_t)hb->n_contigs, nh = (size_t)hb->n_hard, nx = (size_t)hb->n_complex;
    const struct { int which; size_t bytes; } sizes[] = {
        {kdl_ctx::B_REF_START, n * 4}, {kdl_ctx::B_SEQ_OFF, n * 4}, {kdl_ctx::B_L_SEQ, n * 4},
        {kdl_ctx::B_SEQ4, (size_t)hb->seq4_words * 4 + 16}, {kdl_ctx::B_CREAD_OFF, (nc + 1) * 8},
        {kdl_ctx::B_CLEN, nc * 4}, {kdl_ctx::B_CSLOT, nc * 8}, {kdl_ctx::B_CX_IDX, nx * 4}, {kdl_ctx::B_HARD_IDX, nh * 4},
        {kdl_ctx::B_COUNTS, (size_t)n_slots * KDL_NCOL * 4}, {kdl_ctx::B_EVENTS, (size_t)n_events * 16},
        {kdl_ctx::B_CALLS, (size_t)n_slots}, {kdl_ctx::B_FLAG, 16}, {kdl_ctx::B_DIAG, sizeof(kdl_diag)},
        {kdl_ctx::B_TILE_IDX, (size_t)(n_slots / KDL_TILE + 1) * 32},
    };
    for (const auto& z : sizes)
        if ((rc = c->ensure(z.which, z.bytes)) != KDL_OK) return rc;
    // from here on work is enqueued on the context's stream against the caller's host buffers: whatever
    // happens, do not return before the stream has drained
    int32_t flag[4] = {0, 0, 0, 0};
    rc = ctx_consensus_enqueued(c, hb, n_slots, n_events, min_depth_ceil, calls_out, counts_out, ins_events_out,
                                diag_out, flag);
    if (rc != KDL_OK) {
        cudaStreamSynchronize(c->stream);
        if (rc != KDL_ERR_INDEX && rc != KDL_ERR_KEY) {
            c->table_slots = -1;  // the table's contents are unknown after a failed call
            c->ms[0] = c->ms[1] = c->ms[2] = 0.f;
        }
    }
    return rc;
}

int kdl_ctx_last_timing(kdl_ctx* c, float* h2d_ms, float* kernel_ms, float* d2h_ms) {
    if (!c) return KDL_ERR_INVALID_ARG;
    if (h2d_ms) *h2d_ms = c->ms[0];
    if (kernel_ms) *kernel_ms = c->ms[1];
    if (d2h_ms) *d2h_ms = c->ms[2];
    return KDL_OK;
}

}  // extern "C"
