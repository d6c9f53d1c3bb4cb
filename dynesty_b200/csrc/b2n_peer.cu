// b2n_peer.cu -- multi-GPU exchange of finished chains over NVLink peer memory.
//
// the path shards by chains (the reference's pool.map over queue slots,
// sampler.py:717) and every rank needs every finished chain, i.e. an all-gather per queue fill.
// Here the gather is FUSED into the chain kernels: each rank owns an exchange window in its HBM,
// maps the windows of all peers (CUDA IPC; NVSwitch gives every pair full bandwidth), and the
// kernels store every finished chain into all windows (b2n_chain.cuh: peer_put).  The grid's last
// CTA then signals every peer's arrive counter and waits for its own counter to reach
// world x epoch (peer_finish), so "launch complete on this rank" implies "all rows of the fill
// are present in this rank's window" -- no collective, no extra launch, the stores overlap the
// tail of the compute.  This file is the host side: window management and the gather-mode
// plumbing of the batch entry points.
#include "b2n_common.cuh"

static inline uint64_t al256(uint64_t b) { return (b + 255) & ~(uint64_t)255; }

static uint64_t slot_bytes(int64_t R, int n) {
    return 2 * al256((uint64_t)R * n * 8) + al256((uint64_t)R * 8) + 4 * al256((uint64_t)R * 4);
}

static void slot_offsets(int64_t R, int n, int slot, uint64_t off[7]) {
    uint64_t o = B2N_PEER_HDR + (uint64_t)slot * slot_bytes(R, n);
    off[0] = o; o += al256((uint64_t)R * n * 8);
    off[1] = o; o += al256((uint64_t)R * n * 8);
    off[2] = o; o += al256((uint64_t)R * 8);
    for (int k = 0; k < 4; k++) { off[3 + k] = o; o += al256((uint64_t)R * 4); }
}

void b2n_peer_release(b2n_ctx* ctx) {
    PeerState& P = ctx->peer;
    for (int w = 0; w < B2N_MAX_PEERS; w++) {
        if (P.opened[w] && P.base[w]) cudaIpcCloseMemHandle(P.base[w]);
        P.opened[w] = false;
        P.base[w] = nullptr;
    }
    if (P.win) cudaFree(P.win);
    if (P.err_host) cudaFreeHost(P.err_host);
    P = PeerState();
}

extern "C" {

uint64_t b2n_peer_window_bytes(int64_t total_rows, int32_t ndim) {
    return B2N_PEER_HDR + 2 * slot_bytes(total_rows, ndim);
}

int b2n_peer_export(b2n_ctx* ctx, uint64_t bytes, unsigned char* handle) {
    if (!ctx || !handle || bytes < B2N_PEER_HDR) return B2N_ERR_ARG;
    static_assert(sizeof(cudaIpcMemHandle_t) == B2N_PEER_HANDLE_BYTES, "IPC handle size");
    B2N_CUDA(ctx, cudaSetDevice(ctx->device));
    B2N_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    b2n_peer_release(ctx);
    PeerState& P = ctx->peer;
    B2N_CUDA(ctx, cudaMalloc((void**)&P.win, bytes));
    P.win_bytes = bytes;
    B2N_CUDA(ctx, cudaMemsetAsync(P.win, 0, bytes, ctx->stream));
    B2N_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    B2N_CUDA(ctx, cudaDeviceSynchronize());          // header is zero before any peer can store
    B2N_CUDA(ctx, cudaHostAlloc((void**)&P.err_host, 64, cudaHostAllocDefault));
    *P.err_host = 0;
    cudaIpcMemHandle_t h;
    B2N_CUDA(ctx, cudaIpcGetMemHandle(&h, P.win));
    memcpy(handle, &h, sizeof(h));
    return B2N_OK;
}

static int peer_common(b2n_ctx* ctx, int rank, int world) {
    if (!ctx || world < 1 || world > B2N_MAX_PEERS || rank < 0 || rank >= world) return B2N_ERR_ARG;
    if (!ctx->peer.win) return b2n_fail(ctx, B2N_ERR_PEER, "b2n_peer_export must come first");
    ctx->peer.world = world;
    ctx->peer.rank = rank;
    ctx->peer.epoch = 0;
    return B2N_OK;
}

int b2n_peer_import(b2n_ctx* ctx, int32_t rank, int32_t world, const unsigned char* handles) {
    B2N_TRY(peer_common(ctx, rank, world));
    if (!handles) return B2N_ERR_ARG;
    B2N_CUDA(ctx, cudaSetDevice(ctx->device));
    PeerState& P = ctx->peer;
    for (int w = 0; w < world; w++) {
        if (w == rank) { P.base[w] = P.win; continue; }
        cudaIpcMemHandle_t h;
        memcpy(&h, handles + (size_t)w * B2N_PEER_HANDLE_BYTES, sizeof(h));
        void* p = nullptr;
        B2N_CUDA(ctx, cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        P.base[w] = (char*)p;
        P.opened[w] = true;
    }
    return B2N_OK;
}

int b2n_peer_import_raw(b2n_ctx* ctx, int32_t rank, int32_t world, void* const* windows) {
    B2N_TRY(peer_common(ctx, rank, world));
    if (!windows) return B2N_ERR_ARG;
    PeerState& P = ctx->peer;
    for (int w = 0; w < world; w++) {
        if (w != rank && !windows[w]) return B2N_ERR_ARG;
        P.base[w] = (w == rank) ? P.win : (char*)windows[w];
    }
    return B2N_OK;
}

int b2n_peer_rows(b2n_ctx* ctx, int64_t row0, int64_t total_rows) {
    if (!ctx || row0 < 0 || total_rows < 0 || (total_rows > 0 && row0 >= total_rows)) return B2N_ERR_ARG;
    if (total_rows > 0 && ctx->peer.world < 1) return b2n_fail(ctx, B2N_ERR_PEER, "peer windows not imported");
    ctx->peer.row0 = row0;
    ctx->peer.total = total_rows;
    return B2N_OK;
}

int b2n_peer_result(b2n_ctx* ctx, void** window, uint64_t* offsets7) {
    if (!ctx || !ctx->peer.win) return B2N_ERR_ARG;
    if (window) *window = ctx->peer.win;
    if (offsets7) memcpy(offsets7, ctx->peer.off, sizeof(ctx->peer.off));
    return B2N_OK;
}

int b2n_peer_read(b2n_ctx* ctx, uint64_t offset, void* host_dst, uint64_t bytes) {
    if (!ctx || !ctx->peer.win || !host_dst || offset + bytes > ctx->peer.win_bytes) return B2N_ERR_ARG;
    B2N_CUDA(ctx, cudaSetDevice(ctx->device));
    B2N_CUDA(ctx, cudaMemcpyAsync(host_dst, ctx->peer.win + offset, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    B2N_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return B2N_OK;
}

int b2n_peer_check(b2n_ctx* ctx) {
    if (!ctx || !ctx->peer.win) return B2N_ERR_ARG;
    B2N_CUDA(ctx, cudaSetDevice(ctx->device));
    B2N_CUDA(ctx, cudaMemcpyAsync(ctx->peer.err_host, ctx->peer.win + 8, 4, cudaMemcpyDeviceToHost, ctx->stream));
    B2N_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (*ctx->peer.err_host) return b2n_fail(ctx, B2N_ERR_PEER, "a peer never arrived at the exchange (timeout in the kernel)");
    return B2N_OK;
}

}  // extern "C"

int b2n_peer_begin(b2n_ctx* ctx, int n, PeerSet* ps, void** dev7, bool* on) {
    PeerState& P = ctx->peer;
    *on = P.total > 0;
    *ps = PeerSet();
    if (!*on) return B2N_OK;
    if (b2n_peer_window_bytes(P.total, n) > P.win_bytes)
        return b2n_fail(ctx, B2N_ERR_PEER, "exchange window too small for this fill (b2n_peer_window_bytes)");
    P.epoch++;
    slot_offsets(P.total, n, (int)(P.epoch & 1), P.off);
    const uint64_t rowb[7] = {(uint64_t)n * 8, (uint64_t)n * 8, 8, 4, 4, 4, 4};
    for (int k = 0; k < 7; k++) dev7[k] = P.win + P.off[k] + (uint64_t)P.row0 * rowb[k];
    ps->world = P.world;
    ps->rank = P.rank;
    for (int w = 0; w < P.world; w++) ps->base[w] = P.base[w];
    ps->target = (unsigned long long)P.world * P.epoch;
    return B2N_OK;
}

int b2n_peer_end(b2n_ctx* ctx, int n, void* const* user7) {
    PeerState& P = ctx->peer;
    const uint64_t rowb[7] = {(uint64_t)n * 8, (uint64_t)n * 8, 8, 4, 4, 4, 4};
    const cudaMemcpyKind kind = ctx->ptr_mode == B2N_PTR_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
    for (int k = 0; k < 7; k++) {
        if (!user7[k]) continue;
        B2N_CUDA(ctx, cudaMemcpyAsync(user7[k], P.win + P.off[k], (size_t)P.total * rowb[k], kind, ctx->stream));
    }
    if (ctx->ptr_mode == B2N_PTR_HOST)
        B2N_CUDA(ctx, cudaMemcpyAsync(P.err_host, P.win + 8, 4, cudaMemcpyDeviceToHost, ctx->stream));
    return B2N_OK;
}

int b2n_peer_finish(b2n_ctx* ctx, bool on) {
    B2N_TRY(b2n_finish(ctx));
    if (on && ctx->ptr_mode == B2N_PTR_HOST && *ctx->peer.err_host)
        return b2n_fail(ctx, B2N_ERR_PEER, "a peer never arrived at the exchange (timeout in the kernel)");
    return B2N_OK;
}
