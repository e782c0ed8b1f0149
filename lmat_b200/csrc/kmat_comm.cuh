// kmat_comm.cuh -- DB-sharded mode, the exchange itself: NCCL send/recv groups between the ranks of one node, driven from
// C++ (included at the end of kmat_label.cu).  One rank = one GPU = one kmat_ctx over its shard of the table; the ranks may be
// processes (bench.py under torchrun: the unique id travels over torch.distributed once) or threads of one process (the
// read_label binary).  A pass over a rank's reads runs in rounds (SURVEY.md 8(e) mode B):
//
//     kmat_shard_encode -> [counts, query k-mers: ncclSend/ncclRecv group] -> kmat_shard_serve
//                       -> [payload sizes, hit words, list records: ncclSend/ncclRecv group] -> kmat_shard_finish
//
// NCCL is loaded with dlopen("libnccl.so.2") on first use, so libkmat itself has no link-time dependency on it (the CPU-only
// container and a process that already carries torch's NCCL both work); no other communication library is involved.
#include <dlfcn.h>
#include <nccl.h>

#include <functional>

struct KmNccl {
    void *h = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};
static KmNccl *km_nccl() {
    static KmNccl N = [] {
        KmNccl n;
        // an NCCL the process already carries (e.g. the one bundled with torch) wins: a second copy under the same soname would
        // shadow it for everything loaded later.  KMAT_NCCL_LIB names a specific library file.
        if (const char *e = getenv("KMAT_NCCL_LIB")) n.h = dlopen(e, RTLD_NOW | RTLD_LOCAL);
        if (!n.h) n.h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
        if (!n.h) n.h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
        if (!n.h) n.h = dlopen("libnccl.so", RTLD_NOW | RTLD_LOCAL);
        if (!n.h) return n;
        n.GetUniqueId = (decltype(n.GetUniqueId))dlsym(n.h, "ncclGetUniqueId");
        n.CommInitRank = (decltype(n.CommInitRank))dlsym(n.h, "ncclCommInitRank");
        n.CommDestroy = (decltype(n.CommDestroy))dlsym(n.h, "ncclCommDestroy");
        n.GroupStart = (decltype(n.GroupStart))dlsym(n.h, "ncclGroupStart");
        n.GroupEnd = (decltype(n.GroupEnd))dlsym(n.h, "ncclGroupEnd");
        n.Send = (decltype(n.Send))dlsym(n.h, "ncclSend");
        n.Recv = (decltype(n.Recv))dlsym(n.h, "ncclRecv");
        n.GetErrorString = (decltype(n.GetErrorString))dlsym(n.h, "ncclGetErrorString");
        n.ok = n.GetUniqueId && n.CommInitRank && n.CommDestroy && n.GroupStart && n.GroupEnd && n.Send && n.Recv && n.GetErrorString;
        return n; }();
    return &N;
}
#define KM_NCCL(call)                                                                                             \
    do {                                                                                                          \
        ncclResult_t r_ = (call);                                                                                 \
        if (r_ != ncclSuccess) { kmat_set_error("%s failed at %s:%d: %s", #call, __FILE__, __LINE__, km_nccl()->GetErrorString(r_)); return KMAT_ERR_CUDA; } \
    } while (0)

struct kmat_comm {
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1, device = 0;
    unsigned long long *d_cnt = nullptr;        // [2 * world]: counts out | counts in
    unsigned long long *h_cnt = nullptr;        // pinned twin
    uint64_t *d_loc = nullptr; uint64_t cap_loc = 0;                 // offsets of a round's reads, local to the round
    uint64_t *d_inbox = nullptr; uint64_t cap_inbox = 0;             // query k-mers received
    uint32_t *d_reply_in = nullptr; uint64_t cap_reply_in = 0;       // hit words received for this rank's queries
    uint32_t *d_payload_in = nullptr; uint64_t cap_payload_in = 0;   // list records received
    // host-buffer front end (kmat_shard_label_batch): this rank's batch on the device
    char *d_bases = nullptr; uint64_t cap_bases = 0;
    uint64_t *d_offs = nullptr; uint64_t cap_offs = 0;
    kmat_read_result *d_out = nullptr; uint64_t cap_out = 0;
};

extern "C" int kmat_comm_unique_id(unsigned char *id128) {
    if (!id128) return KMAT_ERR_ARG;
    KmNccl *N = km_nccl();
    if (!N->ok) { kmat_set_error("NCCL is not available (dlopen libnccl.so.2: %s)", dlerror() ? "failed" : "symbols missing"); return KMAT_ERR_UNSUPPORTED; }
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    KM_NCCL(N->GetUniqueId(&id));
    memcpy(id128, &id, 128);
    return KMAT_OK;
}
extern "C" void kmat_comm_free(kmat_comm *m) {
    if (!m) return;
    cudaSetDevice(m->device);
    if (m->comm) km_nccl()->CommDestroy(m->comm);
    cudaFree(m->d_cnt); cudaFreeHost(m->h_cnt); cudaFree(m->d_loc); cudaFree(m->d_inbox); cudaFree(m->d_reply_in); cudaFree(m->d_payload_in);
    cudaFree(m->d_bases); cudaFree(m->d_offs); cudaFree(m->d_out);
    delete m;
}
extern "C" int kmat_comm_init(int device, int rank, int world, const unsigned char *id128, kmat_comm **out) {
    if (!out || !id128 || world < 1 || world > KM_MAX_SHARDS || rank < 0 || rank >= world) { kmat_set_error("kmat_comm_init: bad argument"); return KMAT_ERR_ARG; }
    if (kmat_device_count() <= device) { kmat_set_error("CUDA device %d not available", device); return KMAT_ERR_NO_DEVICE; }
    KmNccl *N = km_nccl();
    if (!N->ok) { kmat_set_error("NCCL is not available (libnccl.so.2 could not be loaded)"); return KMAT_ERR_UNSUPPORTED; }
    KM_CUDA(cudaSetDevice(device));
    kmat_comm *m = new kmat_comm();
    m->rank = rank; m->world = world; m->device = device;
    ncclUniqueId id;
    memcpy(&id, id128, 128);
    ncclResult_t r = N->CommInitRank(&m->comm, world, id, rank);
    if (r != ncclSuccess) { kmat_set_error("ncclCommInitRank(rank %d of %d): %s", rank, world, N->GetErrorString(r)); delete m; return KMAT_ERR_CUDA; }
    if (cudaMalloc((void **)&m->d_cnt, 2 * KM_MAX_SHARDS * 8) != cudaSuccess || cudaMallocHost((void **)&m->h_cnt, 2 * KM_MAX_SHARDS * 8) != cudaSuccess) {
        cudaGetLastError(); kmat_comm_free(m); kmat_set_error("kmat_comm_init: out of memory"); return KMAT_ERR_NOMEM;
    }
    *out = m;
    return KMAT_OK;
}

__global__ void km_local_offs_kernel(const uint64_t *__restrict__ offs, uint32_t n_plus_1, uint64_t *out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_plus_1) out[i] = offs[i] - offs[0];
}

// every rank tells every rank one number (all-to-all of one u64 per pair), result on the host
static int km_comm_counts(kmat_comm *m, const uint64_t *send, uint64_t *recv, cudaStream_t st) {
    KmNccl *N = km_nccl();
    for (int j = 0; j < m->world; j++) m->h_cnt[j] = send[j];
    KM_CUDA(cudaMemcpyAsync(m->d_cnt, m->h_cnt, (size_t)m->world * 8, cudaMemcpyHostToDevice, st));
    KM_NCCL(N->GroupStart());
    for (int j = 0; j < m->world; j++) {
        KM_NCCL(N->Send(m->d_cnt + j, 1, ncclUint64, j, m->comm, st));
        KM_NCCL(N->Recv(m->d_cnt + KM_MAX_SHARDS + j, 1, ncclUint64, j, m->comm, st));
    }
    KM_NCCL(N->GroupEnd());
    KM_CUDA(cudaMemcpyAsync(m->h_cnt + KM_MAX_SHARDS, m->d_cnt + KM_MAX_SHARDS, (size_t)m->world * 8, cudaMemcpyDeviceToHost, st));
    KM_CUDA(cudaStreamSynchronize(st));
    for (int j = 0; j < m->world; j++) recv[j] = m->h_cnt[KM_MAX_SHARDS + j];
    return KMAT_OK;
}
// all-to-all of variable-size segments (send: owner 0's first, counts[j] elements to rank j; recv likewise)
template <typename T>
static int km_comm_all_to_all(kmat_comm *m, const T *send, const uint64_t *scnt, T *recv, const uint64_t *rcnt, ncclDataType_t dt, cudaStream_t st) {
    KmNccl *N = km_nccl();
    uint64_t so = 0, ro = 0;
    KM_NCCL(N->GroupStart());
    for (int j = 0; j < m->world; j++) {
        if (scnt[j]) KM_NCCL(N->Send(send + so, scnt[j], dt, j, m->comm, st));
        if (rcnt[j]) KM_NCCL(N->Recv(recv + ro, rcnt[j], dt, j, m->comm, st));
        so += scnt[j]; ro += rcnt[j];
    }
    KM_NCCL(N->GroupEnd());
    return KMAT_OK;
}

// The rounds of one collective pass.  after(r0, r1) (optional) runs on the host after the finish phase of each of THIS rank's
// rounds has been queued (the candidate pairs of a round live in the ctx until the next round starts).
static int km_shard_rounds(kmat_ctx *c, kmat_comm *m, const char *d_bases, const uint64_t *h_offs, const uint64_t *d_offs, uint32_t n_reads,
                           uint32_t round_reads, kmat_read_result *d_out, uint64_t *stats, cudaStream_t st, const std::function<int(uint32_t, uint32_t)> &after) {
    if (c->db->shard_count != m->world || c->db->shard_index != m->rank) { kmat_set_error("sharded pass: the ctx's table is shard %d of %d, the communicator rank %d of %d", c->db->shard_index, c->db->shard_count, m->rank, m->world); return KMAT_ERR_ARG; }
    if (!round_reads) round_reads = 1u << 20;
    const uint64_t round_bases = (1ull << 32) - (1ull << 20);
    const int W = m->world;
    int rc;
    // plan this rank's rounds, then agree on the number of rounds of the pass (the maximum over the ranks)
    std::vector<std::pair<uint32_t, uint32_t>> rounds;
    for (uint32_t r0 = 0; r0 < n_reads;) {
        uint32_t r1 = (uint32_t)std::min<uint64_t>(n_reads, (uint64_t)r0 + round_reads);
        while (r1 > r0 + 1 && h_offs[r1] - h_offs[r0] > round_bases) r1 = r0 + (r1 - r0) / 2;
        rounds.emplace_back(r0, r1);
        r0 = r1;
    }
    uint64_t mine[KM_MAX_SHARDS], theirs[KM_MAX_SHARDS];
    for (int j = 0; j < W; j++) mine[j] = rounds.size();
    if ((rc = km_comm_counts(m, mine, theirs, st)) != KMAT_OK) return rc;
    uint64_t n_rounds = 0;
    for (int j = 0; j < W; j++) n_rounds = std::max(n_rounds, theirs[j]);
    uint64_t lookups = 0, served = 0, pay_words = 0;
    for (uint64_t i = 0; i < n_rounds; i++) {
        const bool have = i < rounds.size();
        const uint32_t r0 = have ? rounds[i].first : 0, r1 = have ? rounds[i].second : 0, n = r1 - r0;
        const uint64_t nb = have ? h_offs[r1] - h_offs[r0] : 0;
        uint32_t max_len = 0;
        for (uint32_t r = r0; r < r1; r++) max_len = std::max<uint32_t>(max_len, (uint32_t)(h_offs[r + 1] - h_offs[r]));
        if ((uint64_t)n + 1 > m->cap_loc) { KM_CUDA(cudaStreamSynchronize(st)); if ((rc = km_grow(&m->d_loc, &m->cap_loc, (uint64_t)n + 1)) != KMAT_OK) return rc; }
        if (n) {
            km_local_offs_kernel<<<(n + 1 + 255) / 256, 256, 0, st>>>(d_offs + r0, n + 1, m->d_loc);
            g_km_launches++;
        } else KM_CUDA(cudaMemsetAsync(m->d_loc, 0, 8, st));
        // home: encode
        const uint64_t *d_q = nullptr;
        uint64_t scnt[KM_MAX_SHARDS] = {0}, rcnt[KM_MAX_SHARDS] = {0}, pcnt[KM_MAX_SHARDS] = {0}, mpcnt[KM_MAX_SHARDS] = {0};
        if ((rc = kmat_shard_encode(c, n ? d_bases + h_offs[r0] : d_bases, m->d_loc, n, nb, max_len, W, &d_q, scnt, st)) != KMAT_OK) return rc;
        // exchange 1: counts, then the query k-mers
        if ((rc = km_comm_counts(m, scnt, rcnt, st)) != KMAT_OK) return rc;
        uint64_t n_in = 0, n_out = 0;
        for (int j = 0; j < W; j++) { n_in += rcnt[j]; n_out += scnt[j]; }
        if (n_in > m->cap_inbox) { KM_CUDA(cudaStreamSynchronize(st)); if ((rc = km_grow(&m->d_inbox, &m->cap_inbox, n_in)) != KMAT_OK) return rc; }
        if ((rc = km_comm_all_to_all<uint64_t>(m, d_q, scnt, m->d_inbox, rcnt, ncclUint64, st)) != KMAT_OK) return rc;
        // owner: serve
        const uint32_t *d_reply = nullptr, *d_payload = nullptr;
        if ((rc = kmat_shard_serve(c, m->d_inbox, rcnt, W, &d_reply, &d_payload, pcnt, st)) != KMAT_OK) return rc;
        // exchange 2: payload sizes, hit words (the split of exchange 1 reversed), list records
        if ((rc = km_comm_counts(m, pcnt, mpcnt, st)) != KMAT_OK) return rc;
        uint64_t n_pay_in = 0;
        for (int j = 0; j < W; j++) n_pay_in += mpcnt[j];
        if (n_out > m->cap_reply_in || n_pay_in + 8 > m->cap_payload_in) {
            KM_CUDA(cudaStreamSynchronize(st));
            if ((rc = km_grow(&m->d_reply_in, &m->cap_reply_in, n_out)) != KMAT_OK) return rc;
            if ((rc = km_grow(&m->d_payload_in, &m->cap_payload_in, n_pay_in + 8)) != KMAT_OK) return rc;
        }
        if ((rc = km_comm_all_to_all<uint32_t>(m, d_reply, rcnt, m->d_reply_in, scnt, ncclUint32, st)) != KMAT_OK) return rc;
        if ((rc = km_comm_all_to_all<uint32_t>(m, d_payload, pcnt, m->d_payload_in, mpcnt, ncclUint32, st)) != KMAT_OK) return rc;
        // home: finish (K3 / K4); asynchronous
        if ((rc = kmat_shard_finish(c, n_out ? m->d_reply_in : nullptr, n_pay_in ? m->d_payload_in : nullptr, mpcnt, W, n ? d_out + r0 : nullptr, st)) != KMAT_OK) return rc;
        lookups += n_out; served += n_in; pay_words += n_pay_in;
        if (have && after && (rc = after(r0, r1)) != KMAT_OK) return rc;
    }
    if (stats) { stats[0] = lookups; stats[1] = served; stats[2] = pay_words; stats[3] = n_rounds; }
    return KMAT_OK;
}

// Labels this rank's device-resident reads against the sharded table.  COLLECTIVE: every rank of the communicator calls it
// (with its own reads, possibly none); inside, ranks keep serving the others' queries until every rank has finished.
// h_offs / d_offs: the n_reads + 1 absolute base offsets on the host and on the device; results to d_out[0 .. n_reads).
// stats (optional, host): [0] unique lookups sent, [1] queries served, [2] list-record words received, [3] rounds.
extern "C" int kmat_shard_label_device(kmat_ctx *c, kmat_comm *m, const char *d_bases, const uint64_t *h_offs, const uint64_t *d_offs, uint32_t n_reads,
                                       uint32_t round_reads, kmat_read_result *d_out, uint64_t *stats, void *stream) {
    if (!c || !m || (n_reads && (!d_bases || !h_offs || !d_offs || !d_out))) { kmat_set_error("kmat_shard_label_device: bad argument"); return KMAT_ERR_ARG; }
    KM_CUDA(cudaSetDevice(c->device));
    return km_shard_rounds(c, m, d_bases, h_offs, d_offs, n_reads, round_reads, d_out, stats, stream ? (cudaStream_t)stream : c->stream, nullptr);
}

// Host buffers in, host buffers out -- kmat_label_batch for a sharded table.  COLLECTIVE like the call above (a rank without
// reads passes n_reads = 0).  The candidate / lineage pairs of every round are copied out before the next round reuses the
// ctx's buffers; cand_off / lin_off index the caller's arrays.  KMAT_ERR_OVERFLOW: *n_cands / *n_lineage = the capacities needed
// (the pass has run to its end, so the ranks stay in step; call again with larger buffers -- collectively).
extern "C" int kmat_shard_label_batch(kmat_ctx *c, kmat_comm *m, const char *bases, const uint64_t *offs, uint32_t n_reads, kmat_read_result *out,
                                      kmat_pair *cands, uint64_t cands_cap, uint64_t *n_cands, kmat_pair *lineage, uint64_t lineage_cap, uint64_t *n_lineage) {
    if (!c || !m || (n_reads && (!bases || !offs || !out))) { kmat_set_error("kmat_shard_label_batch: bad argument"); return KMAT_ERR_ARG; }
    KM_CUDA(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    int rc;
    const uint64_t nb = n_reads ? offs[n_reads] - offs[0] : 0;
    if (n_reads) {
        KM_CUDA(cudaStreamSynchronize(st));
        if ((rc = km_grow(&m->d_bases, &m->cap_bases, nb + 16)) != KMAT_OK) return rc;
        if ((rc = km_grow(&m->d_offs, &m->cap_offs, (uint64_t)n_reads + 1)) != KMAT_OK) return rc;
        if ((rc = km_grow(&m->d_out, &m->cap_out, (uint64_t)n_reads)) != KMAT_OK) return rc;
        KM_CUDA(cudaMemcpyAsync(m->d_bases, bases + offs[0], nb, cudaMemcpyHostToDevice, st));
        KM_CUDA(cudaMemcpyAsync(m->d_offs, offs, ((size_t)n_reads + 1) * 8, cudaMemcpyHostToDevice, st));
    }
    unsigned long long tot_c = 0, tot_l = 0;
    auto after = [&](uint32_t r0, uint32_t r1) -> int {
        KM_CUDA(cudaStreamSynchronize(st));
        unsigned long long cur[2] = {0, 0};
        KM_CUDA(cudaMemcpy(cur, c->d_cursors, 16, cudaMemcpyDeviceToHost));
        if (cur[0] > c->cap_cands || (c->opt.want_lineage && cur[1] > c->cap_lin)) { kmat_set_error("sharded pass: the candidate buffer of a round was too small (%llu pairs)", cur[0]); return KMAT_ERR_NOMEM; }
        KM_CUDA(cudaMemcpy(out + r0, m->d_out + r0, (size_t)(r1 - r0) * sizeof(kmat_read_result), cudaMemcpyDeviceToHost));
        for (uint32_t r = r0; r < r1; r++) { out[r].cand_off += tot_c; out[r].lin_off += tot_l; }
        if (cands && tot_c + cur[0] <= cands_cap && cur[0]) KM_CUDA(cudaMemcpy(cands + tot_c, c->d_cands, (size_t)cur[0] * sizeof(kmat_pair), cudaMemcpyDeviceToHost));
        if (lineage && c->opt.want_lineage && tot_l + cur[1] <= lineage_cap && cur[1]) KM_CUDA(cudaMemcpy(lineage + tot_l, c->d_lin, (size_t)cur[1] * sizeof(kmat_pair), cudaMemcpyDeviceToHost));
        tot_c += cur[0]; tot_l += c->opt.want_lineage ? cur[1] : 0;
        return KMAT_OK;
    };
    // the device arrays are addressed with offsets relative to offs[0]
    std::vector<uint64_t> rel;
    const uint64_t *h_offs = offs;
    if (n_reads && offs[0]) { rel.resize((size_t)n_reads + 1); for (uint32_t i = 0; i <= n_reads; i++) rel[i] = offs[i] - offs[0]; h_offs = rel.data(); KM_CUDA(cudaMemcpyAsync(m->d_offs, rel.data(), rel.size() * 8, cudaMemcpyHostToDevice, st)); KM_CUDA(cudaStreamSynchronize(st)); }
    if ((rc = km_shard_rounds(c, m, m->d_bases, h_offs, m->d_offs, n_reads, 0, m->d_out, nullptr, st, after)) != KMAT_OK) return rc;
    KM_CUDA(cudaStreamSynchronize(st));
    if (n_cands) *n_cands = tot_c;
    if (n_lineage) *n_lineage = tot_l;
    if ((cands && tot_c > cands_cap) || (lineage && c->opt.want_lineage && tot_l > lineage_cap)) { kmat_set_error("candidate buffer too small: need %llu candidate and %llu lineage pairs", tot_c, tot_l); return KMAT_ERR_OVERFLOW; }
    return KMAT_OK;
}
