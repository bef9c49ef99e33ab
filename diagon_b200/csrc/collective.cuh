// Segment-sharded search: the ONE exchange step of a batch. Included by engine.cu only.
//
// Every rank holds a contiguous run of segments on its own GPU and scores the whole query batch on its documents
// (the reference loops over the leaves and shares one collector heap: IndexSearcher.cpp:76-110). Per batch the ranks
// then exchange, per query, their local top k and hit count: pack_results_kernel lays them out as one record of
// (k + 2) 64-bit words per query - k keys (orderable score << 32 | ~doc: descending key order is the collector's
// "score desc, doc asc", TopScoreDocCollector.h:154-164), the count, the hits - ONE ncclAllGather over NVLink moves the
// records, merge_packed_kernel ranks every key among the other ranks' sorted lists and writes the best k and the summed
// hit count back into the engine's result buffers. Nothing else crosses GPUs; postings never do.
//
// NCCL is bound at run time (dlopen of libnccl.so.2: the copy the process already holds, e.g. PyTorch's, or the
// system's), so a single-GPU user of libdiagon_b200.so needs no NCCL at all.
#pragma once

#include <dlfcn.h>
#include <nccl.h>   // types and prototypes only; nothing links against it

namespace {

struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    std::string error;

    static NcclApi& get() {
        static NcclApi api = [] {
            NcclApi a;
            for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
                a.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
                if (a.lib) break;
            }
            if (!a.lib) {
                a.error = std::string("NCCL is not available (dlopen libnccl.so.2: ") + dlerror() + ")";
                return a;
            }
            auto sym = [&](const char* n) {
                void* p = dlsym(a.lib, n);
                if (!p && a.error.empty()) a.error = std::string("NCCL symbol missing: ") + n;
                return p;
            };
            a.GetUniqueId = reinterpret_cast<decltype(a.GetUniqueId)>(sym("ncclGetUniqueId"));
            a.CommInitRank = reinterpret_cast<decltype(a.CommInitRank)>(sym("ncclCommInitRank"));
            a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(sym("ncclCommDestroy"));
            a.AllGather = reinterpret_cast<decltype(a.AllGather)>(sym("ncclAllGather"));
            a.AllReduce = reinterpret_cast<decltype(a.AllReduce)>(sym("ncclAllReduce"));
            a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(sym("ncclGetErrorString"));
            return a;
        }();
        return api;
    }
};

#define NC(expr)                                                                                       \
    do {                                                                                               \
        ncclResult_t _r = (expr);                                                                      \
        if (_r != ncclSuccess)                                                                         \
            return fail("%s failed: %s (%s:%d)", #expr, NcclApi::get().GetErrorString(_r), __FILE__, __LINE__); \
    } while (0)

// one record per query: [k keys][count][hits]
__global__ void pack_results_kernel(const uint64_t* __restrict__ keys, const int32_t* __restrict__ counts,
                                    const int64_t* __restrict__ hits, uint32_t n_queries, int k, uint64_t* __restrict__ out) {
    const uint32_t q = blockIdx.x;
    if (q >= n_queries) return;
    uint64_t* rec = out + static_cast<size_t>(q) * (k + 2);
    for (int i = threadIdx.x; i < k; i += blockDim.x) rec[i] = keys[static_cast<size_t>(q) * k + i];
    if (threadIdx.x == 0) {
        rec[k] = static_cast<uint64_t>(static_cast<uint32_t>(counts[q]));
        rec[k + 1] = static_cast<uint64_t>(hits[q]);
    }
}

// gathered: [rank][query][k + 2]. Every key finds its rank among the other ranks' sorted lists by binary search; keys are
// unique (the ranks hold disjoint docs), so the ranks form a permutation (the collector's order, TopScoreDocCollector.cpp:205-231).
__global__ void merge_packed_kernel(const uint64_t* __restrict__ gathered, int world, uint32_t n_queries, int k,
                                    uint64_t* __restrict__ out_keys, int32_t* __restrict__ out_counts,
                                    int64_t* __restrict__ out_hits) {
    const uint32_t q = blockIdx.x;
    if (q >= n_queries) return;
    const size_t rec = static_cast<size_t>(k) + 2, per_rank = static_cast<size_t>(n_queries) * rec;
    int total = 0;
    int64_t hits = 0;
    for (int p = 0; p < world; ++p) {
        const uint64_t* r = gathered + p * per_rank + q * rec;
        total += static_cast<int>(r[k]);
        hits += static_cast<int64_t>(r[k + 1]);
    }
    const int n_out = min(total, k);
    for (int i = threadIdx.x; i < k; i += blockDim.x)
        if (i >= n_out) out_keys[static_cast<size_t>(q) * k + i] = 0ull;
    for (int e = threadIdx.x; e < world * k; e += blockDim.x) {
        const int p = e / k, i = e % k;
        const uint64_t* mine = gathered + p * per_rank + q * rec;
        if (i >= static_cast<int>(mine[k])) continue;
        const uint64_t key = mine[i];
        int rank = i;
        for (int o = 0; o < world; ++o) {
            if (o == p) continue;
            const uint64_t* other = gathered + o * per_rank + q * rec;
            int lo = 0, hi = static_cast<int>(other[k]);
            while (lo < hi) {   // number of keys of rank o greater than key
                const int mid = (lo + hi) >> 1;
                if (other[mid] > key) lo = mid + 1; else hi = mid;
            }
            rank += lo;
        }
        if (rank < k) out_keys[static_cast<size_t>(q) * k + rank] = key;
    }
    if (threadIdx.x == 0) {
        out_counts[q] = n_out;
        out_hits[q] = hits;
    }
}

}  // namespace

struct dgpu_comm {
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1, device = 0;
};

extern "C" {

int dgpu_comm_unique_id(uint8_t out[DGPU_COMM_ID_BYTES]) {
    NcclApi& n = NcclApi::get();
    if (!n.error.empty()) return fail("%s", n.error.c_str());
    static_assert(sizeof(ncclUniqueId) == DGPU_COMM_ID_BYTES, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    NC(n.GetUniqueId(&id));
    std::memcpy(out, &id, sizeof id);
    return 0;
}

int dgpu_comm_create(const uint8_t id_bytes[DGPU_COMM_ID_BYTES], int rank, int world, int device, dgpu_comm** out) {
    *out = nullptr;
    NcclApi& n = NcclApi::get();
    if (!n.error.empty()) return fail("%s", n.error.c_str());
    if (world < 1 || rank < 0 || rank >= world) return fail("bad rank %d of %d", rank, world);
    CU(cudaSetDevice(device));
    ncclUniqueId id;
    std::memcpy(&id, id_bytes, sizeof id);
    auto* c = new dgpu_comm();
    c->rank = rank;
    c->world = world;
    c->device = device;
    ncclResult_t r = n.CommInitRank(&c->comm, world, id, rank);
    if (r != ncclSuccess) {
        delete c;
        return fail("ncclCommInitRank failed: %s", n.GetErrorString(r));
    }
    *out = c;
    return 0;
}

void dgpu_comm_destroy(dgpu_comm* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->comm) NcclApi::get().CommDestroy(c->comm);
    delete c;
}

int dgpu_comm_rank(const dgpu_comm* c) { return c->rank; }
int dgpu_comm_world(const dgpu_comm* c) { return c->world; }

int dgpu_comm_allreduce_sum_i64(dgpu_comm* c, int64_t* host_inout, size_t n) {
    if (n == 0) return 0;
    NcclApi& api = NcclApi::get();
    CU(cudaSetDevice(c->device));
    int64_t* d = nullptr;
    CU(cudaMalloc(&d, n * sizeof(int64_t)));
    cudaError_t e = cudaMemcpy(d, host_inout, n * sizeof(int64_t), cudaMemcpyHostToDevice);
    ncclResult_t r = ncclSuccess;
    if (e == cudaSuccess) r = api.AllReduce(d, d, n, ncclInt64, ncclSum, c->comm, nullptr);
    if (e == cudaSuccess && r == ncclSuccess) e = cudaStreamSynchronize(nullptr);
    if (e == cudaSuccess && r == ncclSuccess) e = cudaMemcpy(host_inout, d, n * sizeof(int64_t), cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (r != ncclSuccess) return fail("ncclAllReduce failed: %s", api.GetErrorString(r));
    if (e != cudaSuccess) return fail("all-reduce of the index statistics failed: %s", cudaGetErrorString(e));
    return 0;
}

int dgpu_engine_exchange_topk(dgpu_engine* e, dgpu_comm* c, void* stream_v) {
    CU(cudaSetDevice(e->device));
    cudaStream_t stream = stream_v ? static_cast<cudaStream_t>(stream_v) : e->stream;
    if (e->n_queries == 0 || c->world == 1) return 0;
    const size_t rec = static_cast<size_t>(e->k) + 2, mine = static_cast<size_t>(e->n_queries) * rec;
    CU(e->d_packed.ensure(mine));
    CU(e->d_gathered.ensure(mine * static_cast<size_t>(c->world)));
    pack_results_kernel<<<e->n_queries, 128, 0, stream>>>(e->d_keys.p, e->d_counts.p, e->d_hits.p, e->n_queries, e->k, e->d_packed.p);
    CU(cudaGetLastError());
    NC(NcclApi::get().AllGather(e->d_packed.p, e->d_gathered.p, mine * sizeof(uint64_t), ncclUint8, c->comm, stream));
    merge_packed_kernel<<<e->n_queries, 128, 0, stream>>>(e->d_gathered.p, c->world, e->n_queries, e->k, e->d_keys.p,
                                                          e->d_counts.p, e->d_hits.p);
    CU(cudaGetLastError());
    e->launches += 2;
    e->collectives++;
    return 0;
}

}  // extern "C"
