// Host plumbing of a sharded search on one box: the ranks (one process per GPU) divide the parse + compile work of a
// batch. Every rank compiles its slice of the query lines, publishes the compiled descriptors in a POSIX shared-memory
// segment and reads the slices of the others; the descriptors depend on GLOBAL statistics only, so whichever rank compiles
// a query produces the same bytes. No collective, no GPU: atomics in shared memory, two buffers per rank so that a rank
// may publish round n + 1 while others still read round n.
#pragma once

#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstring>
#include <functional>
#include <memory>
#include <string>
#include <thread>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

namespace dgpu {

class ShmExchange {
public:
    static constexpr size_t kSlotBytes = 16u << 20;          // largest slice one rank may publish per round
    static constexpr uint64_t kTooLarge = ~0ull;             // published size: "my slice does not fit" (all ranks fall back)
    static constexpr uint64_t kFailed = ~0ull - 1;           // "my slice failed to compile" (all ranks fail the call)

    // All ranks of one communicator derive the segment's name from the communicator id. Returns null when shared memory
    // cannot be set up (the caller then compiles the whole batch on every rank).
    static std::unique_ptr<ShmExchange> create(const uint8_t id[128], int rank, int world) {
        if (world < 2) return nullptr;
        uint64_t h = 1469598103934665603ull;
        for (int i = 0; i < 128; ++i) h = (h ^ id[i]) * 1099511628211ull;
        char name[64];
        std::snprintf(name, sizeof name, "/dgpu_xch_%016llx", static_cast<unsigned long long>(h));
        const size_t bytes = sizeof(Header) + static_cast<size_t>(world) * 2 * kSlotBytes;
        int fd = -1;
        if (rank == 0) {
            shm_unlink(name);
            fd = shm_open(name, O_CREAT | O_EXCL | O_RDWR, 0600);
            if (fd < 0 || ftruncate(fd, static_cast<off_t>(bytes)) != 0) {
                if (fd >= 0) close(fd);
                return nullptr;
            }
        } else {
            const auto deadline = std::chrono::steady_clock::now() + std::chrono::seconds(60);
            for (;;) {
                fd = shm_open(name, O_RDWR, 0600);
                struct stat st {};
                if (fd >= 0 && fstat(fd, &st) == 0 && static_cast<size_t>(st.st_size) >= bytes) break;
                if (fd >= 0) close(fd);
                fd = -1;
                if (std::chrono::steady_clock::now() > deadline) return nullptr;
                std::this_thread::sleep_for(std::chrono::milliseconds(2));
            }
        }
        void* p = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
        close(fd);
        if (p == MAP_FAILED) return nullptr;
        std::unique_ptr<ShmExchange> x(new ShmExchange());
        x->hdr_ = static_cast<Header*>(p);
        x->bytes_ = bytes;
        x->rank_ = rank;
        x->world_ = world;
        x->name_ = name;
        // (a fresh segment is zero-filled: all counters start at 0)
        x->hdr_->attached.fetch_add(1, std::memory_order_acq_rel);
        // the name is removed once everybody holds a mapping: nothing is left behind if a rank dies later
        const auto deadline = std::chrono::steady_clock::now() + std::chrono::seconds(60);
        while (x->hdr_->attached.load(std::memory_order_acquire) < world) {
            if (std::chrono::steady_clock::now() > deadline) return nullptr;
            std::this_thread::sleep_for(std::chrono::milliseconds(1));
        }
        if (rank == 0) shm_unlink(name);
        return x;
    }

    ~ShmExchange() {
        if (hdr_) munmap(hdr_, bytes_);
    }

    // Round `seq` (1, 2, ...; every rank calls with the same sequence): publishes `n` bytes (or kTooLarge / kFailed as
    // n with data == nullptr) and hands every rank's contribution to `fn` in rank order. Returns false on a timeout.
    bool exchange(uint64_t seq, const void* data, uint64_t n, const std::function<void(int, const uint8_t*, uint64_t)>& fn) {
        const auto deadline = std::chrono::steady_clock::now() + std::chrono::seconds(120);
        auto wait = [&](const std::function<bool()>& ready) {
            for (int spin = 0; !ready(); ++spin) {
                if (spin > 200) std::this_thread::yield();
                if ((spin & 1023) == 1023 && std::chrono::steady_clock::now() > deadline) return false;
            }
            return true;
        };
        // the buffer of this parity was last used in round seq - 2: everybody must have read it
        if (seq > 2 && !wait([&] {
                for (int r = 0; r < world_; ++r)
                    if (hdr_->rank[r].consumed.load(std::memory_order_acquire) + 2 < seq) return false;
                return true;
            }))
            return false;
        Rank& me = hdr_->rank[rank_];
        if (n != kTooLarge && n != kFailed) {
            if (n > kSlotBytes) n = kTooLarge;
            else std::memcpy(slot(rank_, seq), data, n);
        }
        me.size[seq & 1] = n;
        me.published.store(seq, std::memory_order_release);
        for (int r = 0; r < world_; ++r) {
            Rank& o = hdr_->rank[r];
            if (!wait([&] { return o.published.load(std::memory_order_acquire) >= seq; })) return false;
            fn(r, slot(r, seq), o.size[seq & 1]);
        }
        me.consumed.store(seq, std::memory_order_release);
        return true;
    }

    int rank() const { return rank_; }
    int world() const { return world_; }

private:
    struct alignas(64) Rank {
        std::atomic<uint64_t> published;
        std::atomic<uint64_t> consumed;
        uint64_t size[2];
    };
    struct Header {
        std::atomic<int> attached;
        char pad[60];
        Rank rank[64];
    };
    ShmExchange() = default;
    uint8_t* slot(int r, uint64_t seq) {
        return reinterpret_cast<uint8_t*>(hdr_) + sizeof(Header) + (static_cast<size_t>(r) * 2 + (seq & 1)) * kSlotBytes;
    }
    Header* hdr_ = nullptr;
    size_t bytes_ = 0;
    int rank_ = 0, world_ = 1;
    std::string name_;
};

}  // namespace dgpu
