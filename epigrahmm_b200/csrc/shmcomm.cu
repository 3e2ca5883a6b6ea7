// shmcomm.cu -- all-gather of small HOST vectors between the ranks of one node.
//
// The sharded EM exchanges, per iteration, a K x K block operator, a dozen
// sufficient statistics and the flagged counts per column: a few hundred bytes that are produced
// and consumed by host code (the carries are composed on the host, the counts size the uniform
// draw).  Through NCCL each of those costs a pinned copy in, a collective launch, a copy out and a
// stream synchronisation (~0.15 ms measured); through a POSIX shared-memory segment with one
// sequence flag per rank it costs a few microseconds.  NCCL keeps the one exchange that moves real
// data (the all-reduce of the rejection tables in HBM).
//
// Layout of the segment: world flags (one cache line each, a 64-bit epoch), then two banks of
// world slots (bank = epoch parity).  A rank cannot run two epochs ahead of a reader: to finish
// epoch e+1 it needs every flag >= e+1, which a rank only publishes after it has copied epoch e out.
#include <fcntl.h>
#include <sched.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>

#include <cstring>

#include "common.cuh"

struct ehmm_comm {
    int rank, world;
    size_t slot, bytes;
    uint8_t *base;
    uint64_t epoch;
    double timeout_s;
    char name[96];
};

namespace {
constexpr size_t LINE = 64;
double now_s() {
    timespec t;
    clock_gettime(CLOCK_MONOTONIC, &t);
    return (double)t.tv_sec + 1e-9 * (double)t.tv_nsec;
}
uint64_t *flag_of(ehmm_comm *c, int r) { return reinterpret_cast<uint64_t *>(c->base + (size_t)r * LINE); }
uint8_t *slot_of(ehmm_comm *c, int bank, int r) {
    return c->base + (size_t)c->world * LINE + ((size_t)bank * c->world + r) * c->slot;
}
int wait_all(ehmm_comm *c, uint64_t epoch) {
    const double t0 = now_s();
    for (int r = 0; r < c->world; r++) {
        uint64_t spins = 0;
        while (__atomic_load_n(flag_of(c, r), __ATOMIC_ACQUIRE) < epoch) {
            if ((++spins & 0xfff) == 0) {
                if (now_s() - t0 > c->timeout_s) {
                    ehmm::set_error("ehmm_comm: rank %d waited %.0f s for rank %d at epoch %llu", c->rank, c->timeout_s, r,
                                    (unsigned long long)epoch);
                    return EHMM_ERR_STATE;
                }
                sched_yield();
            }
        }
    }
    return EHMM_OK;
}
}  // namespace

extern "C" {

int ehmm_comm_open(const char *name, int rank, int world, int64_t slot_bytes, double timeout_s, ehmm_comm **out) {
    EHMM_REQUIRE(name && out && world >= 1 && rank >= 0 && rank < world && slot_bytes > 0, EHMM_ERR_ARG, "ehmm_comm_open: bad argument");
    EHMM_REQUIRE(strlen(name) < 90 && name[0] == '/', EHMM_ERR_ARG, "ehmm_comm_open: name must start with '/' and be short");
    ehmm_comm *c = new ehmm_comm();
    c->rank = rank;
    c->world = world;
    c->slot = ((size_t)slot_bytes + LINE - 1) / LINE * LINE;
    c->bytes = (size_t)world * LINE + 2 * (size_t)world * c->slot;
    c->epoch = 0;
    c->timeout_s = timeout_s > 0 ? timeout_s : 120.0;
    strcpy(c->name, name);
    int fd = -1;
    const double t0 = now_s();
    if (rank == 0) {
        shm_unlink(name);  // a stale segment of a crashed job
        fd = shm_open(name, O_CREAT | O_EXCL | O_RDWR, 0600);
        if (fd >= 0 && ftruncate(fd, (off_t)c->bytes) != 0) { close(fd); fd = -1; }
    } else {
        while (true) {  // until rank 0 has created and sized it
            fd = shm_open(name, O_RDWR, 0600);
            struct stat st;
            if (fd >= 0 && fstat(fd, &st) == 0 && (size_t)st.st_size == c->bytes) break;
            if (fd >= 0) close(fd);
            fd = -1;
            if (now_s() - t0 > c->timeout_s) break;
            usleep(1000);
        }
    }
    if (fd < 0) {
        ehmm::set_error("ehmm_comm_open: shm_open(%s) failed on rank %d", name, rank);
        delete c;
        return EHMM_ERR_STATE;
    }
    void *p = mmap(nullptr, c->bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (p == MAP_FAILED) {
        ehmm::set_error("ehmm_comm_open: mmap of %zu bytes failed", c->bytes);
        delete c;
        return EHMM_ERR_NOMEM;
    }
    c->base = static_cast<uint8_t *>(p);  // a fresh segment is zero-filled: every flag starts at epoch 0
    // attach barrier (epoch 1), after which nobody needs the name any more
    c->epoch = 1;
    __atomic_store_n(flag_of(c, rank), c->epoch, __ATOMIC_RELEASE);
    int rc = wait_all(c, c->epoch);
    if (rank == 0) shm_unlink(name);
    if (rc != EHMM_OK) {
        munmap(c->base, c->bytes);
        delete c;
        return rc;
    }
    *out = c;
    return EHMM_OK;
}

/* every rank contributes nbytes (<= slot size); out receives world * nbytes in rank order */
int ehmm_comm_allgather(ehmm_comm *c, const void *in, int64_t nbytes, void *out) {
    EHMM_REQUIRE(c && in && out && nbytes >= 0 && (size_t)nbytes <= c->slot, EHMM_ERR_ARG, "ehmm_comm_allgather: bad argument");
    c->epoch++;
    const int bank = (int)(c->epoch & 1);
    memcpy(slot_of(c, bank, c->rank), in, (size_t)nbytes);
    __atomic_store_n(flag_of(c, c->rank), c->epoch, __ATOMIC_RELEASE);
    EHMM_TRY(wait_all(c, c->epoch));
    for (int r = 0; r < c->world; r++) memcpy(static_cast<uint8_t *>(out) + (size_t)r * nbytes, slot_of(c, bank, r), (size_t)nbytes);
    return EHMM_OK;
}

int ehmm_comm_close(ehmm_comm *c) {
    if (!c) return EHMM_OK;
    munmap(c->base, c->bytes);
    delete c;
    return EHMM_OK;
}

}  // extern "C"
