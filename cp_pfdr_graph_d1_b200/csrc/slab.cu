// slab.cu - host side of the peer-memory link between the ranks of a slab-sharded solve (slab.cuh)
#include "slab.cuh"
#include <unistd.h>
#include <cstring>

namespace pfdr {

namespace {

// every rank announces epoch e in every mailbox, then waits until all ranks have announced it
__global__ void slab_barrier_kernel(SlabDev s, unsigned long long epoch)
{
    if (threadIdx.x == 0) {
        __threadfence_system();
        for (int q = 0; q < s.world; q++) st_release_sys(&s.mail[q]->bar[s.rank], epoch);
    }
    if ((int) threadIdx.x < s.world) slab_wait(&s.mail[s.rank]->bar[threadIdx.x], epoch, s.mail[s.rank]);
}

// vals[0..n) <- sum over the ranks of their vals, added in rank order (same bits on every rank)
__global__ void slab_allreduce_kernel(SlabDev s, double *vals, int n, unsigned long long epoch)
{
    SlabMail *me = s.mail[s.rank];
    const int par = (int) (epoch & 1);
    if (threadIdx.x == 0) {
        __threadfence_system();
        for (int q = 0; q < s.world; q++) {
            for (int i = 0; i < n; i++) s.mail[q]->red[par][s.rank][i] = vals[i];
            st_release_sys(&s.mail[q]->red_flag[s.rank], epoch);
        }
    }
    if ((int) threadIdx.x < s.world) slab_wait(&me->red_flag[threadIdx.x], epoch, me);
    __syncthreads();
    if ((int) threadIdx.x < n) {
        double a = 0.0;
        for (int q = 0; q < s.world; q++) a += me->red[par][q][threadIdx.x];
        vals[threadIdx.x] = a;
    }
}

}  // namespace

void SlabLink::allreduce_sum(double *vals, int n)
{
    if (world == 1) return;
    PFDR_REQUIRE(connected && n <= SLAB_NSUM, "slab: all-reduce before connect");
    red_epoch++;
    slab_allreduce_kernel<<<1, 32, 0, ctx().stream>>>(dev(), vals, n, red_epoch);
    ctx().count_launch();
    PFDR_CUDA(cudaGetLastError());
}

// Allocating, exporting, mapping and freeing a multi-GB arena costs ~150 ms per solver (cudaMalloc + cudaIpcOpenMemHandle
// per peer + cudaFree, measured at 252^3 on 2 GPUs) - as much as the iterations of a short solve.  Callers that solve again
// and again with the same shapes (a benchmark's end-to-end leg, a cut-pursuit re-solving on one graph) get the previous
// arena back: one arena per host thread is kept after its solver is gone, and peers' arenas stay mapped, keyed by their
// IPC handle.  Safe because an arena is only ever written by peers between connect() - which follows the exchange of the
// descriptors, i.e. every rank's commit() and its clearing of the arena - and the end of run(), whose last trip waits for
// the neighbours' last stores (simplex_slab_decide / quad_slab_decide).
namespace {
struct ArenaSlot {
    void *ptr = nullptr; size_t bytes = 0; int device = -1;
};
struct PeerMaps {
    struct Entry { char ipc[64]; void *ptr; long long pid; int device; };
    std::vector<Entry> e;
    void *find(const char *ipc) const { for (const Entry &x : e) if (!memcmp(x.ipc, ipc, 64)) return x.ptr; return nullptr; }
    // a process keeps one arena per GPU: a new handle from the same (process, GPU) means the old arena is gone
    void drop_stale(const char *ipc, long long pid, int device)
    {
        for (size_t i = 0; i < e.size();) {
            if (e[i].pid == pid && e[i].device == device && memcmp(e[i].ipc, ipc, 64)) { cudaIpcCloseMemHandle(e[i].ptr); cudaGetLastError(); e.erase(e.begin() + i); }
            else i++;
        }
    }
    void add(const char *ipc, void *ptr, long long pid, int device)
    {
        if (e.size() >= 2 * SLAB_MAXW) { cudaIpcCloseMemHandle(e.front().ptr); cudaGetLastError(); e.erase(e.begin()); }   // oldest mapping goes
        Entry x; memcpy(x.ipc, ipc, 64); x.ptr = ptr; x.pid = pid; x.device = device; e.push_back(x);
    }
};
ArenaSlot &arena_slot() { static thread_local ArenaSlot s; return s; }
PeerMaps &peer_maps() { static thread_local PeerMaps m; return m; }
}  // namespace

// gives the calling thread's cached arena and peer mappings back (a host thread that ends must call it: nothing is
// released implicitly - no CUDA calls from thread-local destructors, which also run while a process is being torn down)
void slab_release_cache()
{
    ArenaSlot &slot = arena_slot();
    if (slot.ptr) { cudaFree(slot.ptr); slot.ptr = nullptr; slot.bytes = 0; }
    PeerMaps &m = peer_maps();
    for (const PeerMaps::Entry &x : m.e) cudaIpcCloseMemHandle(x.ptr);
    m.e.clear();
    cudaGetLastError();
}

SlabLink::~SlabLink()
{
    if (arena) {
        ArenaSlot &slot = arena_slot();
        if (slot.ptr) cudaFree(slot.ptr);
        slot.ptr = arena; slot.bytes = arena_bytes; slot.device = mine.device;
    }
    cudaGetLastError();
}

void SlabLink::reserve(int which, size_t bytes)
{
    PFDR_REQUIRE(which > 0 && which < SLAB_NOFF && !arena, "slab arena: bad reservation");
    if (arena_bytes == 0) arena_bytes = (sizeof(SlabMail) + 255) & ~(size_t) 255;      // the mailbox comes first (offset 0)
    mine.off[which] = arena_bytes;
    arena_bytes += (bytes + 255) & ~(size_t) 255;
}

void SlabLink::commit(int world_, int rank_, long long lo_, long long hi_, long long v0_, long long v1_, long long width_)
{
    PFDR_REQUIRE(world_ >= 1 && world_ <= SLAB_MAXW && rank_ >= 0 && rank_ < world_, "slab: at most 16 ranks");
    world = world_; rank = rank_; lo = lo_; hi = hi_; v0 = v0_; v1 = v1_; width = width_;
    nb_lo = rank > 0 ? rank - 1 : -1;
    nb_hi = rank + 1 < world ? rank + 1 : -1;
    if (arena_bytes == 0) arena_bytes = (sizeof(SlabMail) + 255) & ~(size_t) 255;
    PFDR_CUDA(cudaGetDevice(&mine.device));
    // peer-visible memory must come from cudaMalloc (the stream-ordered pool cannot be exported)
    ArenaSlot &slot = arena_slot();
    if (slot.ptr && slot.bytes == arena_bytes && slot.device == mine.device) { arena = slot.ptr; slot.ptr = nullptr; }
    else {
        if (slot.ptr) { cudaFree(slot.ptr); slot.ptr = nullptr; }
        PFDR_CUDA(cudaMalloc(&arena, arena_bytes));
    }
    PFDR_CUDA(cudaMemsetAsync(arena, 0, arena_bytes, ctx().stream));
    PFDR_CUDA(cudaStreamSynchronize(ctx().stream));
    mine.off[SLAB_MAILBOX] = 0;
    mine.base = (unsigned long long) (uintptr_t) arena;
    mine.pid = (long long) getpid();
    mine.rank = rank; mine.lo = lo; mine.hi = hi; mine.v0 = v0; mine.v1 = v1; mine.width = width; mine.bytes = arena_bytes;
    if (world > 1) {
        cudaIpcMemHandle_t h;
        PFDR_CUDA(cudaIpcGetMemHandle(&h, arena));
        static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t");
        memcpy(mine.ipc, &h, 64);
    }
    blobs.assign(world, SlabBlob{});
    base.assign(world, nullptr);
    opened.assign(world, false);
    blobs[rank] = mine; base[rank] = arena;
    connected = world == 1;
}

void SlabLink::connect(const SlabBlob *all, bool same_process)
{
    PFDR_REQUIRE(arena != nullptr, "slab: connect before commit");
    int dev = 0;
    PFDR_CUDA(cudaGetDevice(&dev));
    for (int q = 0; q < world; q++) {
        if (q == rank) continue;
        blobs[q] = all[q];
        PFDR_REQUIRE(blobs[q].rank == q && blobs[q].width == width, "slab: inconsistent rank table");
        if (same_process || blobs[q].pid == mine.pid) {
            if (blobs[q].device != dev) {
                int can = 0;
                PFDR_CUDA(cudaDeviceCanAccessPeer(&can, dev, blobs[q].device));
                PFDR_REQUIRE(can, "slab: the GPUs of two ranks cannot access each other's memory");
                const cudaError_t e = cudaDeviceEnablePeerAccess(blobs[q].device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) PFDR_CUDA(e);
                cudaGetLastError();
            }
            base[q] = (void *) (uintptr_t) blobs[q].base;
        } else {
            const char *ipc = (const char *) blobs[q].ipc;
            peer_maps().drop_stale(ipc, blobs[q].pid, blobs[q].device);
            void *p = peer_maps().find(ipc);                     // still mapped from an earlier solver of this thread
            if (!p) {
                cudaIpcMemHandle_t h;
                memcpy(&h, ipc, 64);
                PFDR_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
                peer_maps().add(ipc, p, blobs[q].pid, blobs[q].device);
            }
            base[q] = p; opened[q] = true;
        }
    }
    if (nb_lo >= 0) PFDR_REQUIRE(blobs[nb_lo].v1 == v0 && blobs[nb_lo].v1 - blobs[nb_lo].v0 >= v0 - lo, "slab: ranges of neighbouring ranks do not match (or a slab is thinner than the band)");
    if (nb_hi >= 0) PFDR_REQUIRE(blobs[nb_hi].v0 == v1 && blobs[nb_hi].v1 - blobs[nb_hi].v0 >= hi - v1, "slab: ranges of neighbouring ranks do not match (or a slab is thinner than the band)");
    connected = true;
}

SlabDev SlabLink::dev() const
{
    SlabDev d{};
    d.on = world > 1; d.world = world; d.rank = rank; d.nb_lo = nb_lo; d.nb_hi = nb_hi; d.run_id = run_id;
    for (int q = 0; q < world; q++) d.mail[q] = (SlabMail *) base[q];
    return d;
}

void SlabLink::barrier()
{
    if (world == 1) return;
    PFDR_REQUIRE(connected, "slab: not connected");
    bar_epoch++;
    slab_barrier_kernel<<<1, 32, 0, ctx().stream>>>(dev(), bar_epoch);
    ctx().count_launch();
    PFDR_CUDA(cudaGetLastError());
}

template <typename T>
void SlabLink::push(int which, int r, size_t from_elem, size_t count, long long shift_elems)
{
    if (r < 0 || !count) return;
    T *dst = ptr<T>(which, r);
    const T *src = ptr<T>(which);
    PFDR_REQUIRE(dst && src, "slab: array not present in the arena");
    PFDR_CUDA(cudaMemcpyAsync(dst + (long long) from_elem + shift_elems, src + from_elem, count * sizeof(T), cudaMemcpyDeviceToDevice, ctx().stream));
}
template void SlabLink::push<float>(int, int, size_t, size_t, long long);
template void SlabLink::push<double>(int, int, size_t, size_t, long long);

}  // namespace pfdr
