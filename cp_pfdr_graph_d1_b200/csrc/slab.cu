// slab.cu - host side of the peer-memory link between the ranks of a slab-sharded solve (slab.cuh)
#include "slab.cuh"
#include <unistd.h>

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

SlabLink::~SlabLink()
{
    for (size_t q = 0; q < base.size(); q++)
        if (base[q] && opened[q]) cudaIpcCloseMemHandle(base[q]);
    if (arena) cudaFree(arena);
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
    // peer-visible memory must come from cudaMalloc (the stream-ordered pool cannot be exported)
    PFDR_CUDA(cudaMalloc(&arena, arena_bytes));
    PFDR_CUDA(cudaMemsetAsync(arena, 0, arena_bytes, ctx().stream));
    PFDR_CUDA(cudaStreamSynchronize(ctx().stream));
    mine.off[SLAB_MAILBOX] = 0;
    mine.base = (unsigned long long) (uintptr_t) arena;
    mine.pid = (long long) getpid();
    PFDR_CUDA(cudaGetDevice(&mine.device));
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
            cudaIpcMemHandle_t h;
            memcpy(&h, blobs[q].ipc, 64);
            void *p = nullptr;
            PFDR_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
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
