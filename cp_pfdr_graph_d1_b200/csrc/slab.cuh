// slab.cuh - vertex-sharded ("slab") solves of constant-offset graphs over several B200s WITHOUT a collective library on
// the data path.
//
// Rank r owns the contiguous vertex range [v0, v1) of a stencil graph (dia.cuh) - z-slabs of a grid - and holds every
// vertex array over [lo, hi) = [v0 - band, v1 + band) clipped to the graph: owned vertices plus the ghost zones its cut
// edges reach into.  What a neighbour needs of an iteration - the new forward point of the boundary vertices and the
// new copies on the cut edges - is STORED BY THE ITERATION KERNEL ITSELF into the neighbour's ghost zone through
// peer-mapped pointers (NVLink / NVSwitch; cudaIpc handles between processes, plain peer access inside one process):
// no pack / unpack kernels, no send / receive.  Synchronisation is one 64-bit flag per rank pair and iteration:
//   * a rank works on its boundary vertices (within `band` of a neighbour) FIRST; the warp that completes the last
//     boundary chunk, after a system-scope fence, raises the flag "boundary part of iteration k done" in its two
//     neighbours' mailboxes: its stores into their ghost zones have landed, and it no longer reads its own ghost
//     zones - then the rank goes on with its interior, which touches no ghost data, while the flag travels;
//   * the next kernel of a rank starts by waiting (acquire loads on its own mailbox) for the boundary flags of its
//     neighbours - normally raised long before;
//   * when the evolution is monitored, the last CTA also writes the rank's stopping sums and "iteration k completed"
//     into the mailbox of EVERY rank, and the next kernel waits for all of them, adds the sums in rank order and
//     takes the stop / recondition decision: every rank takes the same one at the same iteration, `*it` and Dif[]
//     are exact.
// Buffers are ping/pong by iteration parity, so a rank may run at most one iteration ahead of a neighbour.
// Kernels of different ranks wait on one another: every rank must be on its own GPU (never two ranks on one device).
#pragma once
#include "common.cuh"

namespace pfdr {

constexpr int SLAB_MAXW = 16;
constexpr int SLAB_NSUM = 4;

struct SlabMail {                                   // one per rank, inside its peer-visible arena; slot [q] is written by rank q
    unsigned long long sum_flag[SLAB_MAXW];         // (run id << 32) | iterations completed by rank q
    unsigned long long halo_flag[SLAB_MAXW];        // (run id << 32) | iterations whose BOUNDARY part rank q has completed:
                                                    //   its stores into our ghost zones have landed, it no longer reads its own
    double sums[2][SLAB_MAXW][SLAB_NSUM];           // rank q's stopping sums of iteration k in slot [k & 1] (+ objective part):
                                                    //   q may finish k+1 before a slow reader has looked at k, never k+2
    unsigned long long bar[SLAB_MAXW];              // barrier epochs
    unsigned long long ax_flag[SLAB_MAXW];          // dense A: epoch of rank q's partial products
    unsigned long long red_flag[SLAB_MAXW];         // small all-reduce (preconditioner sums): epoch of rank q's contribution
    double red[2][SLAB_MAXW][SLAB_NSUM];            // ... by epoch parity
    int error;                                      // a wait gave up (peer missing): the solve is aborted
    int pad;
};

template <typename real>
struct SlabPeer {                                   // a neighbour's arrays as seen from this rank (NULL: no such neighbour)
    real *Z[2], *F[2], *X[2];                       // ping/pong: edge copies, forward point, iterate
    size_t plane;                                   // elements between two planes of Z over there
    long long shift;                                // element index over there = element index here + shift
};

struct SlabDev {                                    // kernel-side view of the link
    int on, world, rank, nb_lo, nb_hi, band;
    int lowB, highB;                                // owned vertices within `band` of the rank below / above (0: no such rank)
    int nbchunks;                                   // boundary chunks of this rank per iteration (they come first in its work order)
    unsigned long long run_id;
    SlabMail *mail[SLAB_MAXW];                      // every rank's mailbox (mail[rank] is local memory)
};

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long slab_now()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// spins until *flag >= target; gives up after ~20 s (a rank died or was never launched) and reports it
__device__ __forceinline__ bool slab_wait(const unsigned long long *flag, unsigned long long target, SlabMail *mine)
{
    if (ld_acquire_sys(flag) >= target) return true;
    const unsigned long long t0 = slab_now();
    for (;;) {
        if (ld_acquire_sys(flag) >= target) return true;
        if (*((volatile int *) &mine->error)) return false;
        if (slab_now() - t0 > 20000000000ull) { mine->error = 1; return false; }
        __nanosleep(64);
    }
}
// before a kernel of trip `it` reads ghost-zone data: the neighbours' stores of trip it - 1 have landed (one thread per CTA calls)
__device__ __forceinline__ bool slab_wait_neighbours(const SlabDev &s, int it)
{
    if (it == 0) return true;
    SlabMail *me = s.mail[s.rank];
    const unsigned long long target = (s.run_id << 32) | (unsigned long long) it;
    bool ok = true;
    if (s.nb_lo >= 0) ok = slab_wait(&me->halo_flag[s.nb_lo], target, me) && ok;
    if (s.nb_hi >= 0) ok = slab_wait(&me->halo_flag[s.nb_hi], target, me) && ok;
    return ok;
}
// rank's sums and its "iteration `done` completed" flag into every mailbox (one thread).  One fence, then relaxed flag
// stores: a release store per destination would wait for the NVLink round trip of each in turn.
__device__ __forceinline__ void slab_post(const SlabDev &s, const double (&v)[SLAB_NSUM], unsigned long long done)
{
    for (int q = 0; q < s.world; q++) {
#pragma unroll
        for (int i = 0; i < SLAB_NSUM; i++) s.mail[q]->sums[(done - 1) & 1][s.rank][i] = v[i];
    }
    __threadfence_system();
    for (int q = 0; q < s.world; q++) *((volatile unsigned long long *) &s.mail[q]->sum_flag[s.rank]) = (s.run_id << 32) | done;
}
// a warp has finished `count` boundary chunks of trip `it` (its lanes' stores into the neighbours' ghost zones are issued):
// the warp completing the rank's last one raises the boundary flag at both neighbours.  bdone counts chunks since reset().
__device__ __forceinline__ void slab_boundary_done(const SlabDev &s, int *bdone, int count, int it)
{
    __threadfence_system();                                      // every lane: its peer stores, before the count moves
    __syncwarp();
    if ((threadIdx.x & 31) == 0) {
        const unsigned target = ((unsigned) it + 1u) * (unsigned) s.nbchunks;                  // modulo 2^32, as the counter
        if (atomicAdd((unsigned *) bdone, (unsigned) count) + (unsigned) count == target) {
            __threadfence_system();
            const unsigned long long f = (s.run_id << 32) | (unsigned long long) (it + 1);
            if (s.nb_lo >= 0) *((volatile unsigned long long *) &s.mail[s.nb_lo]->halo_flag[s.rank]) = f;
            if (s.nb_hi >= 0) *((volatile unsigned long long *) &s.mail[s.nb_hi]->halo_flag[s.rank]) = f;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
constexpr int SLAB_NOFF = 12;
struct SlabBlob {                                   // what a rank publishes (pfdr_b200_slab_export): fixed size, plain data
    unsigned char ipc[64];                          // cudaIpcMemHandle_t of the arena
    unsigned long long base;                        // the arena's address in the exporting process
    long long pid;
    int device, rank;
    long long lo, hi, v0, v1;                       // global vertex ranges: held [lo, hi), owned [v0, v1)
    long long width;                                // reals per vertex (K of the simplex family, 1 otherwise)
    unsigned long long off[SLAB_NOFF];              // byte offsets of the arrays inside the arena (0: absent)
    unsigned long long bytes;
};
static_assert(sizeof(SlabBlob) <= 256, "blob size is part of the C ABI (PFDR_B200_SLAB_BLOB)");

enum : int { SLAB_MAILBOX = 0, SLAB_Z0 = 1, SLAB_Z1 = 2, SLAB_F0 = 3, SLAB_F1 = 4, SLAB_T = 5, SLAB_X0 = 6, SLAB_X1 = 7, SLAB_AX = 8 };

struct SlabLink {
    int world = 1, rank = 0, nb_lo = -1, nb_hi = -1;
    long long lo = 0, hi = 0, v0 = 0, v1 = 0, width = 1;
    void *arena = nullptr;
    size_t arena_bytes = 0;
    SlabBlob mine{};
    std::vector<SlabBlob> blobs;                    // every rank's, after connect()
    std::vector<void *> base;                       // every rank's arena as mapped here (NULL: not mapped)
    std::vector<bool> opened;                       // mapped through cudaIpcOpenMemHandle (kept in the thread's mapping cache, slab.cu)
    unsigned long long bar_epoch = 0, red_epoch = 0, run_id = 0;
    bool connected = false;

    ~SlabLink();
    // reserves `bytes` (256-byte aligned) for array `which`; call before commit()
    void reserve(int which, size_t bytes);
    void commit(int world_, int rank_, long long lo_, long long hi_, long long v0_, long long v1_, long long width_);   // allocates + zeroes the arena
    template <typename T> T *ptr(int which, int r = -1) const
    {
        const int q = r < 0 ? rank : r;
        const SlabBlob &b = r < 0 ? mine : blobs[q];
        void *a = r < 0 ? arena : base[q];
        return (a && (which == SLAB_MAILBOX || b.off[which])) ? (T *) ((char *) a + b.off[which]) : nullptr;
    }
    void connect(const SlabBlob *all, bool same_process);
    SlabDev dev() const;
    void barrier();                                 // all ranks: everything enqueued before it on every rank has completed
    void allreduce_sum(double *vals, int n);        // vals[0..n) <- sum over the ranks (rank order), n <= SLAB_NSUM; stream ordered
    // copies `count` elements starting at local element `from` of array `which` into the same global position of rank r's copy
    template <typename T> void push(int which, int r, size_t from_elem, size_t count, long long shift_elems);
};

void slab_release_cache();                          // the calling thread's cached arena / peer mappings (slab.cu)

}  // namespace pfdr
