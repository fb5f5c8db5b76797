// comm.cuh - NCCL plumbing of the sharded (multi-GPU) solvers.  NCCL is opened at run time (dlopen of
// libnccl.so.2: inside a torch process that is torch's own copy) so that single-GPU users of
// libpfdr_b200.so - the MEX / Boost-Python gateways - carry no NCCL dependency.
#pragma once
#include "common.cuh"

namespace pfdr {

// description of this rank's shard = pfdr_b200_shard of include/pfdr_b200.h (mirrors ShardPlan / Neighbor of
// cp_pfdr_graph_d1_b200/shard.py): all index arrays are HOST arrays; *_off have nnb+1 entries
typedef pfdr_b200_shard ShardInfo;

struct Comm {
    void *nccl = nullptr;       // ncclComm_t
    int world = 1, rank = 0;
};
Comm &comm();                    // process-wide communicator (one solve at a time per process)

void comm_unique_id(void *out128);
void comm_init(const void *id128, int world, int rank);
void comm_destroy();

// stream-ordered collectives on ctx().stream
void comm_allreduce_sum(double *buf, int count);
void comm_group_start();
void comm_group_end();
void comm_send_f64(const double *buf, size_t count, int peer);
void comm_recv_f64(double *buf, size_t count, int peer);
template <typename real> void comm_send(const real *buf, size_t count, int peer);
template <typename real> void comm_recv(real *buf, size_t count, int peer);

// a ping/pong pair of device arrays; kernels pick a[it & 1] from the control block (a[1] == a[0] for constants)
template <typename real> struct PingPong {
    real *a[2];
    PingPong() { a[0] = a[1] = nullptr; }
    PingPong(real *x) { a[0] = a[1] = x; }
    PingPong(real *x, real *y) { a[0] = x; a[1] = y; }
};

// Halo engine shared by both solver families: per neighbour, a message is
//   [ va0[send vertices] | va1[send vertices] (optional) | Z[send slots] (optional) ]   (each item `width` reals wide)
// One pack kernel, one grouped ncclSend/ncclRecv, one unpack kernel per exchange.
struct HaloSeg { const int *idx; long long start, n, off; int arr; };   // arr: 0 = va0, 1 = va1, 2 = Z

enum : int { PARITY_ZERO = 0, PARITY_CUR = 1, PARITY_NEXT_IF_RAN = 2 };

enum : int { HALO_VA0 = 1, HALO_VA1 = 2, HALO_Z = 4 };     // what a message carries

template <typename real>
struct Halo {
    int nnb = 0, width = 1;
    std::vector<int> rank, nsv, nsh, nrv, nrh;
    std::vector<size_t> sv_off, sh_off, rv_off, rh_off;
    DevBuf<int> d_sv, d_ss, d_rv, d_rs;           // send vertices / send slots / recv vertices / recv slots
    DevBuf<real> sendbuf, recvbuf;
    struct Layout {                                // segment tables of one message composition (built on first use)
        bool built = false;
        DevBuf<HaloSeg> d_pack, d_unpack;
        int npseg = 0, nuseg = 0;
        long long pack_items = 0, unpack_items = 0;
        std::vector<size_t> soff, roff, scnt, rcnt;
    } layouts[8];

    // hkey: [nslots] half-edge key of every slot (-1 padding); nkeys = 2 * local edges
    void setup(const ShardInfo &s, const int *d_hkey, long long nslots, size_t nkeys, int width_);
    void build(int mask);
    // mask = HALO_VA0 | HALO_VA1 | HALO_Z: which of the arrays travel.  parity selects which of the ping/pong
    // buffers is used, decided on the device from the control block (no host synchronisation)
    // allsums != NULL: the same NCCL group also carries this rank's `nvals` reduction scalars (allsums[rank*nvals..])
    // to every other rank and receives theirs into allsums[r*nvals..] - an all-gather by point-to-point messages, so
    // that an iteration costs one NCCL launch instead of two; every rank then adds the partials in rank order.
    void exchange(int mask, PingPong<real> va0, PingPong<real> va1, PingPong<real> Z, const Ctl *ctl, int parity,
                  double *allsums = nullptr, int nvals = 0);
};

}  // namespace pfdr
