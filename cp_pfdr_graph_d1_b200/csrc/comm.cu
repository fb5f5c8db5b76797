// comm.cu - see comm.cuh
#include "comm.cuh"
#include <dlfcn.h>
#include <nccl.h>

namespace pfdr {

namespace {

struct NcclApi {
    void *h = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi &api()
{
    static NcclApi a;
    if (!a.h) {
        a.h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!a.h) a.h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!a.h) throw Error(PFDR_B200_ERR_NCCL, std::string("cannot open libnccl.so.2: ") + dlerror());
#define LOAD(name) a.name = reinterpret_cast<decltype(a.name)>(dlsym(a.h, "nccl" #name)); \
        if (!a.name) throw Error(PFDR_B200_ERR_NCCL, "libnccl lacks nccl" #name)
        LOAD(GetUniqueId); LOAD(CommInitRank); LOAD(CommDestroy); LOAD(GroupStart); LOAD(GroupEnd);
        LOAD(Send); LOAD(Recv); LOAD(AllReduce); LOAD(GetErrorString);
#undef LOAD
    }
    return a;
}

#define PFDR_NCCL(call)                                                                               \
    do {                                                                                              \
        ncclResult_t r__ = (call);                                                                    \
        if (r__ != ncclSuccess)                                                                       \
            throw Error(PFDR_B200_ERR_NCCL, std::string(#call) + " failed: " + api().GetErrorString(r__)); \
    } while (0)

template <typename real> struct NcclType;
template <> struct NcclType<float> { static constexpr ncclDataType_t v = ncclFloat; };
template <> struct NcclType<double> { static constexpr ncclDataType_t v = ncclDouble; };

__device__ __forceinline__ int halo_parity(const Ctl *ctl, int mode)
{
    if (mode == PARITY_ZERO || !ctl) return 0;
    const int it = ctl->it;
    if (mode == PARITY_NEXT_IF_RAN && ctl->hold == 0 && ctl->state == ST_RUN) return (it + 1) & 1;
    return it & 1;
}

// pack:   out[seg.off + i*width + k] = A[idx[i]*width + k]      unpack: A[idx[i]*width + k] = in[seg.off + i*width + k]
template <typename real, bool PACK>
__global__ void halo_kernel(real *buf, PingPong<real> va0, PingPong<real> va1, PingPong<real> z, const Ctl *ctl,
                            int parity_mode, const HaloSeg *__restrict__ segs, int nsegs, long long total, int width)
{
    const long long t = (long long) blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total * width) return;
    const long long item = t / width;
    const int k = (int) (t - item * width);
    int sidx = 0;
    while (sidx + 1 < nsegs && item >= segs[sidx + 1].start) sidx++;
    const HaloSeg sg = segs[sidx];
    const int par = halo_parity(ctl, parity_mode);
    real *arr = sg.arr == 0 ? va0.a[par] : (sg.arr == 1 ? va1.a[par] : z.a[par]);
    const long long i = item - sg.start;
    const size_t a = (size_t) sg.idx[i] * width + k, b = (size_t) sg.off + (size_t) i * width + k;
    if (PACK) buf[b] = arr[a]; else arr[a] = buf[b];
}
__global__ void key_to_slot_kernel(const int *__restrict__ hkey, long long nslots, int *__restrict__ pos)
{
    const long long s = (long long) blockIdx.x * blockDim.x + threadIdx.x;
    if (s < nslots && hkey[s] >= 0) pos[hkey[s]] = (int) s;
}
__global__ void map_keys_kernel(const int *__restrict__ keys, const int *__restrict__ pos, size_t n, int *__restrict__ out)
{
    const size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = pos[keys[i]];
}

}  // namespace

Comm &comm()
{
    static Comm c;
    return c;
}

void comm_unique_id(void *out128)
{
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is expected to be 128 bytes");
    ncclUniqueId id;
    PFDR_NCCL(api().GetUniqueId(&id));
    memcpy(out128, &id, sizeof(id));
}

void comm_init(const void *id128, int world, int rank)
{
    comm_destroy();
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    ncclComm_t c;
    PFDR_NCCL(api().CommInitRank(&c, world, id, rank));
    comm().nccl = c; comm().world = world; comm().rank = rank;
}

void comm_destroy()
{
    if (comm().nccl) { api().CommDestroy((ncclComm_t) comm().nccl); comm().nccl = nullptr; comm().world = 1; comm().rank = 0; }
}

static ncclComm_t need_comm()
{
    if (!comm().nccl) throw Error(PFDR_B200_ERR_NCCL, "sharded solve without a communicator: call pfdr_b200_comm_init first");
    return (ncclComm_t) comm().nccl;
}

void comm_allreduce_sum(double *buf, int count)
{
    PFDR_NCCL(api().AllReduce(buf, buf, count, ncclDouble, ncclSum, need_comm(), ctx().stream));
    ctx().count_launch();
}
void comm_group_start() { PFDR_NCCL(api().GroupStart()); }
void comm_group_end() { PFDR_NCCL(api().GroupEnd()); ctx().count_launch(); }
template <typename real> void comm_send(const real *buf, size_t count, int peer)
{
    PFDR_NCCL(api().Send(buf, count, NcclType<real>::v, peer, need_comm(), ctx().stream));
}
template <typename real> void comm_recv(real *buf, size_t count, int peer)
{
    PFDR_NCCL(api().Recv(buf, count, NcclType<real>::v, peer, need_comm(), ctx().stream));
}
void comm_send_f64(const double *buf, size_t count, int peer) { comm_send<double>(buf, count, peer); }
void comm_recv_f64(double *buf, size_t count, int peer) { comm_recv<double>(buf, count, peer); }
template void comm_send<float>(const float *, size_t, int);
template void comm_send<double>(const double *, size_t, int);
template void comm_recv<float>(float *, size_t, int);
template void comm_recv<double>(double *, size_t, int);

template <typename real>
void Halo<real>::setup(const ShardInfo &s, const int *d_hkey, long long nslots, size_t nkeys, int width_)
{
    cudaStream_t st = ctx().stream;
    nnb = s.nnb; width = width_;
    rank.assign(s.nb_rank, s.nb_rank + nnb);
    nsv.resize(nnb); nsh.resize(nnb); nrv.resize(nnb); nrh.resize(nnb);
    sv_off.resize(nnb); sh_off.resize(nnb); rv_off.resize(nnb); rh_off.resize(nnb);
    for (int i = 0; i < nnb; i++) {
        nsv[i] = s.sv_off[i + 1] - s.sv_off[i]; nsh[i] = s.sh_off[i + 1] - s.sh_off[i];
        nrv[i] = s.rv_off[i + 1] - s.rv_off[i]; nrh[i] = s.rh_off[i + 1] - s.rh_off[i];
        sv_off[i] = s.sv_off[i]; sh_off[i] = s.sh_off[i]; rv_off[i] = s.rv_off[i]; rh_off[i] = s.rh_off[i];
    }
    const size_t tsv = nnb ? s.sv_off[nnb] : 0, tsh = nnb ? s.sh_off[nnb] : 0, trv = nnb ? s.rv_off[nnb] : 0,
                 trh = nnb ? s.rh_off[nnb] : 0;
    auto up = [&](DevBuf<int> &d, const int *h, size_t n) {
        d.alloc(std::max<size_t>(n, 1));
        if (n) PFDR_CUDA(cudaMemcpyAsync(d.p, h, n * sizeof(int), cudaMemcpyHostToDevice, st));
    };
    DevBuf<int> d_sh, d_rh, pos(std::max<size_t>(nkeys, 1));
    up(d_sv, s.sv, tsv); up(d_rv, s.rv, trv); up(d_sh, s.sh, tsh); up(d_rh, s.rh, trh);
    d_ss.alloc(std::max<size_t>(tsh, 1)); d_rs.alloc(std::max<size_t>(trh, 1));
    if (nslots > 0) PFDR_LAUNCH(key_to_slot_kernel, cdiv((size_t) nslots, 256), 256, 0, d_hkey, nslots, pos.p);
    if (tsh) PFDR_LAUNCH(map_keys_kernel, cdiv(tsh, 256), 256, 0, d_sh.p, pos.p, tsh, d_ss.p);
    if (trh) PFDR_LAUNCH(map_keys_kernel, cdiv(trh, 256), 256, 0, d_rh.p, pos.p, trh, d_rs.p);
    sendbuf.alloc(std::max<size_t>((2 * tsv + tsh) * width, 1));
    recvbuf.alloc(std::max<size_t>((2 * trv + trh) * width, 1));
    PFDR_CUDA(cudaStreamSynchronize(st));
}

template <typename real>
void Halo<real>::build(int mask)
{
    Layout &L = layouts[mask];
    cudaStream_t st = ctx().stream;
    std::vector<HaloSeg> ps, us;
    L.soff.assign(nnb, 0); L.roff.assign(nnb, 0); L.scnt.assign(nnb, 0); L.rcnt.assign(nnb, 0);
    long long pstart = 0, ustart = 0;
    size_t so = 0, ro = 0;
    for (int i = 0; i < nnb; i++) {
        L.soff[i] = so; L.roff[i] = ro;
        for (int a = 0; a < 2; a++) {
            if (!(mask & (a == 0 ? HALO_VA0 : HALO_VA1))) continue;
            if (nsv[i]) { ps.push_back({d_sv.p + sv_off[i], pstart, nsv[i], (long long) so, a}); pstart += nsv[i]; so += (size_t) nsv[i] * width; }
            if (nrv[i]) { us.push_back({d_rv.p + rv_off[i], ustart, nrv[i], (long long) ro, a}); ustart += nrv[i]; ro += (size_t) nrv[i] * width; }
        }
        if (mask & HALO_Z) {
            if (nsh[i]) { ps.push_back({d_ss.p + sh_off[i], pstart, nsh[i], (long long) so, 2}); pstart += nsh[i]; so += (size_t) nsh[i] * width; }
            if (nrh[i]) { us.push_back({d_rs.p + rh_off[i], ustart, nrh[i], (long long) ro, 2}); ustart += nrh[i]; ro += (size_t) nrh[i] * width; }
        }
        L.scnt[i] = so - L.soff[i]; L.rcnt[i] = ro - L.roff[i];
    }
    L.npseg = (int) ps.size(); L.nuseg = (int) us.size();
    L.pack_items = pstart; L.unpack_items = ustart;
    L.d_pack.alloc(std::max<size_t>(ps.size(), 1)); L.d_unpack.alloc(std::max<size_t>(us.size(), 1));
    if (!ps.empty()) PFDR_CUDA(cudaMemcpyAsync(L.d_pack.p, ps.data(), ps.size() * sizeof(HaloSeg), cudaMemcpyHostToDevice, st));
    if (!us.empty()) PFDR_CUDA(cudaMemcpyAsync(L.d_unpack.p, us.data(), us.size() * sizeof(HaloSeg), cudaMemcpyHostToDevice, st));
    PFDR_CUDA(cudaStreamSynchronize(st));      // ps / us die at the end of this scope
    L.built = true;
}

template <typename real>
void Halo<real>::exchange(int mask, PingPong<real> va0, PingPong<real> va1, PingPong<real> Z, const Ctl *ctl, int parity,
                          double *allsums, int nvals)
{
    const int world = comm().world, me = comm().rank;
    if (!nnb) {
        if (allsums && world > 1) {
            comm_group_start();
            for (int r = 0; r < world; r++) if (r != me) {
                comm_send_f64(allsums + (size_t) me * nvals, nvals, r);
                comm_recv_f64(allsums + (size_t) r * nvals, nvals, r);
            }
            comm_group_end();
        }
        return;
    }
    Layout &L = layouts[mask];
    if (!L.built) throw Error(PFDR_B200_ERR_ARG, "halo layout used before build()");
    if (L.pack_items)
        PFDR_LAUNCH((halo_kernel<real, true>), cdiv((size_t) L.pack_items * width, 256), 256, 0, sendbuf.p, va0, va1, Z,
                    ctl, parity, L.d_pack.p, L.npseg, L.pack_items, width);
    comm_group_start();
    for (int i = 0; i < nnb; i++) {
        if (L.scnt[i]) comm_send<real>(sendbuf.p + L.soff[i], L.scnt[i], rank[i]);
        if (L.rcnt[i]) comm_recv<real>(recvbuf.p + L.roff[i], L.rcnt[i], rank[i]);
    }
    if (allsums) {
        for (int r = 0; r < world; r++) if (r != me) {
            comm_send_f64(allsums + (size_t) me * nvals, nvals, r);
            comm_recv_f64(allsums + (size_t) r * nvals, nvals, r);
        }
    }
    comm_group_end();
    if (L.unpack_items)
        PFDR_LAUNCH((halo_kernel<real, false>), cdiv((size_t) L.unpack_items * width, 256), 256, 0, recvbuf.p, va0, va1, Z,
                    ctl, parity, L.d_unpack.p, L.nuseg, L.unpack_items, width);
}

template struct Halo<float>;
template struct Halo<double>;

}  // namespace pfdr
