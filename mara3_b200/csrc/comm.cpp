#include "comm.hpp"
#include <dlfcn.h>
#include <stdexcept>
#include <cuda_runtime.h>
#include <nccl.h>

using namespace m3b;

namespace
{
    struct nccl_api_t
    {
        ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
        ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
        ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
        ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
        ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
        ncclResult_t (*GroupStart)() = nullptr;
        ncclResult_t (*GroupEnd)() = nullptr;
        ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
        const char* (*GetErrorString)(ncclResult_t) = nullptr;
    };

    const nccl_api_t& nccl()
    {
        static nccl_api_t api = []
        {
            void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);       // the copy torch loaded, if any
            if (! h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
            if (! h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
            if (! h) throw std::runtime_error(std::string("mara3_b200: cannot load NCCL (libnccl.so.2): ") + dlerror());
            nccl_api_t a;
            auto sym = [h] (const char* name)
            {
                void* p = dlsym(h, name);
                if (! p) throw std::runtime_error(std::string("mara3_b200: NCCL symbol missing: ") + name);
                return p;
            };
            a.GetUniqueId    = reinterpret_cast<decltype(a.GetUniqueId)>(sym("ncclGetUniqueId"));
            a.CommInitRank   = reinterpret_cast<decltype(a.CommInitRank)>(sym("ncclCommInitRank"));
            a.CommDestroy    = reinterpret_cast<decltype(a.CommDestroy)>(sym("ncclCommDestroy"));
            a.Send           = reinterpret_cast<decltype(a.Send)>(sym("ncclSend"));
            a.Recv           = reinterpret_cast<decltype(a.Recv)>(sym("ncclRecv"));
            a.GroupStart     = reinterpret_cast<decltype(a.GroupStart)>(sym("ncclGroupStart"));
            a.GroupEnd       = reinterpret_cast<decltype(a.GroupEnd)>(sym("ncclGroupEnd"));
            a.AllGather      = reinterpret_cast<decltype(a.AllGather)>(sym("ncclAllGather"));
            a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(sym("ncclGetErrorString"));
            return a;
        }();
        return api;
    }

    void check(ncclResult_t r, const char* what)
    {
        if (r != ncclSuccess) throw std::runtime_error(std::string("NCCL error in ") + what + ": " + nccl().GetErrorString(r));
    }
}

void communicator_t::make_unique_id(unsigned char* out)
{
    static_assert(sizeof(ncclUniqueId) == nccl_unique_id_bytes, "ncclUniqueId size");
    ncclUniqueId id;
    check(nccl().GetUniqueId(&id), "ncclGetUniqueId");
    for (int k = 0; k < nccl_unique_id_bytes; ++k) out[k] = static_cast<unsigned char>(id.internal[k]);
}

communicator_t::communicator_t(int rank, int nranks, const unsigned char* unique_id) : rank_(rank), nranks_(nranks)
{
    ncclUniqueId id;
    for (int k = 0; k < nccl_unique_id_bytes; ++k) id.internal[k] = static_cast<char>(unique_id[k]);
    ncclComm_t c;
    check(nccl().CommInitRank(&c, nranks, id, rank), "ncclCommInitRank");
    comm = c;
}

communicator_t::~communicator_t()
{
    if (comm) nccl().CommDestroy(static_cast<ncclComm_t>(comm));
}

void communicator_t::exchange(const std::vector<const double*>& send, const std::vector<std::size_t>& send_count,
                              const std::vector<double*>& recv, const std::vector<std::size_t>& recv_count, void* cuda_stream)
{
    auto c = static_cast<ncclComm_t>(comm);
    auto s = static_cast<cudaStream_t>(cuda_stream);
    check(nccl().GroupStart(), "ncclGroupStart");
    for (int p = 0; p < nranks_; ++p)
    {
        if (p == rank_) continue;
        if (send_count[p]) check(nccl().Send(send[p], send_count[p], ncclFloat64, p, c, s), "ncclSend");
        if (recv_count[p]) check(nccl().Recv(recv[p], recv_count[p], ncclFloat64, p, c, s), "ncclRecv");
    }
    check(nccl().GroupEnd(), "ncclGroupEnd");
}

void communicator_t::all_gather(const double* send, double* recv, std::size_t count, void* cuda_stream)
{
    check(nccl().AllGather(send, recv, count, ncclFloat64, static_cast<ncclComm_t>(comm), static_cast<cudaStream_t>(cuda_stream)), "ncclAllGather");
}
