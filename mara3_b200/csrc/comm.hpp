/**
 * comm.hpp -- rank-to-rank communication for the multi-GPU path: one process per GPU, NCCL over
 * NVLink / NVSwitch.  The reference has no distributed path (its MPI wrapper is never used by
 * `binary`); this replaces the shared-memory neighbour read of
 * extend() (Mara3 src/subprog_binary_scheme.cpp:132-142) across GPUs.
 *
 * NCCL is bound at run time with dlopen("libnccl.so.2") -- the copy torch has already loaded when
 * the library is driven from Python, the system one otherwise -- so the library has no link-time
 * dependency on it and single-GPU use never touches it.
 */
#pragma once
#include <cstddef>
#include <string>
#include <vector>

namespace m3b
{
    constexpr int nccl_unique_id_bytes = 128;

    class communicator_t
    {
    public:
        /** Fill `out` (128 bytes) with a fresh NCCL unique id; rank 0 calls this and hands it to the others. */
        static void make_unique_id(unsigned char* out);

        /** ncclCommInitRank on the current CUDA device. */
        communicator_t(int rank, int nranks, const unsigned char* unique_id);
        ~communicator_t();
        communicator_t(const communicator_t&) = delete;

        int rank() const { return rank_; }
        int size() const { return nranks_; }

        /** One grouped exchange of doubles: send[p] / recv[p] are device pointers (or null) with counts. */
        void exchange(const std::vector<const double*>& send, const std::vector<std::size_t>& send_count,
                      const std::vector<double*>& recv, const std::vector<std::size_t>& recv_count, void* cuda_stream);

        /** all-gather `count` doubles per rank (device buffers). */
        void all_gather(const double* send, double* recv, std::size_t count, void* cuda_stream);

    private:
        int rank_ = 0, nranks_ = 1;
        void* comm = nullptr;
    };
}
