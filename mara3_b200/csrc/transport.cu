/**
 * transport.cu -- how the ranks of a multi-GPU run talk: set-up of the peer-memory mailboxes (CUDA IPC over NVLink; NCCL carries
 * the handles once and stays as the fallback transport), the guard-zone exchange kernels of the paths that do not do it inside the
 * stage kernel (nested trees, conserve_linear_p = 0, M3B_STAGE=strip, M3B_FUSED_EXCHANGE=0), and the gathers the products use.
 * The device side of extend() across GPUs (Mara3 src/subprog_binary_scheme.cpp:132-142, which reads its neighbours through shared memory).
 */
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>
#include "device_solver_impl.cuh"

using namespace m3b;
using namespace m3b::dev;

namespace
{
    /** Gather (pack = 1) the listed strips of U into the send buffer, or scatter (pack = 0) the receive
     *  buffer into the ghost blocks: the device side of extend() across GPUs (scheme.cpp:132-142). */
    __global__ void __launch_bounds__(128) halo_copy(const halo_entry_dev_t* __restrict__ entries, double* __restrict__ U, size_t FS, int N,
        double* __restrict__ buffer, int pack)
    {
        const halo_entry_dev_t e = entries[blockIdx.x];
        const int cells = e.ni * e.nj;

        for (int k = threadIdx.x; k < 3 * cells; k += blockDim.x)
        {
            int q = k / cells, c = k % cells;
            int i = e.i0 + c / e.nj, j = e.j0 + c % e.nj;
            size_t u = q * FS + (size_t(e.block) * N + i) * N + j;
            if (pack) buffer[e.offset + k] = U[u]; else U[u] = buffer[e.offset + k];
        }
    }

    /**
     * Send side of extend() across GPUs (scheme.cpp:132-142): each CTA copies one strip / corner of an owned
     * block straight into the destination rank's landing buffer (entry.pad = destination rank, entry.offset =
     * position in ITS buffer); the last CTA to finish raises this rank's flag on every destination.
     */
    __global__ void __launch_bounds__(128) halo_push(const halo_entry_dev_t* __restrict__ entries, const double* __restrict__ U, size_t FS, int N,
        peer_table_t peers, int parity, int me, unsigned int dest_mask, unsigned long long counter, int* ticket)
    {
        __shared__ int is_last;
        const halo_entry_dev_t e = entries[blockIdx.x];
        const int cells = e.ni * e.nj;
        double* __restrict__ dst = peers.recv[e.pad][parity] + e.offset;

        for (int k = threadIdx.x; k < 3 * cells; k += blockDim.x)
        {
            int q = k / cells, c = k % cells;
            int i = e.i0 + c / e.nj, j = e.j0 + c % e.nj;
            dst[k] = U[q * FS + (size_t(e.block) * N + i) * N + j];
        }
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0) is_last = atomicAdd(ticket, 1) == int(gridDim.x) - 1;
        __syncthreads();
        if (! is_last) return;
        __threadfence_system();
        if (threadIdx.x < MAX_PEERS && ((dest_mask >> threadIdx.x) & 1u)) store_release_sys(peers.halo_flag[threadIdx.x] + me, counter);
        if (threadIdx.x == 0) *ticket = 0;
    }

    /** Receive side: wait for the source rank's flag (entry.pad = source rank), then scatter its strip into the ghost block. */
    __global__ void __launch_bounds__(128) halo_wait_unpack(const halo_entry_dev_t* __restrict__ entries, double* __restrict__ U, size_t FS, int N,
        const double* __restrict__ landing, const unsigned long long* flags, unsigned long long counter, int* ticket, unsigned long long* ready,
        peer_table_t peers, int me)
    {
        __shared__ int is_last;
        const halo_entry_dev_t e = entries[blockIdx.x];
        if (threadIdx.x == 0) bounded_wait_sys(flags + e.pad, counter, peers, me, e.pad);
        __syncthreads();
        const int cells = e.ni * e.nj;

        for (int k = threadIdx.x; k < 3 * cells; k += blockDim.x)
        {
            int q = k / cells, c = k % cells;
            int i = e.i0 + c / e.nj, j = e.j0 + c % e.nj;
            U[q * FS + (size_t(e.block) * N + i) * N + j] = __ldcg(landing + e.offset + k);     // written by a peer: not through L1
        }
        // the last CTA tells the stage kernel's boundary tiles that every ghost block is in place
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) is_last = atomicAdd(ticket, 1) == int(gridDim.x) - 1;
        __syncthreads();
        if (is_last && threadIdx.x == 0)
        {
            *ticket = 0;
            asm volatile("st.release.gpu.global.u64 [%0], %1;" :: "l"(ready), "l"(counter) : "memory");
        }
    }

}

/** Collective over the ranks: the owned blocks' data (`d_local`: [owned block][doubles_per_block], device memory) of every
 *  rank, concatenated in rank (= global Morton) order into `host_all` on rank 0.  One rank: a plain download. */
void device_solver_t::gather_blocks(const double* d_local, std::size_t doubles_per_block, double* host_all)
{
    M3B_CUDA(cudaSetDevice(device_id));
    auto s = cudaStream_t(stream_);
    if (num_ranks == 1)
    {
        M3B_CUDA(cudaMemcpyAsync(host_all, d_local, size_t(BO) * doubles_per_block * sizeof(double), cudaMemcpyDeviceToHost, s));
        M3B_CUDA(cudaStreamSynchronize(s));
        return;
    }
    const int total = impl->num_global_blocks;
    auto offsets = std::vector<size_t>(num_ranks + 1);
    for (int r = 0; r <= num_ranks; ++r) offsets[r] = size_t((long(total) * r) / num_ranks);       // partition_offsets
    auto send = std::vector<const double*>(num_ranks, nullptr);
    auto recv = std::vector<double*>(num_ranks, nullptr);
    auto send_count = std::vector<size_t>(num_ranks, 0), recv_count = std::vector<size_t>(num_ranks, 0);
    double* d_all = nullptr;

    if (rank_ == 0)
    {
        M3B_CUDA(cudaMalloc(&d_all, std::max<size_t>(1, size_t(total - BO)) * doubles_per_block * sizeof(double)));
        for (int p = 1; p < num_ranks; ++p)
        {
            recv[p] = d_all + (offsets[p] - offsets[1]) * doubles_per_block;
            recv_count[p] = (offsets[p + 1] - offsets[p]) * doubles_per_block;
        }
    }
    else { send[0] = d_local; send_count[0] = size_t(BO) * doubles_per_block; }
    impl->comm->exchange(send, send_count, recv, recv_count, stream_);

    if (rank_ == 0)
    {
        M3B_CUDA(cudaMemcpyAsync(host_all, d_local, size_t(BO) * doubles_per_block * sizeof(double), cudaMemcpyDeviceToHost, s));
        M3B_CUDA(cudaMemcpyAsync(host_all + size_t(BO) * doubles_per_block, d_all, size_t(total - BO) * doubles_per_block * sizeof(double), cudaMemcpyDeviceToHost, s));
    }
    M3B_CUDA(cudaStreamSynchronize(s));
    if (d_all) M3B_CUDA(cudaFree(d_all));
}



void device_solver_t::set_communicator(communicator_t* comm)
{
    impl->comm = comm;
    if (! comm || num_ranks == 1) return;
    const char* transport = std::getenv("M3B_TRANSPORT");
    if (transport && std::string(transport) == "nccl") return;
    if (num_ranks > MAX_PEERS) return;

    // ---- peer-memory transport: every rank maps every other rank's mailbox (CUDA IPC over NVLink); the handles and
    // the landing-buffer layouts travel once through NCCL.  Any failure leaves the NCCL send / recv path in place.
    M3B_CUDA(cudaSetDevice(device_id));
    auto s = cudaStream_t(stream_);
    const size_t flags_bytes = 1024, results_bytes = (size_t(num_ranks) * num_slots * sizeof(stage_result_t) + 255) / 256 * 256;
    const size_t landing = (std::max<size_t>(1, impl->recv_total) * sizeof(double) + 255) / 256 * 256;
    const size_t bytes = flags_bytes + results_bytes + 2 * landing;
    M3B_CUDA(cudaMalloc(&impl->mailbox, bytes));
    M3B_CUDA(cudaMemset(impl->mailbox, 0, bytes));
    M3B_CUDA(cudaMalloc(&impl->d_push_ticket, 2 * sizeof(int)));
    M3B_CUDA(cudaMemset(impl->d_push_ticket, 0, 2 * sizeof(int)));
    M3B_CUDA(cudaMalloc(&impl->d_ready, sizeof(unsigned long long)));
    M3B_CUDA(cudaMemset(impl->d_ready, 0, sizeof(unsigned long long)));

    cudaIpcMemHandle_t mine;
    bool ok = cudaIpcGetMemHandle(&mine, impl->mailbox) == cudaSuccess;
    if (! ok) cudaGetLastError();

    // per rank: [ok flag][64-byte handle as 8 doubles][recv_starts of every source rank]
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t size");
    const size_t words = 1 + 8 + size_t(num_ranks);
    auto send = std::vector<double>(words, 0.0);
    send[0] = ok ? 1.0 : 0.0;
    std::memcpy(&send[1], &mine, sizeof(mine));
    for (int p = 0; p < num_ranks; ++p) send[9 + p] = double(impl->recv_starts_host[p]);
    double* d = nullptr;
    M3B_CUDA(cudaMalloc(&d, (1 + size_t(num_ranks)) * words * sizeof(double)));
    M3B_CUDA(cudaMemcpyAsync(d, send.data(), words * sizeof(double), cudaMemcpyHostToDevice, s));
    comm->all_gather(d, d + words, words, stream_);
    auto all = std::vector<double>(size_t(num_ranks) * words);
    M3B_CUDA(cudaMemcpyAsync(all.data(), d + words, all.size() * sizeof(double), cudaMemcpyDeviceToHost, s));
    M3B_CUDA(cudaStreamSynchronize(s));
    M3B_CUDA(cudaFree(d));

    for (int p = 0; p < num_ranks; ++p) ok = ok && all[size_t(p) * words] == 1.0;
    impl->peer_mailbox.assign(num_ranks, nullptr);
    for (int p = 0; p < num_ranks && ok; ++p)
    {
        if (p == rank_) continue;
        cudaIpcMemHandle_t h;
        std::memcpy(&h, &all[size_t(p) * words + 1], sizeof(h));
        if (cudaIpcOpenMemHandle(&impl->peer_mailbox[p], h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = false; }
    }
    // everybody must agree, or the ranks would wait for each other on different transports
    auto agree = all_gather_scalar(ok ? 1.0 : 0.0);
    for (double a : agree) ok = ok && a == 1.0;
    if (! ok) return;

    for (int p = 0; p < num_ranks; ++p)
    {
        char* base = static_cast<char*>(p == rank_ ? impl->mailbox : impl->peer_mailbox[p]);
        // the landing buffers of rank p have ITS size: only offsets below its recv_total are ever addressed, and the
        // second buffer starts where rank p says -- which this rank learns from p's own layout words
        impl->peers.halo_flag[p]   = reinterpret_cast<unsigned long long*>(base);
        impl->peers.result_flag[p] = reinterpret_cast<unsigned long long*>(base + 512);
        impl->peers.abort_word[p]  = reinterpret_cast<unsigned long long*>(base + 256);
        impl->peers.results[p]     = reinterpret_cast<stage_result_t*>(base + flags_bytes);
        impl->peers.recv[p][0]     = reinterpret_cast<double*>(base + flags_bytes + results_bytes);
        impl->peers.recv[p][1]     = nullptr;       // set below from rank p's landing size
    }
    // landing size of every rank: second all-gather (one double per rank)
    auto sizes = all_gather_scalar(double(landing));
    for (int p = 0; p < num_ranks; ++p)
        impl->peers.recv[p][1] = reinterpret_cast<double*>(reinterpret_cast<char*>(impl->peers.recv[p][0]) + size_t(sizes[p]));

    // send entries re-addressed into the destination's landing buffer: destination p expects this rank's strips at its recv_starts[me]
    auto push = impl->send_entries_host;
    impl->dest_mask = 0;
    for (auto& e : push)
    {
        const int p = e.pad;
        const size_t remote_start = size_t(all[size_t(p) * words + 9 + rank_]);
        e.offset = remote_start + (e.offset - impl->send_starts_host[p]);
        impl->dest_mask |= 1u << p;
    }
    if (! push.empty())
    {
        M3B_CUDA(cudaMalloc(&impl->d_push_entries, push.size() * sizeof(halo_entry_dev_t)));
        M3B_CUDA(cudaMemcpy(impl->d_push_entries, push.data(), push.size() * sizeof(halo_entry_dev_t), cudaMemcpyHostToDevice));
    }
    {
        // how long a rank waits for a peer before it calls the run off (M3B_SPIN_DEADLINE_MS; default 30 s of SM clocks)
        const char* e = std::getenv("M3B_SPIN_DEADLINE_MS");
        int khz = 1965000;
        cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, device_id);
        impl->peers.deadline_cycles = static_cast<long long>((e ? std::atof(e) : 30000.0) * khz);
    }
    impl->peer_transport = true;

    // With at least two full generations of interior tiles ahead of them, the boundary tiles can wait for their guard zones
    // inside the stage kernel while the unpack runs beside it: the exchange latency disappears behind the interior update.
    // (With less interior work the stage kernel could fill every SM with waiting CTAs before the unpack is resident.)
    const int tpb = impl->tile_x ? (N / impl->tile_x) * (N / impl->tile_y) : 0;
    impl->in_kernel_wait = impl->strip && impl->irregular.empty() && impl->num_interior * tpb >= 2 * impl->sm_count * 4;
    if (const char* e = std::getenv("M3B_IN_KERNEL_WAIT")) impl->in_kernel_wait = impl->in_kernel_wait && std::atoi(e) != 0;
}

int device_solver_t::exchange_transport() const
{
    return num_ranks <= 1 ? 0 : (impl->peer_transport ? 2 : 1);
}

std::uint64_t device_solver_t::halo_bytes_per_exchange() const
{
    return impl->halo_bytes_per_exchange;
}

void device_solver_t::exchange_halos(device_field_t& field)
{
    if (num_ranks == 1) return;
    exchange_on(stream_, field);
}

void device_solver_t::exchange_on(void* cuda_stream, device_field_t& field)
{
    if (! impl->comm) throw std::logic_error("exchange_halos: no communicator set");
    auto s = cudaStream_t(cuda_stream);
    M3B_CUDA(cudaSetDevice(device_id));

    if (impl->peer_transport)
    {
        // strips go straight into the neighbours' landing buffers over NVLink; the receive kernel waits on their flags.
        // With in-kernel waiting both kernels run on the exchange stream, beside the stage kernel that follows on `s`
        // (its boundary tiles poll d_ready), so none of the exchange sits on the compute stream.
        const unsigned long long counter = ++impl->exchange_counter;
        const int parity = int(counter & 1);
        auto u = s;
        if (impl->in_kernel_wait && impl->defer_unpack)
        {
            M3B_CUDA(cudaEventRecord(impl->input_ready, s));
            M3B_CUDA(cudaStreamWaitEvent(impl->comm_stream, impl->input_ready, 0));
            u = impl->comm_stream;
        }
        if (impl->num_send_entries)
        {
            halo_push<<<impl->num_send_entries, 128, 0, u>>>(impl->d_push_entries, field.data, cells, N, impl->peers, parity, rank_,
                impl->dest_mask, counter, impl->d_push_ticket);
            ++launches;
        }
        if (impl->num_recv_entries)
        {
            halo_wait_unpack<<<impl->num_recv_entries, 128, 0, u>>>(impl->d_recv_entries, field.data, cells, N, impl->peers.recv[rank_][parity],
                impl->peers.halo_flag[rank_], counter, impl->d_push_ticket + 1, impl->d_ready, impl->peers, rank_);
            ++launches;
        }
        if (u != s) M3B_CUDA(cudaEventRecord(impl->halo_ready, u));
        M3B_CUDA(cudaGetLastError());
        return;
    }
    if (impl->num_send_entries)
    {
        halo_copy<<<impl->num_send_entries, 128, 0, s>>>(impl->d_send_entries, field.data, cells, N, impl->d_send_buffer, 1);
        ++launches;
    }
    impl->comm->exchange(impl->send_ptr, impl->send_count, impl->recv_ptr, impl->recv_count, cuda_stream);

    if (impl->num_recv_entries)
    {
        halo_copy<<<impl->num_recv_entries, 128, 0, s>>>(impl->d_recv_entries, field.data, cells, N, impl->d_recv_buffer, 0);
        ++launches;
    }
    M3B_CUDA(cudaGetLastError());
}

std::vector<double> device_solver_t::all_gather_scalar(double value)
{
    auto out = std::vector<double>(num_ranks, value);
    if (num_ranks == 1) return out;
    auto s = cudaStream_t(stream_);
    double* d = reinterpret_cast<double*>(impl->d_results_all);     // scratch: large enough, idle between steps
    M3B_CUDA(cudaMemcpyAsync(d + num_ranks, &value, sizeof(double), cudaMemcpyHostToDevice, s));
    impl->comm->all_gather(d + num_ranks, d, 1, stream_);
    M3B_CUDA(cudaMemcpyAsync(out.data(), d, num_ranks * sizeof(double), cudaMemcpyDeviceToHost, s));
    M3B_CUDA(cudaStreamSynchronize(s));
    return out;
}

void device_solver_t::gather_results()
{
    if (num_ranks == 1) return;
    auto s = cudaStream_t(stream_);
    const size_t doubles = num_slots * sizeof(stage_result_t) / sizeof(double);
    impl->comm->all_gather(reinterpret_cast<const double*>(impl->d_results_local), reinterpret_cast<double*>(impl->d_results_all), doubles, stream_);
    M3B_CUDA(cudaMemcpyAsync(impl->h_results_all, impl->d_results_all, size_t(num_ranks) * num_slots * sizeof(stage_result_t), cudaMemcpyDeviceToHost, s));
}
