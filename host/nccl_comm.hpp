// nccl_comm.hpp -- macaw::Comm over NCCL (NVLink 5 / NVSwitch inside one box), one process per GPU.
//
// Replaces the MPI calls of MC OpenMCStudy.C:454-607 / OMC eigenvalue.cpp:71-311 for the
// particle-sharded study: MPI_Exscan + MPI_Bcast of bank sizes -> ncclAllGather of one int64;
// MPI_Isend / MPI_Irecv of SourceSite ranges -> grouped ncclSend / ncclRecv between the library's own
// device buffers; MPI_Allreduce / MPI_Reduce of k and tallies -> one ncclAllReduce (NVLS-capable).
// The ncclUniqueId is handed from rank 0 to the others through a file (no MPI in this image).
#pragma once
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <unistd.h>

#include <cuda_runtime_api.h>
#include <nccl.h>

#include "macaw_host.hpp"

namespace macaw
{

class NcclComm : public Comm
{
public:
  NcclComm(int rank, int nranks, int device, const std::string& id_file) : _rank(rank), _n(nranks)
  {
    cuda(cudaSetDevice(device), "cudaSetDevice");
    ncclUniqueId id;
    // The id file carries a launch nonce next to the id (MCW_LAUNCH_NONCE if set, else the launcher's pid, which the
    // ranks of one torchrun / mpirun-style launch share), so a file left behind by an earlier or crashed run is never
    // mistaken for this launch's; rank 0 also removes any old file first and deletes its own once every rank is in.
    unsigned long long nonce = (unsigned long long)getppid();
    if (const char* env = std::getenv("MCW_LAUNCH_NONCE"))
    {
      nonce = 1469598103934665603ull;
      for (const char* c = env; *c; ++c)
        nonce = (nonce ^ (unsigned char)*c) * 1099511628211ull;
    }
    if (rank == 0)
    {
      std::remove(id_file.c_str());
      nccl(ncclGetUniqueId(&id), "ncclGetUniqueId");
      const std::string tmp = id_file + ".tmp";
      FILE* f = std::fopen(tmp.c_str(), "wb");
      if (!f || std::fwrite(&nonce, sizeof(nonce), 1, f) != 1 || std::fwrite(&id, sizeof(id), 1, f) != 1)
        mooseError("cannot write " + tmp);
      std::fclose(f);
      if (std::rename(tmp.c_str(), id_file.c_str()) != 0)
        mooseError("cannot publish " + id_file);
    }
    else
    {
      bool ok = false;
      for (int tries = 0; tries < 6000 && !ok; ++tries)
      {
        unsigned long long seen = 0;
        if (FILE* f = std::fopen(id_file.c_str(), "rb"))
        {
          ok = std::fread(&seen, sizeof(seen), 1, f) == 1 && seen == nonce && std::fread(&id, sizeof(id), 1, f) == 1;
          std::fclose(f);
        }
        if (!ok)
          std::this_thread::sleep_for(std::chrono::milliseconds(10));
      }
      if (!ok)
        mooseError("rank " + std::to_string(rank) + " never saw this launch's NCCL id in " + id_file);
    }
    nccl(ncclCommInitRank(&_comm, nranks, id, rank), "ncclCommInitRank"); // returns once every rank has joined
    if (rank == 0)
      std::remove(id_file.c_str());
    cuda(cudaMalloc(&_d_counts, sizeof(int64_t) * (size_t)(nranks + 1)), "cudaMalloc");
    cuda(cudaStreamCreateWithFlags(&_stream, cudaStreamNonBlocking), "cudaStreamCreate");
  }
  ~NcclComm() override
  {
    if (_comm)
      ncclCommDestroy(_comm);
    cudaFree(_d_counts);
    cudaStreamDestroy(_stream);
  }
  int rank() const override { return _rank; }
  int size() const override { return _n; }

  std::vector<int64_t> allGather(int64_t v) override
  {
    cuda(cudaMemcpyAsync(_d_counts + _n, &v, sizeof(v), cudaMemcpyHostToDevice, _stream), "H2D");
    nccl(ncclAllGather(_d_counts + _n, _d_counts, 1, ncclInt64, _comm, _stream), "ncclAllGather");
    std::vector<int64_t> out((size_t)_n);
    cuda(cudaMemcpyAsync(out.data(), _d_counts, sizeof(int64_t) * (size_t)_n, cudaMemcpyDeviceToHost, _stream), "D2H");
    cuda(cudaStreamSynchronize(_stream), "sync");
    return out;
  }
  void allReduceSum(double* buf, int64_t n, void* stream) override
  {
    nccl(ncclAllReduce(buf, buf, (size_t)n, ncclDouble, ncclSum, _comm, static_cast<cudaStream_t>(stream)), "ncclAllReduce");
  }
  void exchange(const char* send_base, const std::vector<Range>& sends, char* recv_base, const std::vector<Range>& recvs,
                void* stream) override
  {
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    nccl(ncclGroupStart(), "ncclGroupStart");
    for (const auto& r : sends)
    {
      if (r.peer == _rank)
        continue;
      nccl(ncclSend(send_base + r.offset, (size_t)r.bytes, ncclChar, r.peer, _comm, s), "ncclSend");
    }
    for (const auto& r : recvs)
    {
      if (r.peer == _rank)
        continue;
      nccl(ncclRecv(recv_base + r.offset, (size_t)r.bytes, ncclChar, r.peer, _comm, s), "ncclRecv");
    }
    nccl(ncclGroupEnd(), "ncclGroupEnd");
    // the part this rank keeps: a device-to-device copy on the same stream
    for (const auto& a : sends)
      if (a.peer == _rank)
        for (const auto& b : recvs)
          if (b.peer == _rank)
            cuda(cudaMemcpyAsync(recv_base + b.offset, send_base + a.offset, (size_t)a.bytes, cudaMemcpyDeviceToDevice, s), "D2D");
    cuda(cudaStreamSynchronize(s), "sync");
  }

private:
  static void cuda(cudaError_t e, const char* what)
  {
    if (e != cudaSuccess)
      throw Error(MCW_E_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
  }
  static void nccl(ncclResult_t e, const char* what)
  {
    if (e != ncclSuccess)
      throw Error(MCW_E_CUDA, std::string(what) + ": " + ncclGetErrorString(e));
  }
  int _rank, _n;
  ncclComm_t _comm = nullptr;
  int64_t* _d_counts = nullptr;
  cudaStream_t _stream = nullptr;
};

} // namespace macaw
