// NCCL communicator of a context (buoys sharded over ranks); see comm.cu.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstring>
#include <string>

namespace ocp {

bool comm_available(std::string &err);                 // libnccl could be resolved in this process
int comm_nccl_version();                               // 0 when unavailable
bool comm_unique_id(void *id128, std::string &err);    // ncclGetUniqueId (rank 0 calls it, the host distributes it)

class Communicator {
  public:
    Communicator() = default;
    ~Communicator();
    Communicator(const Communicator &) = delete;
    Communicator &operator=(const Communicator &) = delete;
    // collective: every rank of the job calls it with the same id (blocks until all have joined)
    bool init(int nranks, int rank, const void *id128, std::string &err);
    void destroy();
    int size() const { return nranks_; }
    int rank() const { return rank_; }
    // in place, on stream s; no-ops for a single rank
    bool allreduce_sum(double *d_buf, size_t n, cudaStream_t s, std::string &err);
    bool allreduce_sum_i64(long long *d_buf, size_t n, cudaStream_t s, std::string &err);

  private:
    void *comm_ = nullptr;
    int nranks_ = 1, rank_ = 0;
};

}  // namespace ocp
