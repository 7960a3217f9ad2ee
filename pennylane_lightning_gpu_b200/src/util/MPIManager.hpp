// MPIManager with the interface of the reference's util/MPIManager.hpp (getRank / getSize / getSizeNode / Barrier /
// Bcast / Scatter / getTime / getVendor / getVersion), WITHOUT MPI: one process per GPU is started by torchrun (or
// mpirun / srun), rank and size come from the launcher's environment, and all traffic goes over NCCL through the
// C ABI of libqsv_b200.so.  The only thing NCCL cannot do for itself is hand the 128-byte unique id from rank 0 to
// the others; that takes one TCP connection per rank to MASTER_ADDR (the role MPI_Bcast plays in the reference,
// simulator/MPIWorker.hpp:306-320).
#pragma once
#include <arpa/inet.h>
#include <netdb.h>
#include <netinet/in.h>
#include <sys/socket.h>
#include <sys/time.h>
#include <unistd.h>

#include <array>
#include <cstdint>
#include <algorithm>
#include <chrono>
#include <complex>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <tuple>
#include <vector>

#include "Error.hpp"
#include "qsv_b200.h"

namespace Pennylane::MPI {

class MPIManager {
  public:
    MPIManager() {
        rank_ = env_int({"RANK", "OMPI_COMM_WORLD_RANK", "PMI_RANK", "SLURM_PROCID"}, 0);
        size_ = env_int({"WORLD_SIZE", "OMPI_COMM_WORLD_SIZE", "PMI_SIZE", "SLURM_NTASKS"}, 1);
        size_node_ = env_int({"LOCAL_WORLD_SIZE", "OMPI_COMM_WORLD_LOCAL_SIZE", "SLURM_NTASKS_PER_NODE"}, size_);
        local_rank_ = env_int({"LOCAL_RANK", "OMPI_COMM_WORLD_LOCAL_RANK", "SLURM_LOCALID"}, rank_ % size_node_);
        PL_ABORT_IF(size_ < 1 || rank_ < 0 || rank_ >= size_, "invalid rank / world size in the launcher environment");
        PL_ABORT_IF((size_ & (size_ - 1)) != 0, "Processes number is not power of two.");
        // a 1-qubit control register carries the communicator used for Barrier / Bcast / Scatter
        Util::check(qsv_create(1, QSV_C128, local_rank_, &ctrl_));
        std::array<unsigned char, 128> id{};
        if (rank_ == 0) Util::check(qsv_dist_unique_id(id.data()));
        if (size_ > 1) tcp_bcast(id.data(), id.size());
        Util::check(qsv_dist_init(ctrl_, id.data(), rank_, size_));
    }
    MPIManager(const MPIManager &) = delete;
    MPIManager &operator=(const MPIManager &) = delete;
    ~MPIManager() {
        if (ctrl_) {
            qsv_dist_finalize(ctrl_);
            qsv_destroy(ctrl_);
        }
    }

    int getRank() const { return rank_; }
    int getSize() const { return size_; }
    int getSizeNode() const { return size_node_; }
    int getLocalRank() const { return local_rank_; }
    double getTime() const {
        return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
    }
    std::string getVendor() const { return "NCCL"; }
    std::tuple<int, int, int> getVersion() const {
        int v = 0;
        Util::check(qsv_dist_nccl_version(&v));
        return {v / 10000, (v / 100) % 100, v % 100};
    }
    void Barrier() { Util::check(qsv_dist_barrier(ctrl_)); }
    template <class T> void Bcast(T *data, std::size_t count, int root) {
        Util::check(qsv_dist_bcast_bytes(ctrl_, data, count * sizeof(T), root));
    }
    template <class T> void Bcast(std::vector<T> &data, int root) { Bcast(data.data(), data.size(), root); }
    // rank r receives elements [r * recv_count, (r + 1) * recv_count) of root's send buffer
    template <class T> void Scatter(const T *send, T *recv, std::size_t recv_count, int root) {
        Util::check(qsv_dist_scatter_host(ctrl_, send, recv, recv_count * sizeof(T), root));
    }
    // fresh NCCL unique id for a new communicator, identical on all ranks
    std::array<unsigned char, 128> newUniqueId() {
        std::array<unsigned char, 128> id{};
        if (rank_ == 0) Util::check(qsv_dist_unique_id(id.data()));
        Bcast(id.data(), id.size(), 0);
        return id;
    }

  private:
    static int env_int(std::initializer_list<const char *> names, int dflt) {
        for (const char *n : names)
            if (const char *v = std::getenv(n)) return std::atoi(v);
        return dflt;
    }
    static void send_all(int fd, const unsigned char *p, std::size_t n) {
        while (n) {
            const ssize_t k = ::send(fd, p, n, MSG_NOSIGNAL);
            PL_ABORT_IF(k <= 0, "bootstrap: send failed");
            p += k;
            n -= static_cast<std::size_t>(k);
        }
    }
    static void recv_all(int fd, unsigned char *p, std::size_t n) {
        while (n) {
            const ssize_t k = ::recv(fd, p, n, 0);
            PL_ABORT_IF(k <= 0, "bootstrap: receive failed");
            p += k;
            n -= static_cast<std::size_t>(k);
        }
    }
    // Bootstrap of the NCCL unique id over TCP (no MPI): rank 0 listens on the MASTER_ADDR interface, port
    // QSV_BOOTSTRAP_PORT | MASTER_PORT + 1 (+ instance for further communicators), and serves the buffer to size - 1
    // DISTINCT ranks.  A client first sends a 16-byte hello {magic, job token, rank, instance}; the token is a hash of
    // the launcher's job identity (TORCHELASTIC_RUN_ID / SLURM_JOB_ID / QSV_JOB_TOKEN, MASTER_ADDR, MASTER_PORT, world
    // size), so a stray connection (port scan, health probe) or a rank of another job is dropped and the slot stays
    // free.  accept and recv time out (QSV_BOOTSTRAP_TIMEOUT_S, default 120 s): a dead rank is an error, not a hang.
    static std::uint32_t job_token(int size) {
        std::string key;
        for (const char *n : {"QSV_JOB_TOKEN", "TORCHELASTIC_RUN_ID", "SLURM_JOB_ID", "MASTER_ADDR", "MASTER_PORT"})
            if (const char *v = std::getenv(n)) key += std::string(n) + "=" + v + ";";
        key += "world=" + std::to_string(size);
        std::uint32_t h = 2166136261u;  // FNV-1a
        for (unsigned char c : key) h = (h ^ c) * 16777619u;
        return h;
    }
    static void set_timeouts(int fd, int seconds) {
        timeval tv{};
        tv.tv_sec = seconds;
        ::setsockopt(fd, SOL_SOCKET, SO_RCVTIMEO, &tv, sizeof(tv));
        ::setsockopt(fd, SOL_SOCKET, SO_SNDTIMEO, &tv, sizeof(tv));
    }
    void tcp_bcast(unsigned char *buf, std::size_t n) {
        static int instance = 0;  // communicators are created in the same order on every rank (as with MPI_Comm_dup)
        const int inst = instance++;
        const char *addr = std::getenv("MASTER_ADDR");
        const std::string host = addr ? addr : "127.0.0.1";
        const int port = env_int({"QSV_BOOTSTRAP_PORT"}, env_int({"MASTER_PORT"}, 29400) + 1) + inst;
        const int timeout_s = std::max(1, env_int({"QSV_BOOTSTRAP_TIMEOUT_S"}, 120));
        constexpr std::uint32_t MAGIC = 0x51535642u;  // "QSVB"
        const std::uint32_t token = job_token(size_);
        addrinfo hints{}, *res = nullptr;
        hints.ai_family = AF_INET;
        hints.ai_socktype = SOCK_STREAM;
        PL_ABORT_IF(::getaddrinfo(host.c_str(), std::to_string(port).c_str(), &hints, &res) != 0 || res == nullptr,
                    "bootstrap: cannot resolve " + host);
        if (rank_ == 0) {
            const int srv = ::socket(AF_INET, SOCK_STREAM, 0);
            PL_ABORT_IF(srv < 0, "bootstrap: cannot create a socket");
            const int one = 1;
            ::setsockopt(srv, SOL_SOCKET, SO_REUSEADDR, &one, sizeof(one));
            // the MASTER_ADDR interface only (loopback for a single-node torchrun), not INADDR_ANY
            int rc = ::bind(srv, res->ai_addr, res->ai_addrlen);
            ::freeaddrinfo(res);
            PL_ABORT_IF(rc != 0, "bootstrap: cannot bind " + host + ":" + std::to_string(port));
            PL_ABORT_IF(::listen(srv, size_ + 8) != 0, "bootstrap: listen failed");
            set_timeouts(srv, timeout_s);  // SO_RCVTIMEO bounds accept()
            std::vector<bool> served(size_, false);
            int left = size_ - 1;
            const auto deadline = std::chrono::steady_clock::now() + std::chrono::seconds(timeout_s);
            while (left > 0) {
                PL_ABORT_IF(std::chrono::steady_clock::now() > deadline,
                            "bootstrap: " + std::to_string(left) + " rank(s) did not connect within " +
                                std::to_string(timeout_s) + " s");
                const int fd = ::accept(srv, nullptr, nullptr);
                if (fd < 0) continue;  // timeout or interrupted: the deadline check above decides
                set_timeouts(fd, 5);
                std::uint32_t hello[4] = {0, 0, 0, 0};
                std::size_t got = 0;
                while (got < sizeof(hello)) {
                    const ssize_t k = ::recv(fd, reinterpret_cast<unsigned char *>(hello) + got, sizeof(hello) - got, 0);
                    if (k <= 0) break;
                    got += static_cast<std::size_t>(k);
                }
                const bool ok = got == sizeof(hello) && hello[0] == MAGIC && hello[1] == token &&
                                hello[2] >= 1u && hello[2] < static_cast<std::uint32_t>(size_) &&
                                hello[3] == static_cast<std::uint32_t>(inst) && !served[hello[2]];
                if (ok) {
                    bool sent = true;
                    const unsigned char *p = buf;
                    std::size_t todo = n;
                    while (todo && sent) {
                        const ssize_t k = ::send(fd, p, todo, MSG_NOSIGNAL);
                        if (k <= 0) sent = false;
                        else { p += k; todo -= static_cast<std::size_t>(k); }
                    }
                    if (sent) {
                        served[hello[2]] = true;
                        --left;
                    }
                }
                ::close(fd);  // an invalid or duplicate peer is simply dropped; its slot stays free
            }
            ::close(srv);
            return;
        }
        int fd = -1;
        for (int attempt = 0; attempt < timeout_s * 10; ++attempt) {  // wait for rank 0 to come up
            fd = ::socket(AF_INET, SOCK_STREAM, 0);
            if (fd >= 0 && ::connect(fd, res->ai_addr, res->ai_addrlen) == 0) break;
            if (fd >= 0) ::close(fd);
            fd = -1;
            std::this_thread::sleep_for(std::chrono::milliseconds(100));
        }
        ::freeaddrinfo(res);
        PL_ABORT_IF(fd < 0, "bootstrap: cannot reach rank 0 at " + host + ":" + std::to_string(port));
        set_timeouts(fd, timeout_s);
        const std::uint32_t hello[4] = {MAGIC, token, static_cast<std::uint32_t>(rank_), static_cast<std::uint32_t>(inst)};
        send_all(fd, reinterpret_cast<const unsigned char *>(hello), sizeof(hello));
        recv_all(fd, buf, n);
        ::close(fd);
    }

    int rank_{0}, size_{1}, size_node_{1}, local_rank_{0};
    qsv_state *ctrl_{nullptr};
};

}  // namespace Pennylane::MPI
