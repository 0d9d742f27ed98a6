// comm.cuh — NCCL plumbing (one process per GPU).  libnccl is dlopen()ed so the library has no link-time
// dependency on it and shares the copy a host framework (torch.distributed) may already have loaded.
#pragma once
#include "common.cuh"
namespace dpr {
int comm_allreduce_sum_f64(double* buf, size_t count);   // in place, on ctx().stream
int comm_allreduce_sum_i64(long long* buf, size_t count);

// Peer-memory exchange of the per-iteration tables (NVLink, no NCCL on the path): every rank owns a mailbox in its own HBM,
//   tables [2 parities][kMaxPeers sources][kPeerDoubles], flags [2][kMaxPeers][kPeerFlags] (sequence numbers),
// opened by every other rank of the node through CUDA IPC.  reduce_finalize_kernel writes a rank's table into all mailboxes,
// raises the flags and reads the others' tables out of its own mailbox (centres.cu).
constexpr int kMaxPeers = 8;
constexpr size_t kPeerDoubles = (size_t)1 << 19;   // 4 MB per (parity, source): problems x (k*d + k + 1) of one exchange must fit
constexpr int kPeerFlags = 4096;                   // problems per exchange
struct PeerExchange {
    int32_t world, rank;
    unsigned long long seq;          // number of this exchange (the same on every rank); parity = seq & 1
    double* tables[kMaxPeers];       // tables[q]: rank q's mailbox as seen from this rank
    unsigned long long* flags[kMaxPeers];
    int32_t* failed;                 // set when a peer's table did not arrive in time (the host then fails loudly)
};
bool peer_exchange_ready();                       // dpr_comm_peer_attach succeeded for the current communicator
int peer_exchange_next(PeerExchange* px);         // parameters of the next exchange (advances the sequence number)
int peer_exchange_check();                        // fails when an exchange timed out (synchronises)
}
