// comm.h — the communicator of the hot path's one exchange step (the 30-double ICP system per iteration).
// Two transports: NCCL (dlopen at run time; no link-time dependency) and, when every rank can map every
// other rank's mailbox (CUDA IPC over NVLink / NVSwitch peer memory), a direct in-kernel exchange.
#pragma once
#include <cuda_runtime.h>

#include "../../include/open3d_b200.h"

namespace o3db {

static constexpr int kMaxPeers = 16;
static constexpr int kBoxDoubles = 64;           // one 512-byte slot: 30 sums x 2 words (32 data bits : 32-bit sequence tag)

// Device-visible view of a communicator's mailboxes.  box[p] is rank p's mailbox as mapped into THIS
// process: kBoxDoubles doubles per (parity, writer) slot, [2][world] slots.
struct PeerView {
    double* box[kMaxPeers];
    unsigned long long* seq;    // this rank's collective counter (device memory)
    int rank, world;
};

}  // namespace o3db

// peer view of a communicator, or nullptr when the in-kernel exchange is unavailable (NCCL is used then)
const o3db::PeerView* o3db_comm_peer_view(const o3db_comm* comm);
